#!/bin/bash
# GPU session 9: k order (channel chunk outer) + next-tile L2 prefetch: ablations, correctness, sustained forward A/B
timeout 300 python tools/conv_variants.py 8192 > gpurun_out/conv_variants9.json 2> gpurun_out/conv_variants9_err.log; cat gpurun_out/conv_variants9.json; tail -3 gpurun_out/conv_variants9_err.log
(timeout 900 python -m pytest tests/test_gpu_tower.py tests/test_gpu_driver.py -q --timeout 600 2>&1 | tail -5) > gpurun_out/t9_tower.log; cat gpurun_out/t9_tower.log
timeout 300 python tools/bench_tower.py 16384 20 > gpurun_out/tower_bench9_new.log 2>&1; tail -1 gpurun_out/tower_bench9_new.log | cut -c1-500
SGO_CONV_DEBUG=24 timeout 300 python tools/bench_tower.py 16384 20 > gpurun_out/tower_bench9_old.log 2>&1; tail -1 gpurun_out/tower_bench9_old.log | cut -c1-500
SGO_CONV_DEBUG=16 timeout 300 python tools/bench_tower.py 16384 20 > gpurun_out/tower_bench9_nopf.log 2>&1; tail -1 gpurun_out/tower_bench9_nopf.log | cut -c1-500
SGO_CONV_DEBUG=8 timeout 300 python tools/bench_tower.py 16384 20 > gpurun_out/tower_bench9_tapmajor.log 2>&1; tail -1 gpurun_out/tower_bench9_tapmajor.log | cut -c1-500
timeout 300 python tools/bench_tower.py 16384 20 > gpurun_out/tower_bench9_new2.log 2>&1; tail -1 gpurun_out/tower_bench9_new2.log | cut -c1-500
