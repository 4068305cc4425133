#!/bin/bash
# Full GPU validation + bench + ncu evidence (one gpurun call)
cd oracle && make -s && cd ..
(timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15) > gpurun_out/gpu_tests.log; cat gpurun_out/gpu_tests.log
(timeout 900 python bench.py 2> gpurun_out/bench_err.log | tail -1) > gpurun_out/bench_a.json; cat gpurun_out/bench_a.json | cut -c1-600; tail -3 gpurun_out/bench_err.log
(timeout 600 python bench.py --mode b --no-cpu 2> gpurun_out/bench_b_err.log | tail -1) > gpurun_out/bench_b.json; cat gpurun_out/bench_b.json | cut -c1-300
timeout 300 python bench.py --steps 1 --warmup 1 --games 256 --no-cpu > gpurun_out/plain_small.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --games 256 --no-cpu > gpurun_out/ncu_small.log 2>&1
tail -2 gpurun_out/ncu_small.log | cut -c1-200
timeout 120 python tools/bench_tower.py 8192 2 > gpurun_out/plain_tower.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv3x3_pair -s 5 -c 2 -o gpurun_out/prof_conv_pair \
    python tools/bench_tower.py 8192 2 > gpurun_out/ncu_tower.log 2>&1
tail -3 gpurun_out/ncu_tower.log | cut -c1-200
