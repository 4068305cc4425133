#!/bin/bash
# GPU session 11: same-box A/B: shifted-GEMM conv (new) vs real-row tiles with per-tap loads (commit 9b1925b, libsejonggo_b200_rowtile.so)
OLD=$PWD/sejonggo_b200/lib/libsejonggo_b200_rowtile.so
(timeout 900 python -m pytest tests/test_gpu_tower.py tests/test_gpu_driver.py -q --timeout 600 2>&1 | tail -4) > gpurun_out/t11_tower.log; cat gpurun_out/t11_tower.log
for i in 1 2; do
  timeout 300 python tools/bench_tower.py 16384 20 > gpurun_out/tower_bench11_new$i.log 2>&1; tail -1 gpurun_out/tower_bench11_new$i.log | cut -c1-420
  SGO_LIBRARY=$OLD timeout 300 python tools/bench_tower.py 16384 20 > gpurun_out/tower_bench11_old$i.log 2>&1; tail -1 gpurun_out/tower_bench11_old$i.log | cut -c1-420
done
(timeout 600 python bench.py --no-cpu 2> gpurun_out/bench11_new_err.log | tail -1) > gpurun_out/bench11_new.json; cut -c1-160 gpurun_out/bench11_new.json; tail -2 gpurun_out/bench11_new_err.log
(SGO_LIBRARY=$OLD timeout 600 python bench.py --no-cpu 2> gpurun_out/bench11_old_err.log | tail -1) > gpurun_out/bench11_old.json; cut -c1-160 gpurun_out/bench11_old.json; tail -2 gpurun_out/bench11_old_err.log
