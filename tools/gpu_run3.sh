#!/bin/bash
# GPU session 3: smoke, first full bench line, ncu launch list + full capture of the conv kernel
cd oracle && make -s && cd ..
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5) > gpurun_out/smoke.log; cat gpurun_out/smoke.log
(timeout 900 python bench.py 2> gpurun_out/bench_err.log | tail -1) > gpurun_out/bench_r1_a.json; cat gpurun_out/bench_r1_a.json; tail -5 gpurun_out/bench_err.log
(timeout 600 python bench.py --mode b --no-cpu 2> gpurun_out/bench_b_err.log | tail -1) > gpurun_out/bench_r1_b.json; cat gpurun_out/bench_r1_b.json; tail -5 gpurun_out/bench_b_err.log
timeout 300 python bench.py --steps 1 --warmup 1 --games 256 --no-cpu > gpurun_out/plain_small.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 1 --warmup 1 --games 256 --no-cpu > gpurun_out/ncu_small.log 2>&1
tail -2 gpurun_out/ncu_small.log
timeout 120 python tools/bench_tower.py 8192 2 > gpurun_out/plain_tower.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv3x3_tc -s 4 -c 2 -o gpurun_out/prof_conv_r1 \
    python tools/bench_tower.py 8192 2 > gpurun_out/ncu_tower.log 2>&1
tail -3 gpurun_out/ncu_tower.log
ls -la gpurun_out/
