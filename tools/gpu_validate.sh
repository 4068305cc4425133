#!/bin/bash
# One gpurun call: full GPU validation, the round's bench lines and the ncu evidence (profiles/ is filled from gpurun_out/ by hand).
#   gpurun --timeout 3600 -- 'bash tools/gpu_validate.sh <tag>'
T=${1:-x}
cd oracle && make -s && cd ..
(timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15) > gpurun_out/tests_$T.log; cat gpurun_out/tests_$T.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3) > gpurun_out/smoke_$T.log; cat gpurun_out/smoke_$T.log
(timeout 900 python bench.py 2> gpurun_out/bench_${T}_a_err.log | tail -1) > gpurun_out/bench_${T}_a.json; cut -c1-300 gpurun_out/bench_${T}_a.json; tail -3 gpurun_out/bench_${T}_a_err.log
(timeout 600 python bench.py --mode b --no-cpu 2> gpurun_out/bench_${T}_b_err.log | tail -1) > gpurun_out/bench_${T}_b.json; cut -c1-200 gpurun_out/bench_${T}_b.json
(timeout 900 python bench.py --match --no-cpu 2> gpurun_out/bench_${T}_match_err.log | tail -1) > gpurun_out/bench_${T}_match.json; cut -c1-200 gpurun_out/bench_${T}_match.json
(timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/bench_${T}_ref_err.log | tail -1) > gpurun_out/bench_${T}_ref.json; cut -c1-200 gpurun_out/bench_${T}_ref.json
timeout 300 python bench.py --steps 1 --warmup 1 --games 256 --no-cpu > gpurun_out/plain_small_$T.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$T.csv \
    python bench.py --steps 1 --warmup 1 --games 256 --no-cpu > gpurun_out/ncu_small_$T.log 2>&1
tail -2 gpurun_out/ncu_small_$T.log | cut -c1-200
timeout 120 python tools/bench_tower.py 8192 2 > gpurun_out/plain_tower_$T.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv3x3_pair -s 5 -c 2 -o gpurun_out/prof_conv_$T \
    python tools/bench_tower.py 8192 2 > gpurun_out/ncu_tower_$T.log 2>&1
tail -3 gpurun_out/ncu_tower_$T.log | cut -c1-200
ncu -i gpurun_out/prof_conv_$T.ncu-rep --page raw --csv > gpurun_out/prof_conv_${T}_raw.csv 2>/dev/null
python tools/ncu_summary.py < gpurun_out/prof_conv_${T}_raw.csv
du -sh gpurun_out
