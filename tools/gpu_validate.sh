#!/bin/bash
# One gpurun call: full GPU validation and the round's bench lines (ncu evidence: tools/gpu_r02_profiles.sh; N GPUs: tools/gpu_r02_multi.sh).
#   gpurun --timeout 3000 -- 'bash tools/gpu_validate.sh <tag>'
T=${1:-x}
cd oracle && make -s && cd ..
(timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 2>&1 | tail -15) > gpurun_out/tests_$T.log; cat gpurun_out/tests_$T.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3) > gpurun_out/smoke_$T.log; cat gpurun_out/smoke_$T.log
# the driver's line: mode A + other_modes (mode B, match) + cpu_baseline from a child process
(timeout 1500 python bench.py --steps 6 --warmup 3 2> gpurun_out/bench_${T}_err.log | tail -1) > gpurun_out/bench_$T.json; cut -c1-400 gpurun_out/bench_$T.json; tail -3 gpurun_out/bench_${T}_err.log
(timeout 600 python bench.py --impl reference --steps 4 --warmup 1 2> gpurun_out/bench_${T}_ref_err.log | tail -1) > gpurun_out/bench_${T}_ref.json; cut -c1-300 gpurun_out/bench_${T}_ref.json
# whole games with continuous slot turnover, both modes
for m in a b; do (timeout 1100 python bench.py --full-games --mode $m --games 1024 --total 3072 --num-moves 30 2>/dev/null | tail -1) > gpurun_out/full_games_${T}_$m.json; cut -c1-400 gpurun_out/full_games_${T}_$m.json; done
du -sh gpurun_out
