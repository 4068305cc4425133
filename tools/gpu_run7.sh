#!/bin/bash
# GPU session 7: full validation + the round's bench lines + ncu evidence with the real-row conv tiling
cd oracle && make -s && cd ..
(timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15) > gpurun_out/t7_all.log; cat gpurun_out/t7_all.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3) > gpurun_out/smoke7.log; cat gpurun_out/smoke7.log
(timeout 900 python bench.py 2> gpurun_out/bench7_err.log | tail -1) > gpurun_out/bench7_a.json; cut -c1-400 gpurun_out/bench7_a.json; tail -3 gpurun_out/bench7_err.log
(timeout 600 python bench.py --mode b --no-cpu 2> gpurun_out/bench7_b_err.log | tail -1) > gpurun_out/bench7_b.json; cut -c1-300 gpurun_out/bench7_b.json
(timeout 900 python bench.py --match --no-cpu 2> gpurun_out/bench7_match_err.log | tail -1) > gpurun_out/bench7_match.json; cut -c1-300 gpurun_out/bench7_match.json
(timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/bench7_ref_err.log | tail -1) > gpurun_out/bench7_ref.json; cut -c1-300 gpurun_out/bench7_ref.json
timeout 300 python bench.py --steps 1 --warmup 1 --games 256 --no-cpu > gpurun_out/plain_small7.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches7.csv \
    python bench.py --steps 1 --warmup 1 --games 256 --no-cpu > gpurun_out/ncu_small7.log 2>&1
tail -2 gpurun_out/ncu_small7.log | cut -c1-200
timeout 120 python tools/bench_tower.py 8192 2 > gpurun_out/plain_tower7.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv3x3_pair -s 5 -c 2 -o gpurun_out/prof_conv_rr \
    python tools/bench_tower.py 8192 2 > gpurun_out/ncu_tower7.log 2>&1
tail -3 gpurun_out/ncu_tower7.log | cut -c1-200
ncu -i gpurun_out/prof_conv_rr.ncu-rep --page raw --csv > gpurun_out/prof_conv_rr_raw.csv 2>/dev/null
python tools/ncu_summary.py < gpurun_out/prof_conv_rr_raw.csv
ls -la gpurun_out/; du -sh gpurun_out
