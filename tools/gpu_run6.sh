#!/bin/bash
# GPU session 6: real-row tiling of the pair conv (A/B against padded tiling on the same box); ncu of the board/tree kernels
(timeout 900 python -m pytest tests/test_gpu_tower.py tests/test_gpu_driver.py -q --timeout 600 2>&1 | tail -15) > gpurun_out/t6_tower.log; cat gpurun_out/t6_tower.log
timeout 300 python tools/bench_tower.py 8192 20 > gpurun_out/tower_bench_rr.log 2>&1; tail -1 gpurun_out/tower_bench_rr.log | cut -c1-700
SGO_TOWER_PADDED_TILES=1 timeout 300 python tools/bench_tower.py 8192 20 > gpurun_out/tower_bench_padded6.log 2>&1; tail -1 gpurun_out/tower_bench_padded6.log | cut -c1-700
timeout 300 python tools/bench_tower.py 16384 20 > gpurun_out/tower_bench_rr16k.log 2>&1; tail -1 gpurun_out/tower_bench_rr16k.log | cut -c1-700
timeout 600 python tools/prof_kernels.py > gpurun_out/prof_units.json 2> gpurun_out/prof_units_err.log && \
timeout 1500 ncu --set full --clock-control none --profile-from-start off \
    -k regex:'k_(select_a|select_b|expand|backup_a|backup_b|reroot|pick|apply_moves|legal_masks|score|export_planes|export_packed|export_boards|tree_new|tree_valid|records_pack|leaf_gather|stem_im2col|heads_fc|child_stats)' \
    -c 160 -o gpurun_out/prof_tree_kernels python tools/prof_kernels.py > gpurun_out/ncu_tree.log 2>&1
tail -3 gpurun_out/ncu_tree.log | cut -c1-300
ncu -i gpurun_out/prof_tree_kernels.ncu-rep --page raw --csv > gpurun_out/prof_tree_kernels_raw.csv 2>/dev/null
python tools/ncu_summary.py < gpurun_out/prof_tree_kernels_raw.csv > gpurun_out/prof_tree_kernels_summary.txt; cat gpurun_out/prof_tree_kernels_summary.txt
ls -la gpurun_out/
if [ $(stat -c %s gpurun_out/prof_tree_kernels.ncu-rep) -gt 30000000 ]; then rm gpurun_out/prof_tree_kernels.ncu-rep; fi
du -sh gpurun_out
