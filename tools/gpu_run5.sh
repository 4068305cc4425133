#!/bin/bash
# GPU session 5: config-4 match-play bench line; ncu --set full of the board/tree kernels at 1024 games
(timeout 900 python bench.py --match --no-cpu 2> gpurun_out/bench5_match_err.log | tail -1) > gpurun_out/bench5_match.json; cut -c1-1200 gpurun_out/bench5_match.json; tail -3 gpurun_out/bench5_match_err.log
timeout 600 python tools/prof_kernels.py > gpurun_out/prof_units.json 2> gpurun_out/prof_units_err.log && \
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_(select_a|select_b|expand|backup_a|backup_b|reroot|pick|apply_moves|legal_masks|score|export_planes|export_packed|export_boards|tree_new|tree_valid|records_pack|leaf_gather|stem_im2col|heads_fc|child_stats)' \
    -c 160 -o gpurun_out/prof_tree_kernels python tools/prof_kernels.py > gpurun_out/ncu_tree.log 2>&1
cat gpurun_out/prof_units.json; tail -3 gpurun_out/prof_units_err.log; tail -4 gpurun_out/ncu_tree.log | cut -c1-300
ls -la gpurun_out/prof_tree_kernels.ncu-rep
