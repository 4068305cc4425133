#!/bin/bash
# Round-2 ncu evidence in one gpurun call (every ncu pass only after the same command exited 0 without ncu):
#   gpurun --timeout 2400 -- 'bash tools/gpu_r02_profiles.sh r02'
T=${1:-r02}
cd oracle && make -s && cd ..
# 1. launch list of a short bench run (shares of the step per kernel)
timeout 400 python bench.py --steps 1 --warmup 1 --games 256 --no-cpu --no-sub > gpurun_out/plain_small_$T.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/launches_$T.csv \
    python bench.py --steps 1 --warmup 1 --games 256 --no-cpu --no-sub > gpurun_out/ncu_small_$T.log 2>&1
tail -2 gpurun_out/ncu_small_$T.log | cut -c1-200
# 2. the conv kernel at the bench's launch size (16,384 positions): one layer without and one with the skip tensor
timeout 200 python tools/bench_tower.py 16384 2 > gpurun_out/plain_tower_$T.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_conv3x3_pair -s 13 -c 2 -o gpurun_out/prof_conv_$T \
    python tools/bench_tower.py 16384 2 > gpurun_out/ncu_tower_$T.log 2>&1
tail -2 gpurun_out/ncu_tower_$T.log | cut -c1-200
ncu -i gpurun_out/prof_conv_$T.ncu-rep --page raw --csv > gpurun_out/prof_conv_${T}_raw.csv 2>/dev/null
python tools/ncu_summary.py < gpurun_out/prof_conv_${T}_raw.csv
# 3. stem / heads kernels of one forward (im2col, stem GEMM, the three dense-head GEMMs, k_heads_finish)
timeout 900 ncu --set full --clock-control none -k regex:'k_stem_im2col|k_heads_finish|k_conv3x3_pair' -s 12 -c 9 -o gpurun_out/prof_fwd_$T \
    python tools/bench_tower.py 16384 2 > gpurun_out/ncu_fwd_$T.log 2>&1
ncu -i gpurun_out/prof_fwd_$T.ncu-rep --page raw --csv > gpurun_out/prof_fwd_${T}_raw.csv 2>/dev/null
python tools/ncu_summary.py < gpurun_out/prof_fwd_${T}_raw.csv
rm -f gpurun_out/prof_fwd_$T.ncu-rep
# 4. board / tree kernels on a real 1,024-game ply per mode (node pool: k_reroot frees, k_select_* pop)
timeout 400 python tools/prof_kernels.py 1024 > gpurun_out/prof_units_$T.json 2> gpurun_out/prof_units_${T}_err.log && \
timeout 1200 ncu --set full --clock-control none --profile-from-start off -k regex:'^k_(?!conv|stem|heads)' -o gpurun_out/prof_tree_$T \
    python tools/prof_kernels.py 1024 > gpurun_out/ncu_tree_$T.log 2>&1
ncu -i gpurun_out/prof_tree_$T.ncu-rep --page raw --csv > gpurun_out/prof_tree_${T}_raw.csv 2>/dev/null
python tools/ncu_summary.py < gpurun_out/prof_tree_${T}_raw.csv
rm -f gpurun_out/prof_tree_$T.ncu-rep
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
