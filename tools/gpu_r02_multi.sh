#!/bin/bash
# N-GPU checks of the NCCL paths (weight-blob broadcast, record-row gather, launcher):  gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_r02_multi.sh 2 r02'
N=${1:-2}; T=${2:-r02}
P=$((29500 + RANDOM % 400))
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P"
# 1. the launcher on a small configuration: 48 games over N ranks, 8 slots each, records gathered every 4 plies
rm -rf /tmp/sgo_ms && mkdir -p /tmp/sgo_ms && cd /tmp/sgo_ms
(PYTHONPATH=$GRAFT_REPO_ROOT timeout 600 $RUN -m sejonggo_b200.main_selfplay --games 48 --concurrent 8 --sims 32 --mode b --size 9 --blocks 2 --num-moves 10 --gather-every 4 2>&1 | tail -6) > $GRAFT_REPO_ROOT/gpurun_out/main_selfplay_${N}gpu_$T.log
cd $GRAFT_REPO_ROOT
echo "game dirs: $(ls /tmp/sgo_ms/sp_self_play_data/model_1 2>/dev/null | wc -l)  samples: $(find /tmp/sgo_ms/sp_self_play_data -name 'sample.npz' | wc -l)" >> gpurun_out/main_selfplay_${N}gpu_$T.log
cat gpurun_out/main_selfplay_${N}gpu_$T.log
# 2. the bench at N GPUs (per-rank diagnostics in the line)
P=$((P + 1))
(timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 3 --warmup 3 --no-sub 2> gpurun_out/bench_${N}gpu_${T}_err.log | tail -1) > gpurun_out/bench_${N}gpu_$T.json
cat gpurun_out/bench_${N}gpu_$T.json; tail -3 gpurun_out/bench_${N}gpu_${T}_err.log
