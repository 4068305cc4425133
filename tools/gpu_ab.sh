#!/bin/bash
# same-box A/B of the working tree's library against another build of the same ABI:  bash tools/gpu_ab.sh <other.so> <tag>
OLD=$PWD/$1; T=${2:-ab}
(timeout 900 python -m pytest tests/test_gpu_tower.py tests/test_gpu_driver.py -q --timeout 600 2>&1 | tail -4) > gpurun_out/tests_$T.log; cat gpurun_out/tests_$T.log
timeout 300 python tools/conv_variants.py 8192 > gpurun_out/conv_variants_$T.json 2> gpurun_out/conv_variants_${T}_err.log; python - <<PY
import json
d=json.load(open('gpurun_out/conv_variants_$T.json'))
for k,v in d.items(): print('%-28s'%k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items()})
PY
for i in 1 2; do
  timeout 300 python tools/bench_tower.py 16384 20 > gpurun_out/tower_${T}_new$i.log 2>&1; tail -1 gpurun_out/tower_${T}_new$i.log | cut -c1-330
  SGO_LIBRARY=$OLD timeout 300 python tools/bench_tower.py 16384 20 > gpurun_out/tower_${T}_old$i.log 2>&1; tail -1 gpurun_out/tower_${T}_old$i.log | cut -c1-330
done
