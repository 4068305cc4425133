#!/bin/bash
# GPU session 10: shifted-GEMM conv (A slab loaded once per channel chunk, 9 taps by descriptor shift)
(timeout 900 python -m pytest tests/test_gpu_tower.py tests/test_gpu_driver.py -q --timeout 600 -x 2>&1 | tail -12) > gpurun_out/t10_tower.log; cat gpurun_out/t10_tower.log
timeout 300 python tools/conv_variants.py 8192 > gpurun_out/conv_variants10.json 2> gpurun_out/conv_variants10_err.log; python - <<'PY'
import json
d=json.load(open('gpurun_out/conv_variants10.json'))
for k,v in d.items(): print('%-32s'%k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items()})
PY
tail -3 gpurun_out/conv_variants10_err.log
timeout 300 python tools/bench_tower.py 16384 20 > gpurun_out/tower_bench10.log 2>&1; tail -1 gpurun_out/tower_bench10.log | cut -c1-700
