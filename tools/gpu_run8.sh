#!/bin/bash
# GPU session 8: where does a conv layer's time go (ablations) + cuBLAS on the same box
timeout 300 python tools/conv_variants.py 8192 > gpurun_out/conv_variants8.json 2> gpurun_out/conv_variants8_err.log; cat gpurun_out/conv_variants8.json; tail -3 gpurun_out/conv_variants8_err.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.mem,power.draw,power.limit,temperature.gpu --format=csv
