#!/bin/bash
# GPU session 4: native step driver + record packer tests, whole GPU suite, bench with the native step
cd oracle && make -s && cd ..
(timeout 600 python -m pytest tests/test_gpu_driver.py -q --timeout 300 2>&1 | tail -25) > gpurun_out/t4_driver.log; cat gpurun_out/t4_driver.log
(timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15) > gpurun_out/t4_all.log; cat gpurun_out/t4_all.log
(timeout 900 python bench.py --no-cpu 2> gpurun_out/bench4_err.log | tail -1) > gpurun_out/bench4_a.json; cat gpurun_out/bench4_a.json | cut -c1-900; tail -3 gpurun_out/bench4_err.log
(timeout 600 python bench.py --mode b --no-cpu 2> gpurun_out/bench4_b_err.log | tail -1) > gpurun_out/bench4_b.json; cat gpurun_out/bench4_b.json | cut -c1-400
