#!/bin/bash
# round 2, first call: new scale-parity tests first, then the whole suite, then one bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -q --durations=20 > gpurun_out/scale_tests.log 2>&1
tail -40 gpurun_out/scale_tests.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_scale.py > gpurun_out/gpu_tests.log 2>&1
tail -15 gpurun_out/gpu_tests.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
