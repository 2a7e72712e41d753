#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "split_operand or gemm_f16" > gpurun_out/k_tests.log 2>&1; tail -5 gpurun_out/k_tests.log
timeout 600 python -m pytest tests/test_gpu_vae.py -m gpu -q -k "fused_middle" > gpurun_out/mid_tests.log 2>&1; grep -n "Error\|passed\|failed" gpurun_out/mid_tests.log | tail -30
timeout 300 python tools/debug_step.py 1024 20000 50 100 2>&1 | tail -20
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -q > gpurun_out/scale_tests.log 2>&1; grep -n "Error\|passed\|failed" gpurun_out/scale_tests.log | tail -30
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/bench_under_ncu.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/launches.csv gpurun_out/launches.md "r02 wip" ; grep -n "Launch sequence" -A 40 gpurun_out/launches.md
