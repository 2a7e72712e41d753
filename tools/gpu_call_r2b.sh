#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vae.py -m gpu -q -x -k "fused_middle" > gpurun_out/mid_tests.log 2>&1
tail -40 gpurun_out/mid_tests.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_vae.py -m gpu -q -x -k "fused_middle and (one-layer or two-layers)" > gpurun_out/mid_memcheck.log 2>&1
tail -15 gpurun_out/mid_memcheck.log
timeout 600 python -m pytest tests/test_gpu_vae.py tests/test_gpu_model.py tests/test_zz_gpu_reference_graph.py -m gpu -q > gpurun_out/vae_tests.log 2>&1
tail -15 gpurun_out/vae_tests.log
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -q > gpurun_out/scale_tests.log 2>&1
tail -30 gpurun_out/scale_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
