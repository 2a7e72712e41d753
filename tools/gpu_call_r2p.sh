#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/mid_timeline.py 4096 2>&1 | tail -32
timeout 900 python -m pytest tests/test_zz_gpu_reference_graph.py tests/test_gpu_gmvae.py tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/t.log 2>&1; tail -4 gpurun_out/t.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|csr_densify" -s 30 -c 10 -o gpurun_out/prof_gemm -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --no-parity --no-e2e > gpurun_out/ncu_gemm.log 2>&1
tail -3 gpurun_out/ncu_gemm.log | cut -c1-200; ls -la gpurun_out/prof_gemm.ncu-rep
