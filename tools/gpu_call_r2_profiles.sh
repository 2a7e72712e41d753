#!/bin/bash
# Round-2 profile captures (one GPU): launch list of the timed command, ncu --set full of every
# large kernel of one training step, and of the B=100 step.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/r02_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra --no-parity --no-e2e \
    > gpurun_out/r02_bench_under_ncu.log 2>&1
tail -1 gpurun_out/r02_bench_under_ncu.log | cut -c1-200
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:"heads_fused_kernel|vae_mid|gemm_tc_kernel|csr_densify|adam_clip|fused_finish|splitk" -s 60 -c 13 \
    -o gpurun_out/r02_step_full -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --no-parity --no-e2e --no-graph \
    > gpurun_out/r02_ncu_full.log 2>&1
tail -2 gpurun_out/r02_ncu_full.log | cut -c1-200; ls -la gpurun_out/r02_step_full.ncu-rep
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/r02_launches_b100.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra --no-parity --no-e2e --minibatch 100 \
    > gpurun_out/r02_bench_b100_under_ncu.log 2>&1
tail -1 gpurun_out/r02_bench_b100_under_ncu.log | cut -c1-200
