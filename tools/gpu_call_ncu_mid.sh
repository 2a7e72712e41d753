#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vae_mid -s 4 -c 2 -o gpurun_out/prof_mid -f python tools/mid_timeline.py --no-timeline > gpurun_out/ncu_mid.log 2>&1
tail -5 gpurun_out/ncu_mid.log; ls -la gpurun_out/prof_mid.ncu-rep
