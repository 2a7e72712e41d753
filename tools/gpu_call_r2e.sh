#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "split_operand or adam" > gpurun_out/k_tests.log 2>&1; tail -5 gpurun_out/k_tests.log
timeout 600 python -m pytest tests/test_gpu_vae.py -m gpu -q > gpurun_out/vae_tests.log 2>&1; grep -n "^E  .*Error\|passed\|failed\|^FAILED" gpurun_out/vae_tests.log | tail -30
timeout 300 python tools/debug_step.py 1024 20000 50 100 2>&1 | tail -20
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -q > gpurun_out/scale_tests.log 2>&1; grep -n "^E  .*Error\|passed\|failed\|^FAILED" gpurun_out/scale_tests.log | tail -30
timeout 600 python bench.py --no-cpu-baseline --steps 100 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "launches/step", d["gpu_launches"]/d["steps"])
print(d["roofline"]["eager_ms_per_step_by_kernel"]); print(d["parity"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/bench_under_ncu.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/launches.csv gpurun_out/launches.md "r02 wip" ; grep -n "Launch sequence" -A 40 gpurun_out/launches.md
