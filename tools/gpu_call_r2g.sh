#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_eval.py 2>&1 | tail -12
timeout 600 python -m pytest tests/test_gpu_vae.py tests/test_gpu_kernels.py -m gpu -q > gpurun_out/vae_tests.log 2>&1; grep -n "^E  .*Error\|passed\|failed\|^FAILED" gpurun_out/vae_tests.log | tail -30
timeout 300 python tools/mid_timeline.py 2>&1 | tail -32
timeout 600 python bench.py --no-cpu-baseline --steps 100 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "launches/step", d["gpu_launches"]/d["steps"])
print(d["roofline"]["eager_ms_per_step_by_kernel"]); print(d["parity"]["rel_err"], d["parity"]["pass"])
PY
