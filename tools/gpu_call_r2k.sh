#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_packed.py 2>&1 | grep -v "after 2nd" | cut -c1-120
timeout 1500 python -m pytest tests/test_gpu_scale.py -m gpu -q > gpurun_out/scale_tests.log 2>&1; grep -n "^E  .*Error\|passed\|failed\|^FAILED" gpurun_out/scale_tests.log | tail -10
timeout 900 python bench.py --no-cpu-baseline --steps 200 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "launches/step", d["gpu_launches"]/d["steps"])
print(d["parity"])
PY
