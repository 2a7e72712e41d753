#!/bin/bash
# one bench line without the extras (development aid)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_vae.py -m gpu -q -x -k "adam or train or middle" > gpurun_out/v_tests.log 2>&1; grep -n "^E  .*Error\|passed\|failed\|^FAILED" gpurun_out/v_tests.log | tail -5
timeout 900 python bench.py --no-cpu-baseline --no-extra --steps 200 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "launches/step", d["gpu_launches"]/d["steps"], d["parity"]["pass"], "e2e", d["e2e"]["value"])
PY
