#!/bin/bash
mkdir -p gpurun_out
for cfg in "1024 20000 50 100" "128 256 10 64" "333 256 10 100"; do
  echo "=== mid on: $cfg"; timeout 300 python tools/debug_step.py $cfg 2>&1 | tail -14
  echo "=== mid off: $cfg"; SCVAE_MID_FUSED=0 timeout 300 python tools/debug_step.py $cfg 2>&1 | tail -14
done > gpurun_out/debug_step.log 2>&1
cat gpurun_out/debug_step.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-parity --steps 50 > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("ms_per_step", d["ms_per_step"], "launches/step", d["gpu_launches"]/d["steps"], d["roofline"]["eager_ms_per_step_by_kernel"])
PY
