#!/bin/bash
mkdir -p gpurun_out
nproc; free -g | head -2
timeout 900 python bench.py --no-cpu-baseline --no-extra --steps 200 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "launches/step", d["gpu_launches"]/d["steps"], d["parity"]["pass"])
e=d["e2e"]; print(e["value"], e["h2d_bytes_per_step"], e["feeder"], e["host_pack_ms_per_step"])
PY
