#!/bin/bash
# short bench with the driver's step counts (development aid)
mkdir -p gpurun_out
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/bench_k20.json 2> gpurun_out/bench_k20.err; tail -3 gpurun_out/bench_k20.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_k20.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"], "parity", d["parity"]["pass"])
print(d["timed_blocks"])
PY
