#!/bin/bash
# short bench with the driver's step counts (development aid)
mkdir -p gpurun_out
timeout 100 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra --no-cpu-baseline --no-e2e > gpurun_out/bench_k20.json 2> gpurun_out/bench_k20.err; tail -3 gpurun_out/bench_k20.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_k20.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "parity", d["parity"]["pass"], "blocks", len(d["timed_blocks"]["ms"]), min(d["timed_blocks"]["ms"]), max(d["timed_blocks"]["ms"]))
PY
