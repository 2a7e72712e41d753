#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; tail -4 gpurun_out/gpu_tests.log
timeout 1200 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -3 gpurun_out/bench_full.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_full.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "launches/step", d["gpu_launches"]/d["steps"], "parity", d["parity"]["pass"])
print("e2e", d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"])
print("roofline", d["roofline"])
print("cpu", d["cpu_baseline"])
print("clocks", d["clocks"])
for e in d["extra_configs"]:
    print(e["workload"][:40], e["minibatch"], round(e["ms_per_step"],4), round(e["value"]), e["dominant_kernel"], e["parity"]["pass"], e["parity"]["rel_err"])
print(d["evaluate_reconstruction"])
PY
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
