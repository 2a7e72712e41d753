#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/pull_bench.py 2>&1 | tail -2
for f in device host; do
timeout 900 python bench.py --no-cpu-baseline --no-extra --no-parity --steps 200 --feeder $f > gpurun_out/bench_$f.json 2> gpurun_out/bench_$f.err; tail -2 gpurun_out/bench_$f.err
python - $f <<'PY'
import json, sys
d=json.load(open("gpurun_out/bench_%s.json" % sys.argv[1]))
e=d["e2e"]; print(sys.argv[1], "value", d["value"], "e2e", e["value"], e["h2d_bytes_per_step"], e["feeder"])
PY
done
