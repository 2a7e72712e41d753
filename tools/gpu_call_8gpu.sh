#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -2 gpurun_out/bench_n$N.err | cut -c1-300
python - $N <<'PY'
import json, sys
n=sys.argv[1]
d=json.loads(open("gpurun_out/bench_n%s.json" % n).read().strip().splitlines()[-1])
print("N=%s ms_per_step" % n, d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "replicas", d.get("replicas"), "parity", (d.get("parity") or {}).get("pass"))
PY
