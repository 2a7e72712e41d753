#!/bin/bash
mkdir -p gpurun_out
for f in hybrid; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 100 --warmup 5 --no-parity --feeder $f > gpurun_out/bench_n4_$f.json 2> gpurun_out/bench_n4_$f.err; tail -1 gpurun_out/bench_n4_$f.err | cut -c1-200
python - $f <<'PY'
import json, sys
d=json.loads(open("gpurun_out/bench_n4_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("N=4", sys.argv[1], "ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "replicas", d["replicas"]["identical"])
PY
done
