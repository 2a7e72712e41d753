#!/bin/bash
# data-parallel sanity run on two GPUs (no extras, no CPU baseline)
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print("N=2 ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "replicas", d.get("replicas"), d.get("gradient_exchange"), "parity", d["parity"]["pass"])
PY
