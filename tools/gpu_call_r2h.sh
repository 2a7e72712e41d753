#!/bin/bash
mkdir -p gpurun_out
for n in 0 64 74 100; do
  SCVAE_MID_BWD_CTAS=$n timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-parity --steps 100 > gpurun_out/bench_$n.json 2> gpurun_out/bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$n.json"))
print("mid_bwd_ctas $n: ms_per_step", d["ms_per_step"])
PY
done
for g in 84 100 116; do
  SCVAE_MID_BWD_CTAS=0 SCVAE_SIDE_GEMM_CTAS=$g timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-parity --steps 100 > gpurun_out/bench_x.json 2> gpurun_out/bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_x.json"))
print("no share, side gemm ctas $g: ms_per_step", d["ms_per_step"])
PY
done
SCVAE_MID_FUSED=0 timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/bench_old.json 2> gpurun_out/bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_old.json"))
print("old middle path: ms_per_step", d["ms_per_step"], "launches", d["gpu_launches"]/d["steps"], d["parity"]["rel_err"])
PY
