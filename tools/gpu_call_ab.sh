#!/bin/bash
# A/B of engine knobs on one box (development aid)
mkdir -p gpurun_out
B="timeout 300 python bench.py --no-cpu-baseline --no-extra --no-e2e --steps 300"
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "ms_per_step %.4f" % d["ms_per_step"], "parity", d["parity"]["pass"], d["parity"]["rel_err"]["elbo"], "launches", d["gpu_launches"] / d["steps"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
: > gpurun_out/ab.err
for m in 0 127 1 126 0 127; do
  SCVAE_PDL=$m $B > gpurun_out/ab_pdl$m.json 2>> gpurun_out/ab.err; show gpurun_out/ab_pdl$m.json
done
SCVAE_PDL=127 timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_vae.py tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -3
tail -5 gpurun_out/ab.err
