#!/bin/bash
# A/B of engine knobs on one box (development aid)
mkdir -p gpurun_out
B="timeout 600 python bench.py --no-cpu-baseline --no-extra --no-e2e --steps 300"
show() { python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], "ms_per_step %.4f" % d["ms_per_step"], "parity", d["parity"]["pass"], d["parity"]["rel_err"], "clocks", d["clocks"])
print("   ", d["roofline"]["eager_ms_per_step_by_kernel"])
PY
}
SCVAE_DY1_SPLIT=0 timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_vae.py -m gpu -q -x -s 2>&1 | grep -a "gradient error.*ENCODER/1/DENSE\|passed\|failed\|Error" > gpurun_out/ab_dy1_tests.log; cat gpurun_out/ab_dy1_tests.log
SCVAE_DY1_SPLIT=0 $B > gpurun_out/ab_single.json 2>> gpurun_out/ab.err; show gpurun_out/ab_single.json
$B > gpurun_out/ab_split.json 2> gpurun_out/ab.err; show gpurun_out/ab_split.json
tail -3 gpurun_out/ab.err
