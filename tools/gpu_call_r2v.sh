#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vae.py tests/test_gpu_kernels.py -m gpu -q -x -k "middle or packed or fused_training" > gpurun_out/v_tests.log 2>&1; grep -n "^E  .*Error\|passed\|failed\|^FAILED" gpurun_out/v_tests.log | tail -5
for f in device host; do
timeout 900 python bench.py --no-cpu-baseline --no-extra --no-parity --steps 200 --feeder $f > gpurun_out/bench_$f.json 2> gpurun_out/bench_$f.err; tail -2 gpurun_out/bench_$f.err
python - $f <<'PY'
import json, sys
d=json.load(open("gpurun_out/bench_%s.json" % sys.argv[1]))
e=d["e2e"]; print(sys.argv[1], "ms", d["ms_per_step"], "value", d["value"], "e2e", e["value"], e["h2d_bytes_per_step"], e["feeder"])
PY
done
