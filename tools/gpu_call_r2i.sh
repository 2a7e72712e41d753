#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "continuous" > gpurun_out/cont_tests.log 2>&1; grep -n "^E  .*Error\|passed\|failed\|^FAILED" gpurun_out/cont_tests.log | tail -30
timeout 900 python -m pytest tests/test_gpu_vae.py tests/test_gpu_model.py tests/test_gpu_gmvae.py -m gpu -q > gpurun_out/vae_tests.log 2>&1; grep -n "^E  .*Error\|passed\|failed\|^FAILED" gpurun_out/vae_tests.log | tail -30
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -q -k gmvae > gpurun_out/scale_tests.log 2>&1; grep -n "^E  .*Error\|passed\|failed\|^FAILED" gpurun_out/scale_tests.log | tail -10
timeout 900 python bench.py --no-cpu-baseline --steps 100 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "launches/step", d["gpu_launches"]/d["steps"])
print(json.dumps(d["extra_configs"], indent=1)[:3000])
PY
