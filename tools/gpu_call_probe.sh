timeout 300 python tools/wgrad_probe.py 2>&1 | tail -12
