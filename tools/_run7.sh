mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
OCG_PROBE_VARIANTS=31,46,58 timeout 600 python tools/probe.py 1e6 64 > gpurun_out/probe_pot.log 2>&1; grep -E '"variant": (31|46|58),' gpurun_out/probe_pot.log | cut -c1-220
timeout 600 python tools/accuracy.py > gpurun_out/accuracy.log 2>&1; grep -E '"kernel": 0' gpurun_out/accuracy.log | cut -c1-330
