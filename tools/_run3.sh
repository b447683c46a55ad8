mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_direct.py -m gpu -x -q 2>&1 | tail -3
OCG_PROBE_VARIANTS=31,46,47,58,59,60,61,62,63,64,65 timeout 600 python tools/probe.py 1e6 64 > gpurun_out/probe_np.log 2>&1; grep -E '"variant": (31|4[67]|5[89]|6[0-5]),' gpurun_out/probe_np.log | cut -c1-200
