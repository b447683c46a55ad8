set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_direct.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/probe_ops.py > gpurun_out/probe_ops.json 2>&1; cat gpurun_out/probe_ops.json
timeout 600 python tools/probe.py 1e6 64 > gpurun_out/probe.log 2>&1; grep -E '"variant": (31|4[0-9]|5[0-9]),' gpurun_out/probe.log | cut -c1-200
