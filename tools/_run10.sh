mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_direct.py -m gpu -x -q -k "variants" 2>&1 | tail -3
OCG_PROBE_VARIANTS=46,58,66,67,68,69,70 timeout 600 python tools/probe.py 1e6 64 > gpurun_out/probe_ord.log 2>&1; grep -E '"variant": (46|58|6[6-9]|70), "name".*"kernel": 0, "pot": false' gpurun_out/probe_ord.log | cut -c1-200
