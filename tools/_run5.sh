mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_2gpu.log; cat gpurun_out/pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/bridge_multi.py --stars 65536 --steps 5 > gpurun_out/bridge_n2.json 2> gpurun_out/bridge_n2.err; cat gpurun_out/bridge_n2.json; tail -3 gpurun_out/bridge_n2.err

