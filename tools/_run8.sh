mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus 8 --workload c5 --steps 1 --warmup 3 > gpurun_out/bench_c5_n8.json 2> gpurun_out/bench_c5_n8.err; cut -c1-300 gpurun_out/bench_c5_n8.json
timeout 600 $TR --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; cut -c1-300 gpurun_out/bench_n8.json
timeout 600 $TR --master-port 29523 tools/bridge_multi.py --stars 65536 --steps 5 > gpurun_out/bridge_n8.json 2> gpurun_out/bridge_n8.err; cat gpurun_out/bridge_n8.json
