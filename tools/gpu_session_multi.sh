#!/bin/bash
# One multi-GPU box session:  gpurun --gpus N --timeout 900 -- 'bash tools/gpu_session_multi.sh N'   (charged N x the box time)
# Strong-scaling headline, configs[4], the exchange layer's self-check and the sharded BRIDGE step (peer-memory vs NCCL gather).
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
S=_n$N
timeout 300 $TR bench.py --gpus $N --steps 5 > gpurun_out/bench$S.json 2> gpurun_out/bench$S.err
timeout 200 $TR tools/comm_check.py > gpurun_out/comm_check$S.json 2> gpurun_out/comm_check$S.err
timeout 200 $TR tools/bridge_multi.py --graph --steps 20 > gpurun_out/bridge_graph_peer$S.json 2> gpurun_out/bridge_graph_peer$S.err
timeout 200 $TR tools/bridge_multi.py --graph --steps 20 --exchange nccl > gpurun_out/bridge_graph_nccl$S.json 2> gpurun_out/bridge_graph_nccl$S.err
timeout 200 $TR tools/bridge_multi.py --steps 20 > gpurun_out/bridge_peer$S.json 2> gpurun_out/bridge_peer$S.err
timeout 200 $TR tools/bridge_multi.py --graph --steps 10 --integrator hermite > gpurun_out/bridge_graph_hermite$S.json 2> gpurun_out/bridge_graph_hermite$S.err
if [ "$2" = "c5" ]; then
  timeout 400 $TR bench.py --gpus $N --workload c5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5$S.json 2> gpurun_out/bench_c5$S.err
fi
tail -n 2 gpurun_out/bench$S.json gpurun_out/bridge_graph_peer$S.json gpurun_out/bridge_graph_nccl$S.json gpurun_out/comm_check$S.json | cut -c1-600
