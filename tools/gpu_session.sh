#!/bin/bash
# One GPU-box session (round 2): tests, headline bench (+ reference arm), accuracy surveys, secondary benches, launch list, ncu captures.
# Run with:  gpurun --timeout 2400 -- 'bash tools/gpu_session.sh'   (outputs land in gpurun_out/)
# Every profiler run follows a plain run of the same command that exited 0; numbers printed under ncu are never bench values.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
timeout 600 python bench.py --steps 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err || exit 1
timeout 600 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
if [ -z "$SKIP_ACC" ]; then
timeout 600 python tools/accuracy.py > gpurun_out/accuracy.log 2>&1
timeout 300 python tools/accuracy_c0.py > gpurun_out/accuracy_c0.log 2>&1
fi
timeout 600 python tools/bench_extra.py > gpurun_out/bench_extra.log 2>&1 && cp gpurun_out/bench_extra.json gpurun_out/bench_extra_plain.json
timeout 600 python tools/bench_hermite.py > gpurun_out/bench_hermite.log 2>&1
timeout 600 python tools/bench_rbf.py > gpurun_out/bench_rbf.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --n-src 1e6 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:direct_sum_tp --launch-skip 4 -c 1 -o gpurun_out/prof_k1 -f \
    python bench.py --n-src 1e6 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_k1.log 2>&1
# DRAM traffic of the full-size launch (roofline.traffic): only the two byte counters, one launch of the real configs[1] step
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none -k regex:direct_sum_tp \
    --launch-skip 4 -c 1 --csv --log-file gpurun_out/ncu_dram_bench_full.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras \
    > gpurun_out/ncu_dram.log 2>&1
timeout 200 python tools/bench_hermite.py --profile && timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:hermite_tp --launch-skip 3 -c 1 -o gpurun_out/prof_k6 -f python tools/bench_hermite.py --profile > gpurun_out/ncu_k6.log 2>&1
cp gpurun_out/bench_extra_plain.json gpurun_out/bench_extra.json 2>/dev/null
# summaries (run anywhere):  python tools/ncu_summary.py "title=gpurun_out/prof_k1.ncu-rep"
# multi-GPU (gpurun --gpus N):  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541"
#   $TR bench.py --gpus N ; $TR bench.py --gpus N --impl reference ; $TR tools/comm_check.py ; $TR tools/bridge_multi.py --graph [--exchange nccl]
