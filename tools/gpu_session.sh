#!/bin/bash
# One GPU-box session: tests, headline bench (+ reference arm), secondary benches, launch list, ncu captures.
# Run with:  gpurun --timeout 2400 -- 'bash tools/gpu_session.sh'   (outputs land in gpurun_out/)
# Every profiler run follows a plain run of the same command that exited 0; numbers printed under ncu are never bench values.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err || exit 1
timeout 600 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 python tools/bench_extra.py > gpurun_out/bench_extra.log 2>&1 && cp gpurun_out/bench_extra.json gpurun_out/bench_extra_plain.json
timeout 600 python tools/accuracy.py > gpurun_out/accuracy.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --n-src 1e6 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:direct_sum --launch-skip 3 -c 1 -o gpurun_out/prof_k1 -f \
    python bench.py --n-src 1e6 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grid_interp --launch-skip 10 -c 1 -o gpurun_out/prof_k3 -f \
    python tools/bench_extra.py > gpurun_out/ncu_k3.log 2>&1
cp gpurun_out/bench_extra_plain.json gpurun_out/bench_extra.json   # the copy written under ncu is not a measurement
# the reference-algorithm modes: K6 (Hermite force loop) and K7 (kNN + RBF-PHS kick)
timeout 600 python tools/bench_hermite.py > gpurun_out/bench_hermite.log 2>&1
timeout 600 python tools/bench_rbf.py > gpurun_out/bench_rbf.log 2>&1
timeout 200 python tools/bench_hermite.py --profile && timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:hermite_tp --launch-skip 3 -c 1 -o gpurun_out/prof_k6 -f python tools/bench_hermite.py --profile > gpurun_out/ncu_k6.log 2>&1
timeout 200 python tools/bench_rbf.py --profile && timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:rbf_interp --launch-skip 1 -c 1 -o gpurun_out/prof_k7 -f python tools/bench_rbf.py --profile > gpurun_out/ncu_k7.log 2>&1
# summaries (run anywhere):  python tools/ncu_summary.py "title=gpurun_out/prof_k6.ncu-rep" ;  python tools/ncu_lines.py gpurun_out/prof_k7.ncu-rep build/obj/rbf_interp.o rbf_interp_kernel
# multi-GPU (gpurun --gpus N):  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 tools/bridge_multi.py --graph [--integrator hermite]
