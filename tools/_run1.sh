set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
timeout 600 python tools/probe.py 1e6 64 > gpurun_out/probe.log 2>&1; tail -12 gpurun_out/probe.log
timeout 600 python tools/accuracy.py > gpurun_out/accuracy.log 2>&1; cat gpurun_out/accuracy.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:direct_sum --launch-skip 3 -c 1 -o gpurun_out/prof_mf -f python bench.py --n-src 1e6 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_mf.log 2>&1; tail -3 gpurun_out/ncu_mf.log
