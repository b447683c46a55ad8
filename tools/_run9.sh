mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','bridge_step')}); print(d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['cpu_baseline'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-250 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --n-src 1e6 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; wc -l gpurun_out/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grid_interp --launch-skip 10 -c 1 -o gpurun_out/prof_k3 -f python tools/bench_extra.py > gpurun_out/ncu_k3.log 2>&1; tail -2 gpurun_out/ncu_k3.log
