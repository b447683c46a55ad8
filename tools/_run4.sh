mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
timeout 600 python tools/accuracy.py > gpurun_out/accuracy.log 2>&1; grep -E '"variant": (31|40)' gpurun_out/accuracy.log | cut -c1-330
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
timeout 600 python bench.py --variant 46 --no-cpu-baseline > gpurun_out/bench_n1_v46.json 2> gpurun_out/bench_n1_v46.err; cut -c1-200 gpurun_out/bench_n1_v46.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:direct_sum --launch-skip 3 -c 1 --csv --log-file gpurun_out/ncu_dram_full.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_dram_full.log 2>&1; tail -3 gpurun_out/ncu_dram_full.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:direct_sum --launch-skip 3 -c 1 -o gpurun_out/prof_k1_mf -f python bench.py --n-src 1e6 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k1_mf.log 2>&1; tail -2 gpurun_out/ncu_k1_mf.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --n-src 1e6 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; wc -l gpurun_out/launches.csv
