mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/bench_extra.py > gpurun_out/bench_extra.log 2>&1; python - <<'P'
import json
d=json.load(open('gpurun_out/bench_extra.json'))
for k,v in d.items():
    if isinstance(v,dict): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms_median','gbs','frac_of_hbm_peak','ginter_s','pct_fp32_peak','kernel_launches_per_step')})
P
timeout 300 python bench.py --grid 16 --n-src 1e6 --steps 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c1-shaped K1 (16^3 x 1e6):', d['value'], 'G/s', d['pct_fp32_peak'], '%')"
