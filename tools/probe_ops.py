import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oc_nbody_b200 import default_context
ctx = default_context(0)
names = ["FFMA (TFLOP/s)", "FFMA2 (TFLOP/s)", "MUFU.RSQ (G/s)", "FADD2 (TFLOP/s, 2 flop/lane-instr)", "FMUL2", "FADD2 broadcast operand", "FFMA2 broadcast multiplier (TFLOP/s, 4 flop)", "inner-loop mix TPT=2 registers only (TFLOP/s at 20 flop)", "inner-loop mix TPT=1 registers only"]
out = {n: ctx.probe_throughput(i) for i, n in enumerate(names)}
for i, n in ((10, "FFMA2 alone (2 CTA x 256)"), (11, "FFMA2 + 5 LDS.128 per 48"), (12, "FFMA2 + 10 LDS.128 per 48"), (13, "FFMA2 + 8 MUFU per 48"), (14, "FFMA2 + 10 LDS.128 + 8 MUFU per 48")):
    out[n] = ctx.probe_throughput(i)
for i, n in ((20, "FFMA2 three distinct register pairs"), (21, "FFMA2 two distinct pairs + one reused"), (22, "FFMA2 two distinct pairs")):
    out[n] = ctx.probe_throughput(i)
print(json.dumps(out, indent=1))
