"""Static evidence from the built objects: per kernel, the SASS mnemonics that show what the hot loops are made of —
packed FP32 (FFMA2/FADD2/FMUL2), MUFU.RSQ, TMA bulk copies (UBLKCP) with mbarrier transactions (SYNCS), REDUX, broadcast
LDS.128 — plus registers / spills from the ELF.  No GPU needed.
python tools/sass_evidence.py > profiles/r01_sass_evidence.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "build", "obj")
KEYS = ["FFMA2", "FADD2", "FMUL2", "FFMA", "MUFU.RSQ", "MUFU.RCP", "UBLKCP", "SYNCS", "CREDUX", "REDUX", "LDS.128", "LDS.64", "LDS", "STS",
        "DFMA", "DADD", "DMUL", "SHFL", "BAR.SYNC", "LDG", "STG", "STL", "LDL"]
WANT = ("direct_sum_tp_kernel", "direct_sum_kernel", "hermite_tp_kernel", "hermite_small_kernel", "self_gravity_small_kernel",
        "grid_interp_kernel", "rbf_interp_kernel", "near_sum_kernel")


def main():
    for o in sorted(os.listdir(OBJ)):
        if not o.endswith(".o"):
            continue
        path = os.path.join(OBJ, o)
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
        regs = {}
        fn = None
        for line in res.splitlines():
            m = re.match(r"\s*Function (\S+):", line)
            if m:
                fn = m.group(1)
            m = re.search(r"REG:(\d+).*?STACK:(\d+).*?SHARED:(\d+)", line)
            if m and fn:
                regs[fn] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
        demangle = {}
        names = re.findall(r"Function : (\S+)", sass)
        if names:
            dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
            demangle = dict(zip(names, dm))
        blocks = re.split(r"\n\s*Function : ", sass)
        rows = []
        for b in blocks[1:]:
            name = b.split("\n", 1)[0].strip()
            pretty = demangle.get(name, name)
            if not any(w in pretty for w in WANT):
                continue
            cnt = collections.Counter()
            for ins in re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", b):
                for k in KEYS:
                    if ins == k or ins.startswith(k + "."):
                        cnt[k] += 1
                        break
            r = regs.get(name, (0, 0, 0))
            rows.append((pretty, r, cnt))
        if rows:
            print("== %s" % o)
            for pretty, r, cnt in rows:
                short = re.sub(r"\([^()]*\)$", "", pretty).replace("void ", "").replace("(bool)", "").replace("(int)", "").replace(" ", "")
                items = " ".join("%s=%d" % (k, cnt[k]) for k in KEYS if cnt[k])
                print("%-46s regs=%-3d stack=%-3d | %s" % (short[:46], r[0], r[1], items))
            print()


if __name__ == "__main__":
    main()
