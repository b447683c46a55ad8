"""Per-source-line warp-stall samples of one kernel from an .ncu-rep (captured with --import-source on, -lineinfo build):
joins the SASS page of the report with nvdisasm's line table of the built object.
python tools/ncu_lines.py report.ncu-rep build/obj/file.o kernel_substring [top]"""
import csv
import io
import re
import subprocess
import sys
import tempfile
import os


def main():
    rep, obj, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    base = int(rows[2][ia], 16)
    samp = {}
    for r in rows[2:]:
        if len(r) > isamp and r[isamp].isdigit():
            samp[int(r[ia], 16) - base] = (int(r[isamp]), int(r[iex]) if r[iex].isdigit() else 0, r[1].strip())
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    line, infn, per_line, src_file = 0, False, {}, None
    for l in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            infn = kern in m.group(1)
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            src_file, line = m.group(1), int(m.group(2))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
        if m:
            off = int(m.group(1), 16)
            if off in samp:
                key = (os.path.basename(src_file or "?"), line)
                a = per_line.setdefault(key, [0, 0])
                a[0] += samp[off][0]
                a[1] += samp[off][1]
    tot = sum(v[0] for v in per_line.values()) or 1
    src = {}
    print("total samples %d" % tot)
    for (f, ln), (s, ex) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in src:
            try:
                src[f] = open(os.path.join(os.path.dirname(os.path.abspath(obj)), "../../oc_nbody_b200/csrc", f)).read().splitlines()
            except OSError:
                src[f] = []
        text = src[f][ln - 1].strip() if 0 < ln <= len(src[f]) else ""
        print("%5.1f%% %8d inst  %s:%d  %s" % (100.0 * s / tot, ex, f, ln, text[:110]))


if __name__ == "__main__":
    main()
