#!/usr/bin/env python3
"""Attribute a kernel's executed instructions and stall samples (ncu source page) to CUDA source lines.

usage: k1_lineprof.py report.ncu-rep kernel_regex cubin_name_in_lib [min_pct]
Joins `ncu --page source --csv` (SASS order) with `nvdisasm -g` line info of the cubin inside libpoutine_b200.so.
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, kern, cubin = sys.argv[1], sys.argv[2], sys.argv[3]
min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.4
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "poutine_b200", "libpoutine_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and re.search(kern, l)][0]
cur, insts = None, []
for l in dis[start + 1:]:
    if l.startswith("//-----"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        insts.append((cur, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + (sys.argv[5] if len(sys.argv) > 5 else kern)], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
h, rows = r[1], r[2:]
nxt = [i for i, x in enumerate(rows) if x and x[0] == 'Kernel Name']   # several launches matched: keep the first
if nxt:
    rows = rows[:nxt[0]]
ix = {n: i for i, n in enumerate(h)}
assert len(rows) == len(insts), (len(rows), len(insts))
by = collections.defaultdict(lambda: [0, 0, 0])
for (c, s), x in zip(insts, rows):
    by[c][0] += float(x[ix["Instructions Executed"]]); by[c][1] += float(x[ix["# Samples"]]); by[c][2] += 1
tot = sum(v[0] for v in by.values()); ts = sum(v[1] for v in by.values())
print(f"total warp-instructions {tot/1e6:.1f} M, samples {ts:.0f}, SASS lines {len(rows)}")
srcs = {}
for k, v in sorted(by.items(), key=lambda kv: kv[0] or ("", 0)):
    if 100 * v[0] / tot >= min_pct or 100 * v[1] / ts >= min_pct:
        t = ""
        if k and k[0].endswith(".cu"):
            srcs.setdefault(k[0], open(os.path.join(root, "poutine_b200", "csrc", k[0])).read().split("\n"))
            t = srcs[k[0]][k[1] - 1].strip()[:90]
        print(f"{str(k):38s} n={v[2]:4d} exec {v[0]/1e6:7.1f}M {100*v[0]/tot:5.1f}%  samp {100*v[1]/ts:5.1f}%  {t}")
