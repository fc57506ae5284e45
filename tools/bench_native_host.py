#!/usr/bin/env python
"""Files-in/files-out wall clock of the native host (poutine_b200/poutine_b200_host) on a C2-shaped slice, per upload
tile size, with its own phase clocks (--timing).  usage: python tools/bench_native_host.py [n_sites=100000] [tile,tile,...]"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poutine_b200 import synth  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
tiles = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8192, 32768, 131072]
exe = os.path.join(ROOT, "poutine_b200", "poutine_b200_host")
tree = synth.yule_tree(5000, 20262)
B0, B1, V = synth.simulate_planes(tree, S, 7)
sn, lab = synth.simulate_phenotypes(tree, 8)
d = tempfile.mkdtemp(prefix="pb_native_")
fasta = os.path.join(d, "anc.fasta")
with open(fasta, "wb") as f:
    for v0 in range(0, tree.n_nodes, 256):
        v1 = min(tree.n_nodes, v0 + 256)
        ch = synth.planes_to_chars(B0[v0:v1], B1[v0:v1], V[v0:v1], 0, S)
        for v in range(v0, v1):
            f.write(b">" + tree.names[v].encode() + b"\n" + ch[v - v0].tobytes() + b"\n")
with open(os.path.join(d, "t.nwk"), "w") as f:
    f.write(tree.to_newick() + "\n")
with open(os.path.join(d, "p.tsv"), "w") as f:
    for v, l in zip(sn, lab):
        f.write("%s\t%d\n" % (tree.names[int(v)], int(l)))
with open(os.path.join(d, "m.map"), "w") as f:
    for i in range(S):
        f.write("1\ts%d\t0\t%d\n" % (i, 10 * i + 1))
runs = []
for rep, tile in enumerate([tiles[0]] + tiles):        # the first run warms the page cache and the driver; not reported
    t0 = time.perf_counter()
    r = subprocess.run([exe, "-u", "-f", fasta, "-t", os.path.join(d, "t.nwk"), "-p", os.path.join(d, "p.tsv"), "-m", os.path.join(d, "m.map"),
                        "-r", "100000", "--seed", "1", "-d", d, "-o", "native.out", "-X", "--tile-sites", str(tile), "--timing"],
                       capture_output=True, text=True)
    wall = time.perf_counter() - t0
    if rep == 0 and r.returncode == 0:
        continue
    phases = {l[9:38].strip(): float(l[38:].split()[0]) for l in r.stderr.splitlines() if l.startswith("[timing]")}
    runs.append({"tile_sites": tile, "rc": r.returncode, "wall_s_incl_process_start": wall, "phases_ms": phases,
                 "tail": (r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1]})
print(json.dumps({"sites": S, "nodes": tree.n_nodes, "fasta_bytes": os.path.getsize(fasta), "replicates": 100000, "runs": runs}))
for fn in os.listdir(d):
    os.unlink(os.path.join(d, fn))
os.rmdir(d)
