#!/usr/bin/env python
"""Throughput of the host packer (FASTA -> bit planes) and of the files-in/files-out run at a C2-shaped slice.
usage: python tools/bench_packer.py [n_sites=200000] [line_width=0]"""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poutine_b200 import hostio, synth  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
lw = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tree = synth.yule_tree(5000, 20262)
B0, B1, V = synth.simulate_planes(tree, S, 7)
sn, lab = synth.simulate_phenotypes(tree, 8)
d = tempfile.mkdtemp(prefix="pb_pack_")
t0 = time.perf_counter()
fasta = os.path.join(d, "anc.fasta")
with open(fasta, "wb") as f:
    for v0 in range(0, tree.n_nodes, 256):          # decode in row blocks to bound memory
        v1 = min(tree.n_nodes, v0 + 256)
        ch = synth.planes_to_chars(B0[v0:v1], B1[v0:v1], V[v0:v1], 0, S)
        for v in range(v0, v1):
            f.write(b">" + tree.names[v].encode() + b"\n")
            row = ch[v - v0]
            if lw:
                for i in range(0, S, lw):
                    f.write(row[i:i + lw].tobytes() + b"\n")
            else:
                f.write(row.tobytes() + b"\n")
t_write = time.perf_counter() - t0
size = os.path.getsize(fasta)
with open(os.path.join(d, "t.nwk"), "w") as f:
    f.write(tree.to_newick() + "\n")
with open(os.path.join(d, "p.tsv"), "w") as f:
    for v, l in zip(sn, lab):
        f.write("%s\t%d\n" % (tree.names[int(v)], int(l)))
with open(os.path.join(d, "m.map"), "w") as f:
    for i in range(S):
        f.write("1\ts%d\t0\t%d\n" % (i, 10 * i + 1))
t0 = time.perf_counter()
fa = hostio.FastaFile(fasta)
t_index = time.perf_counter() - t0
node_of = {n: v for v, n in enumerate(tree.names)}
nor = np.array([node_of[n] for n in fa.names], dtype=np.int32)
W = (S + 63) // 64
P0, P1, PV = (np.zeros((tree.n_nodes, W), dtype=np.uint64) for _ in range(3))
t0 = time.perf_counter()
fa.pack_tile(nor, 0, S, P0, P1, PV, 0)
t_pack = time.perf_counter() - t0
ok = bool(np.array_equal(PV, V) and np.array_equal(P0 & PV, B0 & V) and np.array_equal(P1 & PV, B1 & V))
fa.close()
out = {"sites": S, "nodes": tree.n_nodes, "fasta_bytes": size, "line_width": lw, "threads": len(os.sched_getaffinity(0)),
       "index_s": t_index, "pack_s": t_pack, "pack_chars_per_s": tree.n_nodes * S / t_pack, "pack_fasta_GBps": size / t_pack / 1e9,
       "planes_GBps": 3 * tree.n_nodes * W * 8 / t_pack / 1e9, "matches_generator": ok, "fasta_write_s": t_write}
try:
    t0 = time.perf_counter()
    res = hostio.run_homoplasy_counter(fasta, os.path.join(d, "t.nwk"), os.path.join(d, "p.tsv"), os.path.join(d, "m.map"),
                                       os.path.join(d, "run.out"), num_replicates=100000, seed=1)
    out["files_in_files_out_s"] = time.perf_counter() - t0
    out["informative"] = res["n_informative"]
except Exception as e:  # no GPU here
    out["files_in_files_out_s"] = None
    out["error"] = str(e)[:200]
# the same run through the native host binary (process start and CUDA context creation included), files byte-compared
exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "poutine_b200", "poutine_b200_host")
if out["files_in_files_out_s"] is not None and os.path.exists(exe):
    import subprocess
    t0 = time.perf_counter()
    r = subprocess.run([exe, "-u", "-f", fasta, "-t", os.path.join(d, "t.nwk"), "-p", os.path.join(d, "p.tsv"), "-m", os.path.join(d, "m.map"),
                        "-r", "100000", "--seed", "1", "-d", d, "-o", "native.out", "-X"], capture_output=True, text=True)
    out["native_host_files_in_files_out_s"] = time.perf_counter() - t0
    out["native_host_rc"] = r.returncode
    out["native_host_stdout_tail"] = r.stdout.strip().splitlines()[-1:] if r.returncode == 0 else r.stderr[-300:]
    if r.returncode == 0:
        out["native_files_equal_python_files"] = all(
            open(os.path.join(d, "native.out") + sfx, "rb").read() == open(os.path.join(d, "run.out") + sfx, "rb").read()
            for sfx in ("", ".sorted_by_a2_maxT", ".sorted_by_a1_maxT"))
print(json.dumps(out))
for fn in os.listdir(d):
    os.unlink(os.path.join(d, fn))
os.rmdir(d)
