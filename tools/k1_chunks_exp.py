#!/usr/bin/env python
"""Sweep kernel time of a small one-plane shard as a function of the edge program's chunk count (PB_K1_TARGET_BLOCKS is read when
the tree / site count is set): python tools/k1_chunks_exp.py --sites 125000 1036 2072 518 259 148"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poutine_b200 as pb
from poutine_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--sites", type=int, default=125000)
ap.add_argument("--config", default="C2")
ap.add_argument("--full", action="store_true")
ap.add_argument("targets", nargs="*")
a = ap.parse_args()
cfg = synth.CONFIGS[a.config]
tree = synth.yule_tree(cfg["n_leaves"], cfg["seed"])
sn, lab = synth.simulate_phenotypes(tree, cfg["seed"] + 1000)
B0, B1, V = synth.simulate_planes(tree, a.sites, cfg["seed"] + 2000 + 7919)
for t in ["default"] + list(a.targets):
    if t != "default":
        os.environ["PB_K1_TARGET_BLOCKS"] = t
    hc = pb.HomoplasyCounter(num_replicates=1000, seed=1, full_planes=a.full)
    hc.set_tree(tree.parent, tree.is_leaf)
    hc.set_phenos(sn, lab)
    hc.set_planes(B0, B1, V, a.sites, compact=not a.full)
    k1 = []
    for i in range(13):
        hc.count_all_homoplasy_events(fetch_sites=False)
        if i >= 3:
            k1.append(hc.timings()["ms_events_kernel"])
    tm = hc.timings()
    gb = tm["k1_algorithmic_bytes"] / 1e9
    print("target_blocks %-8s mode %d  K1 %.4f ms (min %.4f)  %.0f GB/s" % (t, tm["plane_mode"], np.mean(k1), np.min(k1), gb / (np.mean(k1) / 1e3)), flush=True)
    hc.close()
