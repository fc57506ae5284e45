#!/usr/bin/env python
"""Quick resident timing of the C2 workload for kernel experiments (not a bench line):
   PB_LIB=exp/libpb_X.so python tools/k1_quick.py [--full] [--steps N]   -> K1 ms, step ms"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poutine_b200 as pb
from poutine_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--full", action="store_true")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--config", default="C2")
ap.add_argument("--sites", type=int, default=0)
a = ap.parse_args()
cfg = synth.CONFIGS[a.config]
S = a.sites or cfg["n_sites"]
tree = synth.yule_tree(cfg["n_leaves"], cfg["seed"])
sn, lab = synth.simulate_phenotypes(tree, cfg["seed"] + 1000, na_frac=cfg.get("na_frac", 0.0))
B0, B1, V = synth.simulate_planes(tree, S, cfg["seed"] + 2000 + 7919)
hc = pb.HomoplasyCounter(num_replicates=cfg["m"], seed=20260, full_planes=a.full)
hc.set_tree(tree.parent, tree.is_leaf)
hc.set_phenos(sn, lab)
hc.set_planes(B0, B1, V, S, compact=True)
for _ in range(3):
    hc.count_all_homoplasy_events(fetch_sites=False); hc.assoc(fetch=False)
k1 = []
hc.stopwatch_start()
for _ in range(a.steps):
    hc.count_all_homoplasy_events(fetch_sites=False); hc.assoc(fetch=False)
    tm = hc.timings(); k1.append(tm["ms_events_kernel"])
ms = hc.stopwatch_stop_ms() / a.steps
print("%s mode=%d K1 %.4f ms (min %.4f)  events_total %.3f  null_count %.3f  step %.3f ms" %
      (os.environ.get("PB_LIB", "default"), tm["plane_mode"], np.mean(k1), np.min(k1), tm["ms_events_total"], tm["ms_null_count"], ms))
