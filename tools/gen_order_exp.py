#!/usr/bin/env python
"""Resident step time of one workload under several settings of the library's experiment knobs, in ONE process
(the knobs are read with getenv at call time):  python tools/gen_order_exp.py [--config C2] 'PB_GEN_START=k1' 'PB_GEN_HEADSTART_US=0' ..."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poutine_b200 as pb
from poutine_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--config", default="C2")
ap.add_argument("--sites", type=int, default=0)
ap.add_argument("variants", nargs="*")
a = ap.parse_args()
cfg = synth.CONFIGS[a.config]
S = a.sites or cfg["n_sites"]
tree = synth.yule_tree(cfg["n_leaves"], cfg["seed"])
sn, lab = synth.simulate_phenotypes(tree, cfg["seed"] + 1000, na_frac=cfg.get("na_frac", 0.0))
B0, B1, V = synth.simulate_planes(tree, S, cfg["seed"] + 2000 + 7919)
hc = pb.HomoplasyCounter(num_replicates=cfg["m"], seed=20260)
hc.set_tree(tree.parent, tree.is_leaf)
hc.set_phenos(sn, lab)
hc.set_planes(B0, B1, V, S, compact=True)
for var in [""] + list(a.variants) + [""]:
    kv = [x.split("=", 1) for x in var.split(",") if x]
    for k, v in kv:
        os.environ[k] = v
    for _ in range(3):
        hc.count_all_homoplasy_events(fetch_sites=False); hc.assoc(fetch=False)
    hc.stopwatch_start()
    parts = {}
    for _ in range(a.steps):
        hc.count_all_homoplasy_events(fetch_sites=False); hc.assoc(fetch=False)
        tm = hc.timings()
        for k in ("ms_events_kernel", "ms_events_total", "ms_ptable", "ms_null_gen", "ms_null_count", "ms_pvalues", "ms_assoc_total", "ms_gen_kernel"):
            parts[k] = parts.get(k, 0.0) + tm[k] / a.steps
    ms = hc.stopwatch_stop_ms() / a.steps
    print("%-40s step %.3f ms | %s" % (var or "(default)", ms, " ".join("%s %.3f" % (k[3:], v) for k, v in parts.items())), flush=True)
    for k, _ in kv:
        del os.environ[k]
