"""Generate tests/golden/events_refbytecode.json and tests/golden/assoc_refbytecode.json: outputs of the REFERENCE'S OWN
BYTECODE for both entry points of the hot path.

`count_all_homoplasy_events(tree, seg_sites, physical_poss)` (HC.java:1402) is executed from
/root/reference/compiled/Homoplasy_Counter.class (+ coevolution.jar's NewickTree classes) by the bytecode interpreter
in refjvm.py — there is no JVM in the authoring container — on seeded inputs that cover what the path distinguishes:
polytomies, leaves hanging off the root, shuffled node ids, gaps / N / lower-case calls, third alleles at internal
nodes, ties in the allele census, tri- and quad-allelic and all-missing sites, plus one all-valid biallelic alignment
(the one-plane storage mode's input).  Per biallelic site the fixture keeps every Homoplasy_Events field
(HC.java:5531-5550) and both MRCA node-name sets; per case the class counters the reference wrote to its log
(HC.java:1477-1480).  The script refuses to write the fixture unless the C++ oracle reproduces all of it.

Entry point 2, `assoc(all_events, phenos)` (HC.java:2029 -> 2096), runs on the very Homoplasy_Events objects entry point 1
left in the interpreter: subset_by_min_hcount (HC.java:3783), then the live resampling driver (HC.java:3013) with
commons-math 3.6.1's BinomialTest executed from ITS jar, m Replicate.run() calls (HC.java:3229), the point-wise and the
family-wise estimates.  Collections.shuffle draws from a seeded java.util.Random and the replicates run serially (the two
things the JDK leaves open, see refjvm.ReferenceEvents.assoc); every shuffle is recorded as a permutation of the
pheno-file rows, which is what the oracle's and the GPU's replay mode consume.  assoc_refbytecode.json keeps, per
informative site, obs_counts and every Binomial_Test_Stat field (HC.java:5876-5887, doubles as IEEE-754 bit patterns),
the unsorted and sorted maxT arrays and the permutations.  Again the fixture is only written if the oracle reproduces
all of it bit for bit.

This is what pins the oracle (and through it the CUDA path) for rows a1-a11 of SURVEY.md section 8.

run (in the authoring container, needs /root/reference):  python oracle/tools/gen_events_refbytecode_golden.py
"""
import json
import os
import re
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import refjvm  # noqa: E402
import oracle  # noqa: E402
from poutine_b200 import synth  # noqa: E402

CASES = [
    # name, tree kwargs, sites, char kwargs
    dict(name="nasty_small", leaves=14, tseed=701, polytomy=0.3, root_leaf_children=2, S=160, cseed=801, nasty=True,
         m=120, min_hcount=1, na_frac=0.1, no_row_frac=0.1),
    dict(name="nasty_mid", leaves=60, tseed=702, polytomy=0.25, root_leaf_children=1, S=130, cseed=802, nasty=True,
         m=100, min_hcount=1, na_frac=0.05, no_row_frac=0.03),
    dict(name="nasty_polytomies", leaves=45, tseed=703, polytomy=0.6, root_leaf_children=3, S=110, cseed=803, nasty=True,
         m=100, min_hcount=0, na_frac=0.0, no_row_frac=0.05),
    dict(name="binary_no_root_leaves", leaves=80, tseed=704, polytomy=0.0, root_leaf_children=0, S=100, cseed=804, nasty=True,
         m=80, min_hcount=2, na_frac=0.04, no_row_frac=0.0),
    dict(name="biallelic_all_valid", leaves=70, tseed=705, polytomy=0.2, root_leaf_children=2, S=140, cseed=805, nasty=False,
         m=100, min_hcount=1, na_frac=0.0, no_row_frac=0.0),
]


# Second set (python gen_events_refbytecode_golden.py extra -> tests/golden/{events,assoc}_refbytecode_extra.json): degenerate
# topologies and phenotypes the first set does not reach — every leaf on the root (no testable edge at all), a ladder, two
# leaves, nearly every sample without a pheno row or with a missing label, only cases / only controls (background p = 1 / 0:
# the binomial chain's boundary branches), min_hcount above every count but one.  Used by the oracle tests only.
CASES_EXTRA = [
    dict(name="star", kind="star", leaves=12, tseed=711, S=90, cseed=811, nasty=True, m=60, min_hcount=0, na_frac=0.1, no_row_frac=0.1),
    dict(name="caterpillar", kind="caterpillar", leaves=25, tseed=712, S=120, cseed=812, nasty=True, m=80, min_hcount=1, na_frac=0.05,
         no_row_frac=0.05),
    dict(name="two_leaves", kind="star", leaves=2, tseed=713, S=40, cseed=813, nasty=True, m=20, min_hcount=0, na_frac=0.0, no_row_frac=0.0),
    dict(name="few_pheno_rows", leaves=50, tseed=714, polytomy=0.3, root_leaf_children=1, S=120, cseed=814, nasty=True, m=80, min_hcount=1,
         na_frac=0.3, no_row_frac=0.6),
    dict(name="all_cases", leaves=40, tseed=715, polytomy=0.2, root_leaf_children=1, S=100, cseed=815, nasty=False, m=60, min_hcount=1,
         na_frac=0.1, no_row_frac=0.0, labels="all_cases"),
    dict(name="all_controls", leaves=40, tseed=716, polytomy=0.2, root_leaf_children=1, S=100, cseed=816, nasty=False, m=60, min_hcount=1,
         na_frac=0.0, no_row_frac=0.1, labels="all_controls"),
    dict(name="min_hcount_4", leaves=60, tseed=717, polytomy=0.1, root_leaf_children=0, S=150, cseed=817, nasty=False, m=60, min_hcount=4,
         na_frac=0.02, no_row_frac=0.02),
]


def make_tree(c):
    kind = c.get("kind", "random")
    if kind == "star":                      # every leaf hangs off the root: all edges are root edges, never tested (HC.java:1678)
        n = c["leaves"]
        rng = np.random.default_rng(c["tseed"])
        perm = rng.permutation(n + 1)
        parent = np.full(n + 1, -1, dtype=np.int32)
        for old in range(1, n + 1):
            parent[perm[old]] = perm[0]
        return synth._finish_tree(parent, rng)
    if kind == "caterpillar":
        return synth.caterpillar_tree(c["leaves"], c["tseed"])
    return synth.random_tree(c["leaves"], c["tseed"], polytomy=c["polytomy"], root_leaf_children=c["root_leaf_children"])


def bits(x):
    return "%016x" % int(np.float64(x).view(np.uint64))



def names_of(tree):
    return [str(n) for n in tree.names]


def run_case(c):
    tree = make_tree(c)
    S = c["S"]
    if c["nasty"]:
        chars = synth.simulate_chars(tree, S, c["cseed"], nasty=True)
    else:
        B0, B1, V = synth.simulate_planes(tree, S, c["cseed"])
        chars = synth.planes_to_chars(B0, B1, V, 0, S)
    chars = np.asarray(chars, dtype=np.uint8)
    names = names_of(tree)
    assert len(set(names)) == len(names)
    parent = [int(p) for p in tree.parent]
    pos = [int(x) for x in (np.arange(S) * 13 + 7)]
    t0 = time.time()
    ref = refjvm.ReferenceEvents()
    ev, log, _console = ref.run(parent, names, chars.tolist(), pos)
    dt = time.time() - t0
    counts = {k: int(re.search(r"# %s sites: (\d+)" % k, log).group(1)) for k in ("bi-allelic", "monomorphic", "tri-allelic", "quad-allelic")}
    # the oracle must reproduce every field and both MRCA sets
    o = oracle.Oracle(tree.parent, tree.leaf_nodes)
    or_ev, or_cc = o.count_all_homoplasy_events(chars, np.asarray(pos, dtype=np.int32))
    assert (or_cc["bi"], or_cc["mono"], or_cc["tri"], or_cc["quad"]) == (counts["bi-allelic"], counts["monomorphic"], counts["tri-allelic"],
                                                                         counts["quad-allelic"]), (or_cc, counts)
    assert or_cc["other"] == S - sum(counts.values())
    assert len(or_ev) == len(ev)
    name_arr = np.asarray(names)
    for i, (r, e) in enumerate(zip(ev, or_ev)):
        got = (int(e["segsite_id"]), int(e["physical_pos"]), chr(e["allele1"]), chr(e["allele2"]), int(e["a1_count"]), int(e["a2_count"]),
               int(e["a1_extant"]), int(e["a2_extant"]))
        want = (r["segsite_id"], r["physical_pos"], r["allele1"], r["allele2"], r["a1_count"], r["a2_count"], r["a1_extant"], r["a2_extant"])
        assert got == want, (c["name"], i, got, want)
        assert sorted(name_arr[o.mrca(i, 0)].tolist()) == r["a1_mrcas"], (c["name"], i, "a1 set")
        assert sorted(name_arr[o.mrca(i, 1)].tolist()) == r["a2_mrcas"], (c["name"], i, "a2 set")
    print("%-24s N=%3d S=%3d biallelic %3d  %.1f s, %d bytecodes: oracle == reference bytecode" %
          (c["name"], tree.n_nodes, S, len(ev), dt, ref.vm.n_bytecodes))
    ev_case = {"name": c["name"], "names": names, "parent": parent, "leaf_nodes": [int(v) for v in tree.leaf_nodes],
               "physical_pos": pos, "rows": ["".join(map(chr, row)) for row in chars.tolist()],
               "class_counts": {"bi": counts["bi-allelic"], "mono": counts["monomorphic"], "tri": counts["tri-allelic"],
                                "quad": counts["quad-allelic"], "other": S - sum(counts.values())},
               "log": log, "events": ev}

    # ---- entry point 2 on the same interpreter state
    m, mh = c["m"], c["min_hcount"]
    sample_node, label = synth.simulate_phenotypes(tree, c["tseed"] + 50, na_frac=c["na_frac"], no_row_frac=c["no_row_frac"])
    if c.get("labels") == "all_cases":
        label = np.where(label == 2, 2, 1).astype(label.dtype)
    elif c.get("labels") == "all_controls":
        label = np.where(label == 2, 2, 0).astype(label.dtype)
    label_str = [{0: "0", 1: "1"}.get(int(x), "NA") for x in label]
    t0 = time.time()
    n0 = ref.vm.n_bytecodes
    ra = ref.assoc([names[int(v)] for v in sample_node], label_str, m, mh, shuffle_seed=9000 + c["tseed"])
    dt = time.time() - t0
    perms = np.asarray(ra["perms"], dtype=np.uint32)
    for r in range(m):
        assert sorted(perms[r].tolist()) == list(range(len(label)))
    st, mt_s, mt_u = o.assoc(sample_node, label, m, mh, perms=perms)
    assert bits(o.p_success) == bits(ra["p_success"])
    assert len(st) == len(ra["sites"]), (len(st), len(ra["sites"]))
    assert [bits(x) for x in mt_u] == [bits(x) for x in ra["maxT_unsorted"]], "per-replicate min p (maxT) differs"
    assert [bits(x) for x in mt_s] == [bits(x) for x in ra["maxT_sorted"]]
    rows = []
    for srow, r in zip(st, ra["sites"]):
        assert int(srow["segsite_id"]) == r["segsite_id"] and int(srow["a1_in_use"]) == r["a1_in_use"] and int(srow["a2_in_use"]) == r["a2_in_use"]
        assert srow["obs_counts"].tolist() == r["obs_counts"], (c["name"], r["segsite_id"], srow["obs_counts"], r["obs_counts"])
        for k in ("r_a1", "r_a2", "r_maxT_a1", "r_maxT_a2"):
            assert int(srow[k]) == r[k], (c["name"], r["segsite_id"], k, int(srow[k]), r[k])
        row = {k: r[k] for k in ("segsite_id", "a1_in_use", "a2_in_use", "obs_counts", "r_a1", "r_a2", "r_maxT_a1", "r_maxT_a2")}
        for ko, kr in (("obs_p_a1", "obs_binom_pvalue_a1"), ("obs_p_a2", "obs_binom_pvalue_a2"), ("pointwise_a1", "pointwise_pvalue_a1"),
                       ("pointwise_a2", "pointwise_pvalue_a2"), ("familywise_a1", "familywise_pvalue_a1"), ("familywise_a2", "familywise_pvalue_a2")):
            a, b = float(srow[ko]), r[kr]
            assert (a != a and b != b) or bits(a) == bits(b), (c["name"], r["segsite_id"], ko, a, b)
            row[ko] = bits(b)
        rows.append(row)
    print("%-24s assoc: P=%3d m=%3d min_hcount=%d informative %3d  %.1f s, %d bytecodes: oracle == reference bytecode (bit for bit)" %
          ("", len(label), m, mh, len(rows), dt, ref.vm.n_bytecodes - n0))
    as_case = {"name": c["name"], "m": m, "min_hcount": mh, "shuffle_seed": 9000 + c["tseed"],
               "sample_node": [int(v) for v in sample_node], "label": [int(x) for x in label], "p_success": bits(ra["p_success"]),
               "perms": perms.tolist(), "sites": rows, "maxT_unsorted": [bits(x) for x in ra["maxT_unsorted"]],
               "maxT_sorted": [bits(x) for x in ra["maxT_sorted"]]}
    return ev_case, as_case


if __name__ == "__main__":
    extra = len(sys.argv) > 1 and sys.argv[1] == "extra"
    suffix = "_extra" if extra else ""
    both = [run_case(c) for c in (CASES_EXTRA if extra else CASES)]
    out = {"source": "count_all_homoplasy_events executed from /root/reference/compiled/Homoplasy_Counter.class + coevolution.jar by "
                     "oracle/tools/refjvm.py (bytecode interpreter; JDK collections emulated with HashMap's iteration order)",
           "cases": [b[0] for b in both]}
    path = os.path.join(ROOT, "tests", "golden", "events_refbytecode%s.json" % suffix)
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(path, os.path.getsize(path), "bytes")
    out = {"source": "subset_by_min_hcount + resample_all_mutations_binom_test_combined_nulldists_memoization_concurrent executed from "
                     "/root/reference/compiled/*.class + commons-math3-3.6.1.jar by oracle/tools/refjvm.py on the events of "
                     "events_refbytecode%s.json (same case names); replicates serial, Collections.shuffle seeded; doubles as hex bit patterns" % suffix,
           "cases": [b[1] for b in both]}
    path = os.path.join(ROOT, "tests", "golden", "assoc_refbytecode%s.json" % suffix)
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(path, os.path.getsize(path), "bytes")
