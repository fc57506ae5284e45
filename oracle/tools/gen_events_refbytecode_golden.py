"""Generate tests/golden/events_refbytecode.json: outputs of the REFERENCE'S OWN BYTECODE for entry point 1.

`count_all_homoplasy_events(tree, seg_sites, physical_poss)` (HC.java:1402) is executed from
/root/reference/compiled/Homoplasy_Counter.class (+ coevolution.jar's NewickTree classes) by the bytecode interpreter
in refjvm.py — there is no JVM in the authoring container — on seeded inputs that cover what the path distinguishes:
polytomies, leaves hanging off the root, shuffled node ids, gaps / N / lower-case calls, third alleles at internal
nodes, ties in the allele census, tri- and quad-allelic and all-missing sites, plus one all-valid biallelic alignment
(the one-plane storage mode's input).  Per biallelic site the fixture keeps every Homoplasy_Events field
(HC.java:5531-5550) and both MRCA node-name sets; per case the class counters the reference wrote to its log
(HC.java:1477-1480).  The script refuses to write the fixture unless the C++ oracle reproduces all of it.

This is what pins the oracle (and through it the CUDA path) for rows a1-a5 of SURVEY.md section 8.

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
    dict(name="nasty_small", leaves=14, tseed=701, polytomy=0.3, root_leaf_children=2, S=160, cseed=801, nasty=True),
    dict(name="nasty_mid", leaves=60, tseed=702, polytomy=0.25, root_leaf_children=1, S=130, cseed=802, nasty=True),
    dict(name="nasty_polytomies", leaves=45, tseed=703, polytomy=0.6, root_leaf_children=3, S=110, cseed=803, nasty=True),
    dict(name="binary_no_root_leaves", leaves=80, tseed=704, polytomy=0.0, root_leaf_children=0, S=100, cseed=804, nasty=True),
    dict(name="biallelic_all_valid", leaves=70, tseed=705, polytomy=0.2, root_leaf_children=2, S=140, cseed=805, nasty=False),
]


def names_of(tree):
    return [str(n) for n in tree.names]


def run_case(c):
    tree = synth.random_tree(c["leaves"], c["tseed"], polytomy=c["polytomy"], root_leaf_children=c["root_leaf_children"])
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
    return {"name": c["name"], "names": names, "parent": parent, "leaf_nodes": [int(v) for v in tree.leaf_nodes],
            "physical_pos": pos, "rows": ["".join(map(chr, row)) for row in chars.tolist()],
            "class_counts": {"bi": counts["bi-allelic"], "mono": counts["monomorphic"], "tri": counts["tri-allelic"],
                             "quad": counts["quad-allelic"], "other": S - sum(counts.values())},
            "log": log, "events": ev}


if __name__ == "__main__":
    out = {"source": "count_all_homoplasy_events executed from /root/reference/compiled/Homoplasy_Counter.class + coevolution.jar by "
                     "oracle/tools/refjvm.py (bytecode interpreter; JDK collections emulated with HashMap's iteration order)",
           "cases": [run_case(c) for c in CASES]}
    path = os.path.join(ROOT, "tests", "golden", "events_refbytecode.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(path, os.path.getsize(path), "bytes")
