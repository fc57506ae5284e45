"""Generate tests/golden/events_hand.json and tests/golden/c1_slice.npz.

(1) events_hand.json — an 8-leaf tree with one site per quirk of the reference's MRCA walk
    (SURVEY.md §8 a4).  The expected Homoplasy_Events fields are HAND-DERIVED from
    HC.java:1647-1698 / 1837-1888 / 1954-1962 and written below as literals; the script refuses
    to write the fixture unless the oracle reproduces every one of them.
(2) c1_slice.npz — a C1-shaped slice (1,300 leaves x 2,000 sites, 2 % NA labels, replay mode with
    Java-LCG permutations): inputs + the oracle's outputs.  A regression fixture at a size the bytecode interpreter
    cannot reach: it pins the GPU path to the oracle and the oracle to itself over time; the oracle itself is pinned to the
    reference's bytecode by gen_events_refbytecode_golden.py.

run:  python oracle/tools/gen_events_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from poutine_b200 import synth  # noqa: E402

# ((L1,L2)I1,(L3,(L4,L5)I3)I2,L6,(L7,L8)I4)R;
NAMES = ["R", "I1", "I2", "I3", "I4", "L1", "L2", "L3", "L4", "L5", "L6", "L7", "L8"]
PARENT = [-1, 0, 0, 2, 0, 1, 1, 2, 3, 3, 0, 4, 4]
LEAVES = [5, 6, 7, 8, 9, 10, 11, 12]


def col(default="A", **kw):
    c = {n: default for n in NAMES}
    c.update(kw)
    return "".join(c[n] for n in NAMES)


# name, column (node order NAMES), expected: None (not biallelic -> class) or (a1,a2,a1c,a2c,a1x,a2x)
SITES = [
    ("plain homoplasy: two C tips", col(L1="C", L4="C"), ("A", "C", 0, 2, 0, 2)),
    ("mutation on a root-child branch is never tested (HC.java:1674)", col(I1="C", L1="C", L2="C"), ("A", "C", 0, 0, 0, 0)),
    ("N at the root: root is bucketed into a2 (HC.java:1861-1865)", col(R="N", L1="C", L4="C"), ("A", "C", 0, 3, 0, 2)),
    ("parent carries a third allele: its valid children are a1 MRCAs", col(I2="G", L1="C", L7="C"), ("A", "C", 3, 2, 1, 2)),
    ("MRCA node itself carries a third allele: lands in a2", col(I3="G", L1="C"), ("A", "C", 3, 2, 2, 1)),
    ("lower-case leaf is missing: no census, never an MRCA", col(L1="c", L2="C", L4="C"), ("A", "C", 0, 2, 0, 2)),
    ("4-4 tie A/C: first in HashMap order A,C,T,G is major", col(I1="C", I2="C", I3="C", L1="C", L2="C", L3="C", L4="C"), ("A", "C", 2, 0, 1, 0)),
    ("4-4 tie T/G: T precedes G in HashMap order", col(default="T", I1="G", I2="G", I3="G", L1="G", L2="G", L3="G", L4="G"), ("T", "G", 2, 0, 1, 0)),
    ("single minor tip: bucket of size 1 is cleared (HC.java:1954-1962)", col(L1="C"), ("A", "C", 0, 0, 0, 0)),
    ("gap at an internal parent: valid children differ from '-'", col(I3="-", L1="C"), ("A", "C", 3, 0, 2, 0)),
    ("monomorphic", col(), None),
    ("tri-allelic leaves", col(L1="C", L2="G"), None),
    ("quad-allelic leaves", col(L1="C", L2="G", L3="T"), None),
    ("all leaves missing", col(L1="N", L2="N", L3="-", L4="n", L5="N", L6="a", L7="N", L8="N"), None),
    ("minor allele more frequent at internal nodes only: leaves decide", col(default="C", L1="A", L2="A", L3="A", L4="A", L5="A", L6="A", L7="C", L8="C", I4="C"), ("A", "C", 5, 0, 5, 0)),  # L6 hangs off the root: not tested
]


def hand_fixture():
    chars = np.array([[ord(s[1][v]) for s in SITES] for v in range(len(NAMES))], dtype=np.uint8)
    o = oracle.Oracle(PARENT, LEAVES)
    ev, cc = o.count_all_homoplasy_events(chars)
    by_id = {int(e["segsite_id"]): e for e in ev}
    out_sites = []
    for i, (name, column, exp) in enumerate(SITES):
        got = by_id.get(i + 1)
        if exp is None:
            assert got is None, (name, got)
        else:
            tup = (chr(got["allele1"]), chr(got["allele2"]), int(got["a1_count"]), int(got["a2_count"]),
                   int(got["a1_extant"]), int(got["a2_extant"]))
            assert tup == exp, f"oracle disagrees with the hand derivation at site {i + 1} ({name}): {tup} != {exp}"
        out_sites.append({"name": name, "column": column, "expected": list(exp) if exp else None})
    assert cc == dict(mono=1, bi=11, tri=1, quad=1, other=1), cc
    return {"names": NAMES, "parent": PARENT, "leaves": LEAVES, "sites": out_sites,
            "class_counts": cc, "derivation": "hand-derived from HC.java:1647-1698; verified against oracle/oracle.cpp"}


def c1_slice():
    tree = synth.yule_tree(1300, 20261)
    S, m = 2000, 400
    B0, B1, V = synth.simulate_planes(tree, S, 20261 + 2000)
    chars = synth.planes_to_chars(B0, B1, V, 0, S)
    sample_node, label = synth.simulate_phenotypes(tree, 20261 + 1000, na_frac=0.02, no_row_frac=0.01)
    o = oracle.Oracle(tree.parent, tree.leaf_nodes)
    ev, cc = o.count_all_homoplasy_events(chars, np.arange(S, dtype=np.int32) * 7 + 11)
    perms = oracle.java_shuffle_perms(20261, m, len(label))
    st, mt_s, mt_u = o.assoc(sample_node, label, m, 1, perms=perms)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "c1_slice.npz"),
                        parent=tree.parent, is_leaf=tree.is_leaf, leaf_nodes=tree.leaf_nodes,
                        B0=B0, B1=B1, V=V, n_sites=S, sample_node=sample_node, label=label, m=m, perm_seed=20261,
                        events=ev, class_counts=np.array([cc[k] for k in ("mono", "bi", "tri", "quad", "other")]),
                        stats=st, maxT_sorted=mt_s, maxT_unsorted=mt_u, p_success=o.p_success)
    print("c1_slice: biallelic", len(ev), "informative", len(st), "classes", cc)


if __name__ == "__main__":
    fx = hand_fixture()
    with open(os.path.join(ROOT, "tests", "golden", "events_hand.json"), "w") as f:
        json.dump(fx, f, indent=1)
    print("events_hand.json:", len(fx["sites"]), "sites OK")
    c1_slice()
    print({f: os.path.getsize(os.path.join(ROOT, "tests", "golden", f)) for f in os.listdir(os.path.join(ROOT, "tests", "golden"))})
