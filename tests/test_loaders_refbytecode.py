"""The loaders in front of the hot path against the REFERENCE'S OWN BYTECODE.

tests/golden/loaders_refbytecode.json holds what coevolution.jar's NewickTreeTokenizer + NewickTreeReader.readTree() and
compiled/Fasta_Manager.class next() + Fasta_Record.getHeader()/getSequence() — both source-less in the reference — return
for tricky inputs, executed by the class-file interpreter (oracle/tools/gen_loaders_refbytecode_golden.py; no JVM in the
authoring image).  Checked here: hostio.parse_newick (Python host), parse_newick of the native host (poutine_b200_host
--x-parse-tree), and the mmap'ed FASTA index + packer of libpoutine_b200.so (csrc/fasta.cpp) that both hosts use."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

import poutine_b200 as pb
from poutine_b200 import hostio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "poutine_b200", "poutine_b200_host")
with open(os.path.join(ROOT, "tests", "golden", "loaders_refbytecode.json")) as _f:
    GOLD = json.load(_f)


@pytest.fixture(scope="module", autouse=True)
def _built():
    pb.build()


def _f32_bits(x):
    return "%08x" % struct.unpack("<I", struct.pack("<f", float(x)))[0]


@pytest.mark.parametrize("case", GOLD["newick"], ids=[repr(c["text"])[:40] for c in GOLD["newick"]])
def test_newick_python_host_matches_reference_reader(case):
    if "throws" in case:
        with pytest.raises(hostio.DataFormatError) as ei:
            hostio.parse_newick(case["text"])
        assert str(ei.value) == case["message"]          # DataFormatException's and NumberFormatException's own text
        return
    t = hostio.parse_newick(case["text"])
    want = case["nodes"]
    assert list(t.parent) == [n["parent"] for n in want]
    assert t.names == [n["id"] for n in want]
    assert [_f32_bits(x) for x in t.branch_len] == [n["dist_bits"] for n in want]
    assert list(t.leaf_nodes) == case["leaves"]          # NewickTree.buildCaches: left-to-right DFS order


@pytest.mark.parametrize("case", GOLD["newick"], ids=[repr(c["text"])[:40] for c in GOLD["newick"]])
def test_newick_native_host_matches_reference_reader(case, tmp_path):
    p = tmp_path / "t.nwk"
    p.write_text(case["text"], encoding="utf-8")
    r = subprocess.run([EXE, "--x-parse-tree", str(p)], capture_output=True, text=True, timeout=60)
    if "throws" in case:
        assert r.returncode == 255 and r.stderr.strip() == "error: " + case["message"]
        return
    assert r.returncode == 0, r.stderr
    out = r.stdout.split("\n")
    n = int(out[0])
    rows = [l.split("\t") for l in out[1:1 + n]]
    want = case["nodes"]
    assert [int(x[0]) for x in rows] == [w["parent"] for w in want]
    assert [x[4] if x[3] == "1" else None for x in rows] == [w["id"] for w in want]
    assert [x[2] for x in rows] == [w["dist_bits"] for w in want]
    assert [int(x) for x in out[1 + n].split()] == case["leaves"]
    assert [int(x[1]) for x in rows] == [int(i in set(case["leaves"])) for i in range(n)]


@pytest.mark.parametrize("case", GOLD["fasta"], ids=[repr(c["raw_latin1"])[:32] for c in GOLD["fasta"]])
def test_fasta_index_and_packer_match_reference_reader(case, tmp_path):
    """record boundaries, trimmed header text, concatenated (untrimmed) sequence of every record; a file that does not start
    with a header line is refused where Fasta_Manager.next() calls System.exit(-1)"""
    p = tmp_path / "a.fasta"
    p.write_bytes(case["raw_latin1"].encode("latin-1"))
    if case.get("exit"):
        with pytest.raises(OSError):
            hostio.FastaFile(str(p))
        return
    assert "throws" not in case
    fa = hostio.FastaFile(str(p))
    recs = case["records"]
    assert fa.names == [h for h, _ in recs]
    assert [int(x) for x in fa.lengths] == [len(s) for _, s in recs]
    n = len(recs)
    S = max([len(s) for _, s in recs] + [0])
    if n and S:
        # every record's sequence through the tile packer == the char packer on the reference's getSequence() strings
        chars = np.zeros((n, S), dtype=np.uint8)           # beyond a ragged record's end: not a base
        for i, (_, s) in enumerate(recs):
            chars[i, :len(s)] = np.frombuffer(s.encode("latin-1"), dtype=np.uint8)
        B0, B1, V = pb.pack_chars(chars)
        W = (S + 63) // 64
        for tile_w in (W, 1):
            P0, P1, PV = (np.zeros((n, W), dtype=np.uint64) for _ in range(3))
            for w0 in range(0, W, tile_w):
                ns = min(S, 64 * (w0 + tile_w)) - 64 * w0
                t0, t1, tv = (np.zeros((n, tile_w), dtype=np.uint64) for _ in range(3))
                fa.pack_tile(np.arange(n, dtype=np.int32), 64 * w0, ns, t0, t1, tv, 2)
                P0[:, w0:w0 + tile_w], P1[:, w0:w0 + tile_w], PV[:, w0:w0 + tile_w] = t0, t1, tv
            assert np.array_equal(PV, V)
            assert np.array_equal(P0 & PV, B0 & V) and np.array_equal(P1 & PV, B1 & V)
    fa.close()


def test_fixture_covers_what_it_claims():
    nw, fa = GOLD["newick"], GOLD["fasta"]
    assert len(nw) >= 350 and sum("throws" in c for c in nw) >= 15 and sum("nodes" in c for c in nw) >= 20
    assert any(c.get("throws") == "NumberFormatException" for c in nw)
    assert any(any(n["id"] is None for n in c.get("nodes", [])) for c in nw)                 # unnamed nodes
    assert len(fa) >= 200 and sum(bool(c.get("exit")) for c in fa) >= 2
    assert any("\r" in c["raw_latin1"] and "\n" not in c["raw_latin1"] for c in fa)          # bare-CR line ends
