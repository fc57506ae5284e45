"""Host side of call() around the hot path against the REFERENCE'S OWN BYTECODE (SURVEY §8f rows 2 and 3 + the pheno / map
loaders): tests/golden/hostside_refbytecode.json holds what compiled/Homoplasy_Counter.class's nexus_to_newick (HC.java:866),
read_phenos (HC.java:1972), get_physical_positions (HC.java:1207) and output_significance_assessments_binom (HC.java:2137, with
Homoplasy_Events.toString / Binomial_Test_Stat.toString / the Row comparators) produce, executed by the class-file
interpreter (oracle/tools/gen_hostside_refbytecode_golden.py).  Both hosts — poutine_b200/hostio.py and the native
poutine_b200/host/hostio.cpp through the binary's probes — must reproduce it: same Newick text, same entries and totals, same
positions, the three result files byte for byte, and an error wherever the reference's method ends in its System.exit(-1)."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

import poutine_b200 as pb
from poutine_b200 import hostio
from poutine_b200._lib import SITE_DTYPE, STAT_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "poutine_b200", "poutine_b200_host")
with open(os.path.join(ROOT, "tests", "golden", "hostside_refbytecode.json")) as _f:
    GOLD = json.load(_f)


@pytest.fixture(scope="module", autouse=True)
def _built():
    pb.build()


def run(*args):
    return subprocess.run([EXE, *args], capture_output=True, text=True, timeout=60)


def _ids(cases):
    return [repr(c["raw_latin1"])[:36] for c in cases]


@pytest.mark.parametrize("case", GOLD["nexus"], ids=_ids(GOLD["nexus"]))
def test_nexus_to_newick(case, tmp_path):
    p = tmp_path / "annotated_tree.nexus"
    p.write_bytes(case["raw_latin1"].encode("latin-1"))
    r = run("--x-nexus-to-newick", str(p))
    if "exits" in case:
        with pytest.raises(hostio.DataFormatError):
            hostio.nexus_to_newick(str(p))
        assert r.returncode == 255
        return
    assert hostio.nexus_to_newick(str(p)) == case["newick"]
    assert r.returncode == 0 and r.stdout == case["newick"] + "\n"


@pytest.mark.parametrize("case", GOLD["phenos"], ids=_ids(GOLD["phenos"]))
def test_read_phenos(case, tmp_path):
    p = tmp_path / "p.tsv"
    p.write_bytes(case["raw_latin1"].encode("latin-1"))
    r = run("--x-read-phenos", str(p))
    if "exits" in case:
        with pytest.raises(hostio.DataFormatError):
            hostio.read_phenos(str(p))
        assert r.returncode == 255 and "malformed phenotype file" in r.stderr
        return
    want = case["entries"]
    got = hostio.read_phenos(str(p))
    assert got == want
    assert (sum(v == "1" for v in got.values()), sum(v == "0" for v in got.values())) == (case["cases"], case["controls"])
    assert r.returncode == 0
    lines = r.stdout.split("\n")[:-1]
    assert lines[0] == "%d\t%d\t%d" % (case["cases"], case["controls"], len(want) - case["cases"] - case["controls"])
    assert dict(l.split("\t", 1) for l in lines[1:]) == want and len(lines) - 1 == len(want)
    # both hosts list the samples in pheno-file order (first occurrence); the reference's HashMap order is not observable in
    # the results (a uniform shuffle is invariant to the enumeration order) and is kept in the fixture for the replay tooling
    assert [l.split("\t", 1)[0] for l in lines[1:]] == list(got.keys())
    assert sorted(case["hashmap_order"]) == sorted(got.keys())
    if want:
        assert ("cases = %d, controls = %d, samples missing phenotype = %d" %
                (case["cases"], case["controls"], len(want) - case["cases"] - case["controls"])) in case["log"]


@pytest.mark.parametrize("case", GOLD["map"], ids=_ids(GOLD["map"]))
def test_get_physical_positions(case, tmp_path):
    p = tmp_path / "m.map"
    p.write_bytes(case["raw_latin1"].encode("latin-1"))
    r = run("--x-read-map", str(p))
    if "exits" in case:
        with pytest.raises(hostio.DataFormatError):
            hostio.read_map(str(p))
        assert r.returncode == 255 and "malformed map file" in r.stderr
        return
    assert hostio.read_map(str(p)).tolist() == case["positions"]
    assert r.returncode == 0 and [int(x) for x in r.stdout.split()] == case["positions"]


def _dbl(bits_hex):
    return struct.unpack(">d", bytes.fromhex(bits_hex))[0]


@pytest.mark.parametrize("case", GOLD["writer"], ids=[c["name"] for c in GOLD["writer"]])
def test_result_files_match_reference_writer(case, tmp_path):
    rows = case["rows"]
    S = max(r["segsite_id"] for r in rows)
    sites = np.zeros(S, dtype=SITE_DTYPE)
    sites["segsite_id"] = np.arange(1, S + 1)
    pos = np.arange(S) + 1000000                       # positions of the sites without a row: never printed
    st = np.zeros(len(rows), dtype=STAT_DTYPE)
    for i, r in enumerate(rows):
        k = r["segsite_id"] - 1
        sites[k]["site_class"] = 2
        sites[k]["allele1"], sites[k]["allele2"] = ord(r["allele1"]), ord(r["allele2"])
        for f in ("a1_count", "a2_count", "a1_extant", "a2_extant"):
            sites[k][f] = r[f]
        pos[k] = r["physical_pos"]
        st[i]["site_index"], st[i]["segsite_id"] = k, r["segsite_id"]
        st[i]["obs_counts"] = r["obs_counts"]
        for f in ("r_a1", "r_a2", "r_maxT_a1", "r_maxT_a2"):
            st[i][f] = r[f]
        for f in ("obs_p_a1", "obs_p_a2", "pointwise_a1", "pointwise_a2", "familywise_a1", "familywise_a2"):
            st[i][f] = _dbl(r[f])
    want = {"": case["out"], ".sorted_by_a2_maxT": case["sorted_by_a2_maxT"], ".sorted_by_a1_maxT": case["sorted_by_a1_maxT"]}
    assert want[""].count("\n") == len(rows) + 1
    py = str(tmp_path / "py.out")
    hostio.write_outputs(py, sites, st, pos)
    for sfx, text in want.items():
        assert open(py + sfx).read() == text, ("python host", sfx)
    mp = tmp_path / "m.map"
    mp.write_text("".join("1\ts%d\t0\t%d\n" % (i, p) for i, p in enumerate(pos)))
    sites.tofile(tmp_path / "sites.bin")
    st.tofile(tmp_path / "stats.bin")
    r = run("--x-write-records", str(tmp_path / "sites.bin"), str(tmp_path / "stats.bin"), str(mp), str(tmp_path / "cpp.out"))
    assert r.returncode == 0, r.stderr
    for sfx, text in want.items():
        assert open(str(tmp_path / "cpp.out") + sfx).read() == text, ("native host", sfx)


def test_fixture_covers_what_it_claims():
    assert sum("exits" in c for c in GOLD["nexus"]) >= 5 and sum("newick" in c for c in GOLD["nexus"]) >= 5
    assert sum("exits" in c for c in GOLD["phenos"]) >= 3 and sum("exits" in c for c in GOLD["map"]) >= 4
    w = GOLD["writer"]
    assert len(w) == 5 and sum(len(c["rows"]) for c in w) > 400
    assert any("NaN" in c["out"] for c in w) and any("\t-1\t" in c["out"] for c in w)      # sentinel rows (allele not in use)
    # ties in the sort keys exist (the stable double sort matters) and NaN rows are dropped from the sorted files
    assert any(c["sorted_by_a1_maxT"].count("\n") < c["out"].count("\n") for c in w)
    fam = [l.split("\t")[-1] for l in w[0]["sorted_by_a2_maxT"].split("\n")[1:-1]]
    assert len(set(fam)) < len(fam)
