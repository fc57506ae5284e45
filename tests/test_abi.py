"""CPU-side checks of the drop-in boundary: the library builds, loads, exports every symbol the
header declares, the host packer is exact, and a context refuses to exist without an sm_100 GPU
(no CPU fallback).  No GPU compute here."""
import ctypes
import os
import re

import numpy as np
import pytest

import poutine_b200 as pb
from poutine_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    pb.build()
    return _lib.lib()


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "poutine_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pb_[a-z_A-Z0-9]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.pb_version() >= 100


def test_struct_layouts_match_header():
    assert _lib.SITE_DTYPE.itemsize == 32 and _lib.STAT_DTYPE.itemsize == 96
    assert ctypes.sizeof(_lib.PbConfig) == 56
    assert ctypes.sizeof(_lib.PbClassCounts) == 48
    assert ctypes.sizeof(_lib.PbTimings) == 10 * 4 + 10 * 8 + 2 * 4 + 2 * 8


def test_struct_layouts_match_the_c_header_itself(tmp_path):
    """sizes and field offsets as the C compiler sees include/poutine_b200.h == the ctypes / numpy mirrors == the numbers
    INTEGRATION.md gives the Java host (pb_config 56 B, pb_class_counts 48 B, pb_site_out 32 B, pb_stat_out 96 B)"""
    import subprocess
    fields = {"pb_config": [n for n, _ in _lib.PbConfig._fields_], "pb_class_counts": [n for n, _ in _lib.PbClassCounts._fields_],
              "pb_timings": [n for n, _ in _lib.PbTimings._fields_], "pb_site_out": list(_lib.SITE_DTYPE.names),
              "pb_stat_out": list(_lib.STAT_DTYPE.names)}
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "poutine_b200.h"', 'int main(void) {']
    for st, names in fields.items():
        src.append('printf("%s %%zu\\n", sizeof(%s));' % (st, st))
        for n in names:
            src.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (st, n, st, n))
    src.append("return 0; }")
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = str(tmp_path / "layout")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), str(c), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = dict(l.split() for l in subprocess.run([exe], capture_output=True, text=True).stdout.splitlines())
    got = {k: int(v) for k, v in got.items()}
    assert (got["pb_config"], got["pb_class_counts"], got["pb_site_out"], got["pb_stat_out"]) == (56, 48, 32, 96)
    for st, cls in (("pb_config", _lib.PbConfig), ("pb_class_counts", _lib.PbClassCounts), ("pb_timings", _lib.PbTimings)):
        assert got[st] == ctypes.sizeof(cls)
        for n, _ in cls._fields_:
            assert got[st + "." + n] == getattr(cls, n).offset, (st, n)
    for st, dt in (("pb_site_out", _lib.SITE_DTYPE), ("pb_stat_out", _lib.STAT_DTYPE)):
        assert got[st] == dt.itemsize
        for n in dt.names:
            assert got[st + "." + n] == dt.fields[n][1], (st, n)
    # the offsets the FFM stub in INTEGRATION.md writes pb_config through
    assert [got["pb_config." + n] for n in ("device", "mode", "n_replicates", "min_hcount", "seed")] == [0, 4, 8, 16, 24]


def test_packer_matches_reference_char_tests(lib):
    rng = np.random.default_rng(0)
    alphabet = np.frombuffer(b"ACGTacgtN-nRYKM?*", dtype=np.uint8)
    for N, S in [(3, 1), (5, 63), (4, 64), (7, 65), (9, 1000)]:
        chars = alphabet[rng.integers(0, len(alphabet), size=(N, S))]
        B0, B1, V = pb.pack_chars(chars)
        W = (S + 63) // 64
        assert B0.shape == (N, W)
        back = synth.planes_to_chars(B0, B1, V, 0, S)
        valid = np.isin(chars, np.frombuffer(b"ACGT", np.uint8))   # geno == 'A' || 'C' || 'G' || 'T' (HC.java:1600)
        np.testing.assert_array_equal(back[valid], chars[valid])
        assert np.all(back[~valid] == ord("N"))
        # padding bits of the last word are invalid
        if S % 64:
            assert np.all(V[:, -1] >> np.uint64(S % 64) == 0)


def test_packer_large_multithreaded(lib):
    tree = synth.yule_tree(300, 1)
    B0, B1, V = synth.simulate_planes(tree, 5000, 2, leaf_gap=True)
    chars = synth.planes_to_chars(B0, B1, V, 0, 5000)
    P0, P1, PV = pb.pack_chars(chars)
    np.testing.assert_array_equal(PV, V)
    np.testing.assert_array_equal(P0 & PV, B0 & V)
    np.testing.assert_array_equal(P1 & PV, B1 & V)


def test_packer_rejects_bad_arguments(lib):
    a = np.zeros((2, 8), dtype=np.uint8)
    out = np.zeros((2, 1), dtype=np.uint64)
    assert lib.pb_pack_chars(None, 2, 8, 8, _lib.ptr(out), _lib.ptr(out), _lib.ptr(out), 1) != 0
    assert lib.pb_pack_chars(_lib.ptr(a), 2, 80, 8, _lib.ptr(out), _lib.ptr(out), _lib.ptr(out), 1) != 0


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    with pytest.raises(pb.PoutineError) as ei:
        pb.HomoplasyCounter(num_replicates=10)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_create_validates_config(lib):
    cfg = _lib.PbConfig(device=0, mode=7, n_replicates=10, min_hcount=1)
    h = ctypes.c_void_p()
    assert lib.pb_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"invalid config" in lib.pb_last_error(None)
    cfg = _lib.PbConfig(device=0, mode=0, n_replicates=0, min_hcount=1)
    assert lib.pb_create(ctypes.byref(cfg), ctypes.byref(h)) == -1


def test_product_does_not_import_oracle():
    # the oracle is test infrastructure: nothing under poutine_b200/ may reference it
    for dirpath, _, files in os.walk(os.path.join(ROOT, "poutine_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt, f
                assert not re.search(r'#include\s+"[^"]*oracle/', txt), f


# ---- the C ABI from plain C -------------------------------------------------------------------------------
def _build_host_min(tmp_path):
    import subprocess
    exe = str(tmp_path / "host_min")
    libdir = os.path.join(ROOT, "poutine_b200")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "host_min.c"), "-o", exe, "-L" + libdir, "-lpoutine_b200",
                        "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _host_min_inputs():
    """the input examples/host_min.c builds, restated"""
    parent = np.array([-1, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7], dtype=np.int32)
    N, S = 17, 70
    is_leaf = np.ones(N, dtype=np.uint8)
    is_leaf[parent[parent >= 0]] = 0
    chars = np.zeros((N, S), dtype=np.uint8)
    for s in range(S):
        a, b = (ord("C"), ord("T")) if s & 1 else (ord("A"), ord("G"))
        chars[:, s] = a
        for v in range(N):
            if is_leaf[v] and ((v * 7 + s * 3) % 5 == 0 or (v + s) % 11 == 0):
                chars[v, s] = b
        if s % 9 == 4:
            chars[[4, 10, 11], s] = b
    chars[16, 5] = chars[16, 66] = ord("N")
    sample_node = np.flatnonzero(is_leaf).astype(np.int32)
    label = np.array([2 if v == 12 else (v & 1) for v in sample_node], dtype=np.uint8)
    return parent, is_leaf, chars, sample_node, label


def test_header_is_c99_and_plain_c_host_builds(lib, tmp_path):
    """include/poutine_b200.h is valid C99 under -Wall -Wextra -Werror, a plain-C host links against the library, and
    without a GPU it fails loudly at pb_create (no CPU fallback)."""
    import subprocess
    import torch
    exe = _build_host_min(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: the run itself is covered by the gpu test below")
    r = subprocess.run([exe, "50"], capture_output=True, text=True)
    assert r.returncode == 1 and "pb_create failed" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_plain_c_host_matches_python_mirror(lib, tmp_path):
    """examples/host_min.c (C ABI, no Python in the process) prints the same events, statistics and extras as the ctypes
    mirror computes for the same input."""
    import subprocess
    exe = _build_host_min(tmp_path)
    m, seed = 500, 7
    r = subprocess.run([exe, str(m), str(seed)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = r.stdout.splitlines()
    parent, is_leaf, chars, sn, lab = _host_min_inputs()
    hc = pb.HomoplasyCounter(num_replicates=m, seed=seed)
    hc.set_tree(parent, is_leaf)
    hc.set_phenos(sn, lab)
    hc.set_seg_sites(chars)
    ev, cc = hc.count_all_homoplasy_events()
    st, mt = hc.assoc()
    ep, fp = hc.extras()
    assert out[0] == "upload compact=1 all_valid=0"
    assert out[1] == "classes mono=%d bi=%d tri=%d quad=%d none=%d raw_events=%d" % (cc["mono"], cc["bi"], cc["tri"], cc["quad"],
                                                                                     cc["no_allele"], cc["raw_events"])
    site_lines = [l for l in out if l.startswith("site ")]
    assert site_lines == ["site %d %s %s %d %d %d %d" % (e["segsite_id"], chr(e["allele1"]), chr(e["allele2"]), e["a1_count"], e["a2_count"],
                                                         e["a1_extant"], e["a2_extant"]) for e in ev]
    stat_lines = [l for l in out if l.startswith("stat ")]
    assert len(stat_lines) == len(st) and len(st) > 5
    for l, t, e2, f in zip(stat_lines, st, ep, fp):
        want = ("stat %d [%d, %d, %d, %d] r=%d,%d rmaxT=%d,%d obs_p=%.17g,%.17g pointwise=%.17g,%.17g exact=%.17g,%.17g fisher=%.17g"
                % (t["segsite_id"], *t["obs_counts"], t["r_a1"], t["r_a2"], t["r_maxT_a1"], t["r_maxT_a2"], t["obs_p_a1"], t["obs_p_a2"],
                   t["pointwise_a1"], t["pointwise_a2"], e2[0], e2[1], f))
        assert l.replace("-nan", "nan") == want.replace("-nan", "nan"), (l, want)
    assert out[-2] == "maxT[0]=%.17g maxT[m-1]=%.17g" % (mt[0], mt[-1])
    assert out[-1].startswith("plane_mode=2 ")      # the 'N' calls need the V plane: three planes in HBM
    hc.close()
