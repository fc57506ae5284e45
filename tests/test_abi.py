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
    assert ctypes.sizeof(_lib.PbTimings) == 10 * 4 + 8 * 8


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
