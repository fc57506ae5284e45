"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/liboracle.so, the CPU restatement of the reference's live
homoplasy-counting path (see oracle/oracle.cpp for the file:line map into
/root/reference/src/Homoplasy_Counter.java).  Importable only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package `poutine_b200` never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SITE_DTYPE = np.dtype([("segsite_id", "<i4"), ("physical_pos", "<i4"), ("allele1", "<i4"), ("allele2", "<i4"),
                       ("a1_count", "<i4"), ("a2_count", "<i4"), ("a1_extant", "<i4"), ("a2_extant", "<i4")])
STAT_DTYPE = np.dtype([("event_index", "<i4"), ("segsite_id", "<i4"), ("a1_in_use", "<i4"), ("a2_in_use", "<i4"),
                       ("obs_counts", "<i4", (4,)), ("r_a1", "<i4"), ("r_a2", "<i4"), ("r_maxT_a1", "<i4"),
                       ("r_maxT_a2", "<i4"), ("obs_p_a1", "<f8"), ("obs_p_a2", "<f8"), ("pointwise_a1", "<f8"),
                       ("pointwise_a2", "<f8"), ("familywise_a1", "<f8"), ("familywise_a2", "<f8")])
assert SITE_DTYPE.itemsize == 32 and STAT_DTYPE.itemsize == 96


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle.cpp", "commons_math_port.h", "fastmath_tables.inc")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        d, i32, i64, u64, vp = C.c_double, C.c_int32, C.c_int64, C.c_uint64, C.c_void_p
        for name in ("or_fm_exp", "or_fm_log", "or_fm_log1p", "or_log_gamma"):
            getattr(L, name).restype = d
            getattr(L, name).argtypes = [d]
        L.or_log_beta.restype = d; L.or_log_beta.argtypes = [d, d]
        L.or_regularized_beta.restype = d; L.or_regularized_beta.argtypes = [d, d, d]
        L.or_binomial_test.restype = d; L.or_binomial_test.argtypes = [C.c_int, C.c_int, d]
        L.or_binomial_table.restype = None; L.or_binomial_table.argtypes = [C.c_int, d, vp]
        L.or_java_shuffle_perms.restype = None; L.or_java_shuffle_perms.argtypes = [i64, i64, i32, vp]
        L.or_java_random_ints.restype = None; L.or_java_random_ints.argtypes = [i64, i32, i32, vp]
        L.or_philox_labels.restype = None; L.or_philox_labels.argtypes = [u64, u64, i32, i32, i32, vp]
        L.or_create.restype = vp; L.or_create.argtypes = [i32, vp, i32, vp]
        L.or_destroy.restype = None; L.or_destroy.argtypes = [vp]
        L.or_count_all_homoplasy_events.restype = i64; L.or_count_all_homoplasy_events.argtypes = [vp, i64, vp, vp]
        L.or_count_all_homoplasy_events_mt.restype = i64
        L.or_count_all_homoplasy_events_mt.argtypes = [vp, i64, vp, vp, C.c_int]
        L.or_get_events.restype = None; L.or_get_events.argtypes = [vp, vp]
        L.or_get_class_counts.restype = None; L.or_get_class_counts.argtypes = [vp, vp]
        L.or_get_mrca.restype = i64; L.or_get_mrca.argtypes = [vp, i64, C.c_int, vp, i64]
        L.or_assoc.restype = i64
        L.or_assoc.argtypes = [vp, i32, vp, vp, i64, i32, C.c_int, vp, u64, C.c_int]
        L.or_get_stats.restype = None; L.or_get_stats.argtypes = [vp, vp]
        L.or_get_maxT.restype = None; L.or_get_maxT.argtypes = [vp, vp, vp]
        L.or_get_p_success.restype = d; L.or_get_p_success.argtypes = [vp]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def binomial_test(n, k, p):
    """BinomialTest.binomialTest(n, k, p, GREATER_THAN) — HC.java:3050."""
    return lib().or_binomial_test(int(n), int(k), float(p))


def binomial_table(n_max, p):
    out = np.empty((n_max + 1) * (n_max + 2) // 2, dtype=np.float64)
    lib().or_binomial_table(int(n_max), float(p), _p(out))
    return out


def java_shuffle_perms(base_seed, m, P):
    out = np.empty((m, P), dtype=np.uint32)
    lib().or_java_shuffle_perms(int(base_seed), int(m), int(P), _p(out))
    return out


def java_random_ints(seed, n, bound=0):
    out = np.empty(n, dtype=np.int32)
    lib().or_java_random_ints(int(seed), int(n), int(bound), _p(out))
    return out


def philox_labels(seed, rep, P, n_case, n_ctrl):
    out = np.empty(P, dtype=np.uint8)
    lib().or_philox_labels(int(seed), int(rep), int(P), int(n_case), int(n_ctrl), _p(out))
    return out


class Oracle:
    """Mirror of the two reference entry points: count_all_homoplasy_events (HC.java:1402) and
    assoc (HC.java:2029)."""

    def __init__(self, parent, leaf_nodes):
        self.parent = np.ascontiguousarray(parent, dtype=np.int32)
        self.leaf_nodes = np.ascontiguousarray(leaf_nodes, dtype=np.int32)
        self._h = lib().or_create(len(self.parent), _p(self.parent), len(self.leaf_nodes), _p(self.leaf_nodes))
        self.n_events = 0
        self.m = 0

    def __del__(self):
        if getattr(self, "_h", None):
            lib().or_destroy(self._h)
            self._h = None

    def count_all_homoplasy_events(self, seq, physical_poss=None, n_threads=1):
        """seq: (N, S) array of bytes/chars (node-major)."""
        seq = np.ascontiguousarray(seq)
        assert seq.ndim == 2 and seq.shape[0] == len(self.parent) and seq.dtype.itemsize == 1
        S = seq.shape[1]
        pp = None if physical_poss is None else np.ascontiguousarray(physical_poss, dtype=np.int32)
        self.n_events = lib().or_count_all_homoplasy_events_mt(self._h, S, _p(seq), _p(pp), int(n_threads))
        ev = np.empty(self.n_events, dtype=SITE_DTYPE)
        lib().or_get_events(self._h, _p(ev))
        cc = np.empty(5, dtype=np.int64)
        lib().or_get_class_counts(self._h, _p(cc))
        return ev, dict(mono=int(cc[0]), bi=int(cc[1]), tri=int(cc[2]), quad=int(cc[3]), other=int(cc[4]))

    def mrca(self, event_index, allele):
        n = lib().or_get_mrca(self._h, event_index, allele, None, 0)
        out = np.empty(n, dtype=np.int32)
        lib().or_get_mrca(self._h, event_index, allele, _p(out), n)
        return out

    def assoc(self, sample_node, label, m, min_hcount=1, perms=None, seed=0, n_threads=1):
        sample_node = np.ascontiguousarray(sample_node, dtype=np.int32)
        label = np.ascontiguousarray(label, dtype=np.uint8)
        mode = 0 if perms is not None else 1
        if perms is not None:
            perms = np.ascontiguousarray(perms, dtype=np.uint32)
            assert perms.shape == (m, len(label))
        n_inf = lib().or_assoc(self._h, len(label), _p(sample_node), _p(label), int(m), int(min_hcount), mode,
                               _p(perms), int(seed), int(n_threads))
        if n_inf < 0:
            return None, None, None
        st = np.empty(n_inf, dtype=STAT_DTYPE)
        lib().or_get_stats(self._h, _p(st))
        mt_s = np.empty(m, dtype=np.float64)
        mt_u = np.empty(m, dtype=np.float64)
        lib().or_get_maxT(self._h, _p(mt_s), _p(mt_u))
        return st, mt_s, mt_u

    @property
    def p_success(self):
        return lib().or_get_p_success(self._h)
