"""ctypes binding of libpoutine_b200.so (the C ABI in include/poutine_b200.h).

The library is built in-tree by `poutine_b200.build()` / `__graft_entry__.build()`.  There is no
Python or CPU fallback: if the shared object is missing, or no sm_100 GPU is present when a
context is created, the error is raised to the caller.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PB_LIB") or os.path.join(_HERE, "libpoutine_b200.so")   # PB_LIB: experiment builds
_LIB = None

PB_MODE_PHILOX, PB_MODE_REPLAY = 0, 1
PB_FLAG_FORCE_GENERAL_NULL, PB_FLAG_KEEP_RAW_EVENTS, PB_FLAG_FULL_PLANES = 1, 2, 4
PB_PLANES_X, PB_PLANES_FULL = 1, 2
PB_ERR_NO_INFORMATIVE = -3

SITE_DTYPE = np.dtype([("segsite_id", "<i4"), ("site_class", "<i4"), ("allele1", "<i4"), ("allele2", "<i4"),
                       ("a1_count", "<i4"), ("a2_count", "<i4"), ("a1_extant", "<i4"), ("a2_extant", "<i4")])
STAT_DTYPE = np.dtype([("site_index", "<i4"), ("segsite_id", "<i4"), ("a1_in_use", "<i4"), ("a2_in_use", "<i4"),
                       ("obs_counts", "<i4", (4,)), ("r_a1", "<i4"), ("r_a2", "<i4"), ("r_maxT_a1", "<i4"),
                       ("r_maxT_a2", "<i4"), ("obs_p_a1", "<f8"), ("obs_p_a2", "<f8"), ("pointwise_a1", "<f8"),
                       ("pointwise_a2", "<f8"), ("familywise_a1", "<f8"), ("familywise_a2", "<f8")])
assert SITE_DTYPE.itemsize == 32 and STAT_DTYPE.itemsize == 96


class PbConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("mode", C.c_int32), ("n_replicates", C.c_int64), ("min_hcount", C.c_int32),
                ("reserved0", C.c_int32), ("seed", C.c_uint64), ("site_offset", C.c_int64), ("maxt_tau", C.c_double),
                ("flags", C.c_uint32), ("reserved1", C.c_uint32)]


class PbClassCounts(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("n_no_allele", "n_mono", "n_bi", "n_tri", "n_quad", "n_raw_events")]


class PbTimings(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("ms_events_kernel", "ms_events_total", "ms_tally", "ms_ptable", "ms_null_gen",
                                         "ms_null_count", "ms_null_fallback", "ms_allreduce", "ms_pvalues",
                                         "ms_assoc_total")] + \
               [(n, C.c_int64) for n in ("n_launches", "k1_algorithmic_bytes", "n_informative", "n_alleles_in_use",
                                         "n_extant_events", "n_fast_alleles", "n_maxt_fallback_reps", "plane_mode", "eager_words", "n_swept_rows")] + \
               [("ms_gen_kernel", C.c_float), ("gen_shard_frac", C.c_float), ("tailed_words", C.c_int64), ("early_sites", C.c_int64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


EXPORTS = ["pb_create", "pb_destroy", "pb_last_error", "pb_set_tree", "pb_set_labels", "pb_set_sites", "pb_put_planes",
           "pb_put_planes_biallelic", "pb_compact_planes",
           "pb_run_events", "pb_run_assoc", "pb_run_extras", "pb_get_sites", "pb_get_raw_events", "pb_get_event_csr",
           "pb_get_maxT_unsorted", "pb_get_ptable", "pb_get_timings", "pb_stopwatch_mark", "pb_stopwatch_ms", "pb_comm_unique_id", "pb_comm_init",
           "pb_pack_chars", "pb_host_alloc", "pb_host_free", "pb_version",
           "pb_fasta_open", "pb_fasta_close", "pb_fasta_num_records", "pb_fasta_record", "pb_fasta_pack_tile", "pb_fasta_last_error",
           "pb_comm_export", "pb_comm_connect", "pb_comm_export_labels", "pb_comm_connect_labels", "pb_fasta_pack_tile_biallelic", "pb_pack_chars_biallelic",
           "pb_run_events_assoc", "pb_put_planes_async", "pb_put_planes_biallelic_async", "pb_upload_sync", "pb_expect_site_output",
           "pb_group_create", "pb_group_destroy", "pb_group_last_error", "pb_group_size", "pb_group_ctx", "pb_group_set_tree",
           "pb_group_set_labels", "pb_group_set_sites", "pb_group_shard_range", "pb_group_put_planes", "pb_group_put_planes_biallelic",
           "pb_group_run_events", "pb_group_run_assoc", "pb_group_run_extras", "pb_group_get_timings", "pb_group_stopwatch_mark",
           "pb_group_stopwatch_ms", "pb_group_put_planes_async", "pb_group_put_planes_biallelic_async", "pb_group_upload_sync",
           "pb_group_expect_site_output"]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libpoutine_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    src = os.path.join(_HERE, "csrc")
    if force:
        subprocess.run(["make", "-C", src, "clean"], check=True, capture_output=True)
    r = subprocess.run(["make", "-C", src, "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libpoutine_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing — run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
        L.pb_create.argtypes = [C.POINTER(PbConfig), C.POINTER(vp)]
        L.pb_destroy.argtypes = [vp]; L.pb_destroy.restype = None
        L.pb_last_error.argtypes = [vp]; L.pb_last_error.restype = C.c_char_p
        L.pb_set_tree.argtypes = [vp, i32, vp, vp]
        L.pb_set_labels.argtypes = [vp, i32, vp, vp]
        L.pb_set_sites.argtypes = [vp, i64]
        L.pb_put_planes.argtypes = [vp, i64, i64, i64, vp, vp, vp]
        L.pb_put_planes_biallelic.argtypes = [vp, i64, i64, i64, vp, vp, vp]
        L.pb_compact_planes.argtypes = [vp, vp, vp, i64, i64, i64, vp, i64, vp, C.POINTER(i32), i32]
        L.pb_run_events.argtypes = [vp, vp, C.POINTER(PbClassCounts)]
        L.pb_run_assoc.argtypes = [vp, vp, vp, i64, C.POINTER(i64), vp]
        L.pb_run_events_assoc.argtypes = [vp, vp, C.POINTER(PbClassCounts), vp, vp, i64, C.POINTER(i64), vp]
        L.pb_run_extras.argtypes = [vp, vp, vp]
        L.pb_get_sites.argtypes = [vp, vp]
        L.pb_get_raw_events.argtypes = [vp, i64, vp, vp, C.POINTER(i64)]
        L.pb_get_event_csr.argtypes = [vp, vp, vp, i64]
        L.pb_get_maxT_unsorted.argtypes = [vp, vp]
        L.pb_get_ptable.argtypes = [vp, C.POINTER(i32), vp, i64]
        L.pb_get_timings.argtypes = [vp, C.POINTER(PbTimings)]
        L.pb_stopwatch_mark.argtypes = [vp, i32]
        L.pb_stopwatch_ms.argtypes = [vp, C.POINTER(C.c_float)]
        L.pb_comm_unique_id.argtypes = [vp]
        L.pb_comm_init.argtypes = [vp, vp, i32, i32]
        L.pb_pack_chars.argtypes = [vp, i64, i64, i64, vp, vp, vp, i64]
        L.pb_fasta_open.argtypes = [C.c_char_p, C.POINTER(vp)]
        L.pb_fasta_close.argtypes = [vp]; L.pb_fasta_close.restype = None
        L.pb_fasta_num_records.argtypes = [vp]; L.pb_fasta_num_records.restype = i64
        L.pb_fasta_record.argtypes = [vp, i64, C.POINTER(vp), C.POINTER(i64), C.POINTER(i64)]
        L.pb_fasta_pack_tile.argtypes = [vp, vp, i64, i64, vp, vp, vp, i64, i32]
        L.pb_fasta_last_error.argtypes = []; L.pb_fasta_last_error.restype = C.c_char_p
        L.pb_fasta_pack_tile_biallelic.argtypes = [vp, vp, i64, i64, i64, vp, i64, vp, i32]
        L.pb_pack_chars_biallelic.argtypes = [vp, i64, i64, i64, vp, i64, vp, i32]
        L.pb_put_planes_async.argtypes = [vp, i64, i64, i64, vp, vp, vp]
        L.pb_put_planes_biallelic_async.argtypes = [vp, i64, i64, i64, vp, vp, vp]
        L.pb_upload_sync.argtypes = [vp]
        L.pb_expect_site_output.argtypes = [vp, vp]
        L.pb_group_expect_site_output.argtypes = [vp, vp]
        L.pb_comm_export.argtypes = [vp, vp]
        L.pb_comm_connect.argtypes = [vp, vp, i32, i32]
        L.pb_comm_export_labels.argtypes = [vp, vp]
        L.pb_comm_connect_labels.argtypes = [vp, vp]
        L.pb_group_create.argtypes = [C.POINTER(PbConfig), vp, i32, C.POINTER(vp)]
        L.pb_group_destroy.argtypes = [vp]; L.pb_group_destroy.restype = None
        L.pb_group_last_error.argtypes = [vp]; L.pb_group_last_error.restype = C.c_char_p
        L.pb_group_size.argtypes = [vp]; L.pb_group_size.restype = i32
        L.pb_group_ctx.argtypes = [vp, i32]; L.pb_group_ctx.restype = vp
        L.pb_group_set_tree.argtypes = [vp, i32, vp, vp]
        L.pb_group_set_labels.argtypes = [vp, i32, vp, vp]
        L.pb_group_set_sites.argtypes = [vp, i64]
        L.pb_group_shard_range.argtypes = [vp, i32, C.POINTER(i64), C.POINTER(i64)]
        L.pb_group_put_planes.argtypes = [vp, i64, i64, i64, vp, vp, vp]
        L.pb_group_put_planes_biallelic.argtypes = [vp, i64, i64, i64, vp, vp, vp]
        L.pb_group_put_planes_async.argtypes = [vp, i64, i64, i64, vp, vp, vp]
        L.pb_group_put_planes_biallelic_async.argtypes = [vp, i64, i64, i64, vp, vp, vp]
        L.pb_group_upload_sync.argtypes = [vp]
        L.pb_group_run_events.argtypes = [vp, vp, C.POINTER(PbClassCounts)]
        L.pb_group_run_assoc.argtypes = [vp, vp, vp, i64, C.POINTER(i64), vp]
        L.pb_group_run_extras.argtypes = [vp, vp, vp]
        L.pb_group_get_timings.argtypes = [vp, C.POINTER(PbTimings)]
        L.pb_group_stopwatch_mark.argtypes = [vp, i32]
        L.pb_group_stopwatch_ms.argtypes = [vp, C.POINTER(C.c_float)]
        L.pb_host_alloc.argtypes = [C.POINTER(vp), u64]
        L.pb_host_free.argtypes = [vp]
        for n in EXPORTS:
            if n not in ("pb_destroy", "pb_last_error", "pb_fasta_close", "pb_fasta_num_records", "pb_fasta_last_error",
                         "pb_group_destroy", "pb_group_last_error", "pb_group_size", "pb_group_ctx"):
                getattr(L, n).restype = C.c_int
        _LIB = L
    return _LIB


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pinned_empty(shape, dtype):
    """numpy array over cudaHostAlloc memory (for streamed uploads); freed with pinned_free."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    rc = lib().pb_host_alloc(C.byref(p), max(n, 1))
    if rc != 0:
        raise MemoryError(lib().pb_last_error(None).decode())
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr, p


def pinned_free(p):
    lib().pb_host_free(p)
