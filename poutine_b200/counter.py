"""Host-side mirror of the two reference entry points this library replaces.

    Homoplasy_Counter.count_all_homoplasy_events(tree, seg_sites, physical_poss)   HC.java:1402
    Homoplasy_Counter.assoc(all_events, phenos)                                    HC.java:2029

Same names, argument meaning and error behaviour as the reference (which has no plugin API of its
own — both are private methods called from call(), HC.java:299 / 314).  All compute happens in
libpoutine_b200.so on the GPU; this module only flattens/packs inputs and shapes outputs.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import SITE_DTYPE, STAT_DTYPE, PbClassCounts, PbConfig, PbTimings, ptr


class PoutineError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"[pb status {status}] {msg}")
        self.status = status


class NoInformativeSites(PoutineError):
    """'min_hcount = .. is set too high and has filtered out all segregating sites' (HC.java:3866-3869);
    the reference prints that and calls System.exit(-1)."""


def pack_chars(seq: np.ndarray, pitch_words: int | None = None):
    """seg_sites chars (N, S) -> planes (B0, B1, V) uint64[N][W]  (host packer, csrc/packer.cpp)."""
    seq = np.ascontiguousarray(seq)
    assert seq.ndim == 2 and seq.dtype.itemsize == 1
    N, S = seq.shape
    W = (S + 63) // 64
    pw = pitch_words or W
    B0 = np.empty((N, pw), dtype=np.uint64)
    B1 = np.empty((N, pw), dtype=np.uint64)
    V = np.empty((N, pw), dtype=np.uint64)
    rc = _lib.lib().pb_pack_chars(ptr(seq), N, S, seq.strides[0], ptr(B0), ptr(B1), ptr(V), pw)
    if rc != 0:
        raise PoutineError(rc, "pb_pack_chars: invalid arguments")
    return B0, B1, V


def pack_chars_biallelic(seq: np.ndarray, pitch_words: int | None = None, n_threads=0):
    """seg_sites chars (N, S) -> (X, colmask) for put_planes_biallelic(X, None, colmask) in ONE pass, or None when the
    alignment does not qualify (some genotype is not an upper-case ACGT, or a site carries three bases): then use
    pack_chars [+ compact_planes].  Host packer csrc/fasta.cpp pb_pack_chars_biallelic."""
    seq = np.ascontiguousarray(seq)
    assert seq.ndim == 2 and seq.dtype.itemsize == 1
    N, S = seq.shape
    W = (S + 63) // 64
    pw = pitch_words or W
    X = np.empty((N, pw), dtype=np.uint64)
    colmask = np.empty((W, 4), dtype=np.uint64)
    rc = _lib.lib().pb_pack_chars_biallelic(ptr(seq), N, S, seq.strides[0], ptr(X), pw, ptr(colmask), int(n_threads))
    if rc < 0:
        raise PoutineError(rc, "pb_pack_chars_biallelic: invalid arguments")
    return (X[:, :W], colmask) if rc == 1 else None


def compact_planes(B0, B1, V, n_sites, X=None, n_threads=0):
    """Full planes of a column tile -> (X, colmask, all_valid) for put_planes_biallelic, or None when some site of the
    tile carries three or more distinct bases over all nodes (host helper pb_compact_planes, csrc/packer.cpp)."""
    for a in (B0, B1, V):
        assert a.dtype == np.uint64 and a.ndim == 2 and a.strides[1] == 8
    assert B0.shape == B1.shape == V.shape and B0.strides == B1.strides == V.strides
    N, W = B0.shape[0], (n_sites + 63) // 64
    assert B0.shape[1] >= W
    if X is None:
        X = np.empty((N, W), dtype=np.uint64)
    assert X.dtype == np.uint64 and X.shape[0] == N and X.shape[1] >= W and X.strides[1] == 8
    colmask = np.empty((W, 4), dtype=np.uint64)
    av = C.c_int32(0)
    rc = _lib.lib().pb_compact_planes(ptr(B0), ptr(B1), ptr(V), N, int(n_sites), B0.strides[0] // 8, ptr(X), X.strides[0] // 8,
                                      ptr(colmask), C.byref(av), int(n_threads))
    if rc < 0:
        raise PoutineError(rc, "pb_compact_planes: invalid arguments")
    return (X[:, :W], colmask, bool(av.value)) if rc == 1 else None


class HomoplasyCounter:
    """One GPU context.  Usage mirrors Homoplasy_Counter.call() (HC.java:296-314):

        hc = HomoplasyCounter(num_replicates=100000, min_hcount=1)
        hc.set_tree(parent, is_leaf)
        hc.set_planes(B0, B1, V, n_sites)          # or set_seg_sites(chars)
        all_events, class_counts = hc.count_all_homoplasy_events()
        hc.set_phenos(sample_node, label)
        stats, maxT_sorted = hc.assoc()
    """

    def __init__(self, num_replicates=100000, min_hcount=1, device=0, seed=0, replay=False, site_offset=0,
                 maxt_tau=0.0, force_general_null=False, full_planes=False):
        self._h = None
        L = _lib.lib()
        cfg = PbConfig(device=device, mode=_lib.PB_MODE_REPLAY if replay else _lib.PB_MODE_PHILOX,
                       n_replicates=int(num_replicates), min_hcount=int(min_hcount), seed=int(seed),
                       site_offset=int(site_offset), maxt_tau=float(maxt_tau),
                       flags=(_lib.PB_FLAG_FORCE_GENERAL_NULL if force_general_null else 0) |
                             (_lib.PB_FLAG_FULL_PLANES if full_planes else 0))
        h = C.c_void_p()
        rc = L.pb_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise PoutineError(rc, L.pb_last_error(None).decode())
        self._h = h
        self.m = int(num_replicates)
        self.n_sites = 0
        self.n_nodes = 0
        self.n_samples = 0
        self.n_informative = 0
        self._pinned = {}   # name -> (array, handle): page-locked result buffers, reused across calls

    _PREFIX = "pb_"

    def _fn(self, name):
        """C entry point `name` of this object's family (pb_* for one context, pb_group_* for a group)"""
        return getattr(_lib.lib(), self._PREFIX + name)

    def _last_error(self):
        return _lib.lib().pb_last_error(self._h).decode()

    def close(self):
        if self._h is not None:
            if getattr(self, "_owned", True):
                self._fn("destroy")(self._h)
            self._h = None
        for _, p in getattr(self, "_pinned", {}).values():
            _lib.pinned_free(p)
        self._pinned = {}

    def _result_buffer(self, name, n, dtype):
        """page-locked (cudaHostAlloc) result buffer of >= n records; D2H into pageable memory is several
        times slower.  The returned array is a VIEW that the next call of the same entry point overwrites."""
        cur = self._pinned.get(name)
        if cur is None or len(cur[0]) < n:
            registered = name == "sites" and getattr(self, "_expect_sites", False)
            if registered:          # the library must not keep a pointer into the buffer that is about to go
                self._expect_sites = False
                self._check(self._fn("expect_site_output")(self._h, None))
            if cur is not None:
                _lib.pinned_free(cur[1])
            cap = max(int(n), 1)
            self._pinned[name] = _lib.pinned_empty((cap,), dtype)
            if registered:
                self._expect_sites = True
                self._check(self._fn("expect_site_output")(self._h, ptr(self._pinned[name][0])))
        return self._pinned[name][0][:n]

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self._last_error()
            if rc == _lib.PB_ERR_NO_INFORMATIVE:
                raise NoInformativeSites(rc, msg)
            raise PoutineError(rc, msg)

    # ---- inputs --------------------------------------------------------------------------------
    def set_tree(self, parent, is_leaf):
        parent = np.ascontiguousarray(parent, dtype=np.int32)
        is_leaf = np.ascontiguousarray(is_leaf, dtype=np.uint8)
        self._check(self._fn("set_tree")(self._h, len(parent), ptr(parent), ptr(is_leaf)))
        self.n_nodes = len(parent)

    def set_phenos(self, sample_node, label):
        sample_node = np.ascontiguousarray(sample_node, dtype=np.int32)
        label = np.ascontiguousarray(label, dtype=np.uint8)
        self._check(self._fn("set_labels")(self._h, len(label), ptr(sample_node), ptr(label)))
        self.n_samples = len(label)

    def set_sites(self, n_sites):
        self._check(self._fn("set_sites")(self._h, int(n_sites)))
        self.n_sites = int(n_sites)
        if getattr(self, "_expect_sites", False):      # a new site count drops the registration: renew it with a buffer of the new size
            self.expect_site_output()

    def put_planes(self, B0, B1, V, word_begin=0, asynchronous=False):
        """Upload a column tile of the node-major planes (any row pitch).  asynchronous=True: the copies are queued and the
        arrays must stay alive and unchanged until upload_sync()."""
        for a in (B0, B1, V):
            assert a.dtype == np.uint64 and a.ndim == 2 and a.shape[0] == self.n_nodes and a.strides[1] == 8
        assert B0.strides == B1.strides == V.strides
        fn = self._fn("put_planes_async" if asynchronous else "put_planes")
        self._check(fn(self._h, int(word_begin), B0.shape[1], B0.strides[0] // 8, ptr(B0), ptr(B1), ptr(V)))
        if asynchronous:
            self._inflight = getattr(self, "_inflight", []) + [(B0, B1, V)]

    def upload_sync(self):
        """wait until every asynchronous upload has left the host buffers"""
        self._check(self._fn("upload_sync")(self._h))
        self._inflight = []

    def expect_site_output(self, enable=True):
        """Before a streamed upload (after set_sites): tell the library where count_all_homoplasy_events will want its per-site
        records — this object's page-locked `sites` buffer — so that tiles whose sweep and tail ran underneath the upload send
        their records back underneath it too (pb_expect_site_output / pb_group_expect_site_output).  Same results either way."""
        buf = self._result_buffer("sites", self.n_sites, SITE_DTYPE) if enable else None
        self._check(self._fn("expect_site_output")(self._h, ptr(buf)))
        self._expect_sites = bool(enable)

    def put_planes_biallelic(self, X, V, colmask, word_begin=0, asynchronous=False):
        """Compact upload of a column tile whose sites carry at most two bases (see compact_planes): X, optionally V
        (None = every genotype valid) and the per-word-column site masks; expanded on the device.
        One-plane tiles that arrive in column order, cut at multiples of 256 site words, are swept as they land
        (timings()["eager_words"]).  asynchronous=True: see put_planes."""
        assert X.dtype == np.uint64 and X.ndim == 2 and X.shape[0] == self.n_nodes and X.strides[1] == 8
        if V is not None:
            assert V.dtype == np.uint64 and V.shape == X.shape
            if V.strides != X.strides:       # the C ABI takes one row pitch for both planes
                v2 = np.empty((X.shape[0], X.strides[0] // 8), dtype=np.uint64)[:, :X.shape[1]]
                v2[...] = V
                V = v2
        colmask = np.ascontiguousarray(colmask, dtype=np.uint64)
        assert colmask.shape == (X.shape[1], 4)
        fn = self._fn("put_planes_biallelic_async" if asynchronous else "put_planes_biallelic")
        self._check(fn(self._h, int(word_begin), X.shape[1], X.strides[0] // 8, ptr(X), ptr(V), ptr(colmask)))
        if asynchronous:
            self._inflight = getattr(self, "_inflight", []) + [(X, V, colmask)]

    def set_planes(self, B0, B1, V, n_sites, tile_words=None, compact=False):
        """compact=True: every tile that qualifies goes up as one plane (put_planes_biallelic), the others as three."""
        self.set_sites(n_sites)
        W = (n_sites + 63) // 64
        tw = tile_words or W
        for w0 in range(0, W, tw):
            w1 = min(W, w0 + tw)
            b0, b1, v = B0[:, w0:w1], B1[:, w0:w1], V[:, w0:w1]
            c = compact_planes(b0, b1, v, min(n_sites, w1 * 64) - w0 * 64) if compact else None
            if c is not None:
                self.put_planes_biallelic(c[0], None if c[2] else v, c[1], w0)
            else:
                self.put_planes(b0, b1, v, w0)

    def set_seg_sites(self, chars):
        B0, B1, V = pack_chars(chars)
        self.set_planes(B0, B1, V, chars.shape[1])

    # ---- the two reference entry points ----------------------------------------------------------
    def count_all_homoplasy_events(self, fetch_sites=True, compact=True):
        """-> (all_events, class_counts).  all_events = structured array over the BIALLELIC sites in
        site order (the reference's ArrayList<Homoplasy_Events>), fields as Homoplasy_Events.
        compact=False returns the C ABI's own output instead: one record per site (site_class tells which
        are biallelic), as a view of a page-locked buffer that the next call overwrites."""
        cc = PbClassCounts()
        sites = self._result_buffer("sites", self.n_sites, SITE_DTYPE) if fetch_sites else None
        self._check(self._fn("run_events")(self._h, ptr(sites), C.byref(cc)))
        counts = dict(no_allele=cc.n_no_allele, mono=cc.n_mono, bi=cc.n_bi, tri=cc.n_tri, quad=cc.n_quad,
                      raw_events=cc.n_raw_events)
        self.class_counts = counts
        self.sites = sites
        if not fetch_sites:
            return None, counts
        if not compact:
            return sites, counts
        return sites[sites["site_class"] == 2], counts

    def assoc(self, replay_perms=None, fetch=True, want_maxT=True):
        """-> (stats, maxT_sorted).  stats = one record per homoplasically informative site
        (Binomial_Test_Stat + obs_counts)."""
        L = _lib.lib()
        tm = self.timings()
        cap = max(int(tm["n_informative"]), 1)
        stats = self._result_buffer("stats", cap, STAT_DTYPE) if fetch else None
        maxT = self._result_buffer("maxT", self.m, np.float64) if (fetch and want_maxT) else None
        n_inf = C.c_int64(0)
        perms = None
        if replay_perms is not None:
            perms = np.ascontiguousarray(replay_perms, dtype=np.uint32)
            assert perms.shape == (self.m, self.n_samples)
        self._check(self._fn("run_assoc")(self._h, ptr(perms), ptr(stats), cap, C.byref(n_inf), ptr(maxT)))
        self.n_informative = n_inf.value
        return (stats[:n_inf.value] if fetch else None), maxT

    def count_all_homoplasy_events_and_assoc(self, replay_perms=None):
        """The two entry points back to back in one C call (pb_run_events_assoc): the per-site records travel to the host
        underneath the association kernels.  -> (sites (one record per site, see count_all_homoplasy_events(compact=False)),
        class_counts, stats, maxT_sorted); all arrays are views of page-locked buffers the next call overwrites."""
        cc = PbClassCounts()
        sites = self._result_buffer("sites", self.n_sites, SITE_DTYPE)
        stats = self._result_buffer("stats", max(self.n_sites, 1), STAT_DTYPE)      # worst case: every site informative
        maxT = self._result_buffer("maxT", self.m, np.float64)
        n_inf = C.c_int64(0)
        perms = None
        if replay_perms is not None:
            perms = np.ascontiguousarray(replay_perms, dtype=np.uint32)
            assert perms.shape == (self.m, self.n_samples)
        self._check(_lib.lib().pb_run_events_assoc(self._h, ptr(sites), C.byref(cc), ptr(perms), ptr(stats), len(stats), C.byref(n_inf), ptr(maxT)))
        counts = dict(no_allele=cc.n_no_allele, mono=cc.n_mono, bi=cc.n_bi, tri=cc.n_tri, quad=cc.n_quad, raw_events=cc.n_raw_events)
        self.class_counts, self.sites, self.n_informative = counts, sites, n_inf.value
        return sites, counts, stats[:n_inf.value], maxT

    def extras(self):
        """-> (exact_pointwise float64[n_inf, 2], fisher_p float64[n_inf]) after assoc().  Outside the reference's live
        path (parity unpinned): the exact m -> infinity limit of the point-wise estimate under the label-permutation
        null, and R-style two-sided Fisher's exact p of the obs_counts table (dead code in the reference)."""
        n = self.n_informative
        ep = np.empty((n, 2), dtype=np.float64)
        fp = np.empty(n, dtype=np.float64)
        self._check(self._fn("run_extras")(self._h, ptr(ep), ptr(fp)))
        return ep, fp

    # ---- introspection ------------------------------------------------------------------------------
    def timings(self):
        t = PbTimings()
        self._check(self._fn("get_timings")(self._h, C.byref(t)))
        return t.as_dict()

    def stopwatch_start(self):
        self._check(self._fn("stopwatch_mark")(self._h, 0))

    def stopwatch_stop_ms(self):
        """device milliseconds between the two marks (CUDA events on the library's stream)"""
        self._check(self._fn("stopwatch_mark")(self._h, 1))
        ms = C.c_float(0)
        self._check(self._fn("stopwatch_ms")(self._h, C.byref(ms)))
        return float(ms.value)

    def raw_events(self):
        n = C.c_int64(0)
        self._check(_lib.lib().pb_get_raw_events(self._h, 0, None, None, C.byref(n)))
        site = np.empty(n.value, dtype=np.int32)
        node = np.empty(n.value, dtype=np.int32)
        self._check(_lib.lib().pb_get_raw_events(self._h, n.value, ptr(site), ptr(node), C.byref(n)))
        return site, node

    def event_csr(self):
        tm = self.timings()
        n_inf, E = int(tm["n_informative"]), int(tm["n_extant_events"])
        off = np.empty(2 * n_inf + 1, dtype=np.int64)
        idx = np.empty(max(E, 1), dtype=np.int32)
        self._check(_lib.lib().pb_get_event_csr(self._h, ptr(off), ptr(idx), max(E, 1)))
        return off, idx[:E]

    def maxT_unsorted(self):
        out = np.empty(self.m, dtype=np.float64)
        self._check(_lib.lib().pb_get_maxT_unsorted(self._h, ptr(out)))
        return out

    def ptable(self):
        nm = C.c_int32(0)
        self._check(_lib.lib().pb_get_ptable(self._h, C.byref(nm), None, 0))
        total = (nm.value + 1) * (nm.value + 2) // 2
        out = np.empty(total, dtype=np.float64)
        self._check(_lib.lib().pb_get_ptable(self._h, C.byref(nm), ptr(out), total))
        return nm.value, out

    def comm_init(self, id128: bytes, rank: int, world: int):
        """NCCL mode: all-reduce(min) of maxT on the stream (pb_comm_init)"""
        buf = C.create_string_buffer(id128, 128)
        self._check(_lib.lib().pb_comm_init(self._h, buf, rank, world))

    def comm_export(self) -> bytes:
        """peer-memory mode, step 1: this rank's 64-byte handle of its exchange block (pb_comm_export)"""
        buf = C.create_string_buffer(64)
        self._check(_lib.lib().pb_comm_export(self._h, buf))
        return buf.raw

    def comm_connect(self, handles, rank: int, world: int):
        """peer-memory mode, step 2: map every rank's block (handles in rank order); the min over the shards then happens
        inside K4's sort kernel with NVLink loads (pb_comm_connect).  Raises PoutineError(PB_ERR_NCCL) where the peers'
        memory cannot be mapped — fall back to comm_init."""
        blob = b"".join(handles)
        assert len(blob) == 64 * world
        buf = C.create_string_buffer(blob, len(blob))
        self._check(_lib.lib().pb_comm_connect(self._h, buf, rank, world))

    def comm_export_labels(self) -> bytes:
        """label sharing, step 1 (after set_phenos and comm_connect): the 128-byte blob of this rank's label matrices"""
        buf = C.create_string_buffer(128)
        self._check(_lib.lib().pb_comm_export_labels(self._h, buf))
        return buf.raw

    def comm_connect_labels(self, blobs):
        """label sharing, step 2: map every rank's label matrices (blobs in rank order): each rank then draws 1/world of the
        label matrix and pulls the rest from its peers (timings()["gen_shard_frac"])"""
        blob = b"".join(blobs)
        buf = C.create_string_buffer(blob, len(blob))
        self._check(_lib.lib().pb_comm_connect_labels(self._h, buf))

    @classmethod
    def _view(cls, handle, m, n_sites, n_nodes, n_samples):
        """a non-owning wrapper around a context that belongs to a group (introspection only)"""
        o = cls.__new__(cls)
        o._h = C.c_void_p(handle)
        o._owned = False
        o.m, o.n_sites, o.n_nodes, o.n_samples, o.n_informative, o._pinned = m, n_sites, n_nodes, n_samples, 0, {}
        return o


class HomoplasyCounterGroup(HomoplasyCounter):
    """All GPUs of one box behind one host thread (pb_group): the same calls as HomoplasyCounter, the library shards the
    sites over `devices`, runs each shard on its own GPU and merges the outputs (alignment-global segsite_id / site_index).
    Mirrors the single-process Homoplasy_Counter.call() (HC.java:230-319) on a multi-GPU box."""

    _PREFIX = "pb_group_"

    def __init__(self, devices, num_replicates=100000, min_hcount=1, seed=0, replay=False, site_offset=0, maxt_tau=0.0,
                 force_general_null=False, full_planes=False):
        self._h = None
        L = _lib.lib()
        cfg = PbConfig(device=0, mode=_lib.PB_MODE_REPLAY if replay else _lib.PB_MODE_PHILOX, n_replicates=int(num_replicates),
                       min_hcount=int(min_hcount), seed=int(seed), site_offset=int(site_offset), maxt_tau=float(maxt_tau),
                       flags=(_lib.PB_FLAG_FORCE_GENERAL_NULL if force_general_null else 0) |
                             (_lib.PB_FLAG_FULL_PLANES if full_planes else 0))
        dev = np.ascontiguousarray(devices, dtype=np.int32)
        h = C.c_void_p()
        rc = L.pb_group_create(C.byref(cfg), ptr(dev), len(dev), C.byref(h))
        if rc != 0:
            raise PoutineError(rc, L.pb_group_last_error(None).decode())
        self._h = h
        self.devices = [int(d) for d in dev]
        self.m = int(num_replicates)
        self.n_sites = self.n_nodes = self.n_samples = self.n_informative = 0
        self._pinned = {}

    def _last_error(self):
        return _lib.lib().pb_group_last_error(self._h).decode()

    def shard_range(self, i):
        a, b = C.c_int64(0), C.c_int64(0)
        self._check(_lib.lib().pb_group_shard_range(self._h, int(i), C.byref(a), C.byref(b)))
        return a.value, b.value

    def shard(self, i):
        """shard i's context as a non-owning HomoplasyCounter (timings, raw_events, event_csr, maxT_unsorted ... of that GPU)"""
        h = _lib.lib().pb_group_ctx(self._h, int(i))
        assert h
        s0, s1 = self.shard_range(i)
        return HomoplasyCounter._view(h, self.m, s1 - s0, self.n_nodes, self.n_samples)

    def maxT_unsorted(self):
        return self.shard(0).maxT_unsorted()       # pooled over the shards: identical on every GPU

    def ptable(self):
        return self.shard(0).ptable()

    def raw_events(self):
        raise NotImplementedError("per shard: group.shard(i).raw_events()")

    def event_csr(self):
        raise NotImplementedError("per shard: group.shard(i).event_csr()")

    def comm_init(self, *a):
        raise PoutineError(-1, "a group connects its own contexts")

    comm_export = comm_connect = comm_export_labels = comm_connect_labels = comm_init


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = _lib.lib().pb_comm_unique_id(buf)
    if rc != 0:
        raise PoutineError(rc, _lib.lib().pb_last_error(None).decode())
    return buf.raw
