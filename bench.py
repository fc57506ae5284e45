#!/usr/bin/env python
"""bench.py — sites/sec of the homoplasy hot path (count_all_homoplasy_events + assoc) on B200.

    python bench.py --gpus 1 --steps K --warmup W                      # this repo (CUDA via the C ABI)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                               # CPU port of the reference path

One "step" = one full pass of the hot path over the workload: K1 event detection, K2 tallies,
K3 permutation null (m replicates), K4 p-values.  N = 1 runs BASELINE.json configs[1]
(5,000 leaves x 1,000,000 sites, m = 1e5); at N > 1 every rank holds its own 1M-site shard of the
same tree/phenotypes (weak scaling; sites shard, the one collective is the NCCL all-reduce(min) of
maxT[m] inside pb_run_assoc).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "sites/sec (homoplasy count + null resampling)"
CPU_SAMPLE_SITES = 32768      # ~7 s of work for the 16-thread CPU port at C2 (events are linear in sites, the null in alleles)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for n, v in zip(names, c[4:8]):
                if v == "Active":
                    reasons.add(n)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_reference_pass(tree, chars, sn, lab, m, seed, threads, keep=None):
    """One pass of the CPU port over a site sample: (sites/sec, seconds L4, seconds L5).  oracle/ is test
    infrastructure; this is the one place bench.py may execute it (cpu_baseline / --impl reference).
    keep: dict that receives the port's outputs (events, class counts, statistics) for the parity cross-check."""
    import oracle
    o = oracle.Oracle(tree.parent, tree.leaf_nodes)
    t0 = time.perf_counter()
    ev, cc = o.count_all_homoplasy_events(chars, n_threads=threads)
    t1 = time.perf_counter()
    st, _, _ = o.assoc(sn, lab, m, 1, perms=None, seed=seed, n_threads=threads)
    t2 = time.perf_counter()
    if keep is not None:
        keep.update(events=ev, class_counts=cc, stats=st)
    return chars.shape[1] / (t2 - t0), t1 - t0, t2 - t1


def exact_null_check(hc, st, m):
    """the exact point-wise null (pb_run_extras: the m -> infinity limit of r/m under a uniform label
    permutation, by hypergeometric sums — no random numbers involved) as a validator of the Monte-Carlo path that was just
    timed.  Per allele in use, r ~ Binomial(m, p_exact) if K3 (label generator + replicate counts) is right: z = (r - m p) /
    sqrt(m p (1-p)).  Fails the run if any allele leaves the 8-sigma + 4-count band (Poisson-safe for millions of alleles with
    tiny p) or if the mean of z^2 (expectation exactly 1, whatever the discreteness) is off; the 6-sigma count is reported."""
    t0 = time.perf_counter()
    ep, _ = hc.extras()
    t_dev = time.perf_counter() - t0
    zs, n6, n8 = [], 0, 0
    for a, name in enumerate(("a1", "a2")):
        use = st[name + "_in_use"] != 0
        p = ep[use, a]
        r = st["r_" + name][use].astype(np.float64)
        if np.isnan(p).any():
            return {"checked": False, "error": "exact null is NaN for an allele in use"}
        var = m * p * (1.0 - p)
        dev = np.abs(r - m * p)
        n6 += int((dev > 6.0 * np.sqrt(var) + 2.0).sum())
        n8 += int((dev > 8.0 * np.sqrt(var) + 4.0).sum())
        ok = var > 0
        zs.append(((r[ok] - m * p[ok]) ** 2) / var[ok])
        n8 += int((dev[~ok] > 0).sum())            # p exactly 0 or 1: r must be 0 or m
    z2 = np.concatenate(zs) if zs else np.zeros(0)
    mean_z2 = float(z2.mean()) if len(z2) else None
    passed = n8 == 0 and (mean_z2 is None or 0.8 < mean_z2 < 1.25)
    return {"checked": True, "passed": bool(passed), "alleles": int(len(z2)), "outside_6sigma_plus_2": n6, "outside_8sigma_plus_4": n8,
            "mean_z2": mean_z2, "max_abs_z": float(np.sqrt(z2.max())) if len(z2) else None, "seconds": t_dev,
            "what": "Monte-Carlo r of the last timed step against the exact label-permutation null (pb_run_extras, hypergeometric "
                    "sums; parity unpinned: the reference is Monte-Carlo only, HC.java:4812-4838); alleles share the replicates, so "
                    "mean_z2 has expectation 1 but not the variance of independent draws"}


def parity_check(gpu_sites, gpu_stats, ref, n):
    """The GPU's per-site output for the first n sites of the timed input against the CPU port's output on the same sites
    (same Philox seed, same m): Homoplasy_Events fields and, per informative site, in-use flags, observed tallies, observed
    p-values, r counts and point-wise estimates — bit for bit.  (The family-wise columns pool the per-replicate minimum over
    ALL sites of the run, so a slice cannot reproduce them; tests/ cover them on whole inputs.)  -> list of mismatches."""
    bad = []
    g = gpu_sites[:n]
    g = g[g["site_class"] == 2]
    r = ref["events"]
    if len(g) != len(r):
        return ["biallelic sites: gpu %d, cpu port %d" % (len(g), len(r))]
    for f in ("segsite_id", "allele1", "allele2", "a1_count", "a2_count", "a1_extant", "a2_extant"):
        if not np.array_equal(g[f], r[f]):
            bad.append("events." + f)
    gs = gpu_stats[gpu_stats["site_index"] < n]
    rs = ref["stats"]
    if len(gs) != len(rs) or not np.array_equal(gs["segsite_id"], rs["segsite_id"]):
        return bad + ["informative sites: gpu %d, cpu port %d" % (len(gs), len(rs))]
    for f in ("a1_in_use", "a2_in_use", "obs_counts", "r_a1", "r_a2"):
        if not np.array_equal(gs[f], rs[f]):
            bad.append("stats." + f)
    for f in ("obs_p_a1", "obs_p_a2", "pointwise_a1", "pointwise_a2"):
        a, b = gs[f], rs[f]
        ok = ~np.isnan(b)
        if not (np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[ok].view(np.uint64), b[ok].view(np.uint64))):
            bad.append("stats." + f)
    return bad


def workload(args, rank, world):
    from poutine_b200 import synth
    cfg = dict(synth.CONFIGS[args.config])
    S = args.sites or cfg["n_sites"]
    m = args.replicates or cfg["m"]
    tree = synth.yule_tree(cfg["n_leaves"], cfg["seed"])
    sn, lab = synth.simulate_phenotypes(tree, cfg["seed"] + 1000, na_frac=cfg.get("na_frac", 0.0))
    return cfg, S, m, tree, sn, lab


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from poutine_b200 import synth
    cfg, S, m, tree, sn, lab = workload(args, 0, 1)
    n = min(CPU_SAMPLE_SITES, S)
    planes = synth.simulate_planes(tree, n, cfg["seed"] + 2000 + 7919, leaf_gap=cfg.get("leaf_gap", False),
                                   multiallelic_frac=cfg.get("multiallelic_frac", 0.0),
                                   third_allele_internal_frac=cfg.get("third_allele_internal_frac", 0.0))
    chars = synth.planes_to_chars(*planes, 0, n)
    threads = host_cores()
    for _ in range(args.warmup):
        cpu_reference_pass(tree, chars, sn, lab, max(m // 50, 1), 1, threads)   # short warm-up passes (page-in, thread spin-up)
    t0 = time.perf_counter()
    l4 = l5 = 0.0
    for _ in range(args.steps):
        _, a, b = cpu_reference_pass(tree, chars, sn, lab, m, 1, threads)
        l4 += a; l5 += b
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = ("%d-site slice of the %s workload at the full m=%d per step; C++ port of HC.java:1402-1511 + 3013-3186 "
              "(oracle/oracle.cpp) on %d threads (site loop threaded although the reference runs it on one thread); "
              "no JVM in this image, so the Java reference itself cannot run") % (n, args.config, m, threads)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "sites/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64 bit planes / int32 counts / f64 p-values", "data": "synthetic",
            # the b200 arm's config at this N (the workload named is the same; `sample` says what the CPU actually ran)
            "config": dict(bench_config(args, cfg, S, m, tree, max(int(args.gpus), 1)), sites_measured=n,
                           note="the CPU arm runs a %d-site slice of the workload named here, at the full m (see cpu_baseline.sample)" % n),
            "cpu_baseline": {"value": value, "unit": "sites/s", "cores": threads, "kind": "port", "sample": sample,
                             "seconds_events": l4 / args.steps, "seconds_assoc": l5 / args.steps},
            "e2e": {"value": value, "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def bench_config(args, cfg, S, m, tree, world, scaling="weak", launch="one process per GPU"):
    return {"workload": "%s: %d-leaf Yule tree (%d nodes) x %d biallelic sites per GPU, binary phenotype (38%% cases), "
                        "m=%d null replicates, Philox mode" % (args.config, tree.n_leaves, tree.n_nodes, S, m),
            "baseline_config": {"C2": "configs[1]", "C3": "configs[2]", "C1": "configs[0]", "C4": "configs[3]", "C5": "configs[4]"}[args.config],
            "sites_per_gpu": S, "sites_total": S * world, "leaves": tree.n_leaves, "nodes": tree.n_nodes, "replicates": m,
            "parallelism": ("site-sharded x%d (%s scaling, %s), min of maxT[m] over the shards" % (world, scaling, launch)) if world > 1 else "single GPU",
            "l2_policy": "inputs larger than L2 (bit planes per GPU: %.2f GB as one plane when every site is biallelic and every genotype "
                         "valid, %.2f GB as three planes otherwise; L2 is 126 MB)"
                         % (tree.n_nodes * ((S + 63) // 64) * 8 / 1e9, 3 * tree.n_nodes * ((S + 63) // 64) * 8 / 1e9)}


def k3_rooflines(tm, parts, m, P, has_miss, n_shards=1):
    """The two resampling kernels against their own ceilings (ncu captures under profiles/r2_k3*):
    K3a (label generator, sampling spec v3) is bound by warp-instruction issue — an issue-rate ceiling has no live byte count,
    the fraction (issue slots busy over the launch) comes from the committed capture; K3b (replicate counting) is bound by L2 -> SM bandwidth: algorithmic bytes =
    sum over swept alleles of n rows x m/8 bytes, against the L2 read bandwidth measured on this GPU
    (tools/microbench/l2_peak -> profiles/r2_l2_peak.json)."""
    def load(name):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return json.load(f)
        except Exception:
            return {}
    l2 = load("r2_l2_peak.json")
    ceil = load("k3_ceilings.json")
    l2_peak = l2.get("l2_read_gbs_16B_per_lane")
    swept_bytes = float(tm.get("n_swept_rows", 0)) / n_shards * (m / 8.0) * (2 if has_miss else 1)   # per GPU (a group sums its shards)
    out = {"k3a_label_generator": {"bound": "warp-instruction issue: xoshiro128++ stream + four multiply-shift draws per 64-bit word, one thread per replicate",
                                   "draws_per_s": (m * P) / (parts["ms_gen_kernel"] / 1e3) if parts.get("ms_gen_kernel") else None,
                                   "ms": parts.get("ms_gen_kernel"),
                                   "frac": ceil.get("k3_gen_kernel", {}).get("issue_frac"),
                                   "frac_busiest_sm": ceil.get("k3_gen_kernel", {}).get("issue_frac_busiest_sm"),
                                   "frac_source": "ncu smsp__issue_active (profiles/r2_final_k3a_gen_ncu.txt, via profiles/k3_ceilings.json)"},
           "k3b_replicate_count": {"bound": "l2", "unit": "GB/s", "algorithmic_bytes": swept_bytes, "ms": parts.get("ms_null_count"),
                                   "achieved": swept_bytes / (parts["ms_null_count"] / 1e3) / 1e9 if parts.get("ms_null_count") else None,
                                   "peak": l2_peak, "peak_source": "measured: tools/microbench/l2_peak (L2-resident stream, 16 B per lane, L1 bypassed)"}}
    k = out["k3b_replicate_count"]
    k["frac"] = (k["achieved"] / k["peak"]) if (k["achieved"] and k["peak"]) else None
    return out


def run_b200(args):
    import torch
    import poutine_b200 as pb
    from poutine_b200 import _lib, shard as shardmod, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not os.path.exists(pb.LIB_PATH):
        pb.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this path has no CPU fallback (use --impl reference for the CPU port)")
    inproc = args.inproc and world == 1 and args.gpus > 1     # one process, one host thread, pb_group over args.gpus devices
    n_shards = args.gpus if inproc else 1                     # shards this process holds
    n_gpus = args.gpus if inproc else world                   # GPUs of the whole job
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version banner must not precede the JSON line on stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    cfg, S_cfg, m, tree, sn, lab = workload(args, rank, world)
    strong = args.scaling == "strong" and n_gpus > 1
    N = tree.n_nodes
    # sites: weak = S_cfg per GPU; strong = S_cfg over the whole job, contiguous 64-site-word shards
    if strong:
        ranges = [shardmod.site_range(S_cfg, g, n_gpus) for g in range(n_gpus)]
    else:
        W1 = (S_cfg + 63) // 64
        ranges = [(g * W1 * 64, g * W1 * 64 + S_cfg) for g in range(n_gpus)]
    my = list(range(n_gpus)) if inproc else [rank]
    S_total = sum(b - a for a, b in ranges)
    S_mine = sum(ranges[g][1] - ranges[g][0] for g in my)
    t_gen = time.perf_counter()
    t_compact = 0.0
    shards = []          # per shard of this process: dict(word_begin, S, W, Wp, planes (B0,B1,V views) or None, compact (X, colmask, all_valid) or None)
    pinned = []
    cpu_chars = None
    for i, g in enumerate(my):
        s0, s1 = ranges[g]
        Sg = s1 - s0
        W = (Sg + 63) // 64
        Wp = (W + 15) & ~15                                  # the library's device row pitch: lets pb_put_planes do flat copies
        if inproc and not strong and i > 0:
            shards.append(dict(shards[0], word_begin=s0 // 64))   # weak scaling in one process: every shard uploads the same 1M-site block
            continue
        B0, B1, V = synth.simulate_planes(tree, Sg, cfg["seed"] + 2000 + 7919 * (g + 1), leaf_gap=cfg.get("leaf_gap", False),
                                          multiallelic_frac=cfg.get("multiallelic_frac", 0.0),
                                          third_allele_internal_frac=cfg.get("third_allele_internal_frac", 0.0))
        if rank == 0 and g == 0 and n_gpus == 1:
            cpu_chars = synth.planes_to_chars(B0, B1, V, 0, min(CPU_SAMPLE_SITES, Sg))
        sh = dict(word_begin=s0 // 64, S=Sg, W=W, Wp=Wp, planes=None, compact=None)
        keep_planes = n_shards == 1 and not strong           # the three-plane comparison legs run on one shard per process only
        hp = []
        if keep_planes or args.full_planes_only:
            for a in (B0, B1, V):
                h, p = _lib.pinned_empty((N, Wp), np.uint64)
                h[:, :W] = a
                h[:, W:] = 0
                pinned.append(p)
                hp.append(h[:, :W])
            sh["planes"] = hp
        if not args.full_planes_only:
            hX, pX = _lib.pinned_empty((N, Wp), np.uint64)
            hX[:, W:] = 0
            pinned.append(pX)
            t_c = time.perf_counter()
            c = pb.compact_planes(B0, B1, V, Sg, X=hX)
            t_compact += time.perf_counter() - t_c
            if c is not None and not c[2]:                   # some genotype is not a base: the V plane goes up too
                if not hp:
                    hV, pV = _lib.pinned_empty((N, Wp), np.uint64)
                    hV[:, :W] = V
                    hV[:, W:] = 0
                    pinned.append(pV)
                    hp = [None, None, hV[:, :W]]
                    sh["planes_v_only"] = hp
            sh["compact"] = c
            if c is None and not hp:                         # multi-allelic shard without kept planes: keep them after all
                for a in (B0, B1, V):
                    h, p = _lib.pinned_empty((N, Wp), np.uint64)
                    h[:, :W] = a
                    h[:, W:] = 0
                    pinned.append(p)
                    hp.append(h[:, :W])
                sh["planes"] = hp
        del B0, B1, V
        shards.append(sh)
    t_gen = time.perf_counter() - t_gen
    all_compact = all(sh["compact"] is not None for sh in shards)

    if inproc:
        hc = pb.HomoplasyCounterGroup(list(range(n_gpus)), num_replicates=m, min_hcount=1, seed=20260)
    else:
        hc = pb.HomoplasyCounter(num_replicates=m, min_hcount=1, device=local_rank, seed=20260, site_offset=ranges[rank][0])
    hc.set_tree(tree.parent, tree.is_leaf)
    hc.set_phenos(sn, lab)
    hc.set_sites(S_mine)
    exchange = "none (one GPU)"
    if inproc:
        exchange = "peer memory, fused into K4's sort kernel (pb_group, one process); label matrix drawn 1/%d per shard and gathered over NVLink" % n_gpus
    elif world > 1:
        want = os.environ.get("PB_BENCH_EXCHANGE", "p2p")
        ok = 0
        if want == "p2p":
            try:
                handles = [None] * world
                dist.all_gather_object(handles, hc.comm_export())
                hc.comm_connect(handles, rank, world)
                ok = 1
            except pb.PoutineError as e:
                print("bench.py: rank %d: peer-memory exchange unavailable (%s), falling back to NCCL" % (rank, e), file=sys.stderr)
        t = torch.tensor([float(ok)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if t.item() >= 1:
            exchange = "peer memory (cudaIpc), fused into K4's sort kernel"
            if os.environ.get("PB_BENCH_LABEL_SHARING", "1") != "0":
                # the shards sweep the same label matrix: map the matrices too, each rank draws 1/world of it
                ok2 = 0
                try:
                    blobs = [None] * world
                    dist.all_gather_object(blobs, hc.comm_export_labels())
                    hc.comm_connect_labels(blobs)
                    ok2 = 1
                except pb.PoutineError as e:
                    print("bench.py: rank %d: label sharing unavailable (%s)" % (rank, e), file=sys.stderr)
                t2 = torch.tensor([float(ok2)], dtype=torch.float64, device="cuda")
                dist.all_reduce(t2, op=dist.ReduceOp.MIN)
                if t2.item() >= 1:
                    exchange += "; label matrix drawn 1/%d per rank and gathered over NVLink" % world
                else:
                    raise SystemExit("bench.py: label sharing connected on some ranks only")
        else:
            ids = [pb.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            hc.comm_init(ids[0], rank, world)
            exchange = "NCCL all-reduce(min, uint64) on the stream"

    base_word = 0 if inproc else 0       # a single context takes shard-local word coordinates
    TILE = int(args.upload_tile_words)

    def upload(use_compact=True):
        """H2D of every shard this process holds.  Compact shards go up in column tiles cut at the sweep's block width,
        asynchronously: all GPUs' copy engines run at once, and (one-plane mode) each tile is swept underneath the next
        tile's copy while the label generator runs under the whole upload."""
        for sh in shards:
            wb = sh["word_begin"] if inproc else 0
            c = sh["compact"] if use_compact else None
            if c is not None:
                Xs, cm, all_valid = c
                Vs = None if all_valid else (sh.get("planes") or sh.get("planes_v_only"))[2]
                for w0 in range(0, sh["W"], TILE):
                    w1 = min(sh["W"], w0 + TILE)
                    hc.put_planes_biallelic(Xs[:, w0:w1], None if Vs is None else Vs[:, w0:w1], cm[w0:w1], wb + w0, asynchronous=True)
            else:
                hc.put_planes(*sh["planes"], wb, asynchronous=True)
        hc.upload_sync()

    upload()          # biallelic input with every genotype valid stays ONE plane in HBM (k1x_events_kernel), else three

    def step_resident():
        hc.count_all_homoplasy_events(fetch_sites=False)
        hc.assoc(fetch=False)
        return hc.timings()

    # pb_run_events_assoc (one C call, per-site records copied out underneath the association) measured no faster than the
    # two calls on this box (profiles/r2_v3_bench_fused.json vs _twocalls.json): the default stays the reference's two calls
    fused_call = args.fused_call and not inproc
    e2e_parts = {"put_planes": 0.0, "events": 0.0, "assoc": 0.0}
    e2e_steps = []                                       # wall ms of every end-to-end step on this rank (an outlier shows here)

    def step_e2e(use_compact=True):
        t0 = time.perf_counter()
        upload(use_compact)                              # H2D of this step's inputs (pinned host memory)
        t1 = time.perf_counter()
        if fused_call:                                   # both entry points in one C call: the per-site records leave under the association
            ev, cc, st, mt = hc.count_all_homoplasy_events_and_assoc()
            t2 = t3 = time.perf_counter()
        else:
            ev, cc = hc.count_all_homoplasy_events(compact=False)   # D2H: per-site Homoplasy_Events records (page-locked buffer)
            t2 = time.perf_counter()
            st, mt = hc.assoc()                          # D2H: per-informative-site statistics + sorted maxT
            t3 = time.perf_counter()
        e2e_parts["put_planes"] += 1e3 * (t1 - t0); e2e_parts["events"] += 1e3 * (t2 - t1); e2e_parts["assoc"] += 1e3 * (t3 - t2)
        e2e_steps.append(round(1e3 * (t3 - t0), 3))
        return hc.timings(), ev, st, mt

    for _ in range(args.warmup):
        tm = step_resident()

    # ---- timed: inputs resident in HBM ---------------------------------------------------------------
    try:
        gpu_id = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)   # immune to CUDA_VISIBLE_DEVICES remapping
    except Exception:
        gpu_id = str(local_rank)
    sampler = ClockSampler(gpu_id) if rank == 0 else None
    total_sites = S_total
    peak, peak_src = measured_hbm_peak()
    has_miss = bool((np.asarray(lab) == 2).any())

    def timed_resident():
        """K timed steps with the planes resident in HBM -> (device ms, wall ms, launches, per-part ms, roofline of the sweep kernel)"""
        k1_ms, launches, parts = [], 0, {}
        barrier()
        t0 = time.perf_counter()
        hc.stopwatch_start()
        for _ in range(args.steps):
            tm = step_resident()
            k1_ms.append(tm["ms_events_kernel"])
            launches += int(tm["n_launches"])
            for k in ("ms_events_kernel", "ms_events_total", "ms_tally", "ms_ptable", "ms_null_gen", "ms_null_count",
                      "ms_null_fallback", "ms_allreduce", "ms_pvalues", "ms_assoc_total", "ms_gen_kernel"):
                parts[k] = parts.get(k, 0.0) + tm.get(k, 0.0) / args.steps
        dev_ms = hc.stopwatch_stop_ms()
        barrier()
        wall_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
        dev_ms = max_over_ranks(dev_ms)
        # roofline of the HBM-bound counting sweep: algorithmic bytes / mean launch time, both live.
        # (a group reports the slowest shard's time and the summed bytes: divide by the shard count for the per-GPU figure)
        kernel = "k1x_events_kernel" if tm["plane_mode"] == _lib.PB_PLANES_X else "k1_events_kernel"
        k1_avg = float(np.mean(k1_ms))
        alg_bytes = int(tm["k1_algorithmic_bytes"]) // n_shards
        achieved = alg_bytes / (k1_avg / 1e3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
                tj = json.load(f).get(kernel, {})
                if tj.get("sites") == shards[0]["S"] and tj.get("nodes") == N:
                    traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        roofline = {"kernel": kernel, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": k1_avg, "peak_source": peak_src,
                    "frac_of_nominal_8TBps": achieved / 8000.0,
                    "algorithmic_bytes": "plane bits once (N*W*8 B per plane) + one record per (tree row, site word) holding an event; "
                                         "the per-site output (32 B/site) belongs to the tail kernels, not to this launch",
                    "planes_in_hbm": 1 if kernel == "k1x_events_kernel" else 3}
        return dev_ms, wall_ms, launches, parts, roofline, tm

    dev_ms, wall_ms, launches, parts, roofline, tm = timed_resident()
    value = total_sites * args.steps / (dev_ms / 1e3)
    roofline_k3 = k3_rooflines(tm, parts, m, len(lab), has_miss, n_shards)

    # ---- timed: end to end through the public API with host buffers ------------------------------------
    # pb_expect_site_output: the per-site records of tiles whose sweep and tail ran underneath the upload come back underneath it
    # too (a single context's streaming call; a pb_group merges its shards' records itself)
    early_output = not inproc and not args.no_early_output
    if early_output:
        hc.expect_site_output()
    for _ in range(min(args.warmup, 3)):
        step_e2e()
    barrier()
    for k in e2e_parts:
        e2e_parts[k] = 0.0
    del e2e_steps[:]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tm_e, ev, st, mt = step_e2e()
    barrier()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
    # outputs of the last timed step (views of page-locked buffers the next call overwrites): kept for the parity check
    par_outputs = [("headline", np.array(ev[:CPU_SAMPLE_SITES]), np.array(st))]
    exact_null = None
    if rank == 0 and n_gpus == 1 and not args.no_exact_null:       # validator of the Philox-mode null that was just timed
        exact_null = exact_null_check(hc, st, m)
    e2e_breakdown = {k: v / args.steps for k, v in e2e_parts.items()}
    e2e_step_ms = list(e2e_steps)
    eager_words = int(tm_e.get("eager_words", 0))
    h2d_full = sum(3 * N * sh["Wp"] * 8 for sh in shards)     # what pb_put_planes copies (rows padded to the device pitch)
    h2d = h2d_full
    e2e_full = None
    three_planes = None
    if all_compact:
        h2d = sum((1 if sh["compact"][2] else 2) * N * sh["W"] * 8 + sh["W"] * 32 for sh in shards)
    if all_compact and all(sh["planes"] is not None for sh in shards):
        # the same end-to-end step with the plain three-plane upload, reported next to the headline
        step_e2e(False)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            _, ev3, st3, _ = step_e2e(False)
        barrier()
        full_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
        par_outputs.append(("three_planes", np.array(ev3[:CPU_SAMPLE_SITES]), np.array(st3)))
        e2e_full = {"value": total_sites * args.steps / (full_ms / 1e3), "unit": "sites/s", "ms_per_step": full_ms / args.steps,
                    "h2d_bytes_per_step": h2d_full, "upload": "pb_put_planes (three planes)"}
        # the context now holds three planes (the general storage: gaps, multi-allelic sites): the same resident measurement
        for _ in range(args.warmup):
            step_resident()
        f_dev, f_wall, f_launches, f_parts, f_roof, _ = timed_resident()
        launches += f_launches
        three_planes = {"value": total_sites * args.steps / (f_dev / 1e3), "unit": "sites/s", "ms_per_step": f_dev / args.steps,
                        "roofline": f_roof, "breakdown_ms": f_parts, "e2e": e2e_full,
                        "note": "same workload held as three full planes (what gap-carrying or multi-allelic input uses)"}
    clocks = sampler.stop() if sampler else None      # sampled over all timed regions (resident + end to end)
    d2h = S_mine * _lib.SITE_DTYPE.itemsize + len(st) * _lib.STAT_DTYPE.itemsize + m * 8
    e2e_value = total_sites * args.steps / (e2e_ms / 1e3)
    n_draws = sum_over_ranks(tm["n_alleles_in_use"]) * m

    cpu_baseline = None
    host_packer = None
    parity = {"parity_checked": False, "parity_sites": 0}
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        threads = host_cores()
        ref = {}
        v, l4, l5 = cpu_reference_pass(tree, cpu_chars, sn, lab, m, 20260, threads, keep=ref)
        cpu_baseline = {"value": v, "unit": "sites/s", "cores": threads, "kind": "port",
                        "sample": "%d-site slice of this workload at the full m=%d (C++ port oracle/oracle.cpp; events %.2f s + assoc %.2f s)"
                                  % (cpu_chars.shape[1], m, l4, l5)}
        # the timed run's own output (last end-to-end step: `ev`, `st`) against the port on the same first sites
        n_par = cpu_chars.shape[1]
        bad = []
        for name, ev_k, st_k in par_outputs:
            bad += ["%s: %s" % (name, b) for b in parity_check(ev_k, st_k, ref, n_par)]
        parity = {"parity_checked": True, "parity_sites": int(n_par), "parity_informative_sites": int(len(ref["stats"])),
                  "parity_fields": "Homoplasy_Events fields; in-use flags, obs_counts, obs_p, r, point-wise (bit-exact; family-wise needs all sites)",
                  "parity_mismatches": bad, "parity_runs": [p[0] for p in par_outputs]}
        if bad:
            print("bench.py: PARITY FAILURE against the CPU port on the first %d sites: %s" % (n_par, bad), file=sys.stderr)
            sys.exit(3)
        # the host packer that feeds the compact upload, measured on the same character slice: chars -> (X, site masks) in one pass
        t_p = time.perf_counter()
        for _ in range(3):
            pk = pb.pack_chars_biallelic(cpu_chars)
        t_p = (time.perf_counter() - t_p) / 3
        host_packer = {"one_pass": True, "qualifies": pk is not None, "chars_per_s": cpu_chars.size / t_p, "threads": threads,
                       "sample": "%d nodes x %d sites of this workload as characters (the Fasta_Record text the Java host holds)" % cpu_chars.shape,
                       "seconds_for_this_workload_at_this_rate": N * S_cfg / (cpu_chars.size / t_p),
                       "note": "pb_pack_chars_biallelic / pb_fasta_pack_tile_biallelic emit the compact upload format directly: no three-plane "
                               "intermediate and no compaction pass exist on the host path; packing is tile-pipelined with the upload by the hosts"}

    if exact_null is not None and not exact_null.get("passed", False):
        print("bench.py: EXACT-NULL CHECK FAILED: %s" % json.dumps(exact_null), file=sys.stderr)
        sys.exit(4)
    if rank == 0:
        launch = "one process, one host thread (pb_group)" if inproc else "one process per GPU (torchrun)"
        line = {"metric": METRIC, "value": value, "unit": "sites/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
                "dtype": "u64 bit planes / int32 counts / f64 p-values", "data": "synthetic",
                "config": bench_config(args, cfg, shards[0]["S"], m, tree, n_gpus, "strong" if strong else "weak", launch),
                "e2e": {"value": e2e_value, "unit": "sites/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms / args.steps,
                        "breakdown_ms": e2e_breakdown, "ms_steps_rank0": e2e_step_ms, "timer": "host wall clock around the C-ABI calls, max over ranks",
                        "upload": ("pb_put_planes_biallelic_async in %d-word column tiles + pb_upload_sync (%s)" %
                                   (TILE, "X plane + site masks, kept as one plane in HBM" if shards[0]["compact"][2] else
                                    "X + V planes + site masks, expanded to three planes on the device"))
                                  if all_compact else "pb_put_planes_async (three planes)",
                        "swept_under_upload_words": eager_words,
                        "tail_under_upload_words": int(tm_e.get("tailed_words", 0)),
                        "site_records_returned_under_upload": int(tm_e.get("early_sites", 0)),
                        "calls": "pb_run_events_assoc (one call; `events` in breakdown_ms covers both)" if fused_call else "pb_run_events + pb_run_assoc",
                        "compaction_in_timed_region": False,
                        "input_format": "the compact format the one-pass host packer emits (see host_packer); the synthetic generator produces "
                                        "planes, so here it is derived from them once at set-up (%.2f s), outside the timed region like "
                                        "the FASTA packing" % t_compact if all_compact else "three planes",
                        },
                "e2e_from_three_planes": e2e_full,
                "three_planes": three_planes,
                "gpu_launches": launches,
                "roofline": roofline, "roofline_k3": roofline_k3, "cpu_baseline": cpu_baseline, "clocks": clocks,
                "parity_checked": parity["parity_checked"], "parity_sites": parity["parity_sites"], "parity": parity,
                "host_packer": host_packer, "exact_null_check": exact_null,
                "exchange": exchange, "gen_shard_frac": float(tm.get("gen_shard_frac", 1.0)),
                "memoised": "the binomial table f[n][k] (a pure function of n_max and the case fraction) is kept across passes, as the "
                            "reference's own cache does within a run (HC.java:4196-4280); everything else is recomputed every step",
                "replicate_draws_per_sec": n_draws * args.steps / (dev_ms / 1e3),
                "wall_ms_per_step": wall_ms / args.steps,
                "breakdown_ms": parts,
                "counts": {"informative_sites": int(tm["n_informative"]), "alleles_in_use": int(tm["n_alleles_in_use"]),
                           "extant_events": int(tm["n_extant_events"]), "fast_alleles": int(tm["n_fast_alleles"]),
                           "maxt_fallback_reps": int(tm["n_maxt_fallback_reps"])},
                "synth_seconds": t_gen}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()        # no rank frees its exchange block while a peer may still be reading it
    hc.close()
    for p in pinned:
        _lib.pinned_free(p)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--sites", type=int, default=0, help="override sites per GPU (scaled-down runs)")
    ap.add_argument("--replicates", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-early-output", action="store_true", help="e2e leg without pb_expect_site_output (per-site records copied back by pb_run_events)")
    ap.add_argument("--no-exact-null", action="store_true", help="skip the exact-null validation of the Monte-Carlo counts (pb_run_extras)")
    ap.add_argument("--full-planes-only", action="store_true", help="e2e leg uploads three planes even when the input is biallelic")
    ap.add_argument("--fused-call", action="store_true", help="e2e leg: the fused pb_run_events_assoc instead of pb_run_events + pb_run_assoc")
    ap.add_argument("--inproc", action="store_true", help="--gpus N from ONE process / one host thread (pb_group) instead of torchrun ranks")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="weak: the config's sites per GPU (default); strong: the config's sites over all GPUs (default for C3)")
    ap.add_argument("--upload-tile-words", type=int, default=2048, help="column tile of the end-to-end upload (multiple of 256 site words)")
    args = ap.parse_args()
    if args.scaling is None:
        args.scaling = "strong" if args.config == "C3" else "weak"
    if args.warmup < 3 and args.impl == "b200":
        print("bench.py: warning: fewer than 3 warm-up steps", file=sys.stderr)
    sys.exit(run_reference(args) if args.impl == "reference" else run_b200(args))


if __name__ == "__main__":
    main()
