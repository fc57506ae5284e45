"""Seeded synthetic inputs of the BASELINE.md shapes (C1..C5) for tests and bench.

There is no network and the reference ships no data, so every input is generated here
(SURVEY.md §8d): Yule tree, sites with 1+Poisson(1.5) toggling mutations placed on branches
in proportion to Exp(1) branch lengths, true simulated internal states standing in for
treetime's ancestral reconstruction, binary phenotype with 38 % cases.

Two generators:
  * `simulate_planes`  — produces the packed node-major bit planes directly (any size);
  * `simulate_chars`   — small (N, S) char matrices with every quirk the reference's walk is
                         sensitive to (gaps, lowercase, third alleles, N at the root, ...).
`planes_to_chars` decodes a site slice of planes back to the chars the reference host would
have read from ancestral_sequences.fasta (used to feed the CPU oracle a slice of C2/C3).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)  # 2-bit code: A0 C1 G2 T3


@dataclass
class Tree:
    """Flattened NewickTree: what the Java host hands to pb_set_tree (HC.java:1144-1152)."""
    parent: np.ndarray            # int32[N], root = -1
    is_leaf: np.ndarray           # uint8[N]
    leaf_nodes: np.ndarray        # int32[L]  (tree.getLeafNodes() order)
    branch_len: np.ndarray        # float64[N] (parsed by the reference, never used: SURVEY §2)
    names: list = field(default_factory=list)

    @property
    def n_nodes(self):
        return len(self.parent)

    @property
    def n_leaves(self):
        return len(self.leaf_nodes)

    @property
    def root(self):
        return int(np.flatnonzero(self.parent < 0)[0])

    def children(self):
        ch = [[] for _ in range(self.n_nodes)]
        for v, p in enumerate(self.parent):
            if p >= 0:
                ch[p].append(v)
        return ch

    def to_newick(self):
        ch = self.children()
        out = []
        # iterative post-order
        stack = [(self.root, 0)]
        while stack:
            v, i = stack.pop()
            if i == 0 and ch[v]:
                out.append("(")
            if i < len(ch[v]):
                if i > 0:
                    out.append(",")
                stack.append((v, i + 1))
                stack.append((ch[v][i], 0))
            else:
                if ch[v]:
                    out.append(")")
                out.append(self.names[v])
                if self.parent[v] >= 0:
                    out.append(":%.6f" % self.branch_len[v])
        return "".join(out) + ";"


def _finish_tree(parent, rng):
    parent = np.asarray(parent, dtype=np.int32)
    n = len(parent)
    nchild = np.bincount(parent[parent >= 0], minlength=n)
    is_leaf = (nchild == 0).astype(np.uint8)
    # leaves in left-to-right DFS order (NewickTree.buildCaches)
    ch = [[] for _ in range(n)]
    for v, p in enumerate(parent):
        if p >= 0:
            ch[p].append(v)
    root = int(np.flatnonzero(parent < 0)[0])
    leaves, stack = [], [root]
    while stack:
        v = stack.pop()
        if not ch[v]:
            leaves.append(v)
        else:
            stack.extend(reversed(ch[v]))
    bl = rng.exponential(1.0, size=n)
    bl[root] = 0.0
    names = [None] * n
    li = 0
    for v in leaves:
        names[v] = "S%06d" % li
        li += 1
    ii = 0
    for v in range(n):
        if names[v] is None:
            names[v] = "NODE_%07d" % ii
            ii += 1
    return Tree(parent=parent, is_leaf=is_leaf, leaf_nodes=np.asarray(leaves, dtype=np.int32), branch_len=bl, names=names)


def yule_tree(n_leaves: int, seed: int, trifurcating_root: bool = True) -> Tree:
    """Pure-birth topology; bifurcating except a trifurcating root (treetime-style unrooted output).
    Node ids are creation order, so parent[v] < v."""
    rng = np.random.default_rng(seed)
    assert n_leaves >= 3
    parent = [-1]
    tips = []
    for _ in range(3 if trifurcating_root else 2):
        parent.append(0)
        tips.append(len(parent) - 1)
    while len(tips) < n_leaves:
        i = int(rng.integers(len(tips)))
        v = tips[i]
        parent.append(v)
        a = len(parent) - 1
        parent.append(v)
        b = len(parent) - 1
        tips[i] = a
        tips.append(b)
    return _finish_tree(parent, rng)


def random_tree(n_leaves: int, seed: int, polytomy: float = 0.2, root_leaf_children: int = 1) -> Tree:
    """Adversarial topology for parity tests: random polytomies, leaves hanging off the root,
    node ids in random order (parent index may exceed child index)."""
    rng = np.random.default_rng(seed)
    parent = [-1]
    tips = []
    k0 = 2 + int(rng.integers(0, 3))
    for _ in range(k0):
        parent.append(0)
        tips.append(len(parent) - 1)
    protected = set(tips[:min(root_leaf_children, k0 - 1)])  # stay leaves attached to the root
    while len(tips) < n_leaves:
        cand = [t for t in tips if t not in protected]
        v = cand[int(rng.integers(len(cand)))]
        k = 2 + (int(rng.integers(1, 4)) if rng.random() < polytomy else 0)
        tips.remove(v)
        for _ in range(k):
            parent.append(v)
            tips.append(len(parent) - 1)
    parent = np.asarray(parent, dtype=np.int64)
    n = len(parent)
    perm = rng.permutation(n)  # new id of old node
    newp = np.full(n, -1, dtype=np.int32)
    for old in range(n):
        newp[perm[old]] = -1 if parent[old] < 0 else perm[parent[old]]
    return _finish_tree(newp, rng)


def caterpillar_tree(n_leaves: int, seed: int) -> Tree:
    """Maximally unbalanced (ladder) topology: depth = number of leaves.  Stresses the iterative DFS of the
    K1 schedule and the reference's O(L * depth) walks."""
    rng = np.random.default_rng(seed)
    parent = [-1]
    spine = 0
    for i in range(n_leaves - 1):
        parent.append(spine)            # leaf hanging off the spine
        if i < n_leaves - 2:
            parent.append(spine)        # next spine node
            spine = len(parent) - 1
        else:
            parent.append(spine)        # last two leaves share the deepest spine node
    return _finish_tree(parent, rng)


def topo_order(parent: np.ndarray) -> np.ndarray:
    """Nodes ordered so that every parent precedes its children."""
    n = len(parent)
    ch = [[] for _ in range(n)]
    root = -1
    for v, p in enumerate(parent):
        if p >= 0:
            ch[p].append(v)
        else:
            root = v
    order, stack = [], [root]
    while stack:
        v = stack.pop()
        order.append(v)
        stack.extend(ch[v])
    return np.asarray(order, dtype=np.int64)


def simulate_planes(tree: Tree, n_sites: int, seed: int, mean_extra_mut: float = 1.5, leaf_gap: bool = False,
                    multiallelic_frac: float = 0.0, third_allele_internal_frac: float = 0.0):
    """-> (B0, B1, V) uint64[N][W], bit j of word w = site 64*w + j.

    Padding bits beyond n_sites in the last word have V = 0."""
    rng = np.random.default_rng(seed)
    N = tree.n_nodes
    W = (n_sites + 63) // 64
    root = tree.root
    n_mut = 1 + rng.poisson(mean_extra_mut, size=n_sites)
    tot = int(n_mut.sum())
    site_of = np.repeat(np.arange(n_sites, dtype=np.int64), n_mut)
    prob = tree.branch_len / tree.branch_len.sum()
    node_of = rng.choice(N, size=tot, p=prob)
    T = np.zeros((N, W), dtype=np.uint64)
    flat = T.reshape(-1)
    idx = node_of * W + (site_of >> 6)
    bit = np.left_shift(np.uint64(1), (site_of & 63).astype(np.uint64))
    order = np.argsort(idx, kind="stable")
    idx_s, bit_s = idx[order], bit[order]
    # XOR-reduce duplicates of the same (node, word): toggling semantics
    starts = np.flatnonzero(np.r_[True, idx_s[1:] != idx_s[:-1]])
    red = np.bitwise_xor.reduceat(bit_s, starts)
    flat[idx_s[starts]] = red
    for v in topo_order(tree.parent):
        p = tree.parent[v]
        if p >= 0:
            T[v] ^= T[p]
    root_base = rng.integers(0, 4, size=n_sites).astype(np.uint8)
    alt_base = ((root_base + rng.integers(1, 4, size=n_sites)) & 3).astype(np.uint8)

    def pack_bits(b):  # bool/0-1 array[n_sites] -> uint64[W]
        pad = np.zeros(W * 64, dtype=np.uint8)
        pad[:n_sites] = b
        return np.packbits(pad.reshape(W, 64), axis=1, bitorder="little").view(np.uint64).reshape(W)

    r0, r1 = pack_bits(root_base & 1), pack_bits((root_base >> 1) & 1)
    d0, d1 = pack_bits((root_base ^ alt_base) & 1), pack_bits(((root_base ^ alt_base) >> 1) & 1)
    valid = pack_bits(np.ones(n_sites, dtype=np.uint8))
    B0 = (T & d0) ^ r0
    B1 = (T & d1) ^ r1
    del T
    V = np.broadcast_to(valid, (N, W)).copy()
    leaves = tree.leaf_nodes
    if multiallelic_frac > 0:
        # a third (and sometimes fourth) base on a few random leaves of the chosen sites
        ms = np.flatnonzero(rng.random(n_sites) < multiallelic_frac)
        for s in ms:
            others = [b for b in range(4) if b != root_base[s] and b != alt_base[s]]
            k = 1 + int(rng.integers(0, 2))
            for j in range(k):
                for leaf in rng.choice(leaves, size=1 + int(rng.integers(0, 3)), replace=False):
                    w, m = s >> 6, np.uint64(1) << np.uint64(s & 63)
                    b = others[j]
                    B0[leaf, w] = (B0[leaf, w] & ~m) | (m if b & 1 else np.uint64(0))
                    B1[leaf, w] = (B1[leaf, w] & ~m) | (m if b & 2 else np.uint64(0))
    if third_allele_internal_frac > 0:
        internal = np.flatnonzero(tree.is_leaf == 0)
        n_hit = int(third_allele_internal_frac * len(internal))
        for v in rng.choice(internal, size=n_hit, replace=False):
            w = int(rng.integers(W))
            m = rng.integers(0, 2**63, dtype=np.uint64) & rng.integers(0, 2**63, dtype=np.uint64) & rng.integers(0, 2**63, dtype=np.uint64)
            B0[v, w] ^= m  # flips bit 0 of the code at ~1/8 of the word's sites -> usually a third base
    if leaf_gap:
        # ~12.5 % of leaf calls invalid (N / '-' / lowercase all pack to V = 0)
        for leaf in leaves:
            g = rng.integers(0, 2**63, size=W, dtype=np.uint64) | (rng.integers(0, 2, size=W, dtype=np.uint64) << np.uint64(63))
            g &= rng.integers(0, 2**63, size=W, dtype=np.uint64) | (rng.integers(0, 2, size=W, dtype=np.uint64) << np.uint64(63))
            g &= rng.integers(0, 2**63, size=W, dtype=np.uint64) | (rng.integers(0, 2, size=W, dtype=np.uint64) << np.uint64(63))
            V[leaf] &= ~g
    return B0, B1, V


def planes_to_chars(B0, B1, V, site_begin: int, site_end: int) -> np.ndarray:
    """Decode sites [site_begin, site_end) of the planes to the (N, S') char matrix the
    reference host would hold in `seg_sites` (invalid -> 'N')."""
    N = B0.shape[0]
    w0, w1 = site_begin >> 6, (site_end + 63) >> 6

    def unpack(P):
        u8 = np.ascontiguousarray(P[:, w0:w1]).view(np.uint8)
        bits = np.unpackbits(u8, axis=1, bitorder="little")
        return bits[:, site_begin - 64 * w0: site_end - 64 * w0]

    b0, b1, v = unpack(B0), unpack(B1), unpack(V)
    chars = BASES[(b0 | (b1 << 1)).astype(np.intp)]
    chars = np.where(v.astype(bool), chars, np.uint8(ord("N")))
    assert chars.shape == (N, site_end - site_begin)
    return np.ascontiguousarray(chars.astype(np.uint8))


def simulate_chars(tree: Tree, n_sites: int, seed: int, nasty: bool = True) -> np.ndarray:
    """Small (N, S) uint8 char matrix.  With nasty=True it carries every quirk that changes the
    reference's result: gaps 'N'/'-' at leaves and internal nodes (root included),
    lowercase bases, third/fourth alleles at leaves (tri/quad-allelic sites) and at internal
    nodes, monomorphic sites, exact major/minor ties."""
    rng = np.random.default_rng(seed)
    N = tree.n_nodes
    order = topo_order(tree.parent)
    chars = np.empty((N, n_sites), dtype=np.uint8)
    prob = tree.branch_len / tree.branch_len.sum()
    for s in range(n_sites):
        rb = int(rng.integers(4))
        ab = (rb + int(rng.integers(1, 4))) & 3
        k = int(rng.poisson(2.0)) if nasty else 1 + int(rng.poisson(1.5))
        mut = np.zeros(N, dtype=bool)
        if k:
            mut[rng.choice(N, size=k, p=prob)] = True
        st = np.zeros(N, dtype=np.int64)
        for v in order:
            p = tree.parent[v]
            st[v] = 0 if p < 0 else st[p] ^ int(mut[v])
        col = np.where(st == 1, BASES[ab], BASES[rb]).astype(np.uint8)
        if nasty:
            u = rng.random()
            others = [b for b in range(4) if b not in (rb, ab)]
            if u < 0.15:   # third allele on some leaves -> tri-allelic
                for leaf in rng.choice(tree.leaf_nodes, size=min(2, tree.n_leaves), replace=False):
                    col[leaf] = BASES[others[0]]
                if rng.random() < 0.3:  # quad
                    col[rng.choice(tree.leaf_nodes)] = BASES[others[1]]
            elif u < 0.35:  # third allele on internal nodes only (lands in the a2 bucket)
                internal = np.flatnonzero(tree.is_leaf == 0)
                for v in rng.choice(internal, size=min(3, len(internal)), replace=False):
                    col[v] = BASES[others[int(rng.integers(2))]]
            g = rng.random(N)
            col = np.where(g < 0.04, np.uint8(ord("N")), col)
            col = np.where((g >= 0.04) & (g < 0.06), np.uint8(ord("-")), col)
            lower = (g >= 0.06) & (g < 0.09)
            col = np.where(lower, col | 0x20, col).astype(np.uint8)  # 'a','c','g','t','n'
            if rng.random() < 0.1:
                col[tree.root] = ord("N")
            if rng.random() < 0.05:
                col[:] = BASES[rb]  # monomorphic
            if rng.random() < 0.05:
                col[tree.leaf_nodes] = ord("N")  # all leaves missing -> 0 alleles
            if rng.random() < 0.1:  # force an exact major/minor tie among valid leaves
                lv = tree.leaf_nodes
                half = len(lv) // 2
                x, y = rng.choice(4, size=2, replace=False)
                col[lv[:half]] = BASES[x]
                col[lv[half:2 * half]] = BASES[y]
                if len(lv) > 2 * half:
                    col[lv[-1]] = ord("N")
        chars[:, s] = col
    return chars


def simulate_phenotypes(tree: Tree, seed: int, case_frac: float = 0.38, na_frac: float = 0.0, no_row_frac: float = 0.0):
    """-> (sample_node int32[P], label uint8[P]) in pheno-file order; label 1 case "1", 0 control "0",
    2 anything else ("NA").  Leaves dropped by `no_row_frac` have no pheno row at all (allowed:
    HC.java:367-404).  Half the cases come from one clade of ~25 % of the leaves."""
    rng = np.random.default_rng(seed)
    L = tree.n_leaves
    ch = tree.children()
    # subtree leaf counts
    order = topo_order(tree.parent)
    nleaf = tree.is_leaf.astype(np.int64).copy()
    for v in order[::-1]:
        p = tree.parent[v]
        if p >= 0:
            nleaf[p] += nleaf[v]
    target = 0.25 * L
    cand = np.flatnonzero((tree.is_leaf == 0) & (np.arange(tree.n_nodes) != tree.root))
    clade_root = int(cand[np.argmin(np.abs(nleaf[cand] - target))]) if len(cand) else tree.root
    in_clade = np.zeros(tree.n_nodes, dtype=bool)
    stack = [clade_root]
    while stack:
        v = stack.pop()
        in_clade[v] = True
        stack.extend(ch[v])
    leaves = tree.leaf_nodes
    lab = np.zeros(L, dtype=np.uint8)
    clade_leaf = in_clade[leaves]
    n_case_target = int(round(case_frac * L))
    n_clade_cases = min(int(clade_leaf.sum()), n_case_target // 2)
    ci = np.flatnonzero(clade_leaf)
    lab[rng.choice(ci, size=n_clade_cases, replace=False)] = 1
    rest = np.flatnonzero(lab == 0)
    lab[rng.choice(rest, size=n_case_target - n_clade_cases, replace=False)] = 1
    if na_frac > 0:
        lab[rng.random(L) < na_frac] = 2
    keep = rng.random(L) >= no_row_frac if no_row_frac > 0 else np.ones(L, dtype=bool)
    perm = rng.permutation(np.flatnonzero(keep))  # pheno-file row order is arbitrary
    return leaves[perm].astype(np.int32), lab[perm]


CONFIGS = {
    # name: leaves, sites, replicates, extras   (seed = 20260 + config #)
    "C1": dict(n_leaves=1300, n_sites=100_000, m=100_000, seed=20261, na_frac=0.02),
    "C2": dict(n_leaves=5000, n_sites=1_000_000, m=100_000, seed=20262),
    "C3": dict(n_leaves=20000, n_sites=5_000_000, m=100_000, seed=20263),
    "C4": dict(n_leaves=1300, n_sites=1_000_000, m=10_000_000, seed=20264),
    "C5": dict(n_leaves=5000, n_sites=1_000_000, m=100_000, seed=20265, leaf_gap=True, multiallelic_frac=0.05,
               third_allele_internal_frac=0.01),
}


def make_config(name: str, n_sites: int | None = None, site_shard: tuple[int, int] | None = None):
    """Build (tree, (B0,B1,V), sample_node, label, cfg) for a named config.  `n_sites` overrides the
    site count (scaled-down runs); `site_shard=(rank, world)` generates only that rank's
    contiguous 64-site-word range (same tree/phenotypes on every rank)."""
    cfg = dict(CONFIGS[name])
    S = n_sites or cfg["n_sites"]
    tree = yule_tree(cfg["n_leaves"], cfg["seed"])
    sample_node, label = simulate_phenotypes(tree, cfg["seed"] + 1000, na_frac=cfg.get("na_frac", 0.0))
    seed = cfg["seed"] + 2000
    if site_shard is not None:
        rank, world = site_shard
        W = (S + 63) // 64
        w0, w1 = W * rank // world, W * (rank + 1) // world
        S = min(S, 64 * w1) - 64 * w0
        seed += 7919 * (rank + 1)
    planes = simulate_planes(tree, S, seed, leaf_gap=cfg.get("leaf_gap", False),
                             multiallelic_frac=cfg.get("multiallelic_frac", 0.0),
                             third_allele_internal_frac=cfg.get("third_allele_internal_frac", 0.0))
    cfg["n_sites"] = S
    return tree, planes, sample_node, label, cfg


def write_inputs(dirname: str, tree: Tree, chars: np.ndarray, sample_node, label, line_width: int = 0, eol: str = "\n",
                 nexus: bool = False, seed: int = 0):
    """The four files of `poutine.sh -u -f -t -p -m` (README of the reference): ancestral FASTA of ALL nodes (segregating
    sites only), tree with internal labels (Newick, or treetime-style annotated nexus), 2-column pheno TSV, PLINK map.
    -> dict of paths + physical positions."""
    import os
    rng = np.random.default_rng(seed)
    os.makedirs(dirname, exist_ok=True)
    N, S = chars.shape
    paths = {k: os.path.join(dirname, v) for k, v in dict(fasta="ancestral_sequences.fasta", tree="ancestral_tree.newick",
                                                           pheno="pheno.tsv", map="sites.map").items()}
    order = rng.permutation(N)                        # record order is arbitrary in the reference (HashMap by name)
    with open(paths["fasta"], "w", newline="") as f:
        for v in order:
            f.write(">" + tree.names[v] + eol)
            seq = chars[v].tobytes().decode("latin-1")
            if line_width > 0:
                for i in range(0, S, line_width):
                    f.write(seq[i:i + line_width] + eol)
            else:
                f.write(seq + eol)
    nwk = tree.to_newick()
    if nexus:
        paths["tree"] = os.path.join(dirname, "annotated_tree.nexus")
        import re
        annotated = re.sub(r"(:[0-9.eE+-]+)", r'[&mutations="A1C"]\1', nwk, count=5)
        with open(paths["tree"], "w") as f:
            f.write("#NEXUS\nBegin Taxa;\n Dimensions NTax=%d;\nEnd;\nBegin Trees;\n Tree tree1=[&U]%s\nEnd;\n" % (tree.n_leaves, annotated))
    else:
        with open(paths["tree"], "w") as f:
            f.write(nwk + "\n")
    with open(paths["pheno"], "w") as f:
        for v, l in zip(sample_node, label):
            f.write("%s\t%s\n" % (tree.names[int(v)], "1" if l == 1 else "0" if l == 0 else "NA"))
    pos = np.sort(rng.choice(np.arange(1, 50 * S + 2), size=S, replace=False)).astype(np.int64)
    with open(paths["map"], "w") as f:
        for i, p_ in enumerate(pos):
            f.write("1\tsnp%d\t0\t%d\n" % (i + 1, p_))
    paths["physical_pos"] = pos
    return paths
