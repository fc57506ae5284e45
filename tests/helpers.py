"""Test helpers: an INDEPENDENT numpy statement of the event-detection closed form,
used to cross-check the oracle's literal leaf->root walk, plus comparison utilities."""
import numpy as np

ORDER = b"ACTG"  # HashMap<Character,..> iteration order (bucket = c & 15)


def closed_form_events(parent, leaf_nodes, chars):
    """-> list of dicts for biallelic sites + class counts, via
       M = {v : v != root, parent(v) != root, geno(v) in ACGT, geno(parent(v)) != geno(v)} U {root}."""
    parent = np.asarray(parent)
    N, S = chars.shape
    root = int(np.flatnonzero(parent < 0)[0])
    is_leaf = np.zeros(N, bool)
    is_leaf[leaf_nodes] = True
    valid = np.isin(chars, np.frombuffer(b"ACGT", np.uint8))
    par = np.where(parent < 0, 0, parent)
    tested = (np.arange(N) != root) & (parent != root)
    mrca = tested[:, None] & valid & (chars[par] != chars)
    out, classes = [], dict(mono=0, bi=0, tri=0, quad=0, other=0)
    for s in range(S):
        col = chars[:, s]
        cnt = [int(np.sum(col[leaf_nodes] == c)) for c in ORDER]
        present = [i for i in range(4) if cnt[i] > 0]
        k = len(present)
        classes[{1: "mono", 2: "bi", 3: "tri", 4: "quad"}.get(k, "other")] += 1
        if k != 2:
            continue
        f, g = present
        a1, a2 = (ORDER[f], ORDER[g]) if cnt[f] >= cnt[g] else (ORDER[g], ORDER[f])
        m = np.flatnonzero(mrca[:, s])
        b1 = set(int(v) for v in m if col[v] == a1)
        b2 = set(int(v) for v in m if col[v] != a1)
        (b1 if col[root] == a1 else b2).add(root)
        if len(b1) < 2:
            b1 = set()
        if len(b2) < 2:
            b2 = set()
        out.append(dict(segsite_id=s + 1, allele1=a1, allele2=a2, a1_count=len(b1), a2_count=len(b2),
                        a1_extant=sum(is_leaf[v] for v in b1), a2_extant=sum(is_leaf[v] for v in b2),
                        mrca=(sorted(b1), sorted(b2))))
    return out, classes


def assert_stats_equal(gpu, ref, p_rtol=0.0):
    """Compare pb_stat_out records with or_stat_t records (matched by segsite_id)."""
    assert len(gpu) == len(ref), (len(gpu), len(ref))
    np.testing.assert_array_equal(gpu["segsite_id"], ref["segsite_id"])
    for f in ("a1_in_use", "a2_in_use", "obs_counts", "r_a1", "r_a2", "r_maxT_a1", "r_maxT_a2"):
        np.testing.assert_array_equal(gpu[f], ref[f], err_msg=f)
    for f in ("obs_p_a1", "obs_p_a2", "pointwise_a1", "pointwise_a2", "familywise_a1", "familywise_a2"):
        a, b = gpu[f], ref[f]
        np.testing.assert_array_equal(np.isnan(a), np.isnan(b), err_msg=f)
        ok = ~np.isnan(a)
        if p_rtol == 0.0:
            np.testing.assert_array_equal(a[ok].view(np.uint64), b[ok].view(np.uint64), err_msg=f + " (bit-exact)")
        else:
            np.testing.assert_allclose(a[ok], b[ok], rtol=p_rtol, atol=0, err_msg=f)


# ---- exact rational statements of the two "extras" (pb_run_extras; parity unpinned: no reference code can run) ----
def exact_pointwise_null(P, n_case, n_miss, n, p_obs, f):
    """P(f(n - j, k) <= p_obs) under a uniform permutation of the P labels (n_case cases, n_miss missing) over the P
    samples, for an allele with n sampled leaves: j of them draw a missing label, k a case.  f(n, k) -> float."""
    from fractions import Fraction
    from math import comb
    n_ctrl = P - n_case - n_miss
    acc = 0
    for j in range(0, min(n, n_miss) + 1):
        nr = n - j
        for k in range(0, min(nr, n_case) + 1):
            ways = comb(n_miss, j) * comb(n_case, k) * comb(n_ctrl, nr - k)
            if ways and f(nr, k) <= p_obs:
                acc += ways
    return Fraction(acc, comb(P, n))


def fisher_two_sided(a, b, c, d):
    """R's fisher.test p-value for [[a, b], [c, d]]: sum of hypergeometric table probabilities <= P(obs) * (1 + 1e-7)."""
    from fractions import Fraction
    from math import comb
    m, n, k = a + b, c + d, a + c
    lo, hi = max(0, k - n), min(k, m)
    ways = {x: comb(m, x) * comb(n, k - x) for x in range(lo, hi + 1)}
    tot = sum(ways.values())
    acc = sum(w for w in ways.values() if w * 10**7 <= ways[a] * (10**7 + 1))
    return Fraction(acc, tot)
