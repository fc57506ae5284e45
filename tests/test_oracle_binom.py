"""The oracle's binomial chain vs the commons-math3 3.6.1 BYTECODE (golden file produced by
oracle/tools/gen_binom_golden.py running the jar in oracle/tools/minijvm.py).  Bit-exact."""
import json
import math
import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def golden(golden_dir):
    with open(os.path.join(golden_dir, "binom_golden.json")) as f:
        return json.load(f)


def fh(s):
    return float.fromhex(s)


def same(a, b):
    return a == b or (a != a and b != b)


def test_table_matches_jar_bytecode(oracle_mod, golden):
    t = golden["table"]
    p, n_max = fh(t["p"]), t["n_max"]
    ref = np.array([fh(x) for x in t["f"]])
    got = oracle_mod.binomial_table(n_max, p)
    assert got.shape == ref.shape
    np.testing.assert_array_equal(got.view(np.uint64), ref.view(np.uint64))


def test_random_cases_match_jar_bytecode(oracle_mod, golden):
    bad = [(n, k) for n, k, p, f in golden["random"] if not same(oracle_mod.binomial_test(n, k, fh(p)), fh(f))]
    assert not bad, bad[:10]


@pytest.mark.parametrize("name,fn", [("exp", "or_fm_exp"), ("log", "or_fm_log"), ("log1p", "or_fm_log1p"), ("logGamma", "or_log_gamma")])
def test_unary_functions_match_jar_bytecode(oracle_mod, golden, name, fn):
    f = getattr(oracle_mod.lib(), fn)
    bad = [x for x, y in golden["functions"][name] if not same(f(fh(x)), fh(y))]
    assert not bad, bad[:10]


def test_logbeta_and_regularized_beta_match_jar_bytecode(oracle_mod, golden):
    L = oracle_mod.lib()
    bad = [(a, b) for a, b, y in golden["functions"]["logBeta"] if not same(L.or_log_beta(fh(a), fh(b)), fh(y))]
    assert not bad, bad[:10]
    bad = [(x, a, b) for x, a, b, y in golden["functions"]["regularizedBeta"]
           if not same(L.or_regularized_beta(fh(x), fh(a), fh(b)), fh(y))]
    assert not bad, bad[:10]


def test_known_answers(oracle_mod):
    # anchors implied by the reference code path: k = 0 -> 1 - cdf(-1) = 1.0 exactly
    for n in (0, 1, 5, 64, 5000):
        assert oracle_mod.binomial_test(n, 0, 0.379) == 1.0
    # f(n, n) = 1 - (1 - p^n) after double rounding, close to p^n
    for n in (1, 3, 9, 20):
        assert oracle_mod.binomial_test(n, n, 0.379) == pytest.approx(0.379 ** n, rel=1e-9)
    # the value is quantised by 1 - (1 - I): for tiny tails it is a multiple of 2^-53
    v = oracle_mod.binomial_test(40, 40, 0.379)
    assert v == 0.0 or (v / 2.0 ** -53) == int(v / 2.0 ** -53)


def test_tail_matches_exact_sum(oracle_mod):
    # independent check against an exact rational tail sum (not the pin, a sanity net)
    from fractions import Fraction
    p = Fraction(379, 1000)
    for n, k in [(10, 3), (25, 12), (60, 31), (64, 1)]:
        exact = sum(Fraction(math.comb(n, j)) * p ** j * (1 - p) ** (n - j) for j in range(k, n + 1))
        assert oracle_mod.binomial_test(n, k, 0.379) == pytest.approx(float(exact), rel=1e-10, abs=2e-16)


def test_non_increasing_in_k(oracle_mod):
    # the GPU's threshold fast path relies on {k : f(n,k) <= p_obs} being an up-set; it re-checks per
    # allele, here we record that the table is in fact monotone at the synthetic configs' p
    for p in (0.38, 0.379):
        for n in (1, 2, 7, 33, 64, 200):
            row = [oracle_mod.binomial_test(n, k, p) for k in range(n + 1)]
            assert all(row[i + 1] <= row[i] for i in range(n)), (p, n)
