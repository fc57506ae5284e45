"""Generate tests/golden/binom_golden.json by EXECUTING the reference's own arithmetic
dependency (commons-math3-3.6.1.jar bytecode) in oracle/tools/minijvm.py.

This is the pin for the p-value chain: every value below is what the reference's JVM would
compute for BinomialTest.binomialTest(n, k, p, GREATER_THAN) (HC.java:3050) and for the
FastMath/Gamma/Beta functions under it.  Doubles are stored as C99 hex-float strings (exact).

run (authoring container only — needs /root/reference):
    python oracle/tools/gen_binom_golden.py
"""
import json
import math
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from minijvm import CommonsMath  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def hx(d):
    return float(d).hex()


def main():
    cm = CommonsMath()
    rnd = random.Random(20260)
    out = {"source": "commons-math3-3.6.1.jar bytecode via oracle/tools/minijvm.py",
           "call": "BinomialTest.binomialTest(n,k,p,GREATER_THAN)  (HC.java:3050)"}
    # (1) full triangular table for the author's M. tb case fraction (HC.java:177), n <= 64
    p = 0.379
    tab = []
    for n in range(0, 65):
        for k in range(0, n + 1):
            tab.append(hx(cm.binomial_test_greater(n, k, p)))
    out["table"] = {"p": hx(p), "n_max": 64, "layout": "index n*(n+1)/2+k", "f": tab}
    # (2) random (n, k, p) incl. large n and the case fractions the synthetic configs use
    rows = []
    ps = [0.379, 0.38, 0.5, 1900 / 5000.0, 494 / 1274.0]
    for _ in range(400):
        n = rnd.randint(1, 200); k = rnd.randint(0, n)
        pp = rnd.choice(ps + [rnd.random()])
        rows.append([n, k, hx(pp), hx(cm.binomial_test_greater(n, k, pp))])
    for _ in range(120):
        n = rnd.randint(200, 6000); k = rnd.randint(0, n)
        pp = rnd.choice(ps)
        rows.append([n, k, hx(pp), hx(cm.binomial_test_greater(n, k, pp))])
    for n, k, pp in [(0, 0, 0.5), (1, 1, 0.5), (1, 0, 0.5), (9, 9, 0.379), (9, 0, 0.379), (20000, 7700, 0.38),
                     (5000, 5000, 0.38), (5000, 1, 0.38), (1300, 650, 0.38)]:
        rows.append([n, k, hx(pp), hx(cm.binomial_test_greater(n, k, pp))])
    out["random"] = rows
    # (3) function-level vectors
    fn = {}
    xs = [rnd.uniform(-745, 709) for _ in range(150)] + [-745.5, -709.5, -709.0, -708.5, 0.0, 1e-300, 709.7, -800.0]
    fn["exp"] = [[hx(x), hx(cm.exp(x))] for x in xs]
    xs = [math.exp(rnd.uniform(-700, 700)) for _ in range(100)] + [rnd.uniform(0.98, 1.02) for _ in range(60)] + [1.0, 0.99, 1.01, 5e-324, 1e-310]
    fn["log"] = [[hx(x), hx(cm.log(x))] for x in xs]
    xs = [rnd.uniform(-0.999999, 5) for _ in range(100)] + [rnd.uniform(-2e-6, 2e-6) for _ in range(30)] + [-rnd.random() for _ in range(60)]
    fn["log1p"] = [[hx(x), hx(cm.log1p(x))] for x in xs]
    xs = [rnd.uniform(0.01, 12) for _ in range(80)] + [float(i) for i in range(1, 60)] + [rnd.uniform(8, 5000) for _ in range(40)]
    fn["logGamma"] = [[hx(x), hx(cm.log_gamma(x))] for x in xs]
    ab = [(float(rnd.randint(1, 30)), float(rnd.randint(1, 30))) for _ in range(80)] + \
         [(float(rnd.randint(1, 3000)), float(rnd.randint(1, 3000))) for _ in range(80)] + \
         [(float(rnd.randint(1, 12)), float(rnd.randint(900, 5000))) for _ in range(40)]
    fn["logBeta"] = [[hx(a), hx(b), hx(cm.log_beta(a, b))] for a, b in ab]
    xab = [(rnd.random(), float(rnd.randint(1, 80)), float(rnd.randint(1, 80))) for _ in range(120)]
    fn["regularizedBeta"] = [[hx(x), hx(a), hx(b), hx(cm.regularized_beta(x, a, b))] for x, a, b in xab]
    out["functions"] = fn
    path = os.path.join(ROOT, "tests", "golden", "binom_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes; bytecodes executed:", cm.vm.n_bytecodes)


if __name__ == "__main__":
    main()
