"""Development probe: chained kernel (lanes_per_poly = L) against the oracle, per polynomial.
usage: chain_debug.py L degs [strict] [solver] [small] [reuse]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import amrvw_jl_b200 as A
from oracle import oracle as orc
from conftest import match_roots, root_condition_tolerance

L = int(sys.argv[1]) if len(sys.argv) > 1 else 256
degs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1500, 1500, 1500, 900, 900]
strict = int(sys.argv[3]) if len(sys.argv) > 3 else 0
solver = int(sys.argv[4]) if len(sys.argv) > 4 else 0
small = int(sys.argv[5]) if len(sys.argv) > 5 else 0
reuse = int(sys.argv[6]) if len(sys.argv) > 6 else 0
rows = [orc.fill_coeffs(5, p, 2, d + 1) for p, d in enumerate(degs)]
c = A.Context(seed=0, strict=bool(strict), schedule=A.SCHEDULE_PIPELINED, lanes_per_poly=L, shift_solver=solver, small_window=small, shift_reuse=reuse)
got = A.roots(rows, ctx=c)
print(sys.argv[1:], c.last_stats)
for p, (r, g) in enumerate(zip(rows, got)):
    w, _ = orc.roots(r)
    d, idx = match_roots(g, w)
    tol = root_condition_tolerance(r, w[idx])
    bad = np.where(d > tol)[0]
    print(L, p, len(r) - 1, "ratio %.3g" % np.max(d / tol), "bad", len(bad), "zeros", np.sum(g == 0), "nan", np.sum(~np.isfinite(g.view(np.float64))))
    for k in bad[:24]:
        print("   slot", k, "got", g[k], "|got|", abs(g[k]), "nearest oracle", w[idx[k]], "dist %.3g" % d[k])
