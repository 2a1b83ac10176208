"""Development probe: a ragged batch (degrees uniform in [lo, hi]) through the C ABI, device time of the call.
usage: ragged_bench.py kind B lo hi   (kind: rds | css | pencil)"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import amrvw_jl_b200 as A
from amrvw_jl_b200 import synth

kind, B, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
rng = np.random.default_rng(3)
degs = rng.integers(lo, hi + 1, size=B)
rows = []
for p, d in enumerate(degs):
    if kind == "css":
        r = synth.fill_coeffs(9, np.array([p]), 1, 2 * (int(d) + 1))[0]
        rows.append(r[0::2] + 1j * r[1::2])
    else:
        rows.append(synth.fill_coeffs(9, np.array([p]), 0, int(d) + 1)[0])
ctx = A.Context(seed=0)
best = 1e30
for rep in range(3):
    if kind == "pencil":
        pairs = [A.basic_pencil(r) for r in rows]
        got = A.roots([v for v, _ in pairs], [w for _, w in pairs], ctx=ctx)
    else:
        got = A.roots(rows, ctx=ctx)
    best = min(best, ctx.last_stats["device_ms"])
print(kind, "B", B, "degrees", lo, hi, "device ms %.1f" % best, "lanes", ctx.last_stats["lanes"],
      "unconverged", ctx.last_stats["unconverged"], "sum|roots| %.6f" % sum(float(np.abs(g).sum()) for g in got))
