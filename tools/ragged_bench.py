"""Development probe: a ragged batch (degrees uniform in [lo, hi]) through the C ABI, device time of the call.
usage: ragged_bench.py kind B lo hi [lanes]   (kind: rds | css | pencil; lanes pins one lane count = no degree classes)"""
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
lanes = int(sys.argv[5]) if len(sys.argv) > 5 else 0
ctx = A.Context(seed=0, lanes_per_poly=lanes, schedule=2 if lanes else 0)
import time
if kind == "pencil":
    pairs = [A.basic_pencil(r) for r in rows]
    vs, ws = [v for v, _ in pairs], [w for _, w in pairs]
    flat = np.concatenate(vs); wflat = np.concatenate(ws); lens = [len(v) for v in vs]
else:
    flat = np.concatenate(rows); wflat = None; lens = [len(r) for r in rows]
off = np.zeros(len(lens) + 1, dtype=np.int64); np.cumsum(lens, out=off[1:])
out = ctx.roots_csr(flat, off, wflat, want_stats=True)          # warm-up (allocations), with the statistics
st = dict(ctx.last_stats)
best = 1e30
for rep in range(3):                                             # wall clock of the host-buffer call (H2D + kernels + D2H), no statistics:
    t0 = time.perf_counter()                                     # with them the degree classes run one after the other
    out = ctx.roots_csr(flat, off, wflat, want_stats=False)
    best = min(best, time.perf_counter() - t0)
print(kind, "B", B, "degrees", lo, hi, "lanes option", lanes, "host-buffer call ms %.1f" % (1e3 * best), "lanes (max over classes)", st["lanes"], "schedule", st["schedule_used"],
      "unconverged", st["unconverged"], "sum|roots| %.6f" % float(np.abs(out[0]).sum()))
