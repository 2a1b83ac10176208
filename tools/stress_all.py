"""Development probe: Horner backward error of every polynomial of a batch, any variant, default options.
usage: stress_all.py kind B n law seed   (kind: rds | css | pencil | pencil_half)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import amrvw_jl_b200 as A
from oracle import oracle as orc
from conftest import pencil_split
from test_gpu_parity import horner_backward_error

kind, B, n, law, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
cx = kind == "css"
rows = [orc.fill_coeffs(seed, p, law, (2 if cx else 1) * (n + 1)) for p in range(B)]
if cx:
    rows = [r[0::2] + 1j * r[1::2] for r in rows]
ctx = A.Context(seed=0)
if kind.startswith("pencil"):
    pairs = [A.basic_pencil(r) if kind == "pencil" else pencil_split(r, n // 2) for r in rows]
    got = A.roots([v for v, _ in pairs], [w for _, w in pairs], ctx=ctx)
else:
    got = A.roots(rows, ctx=ctx)
be = np.array([horner_backward_error(r, g).max() for r, g in zip(rows, got)])
bad = np.where(~(be < 1e-10))[0]
print(kind, "B", B, "n", n, "law", law, "seed", seed, "lanes", ctx.last_stats["lanes"], "unconverged", ctx.last_stats["unconverged"], "exceptional", ctx.last_stats["exceptional"],
      "worst be %.3g" % np.nanmax(be), "bad", list(bad[:10]), "of", len(bad), "ms %.1f" % ctx.last_stats["device_ms"], flush=True)
