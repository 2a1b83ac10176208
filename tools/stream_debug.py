"""Development probe: strict streamed kernel against the oracle's sequential multishift schedule, polynomial by polynomial."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import amrvw_jl_b200 as A
from oracle import oracle as orc

L, reuse, small = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
sched = int(sys.argv[4]) if len(sys.argv) > 4 else 3
def batch(seed, B, n, kind=0):
    return [orc.fill_coeffs(seed, p, kind, n + 1) for p in range(B)]
rows = batch(31, 40, 150) + batch(32, 30, 700) + batch(35, 6, 1300) + [np.array([24.0, -50.0, 35.0, -10.0, 1.0]), np.array([720.0, -1764.0, 1624.0, -735.0, 175.0, -21.0, 1.0]), np.zeros(0), np.array([1.0, 2.0, 1.0])]
flat = np.concatenate(rows); off = np.zeros(len(rows) + 1, dtype=np.int64); np.cumsum([len(r) for r in rows], out=off[1:])
out, roff, nr, tot, pp = orc.roots_csr(flat, off, sched=1, L=L, small=small, reuse=reuse, dense=1, per_poly=True)
c = A.Context(seed=0, strict=True, schedule=sched, lanes_per_poly=L, shift_reuse=reuse, small_window=small)
for rep in range(2):
    got = A.roots(rows, ctx=c)
    bad = [p for p, g in enumerate(got) if not np.array_equal(g.view(np.float64), out[roff[p]:roff[p] + nr[p]].view(np.float64))]
    print("batch run", rep, "differing polynomials:", bad, "stats", {k: c.last_stats[k] for k in ("fat_steps", "turnovers", "sweeps")}, "oracle", {k: tot[k] for k in ("fat_steps", "turnovers", "sweeps")})
for p in bad:
    print("poly", p, "got", got[p][:8], "want", out[roff[p]:roff[p] + nr[p]][:8], "oracle stats", pp[p])
for p in sorted(set(bad + [40, 64])):
    g = A.roots([rows[p]], ctx=c)[0]
    w = out[roff[p]:roff[p] + nr[p]]
    st = c.last_stats
    print("alone", p, "deg", len(rows[p]) - 1, "equal", np.array_equal(g.view(np.float64), w.view(np.float64)), "maxdiff %.3g" % np.max(np.abs(np.sort_complex(g) - np.sort_complex(w))),
          "gpu fat/turn/sweeps", st["fat_steps"], st["turnovers"], st["sweeps"], "oracle", pp[p]["fat_steps"], pp[p]["turnovers"], pp[p]["sweeps"], "exc", pp[p]["exceptional"])
