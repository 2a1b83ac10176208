"""Development probe: Horner backward error of every polynomial of a batch for several lanes_per_poly."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import amrvw_jl_b200 as A
from oracle import oracle as orc
from test_gpu_parity import horner_backward_error

Ls = [int(x) for x in sys.argv[1].split(",")]
B = int(sys.argv[2]); n = int(sys.argv[3]); kind = int(sys.argv[4]); seed = int(sys.argv[5]) if len(sys.argv) > 5 else 5
strict = int(sys.argv[6]) if len(sys.argv) > 6 else 0
rows = [orc.fill_coeffs(seed, p, kind, n + 1) for p in range(B)]
for L in Ls:
    c = A.Context(seed=0, strict=bool(strict), schedule=A.SCHEDULE_PIPELINED, lanes_per_poly=L)
    got = A.roots(rows, ctx=c)
    be = np.array([horner_backward_error(r, g).max() for r, g in zip(rows, got)])
    bad = np.where(be > 1e-10)[0]
    print("L", L, "B", B, "n", n, "kind", kind, "strict", strict, "worst be %.3g" % be.max(), "bad polys", list(bad[:20]), "of", len(bad), c.last_stats["device_ms"], flush=True)
    c.close()
