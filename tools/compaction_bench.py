"""Effect of the unit compaction (rounds.cuh) on a batch with more units than resident warps.  Usage: python tools/compaction_bench.py [B] [degree]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import amrvw_jl_b200 as A
from amrvw_jl_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
rows = synth.fill_coeffs(2, np.arange(B), 0, n + 1)
flat = np.ascontiguousarray(rows).reshape(-1); off = np.arange(B + 1, dtype=np.int64) * (n + 1)
for flag in ("0", "1", "0", "1"):
    os.environ["AMRVW_COMPACT"] = flag
    c = A.Context(seed=0, lanes_per_poly=16, schedule=A.SCHEDULE_PIPELINED)
    c.roots_csr(flat, off, want_stats=False)
    best = 1e9
    for _ in range(2):
        c.roots_csr(flat, off, want_stats=True)
        best = min(best, c.last_stats["device_ms"])
    print("batch %d degree %d compaction %s: kernels %.1f ms, polynomials moved %d, unconverged %d" % (B, n, flag, best, c.debug_counters()[47], c.last_stats["unconverged"]), flush=True)
    c.close()
