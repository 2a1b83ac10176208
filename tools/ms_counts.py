"""Counts on the CPU (oracle, sequential multishift schedule): executed / useful turnovers, fat steps and lock-steps per
polynomial for (bulges per fat step L, bulges per shift pair) -- the table behind the choice of schedules in DESIGN.md 4.2f.
pipe_steps = sum over fat steps of (window + 3 (L - 1)): lock-steps of one CTA per polynomial (chain.cuh);
stream_steps = turnovers / (7 L): lock-steps when the array never drains (stream.cuh).

    python tools/ms_counts.py 3000 8        # degree, polynomials (config-2 inputs)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import oracle as orc  # noqa: E402
from amrvw_jl_b200 import synth  # noqa: E402

n, B = int(sys.argv[1]), int(sys.argv[2])
law, seed = (2, 5) if n >= 10000 else (0, 2)
rows = synth.fill_coeffs(seed, np.arange(B), law, n + 1)
flat = rows.reshape(-1); off = np.arange(B + 1, dtype=np.int64) * (n + 1)
_, _, _, st0 = orc.roots_csr(flat, off)
print(f"degree {n}, {B} polynomials: reference schedule {st0['turnovers'] / B:.4g} turnovers, {st0['sweeps'] / B:.0f} sweeps per polynomial")
print("%5s %5s %6s %8s %6s %11s %13s %9s" % ("L", "reuse", "shifts", "exec/use", "fat", "pipe_steps", "stream_steps", "hqr_cost"))
for L, reuse in [(16, 1), (16, 2), (32, 2), (32, 3), (32, 4), (64, 2), (64, 3), (64, 4), (128, 4), (128, 5), (128, 6), (128, 8), (256, 8)]:
    m = 2 * L // reuse
    if m > 64:
        continue
    small = max(m + 8, 24)
    _, _, _, st = orc.roots_csr(flat, off, sched=1, L=L, small=small, reuse=reuse, dense=1)
    print("%5d %5d %6d %8.3f %6.0f %11.0f %13.0f %9.1f" % (L, reuse, m, st["turnovers"] / st0["turnovers"], st["fat_steps"] / B, st["pipe_steps"] / B,
                                                           st["turnovers"] / B / 7 / L, st["fat_steps"] / B * m ** 3 / 64 ** 3), flush=True)
