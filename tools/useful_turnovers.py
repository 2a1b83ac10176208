"""Writes profiles/useful_turnovers.json: the reference schedule's turnover count per polynomial (the oracle's
count on the benchmark's own inputs -- what `roofline.achieved` calls useful) for every BASELINE
config, so bench lines that do not time the CPU leg themselves (N > 1) use the same honest figure.

    python tools/useful_turnovers.py            # CPU only; a few minutes on 8 cores
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import oracle as orc  # noqa: E402

SAMPLES = {"c1": 1000, "c2": 64, "c3": 64, "c4": 128, "c5": 8}

out = {}
for name, (kind, batch, degree, law, seed) in sorted(bench.CONFIGS.items()):
    k = SAMPLES[name]
    flat, ws, offsets, _ = bench.make_inputs(kind, k, degree, law, seed, 0)
    if kind == "css_c64":
        flat = flat.view(np.complex128)
    t0 = time.time()
    _, _, _, st = orc.roots_csr(flat, offsets, ws=ws, seed=0)
    key = f"{kind}:{degree}:{law}:{seed}"
    out[key] = {"config": name, "polynomials": k, "turnovers_per_poly": st["turnovers"] / k, "sweeps_per_poly": st["sweeps"] / k,
                "unconverged": st["unconverged"], "seconds": round(time.time() - t0, 2)}
    print(key, out[key], flush=True)
with open(os.path.join(ROOT, "profiles", "useful_turnovers.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
