"""Development aid: one config-2-style run (B polynomials of degree n) with the warp-cycle profile."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import amrvw_jl_b200 as A
from amrvw_jl_b200 import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
kw = {}
for a in sys.argv[3:]:
    k, v = a.split("=")
    kw[k] = int(v)
ctx = A.Context(device=0, seed=0, **kw)
rows = synth.fill_coeffs(2, np.arange(B), 0, n + 1)
coeffs = np.ascontiguousarray(np.concatenate(list(rows)))
offsets = np.arange(B + 1, dtype=np.int64) * (n + 1)
for rep in range(2):
    out = ctx.roots_csr(coeffs, offsets, want_stats=True)
st = ctx.last_stats
pf = ctx.last_profile()
w = max(pf["warps"], 1)
tot = pf["cycles_total"]
print(json.dumps({"B": B, "n": n, "opts": kw, "device_ms": st["device_ms"], "turnovers": st["turnovers"], "fat_steps": st["fat_steps"], "sweeps": st["sweeps"]}))
print("  warps %d  mean lifetime %.1f Mcyc  max %.1f Mcyc" % (w, tot / w / 1e6, pf["cycles_max_warp"] / 1e6))
for k in ("cycles_setup", "cycles_phase_a", "cycles_lean", "cycles_general", "cycles_epilogue"):
    print("  %-16s %5.1f%%" % (k, 100.0 * pf[k] / tot))
print("  lean steps %d (%.0f cyc/step)  general steps %d (%.0f cyc/step)  lean calls %d" % (
    pf["steps_lean"], pf["cycles_lean"] / max(pf["steps_lean"], 1), pf["steps_general"], pf["cycles_general"] / max(pf["steps_general"], 1), pf["lean_calls"]))
print("  per polynomial: shift sub-solves %.1f Mcyc, small windows %.1f Mcyc (%d windows); warp lifetime %.1f Mcyc" % (
    pf["leader_cycles_subsolve"] / B / 1e6, pf["leader_cycles_small"] / B / 1e6, pf["small_windows"] // B, tot / w / 1e6))
