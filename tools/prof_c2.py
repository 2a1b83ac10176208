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
print(json.dumps({"B": B, "n": n, "opts": kw, "device_ms": st["device_ms"], "turnovers": st["turnovers"], "fat_steps": st["fat_steps"], "sweeps": st["sweeps"], "launches": st["launches"]}))
print("  kernel ms: set-up %.2f  units %.2f  small windows %.2f  finish %.2f" % (
    pf["us_setup"] / 1e3, pf["us_main"] / 1e3, pf["us_small"] / 1e3, pf["us_finish"] / 1e3))
tot = max(pf["cycles_units"], 1)
print("  unit kernel: busy warp-cycles %.3e (phase A %.1f%%, lean %.1f%%, general %.1f%%); lean steps %d (%.0f cyc/step)  general steps %d (%.0f cyc/step)  visits %d" % (
    tot, 100.0 * pf["cycles_phase_a"] / tot, 100.0 * pf["cycles_lean"] / tot, 100.0 * pf["cycles_general"] / tot,
    pf["steps_lean"], pf["cycles_lean"] / max(pf["steps_lean"], 1), pf["steps_general"], pf["cycles_general"] / max(pf["steps_general"], 1), pf["unit_visits"]))
print("  phase A split: driver %.1f%% of busy cycles, shifts %.1f%%; waiting in pop: %.3e warp-cycles (%.1f%% of busy)" % (
    100.0 * pf["cycles_driver"] / tot, 100.0 * (pf["cycles_phase_a"] - pf["cycles_driver"]) / tot, pf["cycles_pop"], 100.0 * pf["cycles_pop"] / tot))
print("  chase loops (lean + general) %.1f%% of busy; segment overhead (load/park/push) %.1f%%; creation + last-position steps, lane-serialised upper bound %.1f%%" % (
    100.0 * pf["cycles_loop"] / tot, 100.0 * (tot - pf["cycles_loop"] - pf["cycles_phase_a"]) / tot, 100.0 * pf["lane_cycles_special"] / tot))
