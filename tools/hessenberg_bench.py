"""Timing of the Hessenberg eigenvalue entry on the GPU box: matrices per second through the C ABI with HOST
buffers (H2D/D2H inside), the kernel's device time, and beside it the oracle (one core) and LAPACK (numpy) on a sample.
Usage: python tools/hessenberg_bench.py [out.json] [--cases n:batch:complex,...]   (e.g. --cases 128:2368:0 for a profiler run)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import amrvw_jl_b200 as A          # noqa: E402
from oracle import oracle as orc   # noqa: E402  (CPU baseline only)


def main():
    argv = sys.argv[1:]
    only = None
    if "--cases" in argv:
        k = argv.index("--cases")
        only = tuple((int(a), int(b), bool(int(c))) for a, b, c in (x.split(":") for x in argv[k + 1].split(",")))
        del argv[k:k + 2]
    out_path = argv[0] if argv else None
    ctx = A.Context(seed=0)
    rng = np.random.default_rng(3)
    rows = []
    from importlib import import_module
    pack = import_module("amrvw_jl_b200.roots").pack_matrices
    warp_ctx = A.Context(seed=0, lanes_per_poly=32)        # keeps one warp per matrix for small matrices too (comparison)
    cases = ((8, 40000, False), (15, 20000, False), (15, 20000, True), (24, 12000, False), (32, 8000, False), (64, 4000, False), (64, 2000, True), (96, 2368, False), (128, 2368, False), (100, 1184, True), (256, 1184, False))
    for n, B, cx in (only or cases):
        mats = [np.triu(rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if cx else 0.0), -1) for _ in range(B)]
        flat, dims = pack(mats)
        ctx.eigvals_hessenberg(mats[:64])
        best = 1e30; dev = 0.0
        for _ in range(3):
            t0 = time.perf_counter()
            out, eoff, nunc = ctx.eigvals_hessenberg_packed(flat, dims)      # the C call with host buffers: H2D, kernel, D2H
            dt = time.perf_counter() - t0
            if dt < best:
                best = dt; dev = ctx.last_stats["device_ms"]
        got = [out[eoff[p]:eoff[p + 1]] for p in range(4)]
        st = ctx.last_stats
        warp_ms = None
        if st["lanes"] == 2:
            warp_ctx.eigvals_hessenberg_packed(flat, dims)
            warp_ctx.eigvals_hessenberg_packed(flat, dims)
            warp_ms = warp_ctx.last_stats["device_ms"]
        ns = max(2, min(B, int(2e6 / n ** 3) + 2))
        t0 = time.perf_counter()
        for M in mats[:ns]:
            orc.eigvals_hessenberg(M)
        t_orc = (time.perf_counter() - t0) / ns
        t0 = time.perf_counter()
        for M in mats[:ns]:
            np.linalg.eigvals(M)
        t_lap = (time.perf_counter() - t0) / ns
        err = max(float(np.abs(np.sort_complex(g) - np.sort_complex(np.linalg.eigvals(M))).max()) for g, M in zip(got[:4], mats[:4]))
        row = {"n": n, "batch": B, "complex": cx, "gpu_call_ms": best * 1e3, "gpu_kernel_ms": dev, "matrices_per_s_e2e": B / best,
               "matrices_per_s_kernel": B / (dev * 1e-3), "r_in_shared_memory": bool(st["lanes"]), "thread_per_matrix": st["lanes"] == 2, "warp_per_matrix_kernel_ms": warp_ms, "sweeps": st["sweeps"], "unconverged": int(nunc.sum()),
               "oracle_one_core_matrices_per_s": 1.0 / t_orc, "lapack_one_core_matrices_per_s": 1.0 / t_lap, "max_abs_diff_vs_lapack_first4": err}
        print(json.dumps(row), flush=True)
        rows.append(row)
    if out_path:
        with open(out_path, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
