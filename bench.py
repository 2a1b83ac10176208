#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched core-chasing root finder (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # ours; N > 1 under torchrun
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference algorithm on host cores

A "step" is one pass of the hot path over one batch: rooting `batch` polynomials of degree `degree`
(default: BASELINE config 2, the README statistics workload -- 3000 N(0,1) polynomials of degree
3000, Float64 real double shift).  Polynomials are independent, so at N GPUs that ONE batch is split
contiguously by polynomial (`synth.shard_range`: 375 each at N = 8, the north-star configuration) with
no collective: "scaling": "strong".  `--weak` gives every rank a full batch of its own instead.

Printed JSON (one line, rank 0): metric polynomials/s (`value`: inputs resident in HBM, CUDA-event
timed; `e2e`: through the C ABI with pinned HOST buffers, H2D + kernels + D2H inside the timed
region), plus `roofline` (FP64 vector pipe -- measured live with a DFMA probe because
MEASURED_PEAKS.json has no FP64 entry -- and the HBM figure next to it), `cpu_baseline` (the oracle =
the reference's algorithm and schedule, OpenMP over polynomials on this box's host cores, bounded
sample), clocks sampled during the timed region, and the kernel launch count.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (kind, batch, degree, input law, seed)   -- BASELINE.json configs[0..4]
    "c1": ("rds_f64", 1000, 100, 0, 1),
    "c2": ("rds_f64", 3000, 3000, 0, 2),
    "c3": ("css_c64", 2000, 1000, 1, 3),
    "c4": ("pencil_f64", 1000, 500, 0, 4),
    "c5": ("rds_f64", 64, 10000, 2, 5),     # 512 over 8 GPUs = 64 per GPU
}
FLOP_PER_TURNOVER = {"rds_f64": 37.0, "css_c64": 94.0, "pencil_f64": 37.0}
BYTES_PER_TURNOVER = {"rds_f64": 96.0 / 7, "css_c64": 208.0 / 3, "pencil_f64": 160.0 / 11}
# reference-schedule turnovers per polynomial (the oracle's count: what "useful" means) for runs that do
# not time the CPU leg themselves (N > 1): profiles/useful_turnovers.json, written by tools/useful_turnovers.py
USEFUL_FILE = os.path.join(ROOT, "profiles", "useful_turnovers.json")
# dram bytes of the dominant kernel per launch from the committed `ncu --set full` capture of the same command
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic.json")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch (development only)")
    ap.add_argument("--degree", type=int, default=0, help="override the degree (development only)")
    ap.add_argument("--schedule", type=int, default=0, help="0 auto, 1 reference one-bulge, 2 pipelined")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--reuse", type=int, default=0)
    ap.add_argument("--small", type=int, default=0)
    ap.add_argument("--stream-ctas", type=int, default=0, help="streamed schedule: resident CTAs per SM (development)")
    ap.add_argument("--lib", default=None, help="another build of the library (development: kernel variants under build/)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="polynomials in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--weak", action="store_true", help="every rank roots a full batch of its own (round-1 behaviour) instead of a share of one batch")
    return ap.parse_args()


def make_inputs(kind, batch, degree, law, seed, first_id):
    from amrvw_jl_b200 import synth
    ids = np.arange(first_id, first_id + batch)
    if kind == "css_c64":
        raw = synth.fill_coeffs(seed, ids, law, 2 * (degree + 1))
        flat = np.ascontiguousarray(raw).reshape(-1)          # interleaved (re, im)
        ncoef = degree + 1
    else:
        flat = np.ascontiguousarray(synth.fill_coeffs(seed, ids, law, degree + 1)).reshape(-1)
        ncoef = degree + 1
    if kind == "pencil_f64":                                   # basic_pencil split (companion.jl:44-54)
        a = flat.reshape(batch, degree + 1)
        vs = np.ascontiguousarray(a[:, :-1]).reshape(-1)
        ws = np.zeros((batch, degree)); ws[:, -1] = a[:, -1]
        offsets = np.arange(batch + 1, dtype=np.int64) * degree
        return vs, ws.reshape(-1), offsets, degree
    offsets = np.arange(batch + 1, dtype=np.int64) * ncoef
    return flat, None, offsets, ncoef


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "power_w_max": max(float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()), "reasons": sorted(reasons), "samples": len(self.rows)}


def cpu_baseline(kind, degree, law, seed, sample, threads=0):
    """The reference's algorithm and schedule (the oracle port) on this box's host cores."""
    from oracle import oracle as orc            # allowed here: cpu_baseline / --impl reference legs only
    # all the host cores this process may run on: torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    # leave the reference arm at N > 1 on a single core (observed: 1.4 instead of 20.7 polynomials/s)
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    cores = max(orc.num_threads(), avail) if threads <= 0 else threads
    if sample <= 0:   # ~12 s of wall time: one polynomial costs about k * degree^2 seconds on one core
        per_poly_s = {"rds_f64": 1.6e-7, "css_c64": 4.7e-7, "pencil_f64": 3.5e-7}[kind] * degree * degree
        sample = max(cores, min(4096, int(cores * max(1.0, 12.0 / max(per_poly_s, 1e-9)))))
    flat, ws, offsets, _ = make_inputs(kind, sample, degree, law, seed, 0)
    if kind == "css_c64":
        flat = flat.view(np.complex128)
    t0 = time.perf_counter()
    out, roff, nr, st = orc.roots_csr(flat, offsets, ws=ws, seed=0, threads=cores)
    dt = time.perf_counter() - t0
    return {"value": sample / dt, "unit": "polynomials/s", "cores": cores, "kind": "port", "value_per_core": sample / dt / cores,
            "sample": f"{sample} polynomials of degree {degree} (same generator and seed as the GPU batch), {dt:.2f} s wall, OpenMP dynamic over polynomials",
            "turnovers_per_s": st["turnovers"] / dt, "turnovers_per_poly": st["turnovers"] / sample, "seconds": dt}


def main():
    args = parse()
    kind, batch, degree, law, seed = CONFIGS[args.config]
    if args.batch:
        batch = args.batch
    if args.degree:
        degree = args.degree
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference" and world == 1 and args.gpus > 1:
        world = args.gpus          # launched without torchrun: describe the same N-GPU workload our arm runs
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from amrvw_jl_b200 import synth as _synth
    if args.config == "c5" and not args.batch and not args.weak:
        batch = 512                                    # BASELINE config 5 is 512 polynomials over 8 GPUs
    if args.weak:
        lo, hi, total_batch = rank * batch, (rank + 1) * batch, batch * world
    else:
        lo, hi = _synth.shard_range(batch, rank, world)
        total_batch = batch
    share = hi - lo
    workload = (f"{args.config}: {kind} batch={total_batch} degree={degree} law={['N(0,1)', 'U[0,1)', 'U[0,1)-1/2'][law]} seed={seed}, "
                + (f"{batch} per GPU" if args.weak else f"split by polynomial over {world} GPU(s): {batch // world}-{-(-batch // world)} each"))
    config = {"workload": workload, "kind": kind, "batch": total_batch, "batch_per_gpu": share, "degree": degree, "parallelism": f"shard-by-polynomial x{world}",
              "l2": "per-step working set (rotator chains, rewritten by the set-up every step) exceeds the 126 MB L2 at every N for c2 (54 MB of chains per 375 polynomials x 3 chains + roots); no flush between steps"}
    scaling = "weak" if args.weak else "strong"

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        runs = []
        sample = args.cpu_sample
        for i in range(args.warmup + args.steps):
            cb = cpu_baseline(kind, degree, law, seed, sample)
            if sample <= 0:
                sample = int(cb["sample"].split()[0])
            if i >= args.warmup:
                runs.append(cb)
            if i == 0 and cb["seconds"] * (args.warmup + args.steps) > 240:   # keep the whole run within minutes
                sample = max(cb["cores"], int(sample * 240 / (cb["seconds"] * (args.warmup + args.steps))))
        secs = sum(r["seconds"] for r in runs)
        n = sum(int(r["sample"].split()[0]) for r in runs)
        val = n / secs
        line = {"impl": "reference", "metric": "polynomials/sec", "value": val, "unit": "polynomials/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * secs / len(runs), "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "polynomials/s", "cores": runs[-1]["cores"], "kind": "port", "sample": runs[-1]["sample"],
                                 "value_per_core": val / runs[-1]["cores"],
                                 "turnovers_per_s": sum(r["turnovers_per_s"] * r["seconds"] for r in runs) / secs},
                "e2e": {"value": val, "unit": "polynomials/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "Julia is not available in this image: the reference arm is the C++ oracle (same algorithm, same one-bulge schedule); each step is a bounded sample of the workload"}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    import torch.distributed as dist
    import amrvw_jl_b200 as A

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: this framework has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")      # host-side barriers: a rank waiting in an NCCL barrier keeps a kernel spinning on its GPU

    batch = share                                      # from here on: this rank's polynomials
    flat, ws, offsets, ncoef = make_inputs(kind, batch, degree, law, seed, lo)
    total = int(offsets[-1])
    ctx = A.Context(device=local_rank, seed=seed, schedule=args.schedule, lanes_per_poly=args.lanes, shift_reuse=args.reuse, small_window=args.small, stream_ctas_per_sm=args.stream_ctas, lib_path=args.lib)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    d_c = torch.from_numpy(flat).to(dev)
    d_w = torch.from_numpy(ws).to(dev) if ws is not None else None
    d_o = torch.from_numpy(offsets).to(dev)
    nroot_slots = total                               # polynomial p's roots start at complex slot offsets[p]
    d_r = torch.zeros(2 * (total + 1), dtype=torch.float64, device=dev)
    d_n = torch.zeros(batch, dtype=torch.int32, device=dev)
    d_u = torch.zeros(batch, dtype=torch.int32, device=dev)

    def step(stats=False):
        return ctx.roots_dev(kind, d_c.data_ptr(), d_o.data_ptr(), batch, total, ncoef, d_r.data_ptr(), d_n.data_ptr(), d_u.data_ptr(),
                             d_ws=d_w.data_ptr() if d_w is not None else None, want_stats=stats)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = ctx.fp64_peak_tflops()
    st = None
    for _ in range(max(args.warmup, 1)):
        st = step(stats=True)           # warm-up also yields the per-step counters (launches, turnovers)
    torch.cuda.synchronize()
    prof = ctx.last_profile() if args.schedule != 1 else None   # per-kernel CUDA-event times (+ warp-cycle counters, RDS) of the last warm-up step
    dv = ctx.debug_counters() if args.schedule != 1 else None
    n_unconv = int(d_u.sum().item())
    nroots_ok = int(d_n.sum().item())

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local_rank) as clk:
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        barrier()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = total_batch * args.steps / (ms * 1e-3)

    # ---- end to end through the C ABI with pinned host buffers
    e2e = None
    if not args.no_e2e:
        h_c = torch.from_numpy(flat).pin_memory()
        h_w = torch.from_numpy(ws).pin_memory() if ws is not None else None
        h_o = torch.from_numpy(offsets).pin_memory()
        h_r = torch.zeros(2 * (total + 1), dtype=torch.float64).pin_memory()
        h_n = torch.zeros(batch, dtype=torch.int32).pin_memory()
        h_u = torch.zeros(batch, dtype=torch.int32).pin_memory()
        lib = A.load(path=args.lib)
        import ctypes
        P = ctypes.c_void_p

        def e2e_step():
            if kind == "pencil_f64":
                rc = lib.amrvw_roots_pencil_f64(ctx._h, P(h_c.data_ptr()), P(h_w.data_ptr()), P(h_o.data_ptr()), batch, P(h_r.data_ptr()), P(h_n.data_ptr()), P(h_u.data_ptr()), None)
            elif kind == "css_c64":
                rc = lib.amrvw_roots_css_c64(ctx._h, P(h_c.data_ptr()), P(h_o.data_ptr()), batch, P(h_r.data_ptr()), P(h_n.data_ptr()), P(h_u.data_ptr()), None)
            else:
                rc = lib.amrvw_roots_rds_f64(ctx._h, P(h_c.data_ptr()), P(h_o.data_ptr()), batch, P(h_r.data_ptr()), P(h_n.data_ptr()), P(h_u.data_ptr()), None)
            assert rc == 0, lib.amrvw_last_error(ctx._h)
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()                  # synchronous: returns after the D2H copy of the roots
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        h2d = flat.nbytes + (ws.nbytes if ws is not None else 0) + offsets.nbytes
        d2h = 16 * nroot_slots + 8 * batch
        e2e = {"value": total_batch * args.steps / dt, "unit": "polynomials/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "bytes_are": "per rank (every rank copies its own share)", "roots_checksum": float(h_r[: 2 * nroot_slots].abs().sum().item()),
               "call": "amrvw_roots_*_f64/c64 with pinned host buffers, one single-device context per rank"}
        # ---- the same batch through ONE call of ONE process: a multi-device context over all N GPUs (amrvw_create_multi)
        if world > 1 and not args.weak and kind == "rds_f64":
            barrier()
            torch.cuda.synchronize()
            dist.barrier(group=host_group)            # every GPU idle from here on: rank 0 drives all of them from one process
            if rank == 0:
                try:
                    fa, _, oa, _ = make_inputs(kind, total_batch, degree, law, seed, 0)
                    g_c = torch.from_numpy(fa).pin_memory(); g_o = torch.from_numpy(oa).pin_memory()
                    g_r = torch.zeros(2 * (int(oa[-1]) + 1), dtype=torch.float64).pin_memory()
                    g_n = torch.zeros(total_batch, dtype=torch.int32).pin_memory(); g_u = torch.zeros(total_batch, dtype=torch.int32).pin_memory()
                    mctx = A.Context(devices=list(range(world)), seed=seed, schedule=args.schedule, lanes_per_poly=args.lanes, shift_reuse=args.reuse, small_window=args.small)

                    def one_call():
                        rc = lib.amrvw_roots_rds_f64(mctx._h, P(g_c.data_ptr()), P(g_o.data_ptr()), total_batch, P(g_r.data_ptr()), P(g_n.data_ptr()), P(g_u.data_ptr()), None)
                        assert rc == 0, lib.amrvw_last_error(mctx._h)
                    one_call()
                    t0 = time.perf_counter()
                    for _ in range(args.steps):
                        one_call()
                    dt1 = time.perf_counter() - t0
                    e2e["one_process_one_call"] = {"value": total_batch * args.steps / dt1, "unit": "polynomials/s", "devices": world,
                                                   "h2d_bytes_per_step": int(fa.nbytes + oa.nbytes), "d2h_bytes_per_step": int(16 * int(oa[-1]) + 8 * total_batch),
                                                   "unconverged": int(g_u.sum().item()),
                                                   "call": "one amrvw_roots_rds_f64 call on an amrvw_create_multi context: split, per-device streams, gather inside the library"}
                    mctx.close()
                except Exception as ex:      # never lose the main line to the extra measurement
                    e2e["one_process_one_call"] = {"error": repr(ex)}
            dist.barrier(group=host_group)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- CPU baseline (rank 0, N = 1 only) and roofline
    cb = None
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(kind, degree, law, seed, args.cpu_sample)
    ms_step = ms / args.steps
    executed = st["turnovers"]                       # this rank's share
    # "useful" = the reference schedule's turnover count on the same inputs, at every N: from this run's CPU sample
    # when it has one, else from the committed oracle counts (never from the executed count)
    useful_per_poly, useful_src = None, None
    if cb:
        useful_per_poly, useful_src = cb["turnovers_per_poly"], "CPU sample of this run (oracle, reference schedule)"
    else:
        try:
            tab = json.load(open(USEFUL_FILE))
            key = f"{kind}:{degree}:{law}:{seed}"
            useful_per_poly, useful_src = float(tab[key]["turnovers_per_poly"]), f"profiles/useful_turnovers.json[{key}] (oracle, {tab[key]['polynomials']} polynomials)"
        except Exception:
            pass
    useful = useful_per_poly * batch if useful_per_poly else None
    # dominant kernel: the persistent unit kernel (phase A + systolic chase) for the pipelined RDS path, the
    # one iteration kernel otherwise; its duration is the CUDA-event time of that launch in the last warm-up step
    lanes = int(st.get("lanes", 0))
    if prof and prof["us_main"] > 0:
        kernel_ms = prof["us_main"] / 1e3
        kernel_name = {"rds_f64": (f"stream_kernel<{lanes // 32},MH> ({lanes // 32} chained warps per CTA fed by a stream of fat steps of different polynomials, {lanes} bulges each; phase A on service warps)" if st.get("schedule_used") == 3
                                   else f"chain_kernel<{lanes // 32},MH> (one CTA of {lanes // 32} chained warps per polynomial: phase A + systolic chase of {lanes} bulges)") if lanes > 32
                       else f"unit_kernel<{lanes},MH> (phase A + systolic chase behind the unit queue)",
                       "css_c64": f"css_unit_kernel<{lanes},MH> (complex single shift: phase A + systolic chase behind the unit queue)",
                       "pencil_f64": f"pen_unit_kernel<{lanes},MH> (companion pencil, five chains: phase A + systolic chase behind the unit queue)"}[kind]
    else:
        kernel_ms, kernel_name = st["device_ms"], "roots_reference_kernel / roots_pencil_reference_kernel (one polynomial per thread)"
    tf = useful * FLOP_PER_TURNOVER[kind] / (kernel_ms * 1e-3) / 1e12 if useful else None
    hbm_peak, hbm_src = 6650.0, "fallback"
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, hbm_src = float(mp["hbm_gbs"]), "measured"
    except Exception:
        pass
    gbs = useful * BYTES_PER_TURNOVER[kind] / (kernel_ms * 1e-3) / 1e9 if useful else None
    # dram__bytes_read.sum + dram__bytes_write.sum of that kernel per launch: from the committed ncu --set full capture of
    # the same command (profiles/ncu_traffic.json names the .csv it was read from), only for the launch it was captured on
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(TRAFFIC_FILE)).get(f"{args.config}:{batch}:{degree}:{lanes}")
        if tr and not args.reuse and not args.small:
            traffic, traffic_src = float(tr["dram_bytes"]), tr["source"]
    except Exception:
        pass
    roofline = {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak if tf else None, "traffic": traffic, "traffic_source": traffic_src,
                "per": "GPU (rank 0's share of the batch, its kernel time, one GPU's peak): the same quantity at every N",
                "traffic_note": "bytes per launch (ncu dram read+write); the reference-schedule algorithmic bytes (96 B per chase step per bulge) are ~12x larger: the systolic chase streams each rotator once per L bulges",
                "peak_source": "DFMA probe measured live in this run (MEASURED_PEAKS.json has no FP64 entry); nominal 37.2",
                "kernel": kernel_name, "kernel_ms": kernel_ms, "kernel_share_of_step": kernel_ms / st["device_ms"] if st["device_ms"] else None,
                "useful_turnovers_per_step": useful, "executed_turnovers_per_step": executed, "exec_over_useful": executed / useful if useful else None,
                "useful_source": useful_src,
                "useful_definition": "the reference schedule's turnover count on the same inputs (oracle mean per polynomial x this rank's polynomials) x 37 flop (94 complex)",
                "hbm": {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak if gbs else None, "peak_source": hbm_src + " copy bandwidth",
                        "note": "algorithmic bytes per chase step: RDS 96 B / 7 turnovers, CSS 208 B / 3, pencil 160 B / 11 (SURVEY 8d); the path is not HBM-bound"}}
    if prof and prof["us_main"] > 0:
        roofline["kernel_times_ms"] = {"setup": prof["us_setup"] / 1e3, "main": prof["us_main"] / 1e3, "small_windows": prof["us_small"] / 1e3, "finish": prof["us_finish"] / 1e3}
    if prof and st.get("schedule_used") == 3:
        # stream.cuh accounting (feeding threads): lock-steps, with an empty array, with bulges in flight but nothing to feed
        steps = max(prof["steps_lean"], 1)
        roofline["stream_profile"] = {"chase_cycles_per_lock_step": prof["cycles_units"] / steps, "lean_share_of_lock_steps": prof["cycles_lean"] / steps,
                                      "lock_steps": prof["steps_lean"], "empty_share": prof["steps_general"] / steps, "starved_share": prof["unit_visits"] / steps,
                                      "jobs": prof["lean_calls"], "phase_a_count": prof["cycles_driver"], "phase_a_cycles_each": prof["cycles_phase_a"] / max(prof["cycles_driver"], 1),
                                      "lean_cycles_per_step": prof["cycles_general"] / max(prof["cycles_lean"], 1), "nap_share_of_cycles": prof["cycles_pop"] / max(prof["cycles_units"], 1),
                                      "other_cycles_per_step": (prof["cycles_units"] - prof["cycles_general"] - prof["cycles_pop"]) / max(steps - prof["cycles_lean"] - prof["steps_general"], 1),
                                      "cta_last_busy_ms_sum": prof["cycles_loop"] / 1e3, "cta_first_busy_ms_sum": prof["lane_cycles_special"] / 1e3}
        if dv and any(dv[16:31]):
            roofline["stream_profile"]["plain_generic_step_phases"] = {nm: {ph: dv[16 + 5 * i + k] / max(dv[16 + 5 * i + 4], 1) for k, ph in enumerate(["ingest", "compute", "retire_control", "barrier_wait"])} for i, nm in enumerate(["thread0", "thread40", "last_thread"])}
        if dv and any(dv[8:16]):
            roofline["stream_profile"]["generic_step_cycles_by_special"] = {k: (dv[i] / dv[8 + i] if dv[8 + i] else None, dv[8 + i]) for i, k in enumerate(["none", "create", "last_a", "create+?", "last_b", "create+last_b", "last_a+last_b", "all"])}
    elif prof and prof["cycles_units"] > 0:
        busy = max(prof["cycles_units"], 1)
        roofline["warp_cycle_shares"] = {"phase_a": prof["cycles_phase_a"] / busy, "lean_chase": prof["cycles_lean"] / busy, "fill_drain_chase": prof["cycles_general"] / busy,
                                         "queue_wait_vs_busy": prof["cycles_pop"] / busy, "lean_cycles_per_step": prof["cycles_lean"] / max(prof["steps_lean"], 1)}
    line = {"metric": "polynomials/sec", "value": value, "unit": "polynomials/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config, "turnovers_per_sec": (useful_per_poly * total_batch / (ms_step * 1e-3)) if useful_per_poly else None, "roofline": roofline, "cpu_baseline": cb, "e2e": e2e,
            "gpu_launches": int(st["launches"] * args.steps), "gpu_launches_are": "this rank's kernels inside the timed region", "clocks": clk.summary(),
            "stats_per_step": {k: st[k] for k in ("turnovers", "sweeps", "exceptional", "unconverged", "fat_steps", "launches", "lanes", "schedule_used")},
            "unconverged": n_unconv, "roots_found": nroots_ok}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
