// rounds.cuh -- the pipelined multishift RDS path as ROUNDS of two kernels (the north-star path).
//
//   set-up kernel      one thread per polynomial: trim, closed forms, companion factorization
//   repeat until no polynomial is iterating:
//     phase A kernel   one thread per polynomial: the driver of AMRVW_algorithm.jl:10-138 up to the
//                      next fat step -- deflation bookkeeping, 1x1 / 2x2 extraction, shifts from the
//                      dense trailing block (shifts.cuh); small decoupled windows are only QUEUED
//     phase B kernel   one group of L lanes per polynomial: the systolic chase of L bulges
//                      (pipeline.cuh: lean loop + fill/drain loop)
//   small-window kernel  one thread per queued window: the reference's one-bulge iteration.  A window
//                      below a deflated rotator is decoupled from everything above it (no turnover
//                      of the upper window touches its rotators), so solving all of
//                      them at the end, in parallel, gives bit-identical eigenvalues
//   finish kernel      one thread per polynomial: sort, zero roots, statistics
//
// Why kernels and not one fused kernel (the round-1 design, 0.93 s on config 2):
//   * the serial phases ran on ONE lane of a 32-lane warp while the other 31 waited, 28% of every
//     warp's life (profiles/r01_warp_cycles.txt), and shared the FP64 pipe with the chase warps;
//     one thread per polynomial runs 32 of them per warp-instruction, off the chase's critical path;
//   * ptxas stacks the registers of a callee on top of what its caller keeps live, so inside the
//     fused kernel the systolic loop was scheduled into 93 registers (1119 cycles per step alone on
//     an SMSP) where it wants ~115 of its own (853 cycles: tools/ubench_lean.cu); here it is the
//     body of its own kernel;
//   * the rounds cost one grid-wide join per fat step: sum over rounds of the largest window is
//     1.04x the slowest polynomial's own step count (oracle trace, n = 1000), i.e. what the fused
//     kernel already paid to its slowest warp.
#pragma once
#include "pipeline.cuh"
#include <cstdio>
#include <cstdlib>

namespace amrvw {

// per-polynomial record of the rounds schedule (global memory)
struct PolyRec {
    LeaderState ls;
    PolyStats st;
    long long serial_turnovers;   // turnovers of the serial phases
    long long chase_turnovers;    // turnovers of phase B
    int N, K;                     // trimmed degree, number of trimmed zero roots
    int mode, nb;                 // pending fat step: 0 none, 1 shifted, 2 exceptional; bulges
    int iterating;
    int pad;
};
typedef SmallWinRec SmallWin;

// counters[0]: polynomials that got a fat step from the last phase A launch
// counters[1]: small windows queued
// counters[2..]: reserved
// counters[2]: sticky abort flag of the unit queue (a warp waited ~10 s for work: see queue_pop)
// counters[3]: compaction of the real unit kernel: polynomial (+1) waiting for a unit with a free lane group, 0 = none
// counters[4]: polynomials adopted so far (diagnostic)
enum { CNT_ACTIVE = 0, CNT_SMALL = 1, CNT_ABORT = 2, CNT_ORPHAN = 3, CNT_ADOPTED = 4, CNT_WORDS = 8 };

// ---------------------------------------------------------------- set-up
__global__ void __launch_bounds__(64) rounds_setup_kernel(Problem<double> pr) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= pr.nbatch) return;
    PolyRec rec;
    rec.ls.kk = 0; rec.ls.it_max = 0; rec.ls.ord = 0; rec.ls.unconv = 0; rec.ls.sp = 0; rec.ls.pc_sub = 0; rec.ls.pc_small = 0; rec.ls.pn_small = 0;
    rec.ls.ctr = {0, 1, 0, AMRVW_IT_COUNT, -1};
    rec.st = {0, 0, 0, 0, 0};
    rec.serial_turnovers = 0; rec.chase_turnovers = 0;
    rec.N = 0; rec.K = 0; rec.mode = 0; rec.nb = 0; rec.iterating = 0; rec.pad = 0;
    const long long off = pr.offsets[p];
    const int len = (int)(pr.offsets[p + 1] - off);
    cplx* out = pr.roots + off;
    int first, last;
    trim_zeros(pr.coeffs + off, len, first, last);
    if (first < 0) {
        finish_poly(pr, p, out, 0, 0, 0, rec.st);
    } else {
        const double* ps = pr.coeffs + off + first;
        const int ncoef = last - first + 1;
        rec.K = first;
        if (ncoef <= 3) {
            const int cnt = solve_simple(ps, ncoef, out);
            finish_poly(pr, p, out, cnt, rec.K, 0, rec.st);
        } else {
            const int N = ncoef - 1;
            Fact<double, false> F = make_fact<double, false>(pr, p, N);
            factor_companion(F, ps);
            for (int i = 0; i < N; ++i) out[i] = cplx(0.0, 0.0);
            rec.N = N;
            rec.ls.ctr = {0, 1, N - 1, AMRVW_IT_COUNT, -1};
            rec.ls.it_max = AMRVW_IT_MAX(N - 1);
            rec.iterating = 1;
        }
    }
    pr.rec[p] = rec;
}

// ---------------------------------------------------------------- unit formation by degree
// keys[p] = (maxN - N_p) << 32 | p, sorted ascending by one CTA with the all-ascending bitonic network of the
// finish kernel (slots >= n act as +infinity); order[i] = polynomial with the i-th largest trimmed degree,
// ties in batch order -- the identity for a batch of equal degrees.
__global__ void __launch_bounds__(1024) order_by_degree_kernel(const PolyRec* rec, const long long n, const int maxN, unsigned long long* keys, int* order) {
    for (long long p = threadIdx.x; p < n; p += blockDim.x) keys[p] = ((unsigned long long)(unsigned)(maxN - rec[p].N) << 32) | (unsigned long long)p;
    __syncthreads();
    long long np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (long long k = 2; k <= np2; k <<= 1) {
        for (long long j = k >> 1; j >= 1; j >>= 1) {
            const bool first = (j == (k >> 1));
            for (long long t = threadIdx.x; t < (np2 >> 1); t += blockDim.x) {
                const long long i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const long long l = first ? (i ^ (k - 1)) : (i | j);
                if (l < n && i < n) {
                    const unsigned long long x = keys[i], y = keys[l];
                    if (y < x) { keys[i] = y; keys[l] = x; }
                }
            }
            __syncthreads();
        }
    }
    for (long long p = threadIdx.x; p < n; p += blockDim.x) order[p] = (int)(keys[p] & 0xffffffffull);
}
// The same network for batches too large for one CTA: one launch per comparator stage over the whole grid
// (log2(n) (log2(n) + 1) / 2 launches: 210 for a million polynomials, a few microseconds each).
__global__ void order_keys_kernel(const PolyRec* rec, const long long n, const int maxN, unsigned long long* keys) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) keys[p] = ((unsigned long long)(unsigned)(maxN - rec[p].N) << 32) | (unsigned long long)p;
}
__global__ void order_stage_kernel(unsigned long long* keys, const long long n, const long long k, const long long j) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    const long long l = (j == (k >> 1)) ? (i ^ (k - 1)) : (i | j);
    if (l < n && i < n) {
        const unsigned long long x = keys[i], y = keys[l];
        if (y < x) { keys[i] = y; keys[l] = x; }
    }
}
__global__ void order_extract_kernel(const unsigned long long* keys, const long long n, int* order) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) order[p] = (int)(keys[p] & 0xffffffffull);
}
static cudaError_t launch_order_by_degree(const PolyRec* rec, const long long n, const int maxN, unsigned long long* keys, int* order, cudaStream_t st, int* launches) {
    if (n <= 8192) {
        order_by_degree_kernel<<<1, 1024, 0, st>>>(rec, n, maxN, keys, order);
        *launches += 1;
        return cudaGetLastError();
    }
    const int threads = 256;
    order_keys_kernel<<<(int)((n + threads - 1) / threads), threads, 0, st>>>(rec, n, maxN, keys);
    long long np2 = 1;
    while (np2 < n) np2 <<= 1;
    const int blocks = (int)(((np2 >> 1) + threads - 1) / threads);
    for (long long k = 2; k <= np2; k <<= 1)
        for (long long j = k >> 1; j >= 1; j >>= 1) { order_stage_kernel<<<blocks, threads, 0, st>>>(keys, n, k, j); *launches += 1; }
    order_extract_kernel<<<(int)((n + threads - 1) / threads), threads, 0, st>>>(keys, n, order);
    *launches += 2;
    return cudaGetLastError();
}

// ---------------------------------------------------------------- records, coherently
// PolyRec is read and written by whichever warp holds the unit: whole-record ld.cg / st.cg copies.
static_assert(sizeof(PolyRec) % 8 == 0, "PolyRec is copied in 8-byte words");
AMRVW_DI PolyRec load_rec(const PolyRec* g) {
    PolyRec r;
    const long long* src = reinterpret_cast<const long long*>(g);
    long long* dst = reinterpret_cast<long long*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(PolyRec) / 8); ++i) dst[i] = __ldcg(src + i);
    return r;
}
AMRVW_DI void store_rec(PolyRec* g, const PolyRec& r) {
    long long* dst = reinterpret_cast<long long*>(g);
    const long long* src = reinterpret_cast<const long long*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(PolyRec) / 8); ++i) __stcg(dst + i, src[i]);
}

// ---------------------------------------------------------------- the unit queue
// Work unit = one warp's worth of polynomials (G = 32/L groups).  A unit is either waiting for its
// phase A (park.t < 0) or in the middle of a fat step (park.t = next lock-step).  Units circulate
// through a device-side FIFO: a warp pops a unit, runs phase A if it is due, chases SEG lock-steps,
// parks the lanes' 24 doubles in global memory and pushes the unit back, until its polynomials are
// done.  Why a queue: 1500 units on 592 SMSPs leave some SMSPs with 3 resident warps and some with
// 2; an SMSP delivers the same ~1.67 steps per kilocycle either way (the FP64 pipe is ~72% busy with
// two warps), so warps advance at 1/1772 or 1/1213 steps per cycle.  Binding a fat step (or a whole
// polynomial) to a warp makes everything wait for the slow SMSPs -- 27% of the chase time in the
// round-per-kernel version of this file, 20% in the fused round-1 kernel (profiles/).  Through the
// queue every unit sees the average speed and no SMSP runs dry before the last units finish.
struct RoundQueue {
    unsigned head, tail, remaining, pad;
};
struct UnitPark {           // mutable per-unit state between visits; accessed with ld.cg / st.cg only
    int t;                  // next lock-step of the pending fat step; < 0: phase A is due
    int sp[8];              // deflation-stack sizes of the unit's polynomials
    int nturn[8];           // turnovers so far in this fat step
    int poly[8];            // the polynomials of the unit's lane groups, -1 = none (unit_kernel: a group whose polynomial has finished
                            // adopts another unit's leftover polynomial, see the compaction at the top of phase A)
    int pad[3];
};

// Pop a unit: >= 0 its id, -2 no work left, -3 the queue was aborted.  A warp that has waited ~10 s for its
// ticket's slot (something upstream is broken: a lost unit would otherwise hang the device) raises the sticky
// abort flag; every other warp sees it in its own wait loop and leaves too, and the host turns the flag into
// AMRVW_ERR_QUEUE_TIMEOUT instead of returning roots that silently miss the lost units' polynomials.
AMRVW_DI int queue_pop(RoundQueue* q, int* slots, const unsigned mask, unsigned* counters) {
    const unsigned k = atomicAdd(&q->head, 1u);
    int* slot = slots + (k & mask);
    for (unsigned spins = 0;; ++spins) {
        const int u = atomicExch(slot, -1);
        if (u >= 0) return u;
        if (*reinterpret_cast<volatile unsigned*>(&q->remaining) == 0u) return -2;
        if (*reinterpret_cast<volatile unsigned*>(counters + CNT_ABORT) != 0u) return -3;
        if (spins > (1u << 24)) { atomicExch(counters + CNT_ABORT, 1u); return -3; }
        __nanosleep(500);
    }
}

AMRVW_DI void st_cg_rot(Rot<double>* p, const Rot<double> r) { __stcg(reinterpret_cast<double2*>(p), make_double2(r.c, r.s)); }

// queue every unit that has a polynomial to iterate; one thread per unit
template <int L, class S>
__global__ void unit_enqueue_kernel(Problem<S> pr, RoundQueue* q, int* slots, unsigned mask, UnitPark* park) {
    constexpr int G = 32 / L;
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nunits = (pr.nbatch + G - 1) / G;
    if (u >= nunits) return;
    bool any = false;
    for (int g = 0; g < G; ++g) {
        const long long p = unit_poly(pr, u * G + g);
        if (p < pr.nbatch && pr.rec[p].iterating) any = true;
    }
    if (!any) return;
    UnitPark up;
    up.t = -1;
    for (int g = 0; g < 8; ++g) { up.sp[g] = 0; up.nturn[g] = 0; up.poly[g] = -1; }
    for (int g = 0; g < G; ++g) {
        const long long p = unit_poly(pr, u * G + g);
        if (p < pr.nbatch && pr.rec[p].iterating) up.poly[g] = (int)p;
    }
    up.pad[0] = up.pad[1] = up.pad[2] = 0;
    park[u] = up;
    __threadfence();
    atomicAdd(&q->remaining, 1u);
    const unsigned k = atomicAdd(&q->tail, 1u);
    atomicExch(slots + (k & mask), (int)u);
}

// shared-memory scratch of one warp for the shift computation
template <int MH> struct __align__(16) ShiftScratch {
    Rot<double> q[MH + 3], b[MH + 3], c[MH + 3];
    cplx e[MH + 3];
    double H[MH * (MH + 1)];
    double isg[MH + 2];
    double2 sn[12];
};

// ---------------------------------------------------------------- phase A of a unit
// Part 1, the driver (AMRVW_algorithm.jl:10-138): lane 0 of each group, up to the next fat step --
// deflation bookkeeping, 1x1 / 2x2 extraction, queueing of small windows.
AMRVW_NI int unit_driver(const Problem<double>& pr, const PipeParams pp, const int L, const long long p) {
    PolyRec rec = load_rec(pr.rec + p);
    if (!rec.iterating) return 0;
    Fact<double, false> F = make_fact<double, false>(pr, p, rec.N);
    GroupScratch sc;
    sc.q = sc.b = sc.c = nullptr; sc.e = nullptr; sc.H = nullptr; sc.isg = nullptr;
    cplx* out = pr.roots + pr.offsets[p];
    const int* dstack = pr.dstack + (pr.offsets[p] + (long long)pr.extra * p);
    int nb = 0;
    SmallSink sink;
    sink.list = pr.swlist; sink.count = pr.counters + CNT_SMALL; sink.poly = (int)p;
    const int mode = phase_a(F, rec.ls, out - 1, sc, nullptr, L, pp, nb, pr.seed, (uint64_t)p, rec.st, dstack, nullptr, 1, sink);
    rec.serial_turnovers += F.turnovers;
    rec.mode = mode; rec.nb = nb;
    if (mode == 0) rec.iterating = 0;
    store_rec(pr.rec + p, rec);
    return mode;
}
// Part 2, the shifts of a pending fat step (mode 3), each lane group on its own polynomial: stage the
// trailing block's rotators, write the block out densely, Hessenberg QR with the lanes sharing rows
// and columns (shifts.cuh), pair the shifts.
// L: lanes sharing the work; LB: bulges of a fat step (= L in the unit kernel, 32 NW in the chain kernel)
template <int MH>
AMRVW_NI void unit_shifts(const Problem<double>& pr, const PipeParams pp, const int L, const long long p, ShiftScratch<MH>& ws, const int lane, const unsigned gmask, const int Lb = 0) {
    constexpr int LDH = MH + 1;
    const int LB = Lb ? Lb : L;
    PolyRec rec = load_rec(pr.rec + p);
    const int m = rec.nb, stop = rec.ls.ctr.stop_index, k0 = stop - m, M = m + 1;
    Fact<double, false> F = make_fact<double, false>(pr, p, rec.N);
    __syncwarp(gmask);
    // staged copy with global indices: slot t <-> index k0 + t
    for (int t = lane; t <= m + 1; t += L) {
        ws.q[t] = ldcg_rot(F.Q + k0 + t); ws.b[t] = ldcg_rot(F.B + k0 + t); ws.c[t] = ldcg_rot(F.Ct + k0 + t);
    }
    __syncwarp(gmask);
    if (lane == 0) {
        const double sg = sign_(ws.q[0].c);
        ws.q[0] = make_rot<double>(sg == 0.0 ? 1.0 : sg, 0.0);
    }
    __syncwarp(gmask);
    bool ok = false;
    if (pp.dense) {
        dense_trailing_block_warp(ws.q - k0, ws.b - k0, ws.c - k0, k0, stop, ws.H, LDH, ws.isg, M, lane, gmask, L);
        ok = hqr_eigenvalues_warp(ws.H, LDH, M, ws.e, lane, gmask, L);
    }
    __syncwarp(gmask);
    if (lane == 0) {
        if (!ok) {   // core-chasing iteration on the staged copy (also AMRVW_SHIFTS_CHASE)
            Fact<double, false> W = F;
            W.Q = ws.q - k0; W.B = ws.b - k0; W.Ct = ws.c - k0; W.turnovers = 0;
            rec.serial_turnovers += chase_shifts(W, k0, stop, ws.e, pr.seed, (uint64_t)p, rec.ls.ord, ws.sn, 1);
        }
        rec.nb = pair_shifts(ws.e, m, pr.shifts + p * LB, LB, pp.reuse);
        rec.mode = 1;
        store_rec(pr.rec + p, rec);
    }
    __syncwarp(gmask);
}

// ---------------------------------------------------------------- the persistent kernel
// resident CTAs (= warps) per SM the register allocation is held to.  Config 2 has 1500 units: 11 per SM
// (186 registers, 1628 slots) keeps every unit resident and runs the lean loop at 1538 cycles per step, 12
// (168 registers) at 1583 -- 506 vs 517 ms; 8-10 leave units queueing, 14 spills (gpurun_out/sweep68/69.log)
#ifndef AMRVW_PIPE_MINBLOCKS
#define AMRVW_PIPE_MINBLOCKS 11
#endif
template <int L, int MH>
__global__ void __launch_bounds__(32, AMRVW_PIPE_MINBLOCKS) unit_kernel(Problem<double> pr, PipeParams pp, RoundQueue* q, int* slots, unsigned mask, UnitPark* park, StepState* parked,
                                                                        int seg_steps, int compact) {
    constexpr int G = 32 / L;
    constexpr int NT = 32;
    __shared__ __align__(16) double2 smem_snap[12 * NT + G * 3 * AMRVW_RING];
    __shared__ ShiftScratch<MH> shift_scratch[G];
    const int lane32 = threadIdx.x & 31;
    const int grp = lane32 / L, lane = lane32 % L;
    double2* const snap_t = smem_snap + threadIdx.x;                                   // [12][NT] step snapshots
    const unsigned stage_sa = (unsigned)__cvta_generic_to_shared(smem_snap + 12 * NT + grp * (3 * AMRVW_RING));   // look-ahead ring of lane 0
    const unsigned snap_sa = (unsigned)__cvta_generic_to_shared(snap_t);
    long long pc_all = 0, pc_lean = 0, pc_a = 0, pc_drv = 0, pc_pop = 0, pc_special = 0, pc_loop = 0;
    int pn_lean = 0, pn_all = 0, pn_calls = 0, pn_units = 0;

    for (;;) {
        // ---------------- pop a unit
        const long long pc_p0 = clock64();
        int u = -1;
        if (lane32 == 0) {
            u = queue_pop(q, slots, mask, pr.counters);
        }
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u < 0) break;
        __threadfence();
        const long long pc_u0 = clock64();
        pc_pop += pc_u0 - pc_p0;
        UnitPark* const up = park + u;
        long long p = pr.nbatch;      // the polynomial of this lane group (nbatch = none)
        {
            const int mapped = __ldcg(&up->poly[grp]);
            if (mapped >= 0) p = mapped;
        }
        int t0 = __ldcg(&up->t);

        // ---------------- phase A when due
        if (t0 < 0) {
            if (G == 2 && compact) {
                // Compaction (two polynomials per warp): a lane group whose polynomial has finished would idle until its
                // partner is done too.  One polynomial may wait in a device-wide slot: a unit with a free group takes it;
                // a half-empty unit that finds the slot empty leaves its own polynomial there and retires, so two half-empty
                // units become one full one.  `remaining` counts live units, never 0 while the slot is taken: a donor that turns
                // out to be the last unit takes its polynomial back, a retiring unit looks into the slot once more.
                // A polynomial's own sequence of phase A / fat steps does not depend on the unit that hosts it.
                unsigned* const orphan = pr.counters + CNT_ORPHAN;
                const unsigned heldm = __ballot_sync(0xffffffffu, lane == 0 && p < pr.nbatch);
                const int nheld = __popc(heldm);
                if (nheld < 2) {
                    const int mine = nheld == 1 ? (int)__shfl_sync(0xffffffffu, (int)p, (heldm & 1u) ? 0 : L) : -1;
                    int act = 0, x = 0;      // act 1: adopt polynomial x - 1, 2: this unit is gone
                    if (lane32 == 0) {
                        x = (int)atomicExch(orphan, 0u);
                        if (x > 0) act = 1;
                        else if (nheld == 1) {
                            // nothing waits: offer ours -- while enough units are alive that a free group turns up soon
                            if (*reinterpret_cast<volatile unsigned*>(&q->remaining) >= 32u && atomicCAS(orphan, 0u, (unsigned)mine + 1u) == 0u) {
                                __threadfence();
                                const unsigned old = atomicSub(&q->remaining, 1u);
                                act = 2;
                                if (old <= 1u && atomicCAS(orphan, (unsigned)mine + 1u, 0u) == (unsigned)mine + 1u) { atomicAdd(&q->remaining, 1u); act = 0; }
                            }
                        } else {
                            // nothing here, nothing waiting: retire, then look once more (a donor may have counted on this unit)
                            atomicSub(&q->remaining, 1u);
                            __threadfence();
                            x = (int)atomicExch(orphan, 0u);
                            if (x > 0) { atomicAdd(&q->remaining, 1u); act = 1; } else act = 2;
                        }
                        if (act == 1) {
                            __stcg(&up->poly[(heldm & 1u) ? 1 : 0], x - 1);
                            atomicAdd(pr.counters + CNT_ADOPTED, 1u);
                        } else if (act == 2) {
                            __stcg(&up->poly[0], -1); __stcg(&up->poly[1], -1);
                        }
                    }
                    act = __shfl_sync(0xffffffffu, act, 0); x = __shfl_sync(0xffffffffu, x, 0);
                    if (act == 2) continue;
                    if (act == 1) {
                        __threadfence();
                        if (grp == ((heldm & 1u) ? 1 : 0)) p = x - 1;
                    }
                }
            }
            int m0 = 0;
            if (lane == 0 && p < pr.nbatch) {
                m0 = unit_driver(pr, pp, L, p);
                if (m0 == 0) __stcg(&up->poly[grp], -1);      // finished: the group is free from the next phase A on
            }
            __syncwarp();
            pc_drv += clock64() - pc_u0;
            {   // the groups compute their shifts side by side
                const int mg = __shfl_sync(0xffffffffu, m0, grp * L);
                const unsigned gmask = (L == 32 ? 0xffffffffu : ((1u << L) - 1u)) << (grp * L);
                if (mg == 3) unit_shifts<MH>(pr, pp, L, p, shift_scratch[grp], lane, gmask);
                __syncwarp();
            }
            pc_a += clock64() - pc_u0;
            if (!__any_sync(0xffffffffu, m0 != 0)) {   // every polynomial of the unit is finished
                if (G == 2 && compact) {
                    // come back once more as an empty unit: the compaction above retires it or gives it a waiting polynomial
                    if (lane32 == 0) {
                        __stcg(&up->t, -1);
                        __threadfence();
                        const unsigned k = atomicAdd(&q->tail, 1u);
                        atomicExch(slots + (k & mask), u);
                    }
                } else {
                    if (lane32 == 0) atomicSub(&q->remaining, 1u);
                }
                continue;
            }
            if (lane == 0) {
                int sp0 = 0;
                if (m0 != 0) sp0 = __ldcg(&pr.rec[p].ls.sp);
                __stcg(&up->sp[grp], sp0); __stcg(&up->nturn[grp], 0);
            }
            __syncwarp();
            t0 = 0;
        }

        // ---------------- the unit's fat steps (every lane of a group reads the same record)
        int mode = 0, nb = 0, start = 1, stop = 0, ord0 = 0, N = 0;
        if (p < pr.nbatch) {
            const PolyRec* r = pr.rec + p;
            if (__ldcg(&r->iterating)) {
                mode = __ldcg(&r->mode); nb = __ldcg(&r->nb); start = __ldcg(&r->ls.ctr.start_index); stop = __ldcg(&r->ls.ctr.stop_index);
                ord0 = __ldcg(&r->ls.ord); N = __ldcg(&r->N);
            }
        }
        const bool act = mode != 0;
        const int T = act ? (stop - start + 3 * L + 1) : 0;
        const int Tmax = __reduce_max_sync(0xffffffffu, T);
        const int t1 = min(Tmax, t0 + seg_steps);
        int sp = __ldcg(&up->sp[grp]);
        int nturn = 0;

        Fact<double, false> F;
        F.Q = F.B = F.Ct = F.WB = F.WCt = nullptr; F.DQ = F.DR = F.WD = nullptr; F.N = N; F.turnovers = 0;
        int* dstack = nullptr;
        if (act) {
            F = make_fact<double, false>(pr, p, N);
            dstack = pr.dstack + (pr.offsets[p] + (long long)pr.extra * p);
        }
        Rot<double>* const FQ = F.Q;
        Rot<double>* const FB = F.B;
        Rot<double>* const FC = F.Ct;

        Rot<double> q0, q1, q2, b0, b1, b2, c0, c1, c2, U, V, W;
        if (t0 == 0) {
            q0 = q1 = q2 = b0 = b1 = b2 = c0 = c1 = c2 = U = V = W = make_rot<double>(1.0, 0.0);
        } else {
            const StepState* z = parked + ((long long)u * 32 + lane32);
            q0 = ldcg_rot(&z->q0); q1 = ldcg_rot(&z->q1); q2 = ldcg_rot(&z->q2);
            b0 = ldcg_rot(&z->b0); b1 = ldcg_rot(&z->b1); b2 = ldcg_rot(&z->b2);
            c0 = ldcg_rot(&z->c0); c1 = ldcg_rot(&z->c1); c2 = ldcg_rot(&z->c2);
            U = ldcg_rot(&z->U); V = ldcg_rot(&z->V); W = ldcg_rot(&z->W);
        }
        const bool has_bulge = act && lane < nb;
        if (act && lane == 0) {
            for (int a = 0; a < AMRVW_RING - 1; ++a) {
                const int nidx = min(start + t0 + a, stop + 1);
                const unsigned sg = stage_sa + ((t0 + a) & (AMRVW_RING - 1)) * 48;
                cp_async16_sa(sg, FQ + nidx); cp_async16_sa(sg + 16, FB + nidx); cp_async16_sa(sg + 32, FC + nidx);
                cp_async_commit();
            }
        }
        const long long pc_lp0 = clock64();
        for (int t = t0; t < t1; ++t) {
            {   // every lane of the warp at a regular position (or idle for good): the lean loop
                const bool g_dead = !act || t >= T;
                const bool g_main = act && nb == L && t >= 3 * L && t <= stop - start;
                if (__all_sync(0xffffffffu, g_dead || g_main) && __any_sync(0xffffffffu, g_main)) {
                    const int te = min(t1, __reduce_min_sync(0xffffffffu, g_main ? stop - start + 1 : 0x7fffffff));
                    const long long pc_l0 = clock64();
                    StepState z = {q0, q1, q2, b0, b1, b2, c0, c1, c2, U, V, W};
                    const int role = (g_main ? 1 : 0) | (lane == 0 ? 2 : 0) | (lane == L - 1 ? 4 : 0);
                    sp = lean_chase<L, NT>(z, FQ, FB, FC, stage_sa, snap_sa, dstack, sp, start, t, te, role);
                    q0 = z.q0; q1 = z.q1; q2 = z.q2; b0 = z.b0; b1 = z.b1; b2 = z.b2; c0 = z.c0; c1 = z.c1; c2 = z.c2; U = z.U; V = z.V; W = z.W;
                    if (g_main) nturn += 7 * (te - t);
                    pc_lean += clock64() - pc_l0; pn_lean += te - t; pn_calls += 1;
                    t = te - 1;
                    continue;
                }
            }
            const int tp = t - 3 * lane;
            // ingest: what the previous lane finished last step (lane 0: from memory)
            Rot<double> iq = shfl_up_rot(q0, L), ib = shfl_up_rot(b0, L), ic = shfl_up_rot(c0, L);
            if (lane == 0) {
                cp_async_wait_ring();
                const unsigned sg = stage_sa + (t & (AMRVW_RING - 1)) * 48;
                iq = lds_rot(sg); ib = lds_rot(sg + 16); ic = lds_rot(sg + 32);
                if (act) {
                    const int nidx = min(start + t + AMRVW_RING - 1, stop + 1);
                    const unsigned sn2 = stage_sa + ((t + AMRVW_RING - 1) & (AMRVW_RING - 1)) * 48;
                    cp_async16_sa(sn2, FQ + nidx); cp_async16_sa(sn2 + 16, FB + nidx); cp_async16_sa(sn2 + 32, FC + nidx);
                }
                cp_async_commit();
            }
            q0 = q1; q1 = q2; q2 = iq;
            b0 = b1; b1 = b2; b2 = ib;
            c0 = c1; c1 = c2; c2 = ic;
            const int i = start + tp - 2;   // position of this lane's bulge; window = indices i..i+2
            if (has_bulge && tp >= 2 && i <= stop - 1) {
                if (tp == 2) {
                    const long long pc_s0 = clock64();
                    StepState z = {q0, q1, q2, b0, b1, b2, c0, c1, c2, U, V, W};
                    create_and_absorb_at(z, F, start, stop, mode, pr.shifts + p * L, lane, pr.seed, (uint64_t)p, ord0 + lane);
                    q0 = z.q0; q1 = z.q1; U = z.U; V = z.V; W = z.W;
                    nturn += 1;
                    pc_special += clock64() - pc_s0;
                }
                if (i <= stop - 2) {
                    // regular position: the 7 turnovers of one chase step (RDS.jl:42-70, 97-108) as ONE
                    // straight-line block; a degenerate turnover (flagged, essentially never) redoes the
                    // step on the general path from the shared-memory snapshot
#ifndef AMRVW_STRICT
                    AMRVW_CHASE_REGULAR(snap_t, NT);
#else
                    AMRVW_CHASE_GENERAL();
#endif
                } else {
                    const long long pc_s0 = clock64();
                    StepState z = {q0, q1, q2, b0, b1, b2, c0, c1, c2, U, V, W};
                    chase_step_final_at(z, F, stop);
                    q0 = z.q0; q1 = z.q1; q2 = z.q2; b0 = z.b0; b1 = z.b1; b2 = z.b2; c0 = z.c0; c1 = z.c1; c2 = z.c2; U = z.U; V = z.V; W = z.W;
                    pc_special += clock64() - pc_s0;
                }
                nturn += 7;
            }
            // retire: the last lane writes index i back
            if (lane == L - 1 && act && i >= start && i <= stop + 1) {
                stg_rot(FQ + i, q0); stg_rot(FB + i, b0); stg_rot(FC + i, c0);
                if (i <= stop && fabs(q0.s) <= AMRVW_EPS) dstack[sp++] = i;   // deflatable: record for phase A
            }
        }
        pc_loop += clock64() - pc_lp0;
        if (lane == 0) cp_async_wait_all();   // drain look-ahead requests past the end of the segment
        sp = __shfl_sync(0xffffffffu, sp, grp * L + L - 1);
        for (int d = L / 2; d >= 1; d >>= 1) nturn += __shfl_down_sync(0xffffffffu, nturn, d, L);
        pn_all += t1 - t0; pn_units += 1;
        if (t1 < Tmax) {
            // ---------------- park the lanes' state
            StepState* z = parked + ((long long)u * 32 + lane32);
            st_cg_rot(&z->q0, q0); st_cg_rot(&z->q1, q1); st_cg_rot(&z->q2, q2);
            st_cg_rot(&z->b0, b0); st_cg_rot(&z->b1, b1); st_cg_rot(&z->b2, b2);
            st_cg_rot(&z->c0, c0); st_cg_rot(&z->c1, c1); st_cg_rot(&z->c2, c2);
            st_cg_rot(&z->U, U); st_cg_rot(&z->V, V); st_cg_rot(&z->W, W);
            if (lane == 0) { __stcg(&up->sp[grp], sp); __stcg(&up->nturn[grp], __ldcg(&up->nturn[grp]) + nturn); }
            if (lane32 == 0) __stcg(&up->t, t1);
        } else {
            // ---------------- fat step through: back to the record, phase A is due
            if (act && lane == 0) {
                PolyRec* r = pr.rec + p;
                PolyRec rec = load_rec(r);
                rec.ls.sp = sp;
                rec.st.sweeps += nb;
                rec.chase_turnovers += nturn + __ldcg(&up->nturn[grp]);
                if (mode == 2) { rec.st.exceptional += nb; rec.ls.ord += nb; rec.ls.ctr.it_count = AMRVW_IT_COUNT - 1; }
                store_rec(r, rec);
            }
            if (lane32 == 0) __stcg(&up->t, -1);
        }
        // ---------------- hand the unit back
        __threadfence();
        __syncwarp();
        if (lane32 == 0) {
            const unsigned k = atomicAdd(&q->tail, 1u);
            atomicExch(slots + (k & mask), u);
        }
        pc_all += clock64() - pc_u0;
    }
    if (pr.prof && lane32 == 0) {
        atomicAdd(pr.prof + 0, (unsigned long long)pc_all);
        atomicAdd(pr.prof + 1, (unsigned long long)pc_a);
        atomicAdd(pr.prof + 2, (unsigned long long)pc_lean);
        atomicAdd(pr.prof + 3, (unsigned long long)(pc_all - pc_lean - pc_a));
        atomicAdd(pr.prof + 4, (unsigned long long)pn_lean);
        atomicAdd(pr.prof + 5, (unsigned long long)(pn_all - pn_lean));
        atomicAdd(pr.prof + 6, (unsigned long long)pn_calls);
        atomicAdd(pr.prof + 7, (unsigned long long)pn_units);
        atomicAdd(pr.prof + 10, (unsigned long long)pc_drv);
        atomicAdd(pr.prof + 13, (unsigned long long)pc_pop);
        atomicAdd(pr.prof + 14, (unsigned long long)pc_loop);
    }
    if (pr.prof) {   // per-lane time in bulge creation / last-position steps (the lanes take them one at a time)
        unsigned long long v = (unsigned long long)pc_special;
        for (int d = 16; d >= 1; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
        if (lane32 == 0) atomicAdd(pr.prof + 15, v);
    }
}

// ---------------------------------------------------------------- small windows
// SM: largest window (rotators) this instantiation stages in per-thread local memory
template <int SM>
__global__ void __launch_bounds__(64) rounds_small_kernel(Problem<double> pr) {
    const unsigned w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= pr.counters[CNT_SMALL]) return;
    const SmallWin sw = reinterpret_cast<const SmallWin*>(pr.swlist)[w];
    const long long p = sw.poly;
    const PolyRec* r = pr.rec + p;
    Fact<double, false> F = make_fact<double, false>(pr, p, r->N);
    cplx* E = pr.roots + pr.offsets[p] - 1;
    Rot<double> sq[SM + 3], sb[SM + 3], scc[SM + 3];
    double2 sn[12];
    GroupScratch sc;
    sc.q = sq; sc.b = sb; sc.c = scc; sc.e = nullptr; sc.H = nullptr; sc.isg = nullptr;
    Fact<double, false> W = stage_window(F, sc, sw.lo - 1, sw.hi);
    Counter c2 = {sw.lo - 1, sw.lo, sw.hi, sw.it_count, -1};
    PolyStats st = {0, 0, 0, 0, 0};
    int ord = 1000000 + sw.lo;   // exceptional shifts of a deferred window: ordinal keyed by the window, not by time
    const WindowStep wstep = {sn, 1};
    const int left = drive_window(W, sw.lo, sw.hi, c2, E, AMRVW_IT_MAX(r->N - 1), pr.seed, (uint64_t)p, ord, st, wstep);
    PolyRec* rw = pr.rec + p;
    atomicAdd((unsigned long long*)&rw->serial_turnovers, (unsigned long long)W.turnovers);
    atomicAdd(&rw->st.sweeps, st.sweeps);
    if (st.exceptional) atomicAdd(&rw->st.exceptional, st.exceptional);
    if (left) atomicAdd(&rw->ls.unconv, left);
}

// ---------------------------------------------------------------- finish
// One CTA per polynomial: sort (LinearAlgebra.sorteig!, AMRVW_algorithm.jl:134: by (real, imag) in
// isless order), zero roots, statistics.  The sort is a bitonic network in its all-ascending form
// (first step of every merge pairs i with i ^ (2^(k+1) - 1), the others i with i ^ 2^j; every
// comparator puts the smaller element at the lower index), so slots >= n act as +infinity without
// being stored; it runs in shared memory when the polynomial fits, otherwise in place in global
// memory.  eig_less is a total order, so the result is the one the serial heap sort gives.
#define AMRVW_SORT_SMEM 4096
template <class S>
__global__ void __launch_bounds__(256) rounds_finish_kernel(Problem<S> pr, int smem_elems) {
    extern __shared__ __align__(16) unsigned char sort_smem[];
    cplx* tile = reinterpret_cast<cplx*>(sort_smem);
    const long long p = blockIdx.x;
    const PolyRec* recp = pr.rec + p;
    const int n = recp->N;
    if (n <= 0) return;   // closed forms were finished by the set-up kernel
    cplx* out = pr.roots + pr.offsets[p];
    cplx* a = out;
    const bool in_smem = n <= smem_elems;
    if (in_smem) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) tile[i] = out[i];
        a = tile;
    }
    __syncthreads();
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j >= 1; j >>= 1) {
            const bool first = (j == (k >> 1));
            for (int t = threadIdx.x; t < (np2 >> 1); t += blockDim.x) {
                // t-th comparator of this step: lower index i has bit j clear
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = first ? (i ^ (k - 1)) : (i | j);
                if (l < n && i < n) {
                    const cplx x = a[i], y = a[l];
                    if (eig_less(y, x)) { a[i] = y; a[l] = x; }
                }
            }
            __syncthreads();
        }
    }
    if (in_smem) for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = tile[i];
    if (threadIdx.x == 0) {
        PolyRec rec = *recp;
        rec.st.turnovers = rec.serial_turnovers + rec.chase_turnovers;
        rec.st.unconverged = rec.ls.unconv + (rec.ls.ctr.stop_index >= 1 ? rec.ls.ctr.stop_index + 1 : 0);
        finish_poly(pr, p, out, rec.N, rec.K, rec.st.unconverged, rec.st);
    }
}

// ---------------------------------------------------------------- host side
struct RoundsTimes { float setup_ms, main_ms, small_ms, finish_ms; };
struct QueueBufs { RoundQueue* q; int* slots; unsigned mask; UnitPark* park; StepState* parked; };

// Chooses L (bulges in flight per polynomial) from the degree and the batch: enough lanes to fill the
// GPU, not more than the windows can feed (a fat step needs a window > small >= 2L/reuse).
// nbatch: the load the lane count is chosen for (the batch, or more when other degree classes share the device);
// nresident: the polynomials that will really hold CTAs (chain or stream)
static void choose_pipe_params(const amrvw_options& opt, long long nbatch, int max_len, int sm_count, int& L, PipeParams& pp, long long nresident = -1) {
    if (nresident < 0) nresident = nbatch;
    const int degree = max_len - 1;
    L = opt.lanes_per_poly;
    if (L == 0) {
        if (degree < 48) L = 4;
        else if (degree < 200) L = 8;
        else if (degree < 1200 && nbatch * 16 > (long long)sm_count * 512) L = 8;
        else L = 16;
        // measured on B200 (gpurun_out/sweep58.log, sweep59.log; degree 3000): 16 lanes win from ~2000 polynomials
        // up (400 vs 352 ms at 1500, 435 vs 423 at 2000, 455 vs 495 at 2400), 32 lanes below that, and four chained
        // warps (chain.cuh, 128 bulges) while every CTA is resident, <= 3 per SM (226 vs 252 ms at 444 polynomials,
        // 348 vs 255 at 520; config 5: 800 vs 1720 ms); at degree 1500 the chain loses (105 vs 95 ms at 400)
        if (degree >= 1200 && nbatch * 32 <= (long long)sm_count * 410) L = 32;
        // round 2 (gpurun_out/r2b_sweep.txt, r2k_sweep.txt; degree 3000): 375 polynomials 205-211 ms streamed (stream.cuh) against
        // 345 ms with one CTA per polynomial (two CTAs fit an SM, the 79 polynomials beyond 296 wait for a second wave);
        // 750: 303 ms streamed, 247 ms with 32 lanes on the unit queue; 1500: 562 vs 333 ms
        // streamed: 64 and 128 bulges tie at 375 polynomials (206 / 205 ms), 64 wins above (750: 267 / 303 ms) and executes
        // 1.75x instead of 2.39x the reference's turnovers
        if (L == 32 && degree >= 2500 && nbatch * 2 <= (long long)sm_count * 7) L = nresident > (long long)sm_count * 2 ? 64 : 128;
    }
    // chained warps: one CTA per polynomial (chain.cuh) while there are fewer polynomials than SMs can hold two of;
    // above that the CTAs take the fat steps of all polynomials as one stream (stream.cuh): no fill/drain, phase A hidden
    pp.stream = 0;
    pp.stream_ctas = opt.stream_ctas_per_sm;
    if (opt.schedule == AMRVW_SCHEDULE_STREAMED && opt.lanes_per_poly == 0) L = 128;
    if (opt.schedule == AMRVW_SCHEDULE_STREAMED && (L == 16 || L == 32)) pp.stream = 1;     // sub-warp pipelines (stream.cuh, second half)
    if (L == 128 || L == 64) {
        if (opt.schedule == AMRVW_SCHEDULE_STREAMED) pp.stream = 1;
        else if (opt.schedule == AMRVW_SCHEDULE_AUTO && opt.lanes_per_poly == 0 && nresident > (long long)sm_count * 2) pp.stream = 1;     // more polynomials than resident chain CTAs
    }
    pp.reuse = opt.shift_reuse ? opt.shift_reuse : (L == 32 ? 4 : (pp.stream && L >= 64 ? L / 16 : 2));   // 32 lanes: 16 shifts x 4 beat 32 x 2 by 4-8% (gpurun_out/sweep63.log)
    if (pp.reuse > L / 2) pp.reuse = L / 2 > 0 ? L / 2 : 1;
    while (2 * L / pp.reuse > 64) pp.reuse *= 2;     // the largest dense trailing block is 64 x 64
    const int mshift = 2 * L / pp.reuse;
    pp.small = opt.small_window ? opt.small_window : (mshift + 8 > 24 ? mshift + 8 : 24);
    if (pp.small < mshift) pp.small = mshift;      // a fat step needs delta - 1 >= m
    if (pp.small > 96) pp.small = 96;
    pp.ms = (mshift > pp.small ? mshift : pp.small) + 4;
    pp.dense = opt.shift_solver != AMRVW_SHIFTS_CHASE;
    pp.mh = mshift;
}

template <int L, int MH>
static cudaError_t launch_units(const Problem<double>& pr, const PipeParams& pp, const QueueBufs& qb, int seg_steps, int sm_count, cudaStream_t st) {
    constexpr int G = 32 / L;
    const long long nunits = (pr.nbatch + G - 1) / G;
    unit_enqueue_kernel<L, double><<<(int)((nunits + 127) / 128), 128, 0, st>>>(pr, qb.q, qb.slots, qb.mask, qb.park);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, unit_kernel<L, MH>, 32, 0)) != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const long long resident = (long long)sm_count * per_sm;     // every warp of the grid must be resident: they wait on each other
    const int blocks = (int)(nunits < resident ? nunits : resident);
    // compaction of half-empty units frees warps for units that wait in the queue; when every unit is resident anyway it only
    // costs (config 2, 1500 units on 1628 resident warps: 509.8 -> 513.1 ms), so it is on when units outnumber the resident warps
    int compact = (G == 2 && nunits > resident) ? 1 : 0;
    if (const char* ev = std::getenv("AMRVW_COMPACT")) compact = (G == 2 && std::atoi(ev) != 0) ? 1 : 0;      // development: force on / off
    unit_kernel<L, MH><<<blocks, 32, 0, st>>>(pr, pp, qb.q, qb.slots, qb.mask, qb.park, qb.parked, seg_steps, compact);
    return cudaGetLastError();
}
template <int L>
static cudaError_t launch_units_mh(const Problem<double>& pr, const PipeParams& pp, const QueueBufs& qb, int seg_steps, int sm_count, cudaStream_t st) {
    // the trailing block has order 2L/reuse <= 2L: only those scratch sizes exist
    if (pp.mh <= 16) return launch_units<L, 16>(pr, pp, qb, seg_steps, sm_count, st);
    if constexpr (L >= 16) {
        if (pp.mh <= 32) return launch_units<L, 32>(pr, pp, qb, seg_steps, sm_count, st);
    }
    if constexpr (L >= 32) return launch_units<L, 64>(pr, pp, qb, seg_steps, sm_count, st);
    return cudaErrorInvalidValue;
}

// chain.cuh: one CTA of NW chained warps per polynomial (small batches of large polynomials)
template <int NW, int MH> static cudaError_t launch_chain(const Problem<double>& pr, const PipeParams& pp, int sm_count, cudaStream_t st);
// stream.cuh: chained warps fed by a stream of fat steps of different polynomials
struct StreamBufs;
template <int NW, int MH> static cudaError_t launch_stream(const Problem<double>& pr, const PipeParams& pp, const StreamBufs& sb, int sm_count, cudaStream_t st);
template <int L, int MH, int NW, int NS> static cudaError_t launch_ustream(const Problem<double>& pr, const PipeParams& pp, const StreamBufs& sb, int sm_count, cudaStream_t st);

// Queues the whole schedule on `st`: set-up, the persistent unit kernel, small windows, finish.
// Nothing here waits for the device unless `times` is asked for.
static cudaError_t read_round_times(cudaEvent_t* ev, RoundsTimes* times, cudaStream_t st) {
    cudaError_t e;
    cudaEventRecord(ev[4], st);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    cudaEventElapsedTime(&times->setup_ms, ev[0], ev[1]);
    cudaEventElapsedTime(&times->main_ms, ev[1], ev[2]);
    cudaEventElapsedTime(&times->small_ms, ev[2], ev[3]);
    cudaEventElapsedTime(&times->finish_ms, ev[3], ev[4]);
    return cudaSuccess;
}
static cudaError_t launch_rounds_rds(Problem<double> pr, const amrvw_options& opt, int max_len, long long total, int sm_count, cudaStream_t st, int* launches, cudaEvent_t* ev,
                                     RoundsTimes* times, const QueueBufs& qb, int* lanes_out = nullptr, unsigned long long* order_keys = nullptr, const StreamBufs* sb = nullptr) {
    int L; PipeParams pp;
    choose_pipe_params(opt, pr.load_hint > pr.nbatch ? pr.load_hint : pr.nbatch, max_len, sm_count, L, pp, pr.nbatch);
    if (lanes_out) *lanes_out = L;
    cudaError_t e;
    const int nb64 = (int)((pr.nbatch + 63) / 64);
    if ((e = cudaMemsetAsync(pr.counters, 0, CNT_WORDS * sizeof(unsigned), st)) != cudaSuccess) return e;
    if (L <= 32 && !(pp.stream && sb)) {    // the unit queue (chained warps and streams have none)
        if ((e = cudaMemsetAsync(qb.q, 0, sizeof(RoundQueue), st)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(qb.slots, 0xFF, (size_t)(qb.mask + 1) * sizeof(int), st)) != cudaSuccess) return e;
    }
    const int seg_steps = 512;   // lock-steps a warp chases before it parks the unit and takes the next one (128: 549 ms, 512: 530 ms on config 2)
    if (times) cudaEventRecord(ev[0], st);
    rounds_setup_kernel<<<nb64, 64, 0, st>>>(pr);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (pr.order) {
        if ((e = launch_order_by_degree(pr.rec, pr.nbatch, max_len, order_keys, const_cast<int*>(pr.order), st, launches)) != cudaSuccess) return e;
    }
    if (times) cudaEventRecord(ev[1], st);
    switch (L) {
        case 4: e = launch_units_mh<4>(pr, pp, qb, seg_steps, sm_count, st); break;
        case 8: e = launch_units_mh<8>(pr, pp, qb, seg_steps, sm_count, st); break;
        case 16:
            if (pp.stream && sb) e = pp.mh <= 16 ? launch_ustream<16, 16, 4, 1>(pr, pp, *sb, sm_count, st) : launch_ustream<16, 32, 4, 1>(pr, pp, *sb, sm_count, st);
            else e = launch_units_mh<16>(pr, pp, qb, seg_steps, sm_count, st);
            break;
        case 32:
            if (pp.stream && sb) e = pp.mh <= 16 ? launch_ustream<32, 16, 4, 1>(pr, pp, *sb, sm_count, st) : (pp.mh <= 32 ? launch_ustream<32, 32, 4, 1>(pr, pp, *sb, sm_count, st) : cudaErrorInvalidValue);
            else e = launch_units_mh<32>(pr, pp, qb, seg_steps, sm_count, st);
            break;
        case 64:
            if (pp.stream && sb) e = pp.mh <= 32 ? launch_stream<2, 32>(pr, pp, *sb, sm_count, st) : launch_stream<2, 64>(pr, pp, *sb, sm_count, st);
            else e = pp.mh <= 32 ? launch_chain<2, 32>(pr, pp, sm_count, st) : launch_chain<2, 64>(pr, pp, sm_count, st);
            break;
        case 128:
            if (pp.stream && sb) e = pp.mh <= 32 ? launch_stream<4, 32>(pr, pp, *sb, sm_count, st) : launch_stream<4, 64>(pr, pp, *sb, sm_count, st);
            else e = pp.mh <= 32 ? launch_chain<4, 32>(pr, pp, sm_count, st) : launch_chain<4, 64>(pr, pp, sm_count, st);
            break;
        case 256: e = launch_chain<8, 64>(pr, pp, sm_count, st); break;
        default: e = launch_units_mh<32>(pr, pp, qb, seg_steps, sm_count, st); break;
    }
    if (e != cudaSuccess) return e;
    if (times) cudaEventRecord(ev[2], st);
    {   // deferred small windows: the grid covers the queue's capacity, the kernel reads the count
        const long long cap = total / 3 + pr.nbatch + 16;
        const int blocks = (int)((cap + 63) / 64);
        if (pp.small <= 32) rounds_small_kernel<32><<<blocks, 64, 0, st>>>(pr);
        else rounds_small_kernel<96><<<blocks, 64, 0, st>>>(pr);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if (times) cudaEventRecord(ev[3], st);
    {
        const int smem_elems = max_len - 1 <= AMRVW_SORT_SMEM ? (max_len > 1 ? max_len - 1 : 1) : AMRVW_SORT_SMEM;
        const size_t smem = (size_t)smem_elems * sizeof(cplx);
        if ((e = cudaFuncSetAttribute(rounds_finish_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        rounds_finish_kernel<double><<<(int)pr.nbatch, 256, smem, st>>>(pr, smem_elems);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    *launches += (pp.stream && sb) ? 6 : (L > 32 ? 4 : 5);
    if (times) return read_round_times(ev, times, st);
    return cudaSuccess;
}

}  // namespace amrvw
