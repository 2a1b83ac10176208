// chain.cuh -- the pipelined RDS path for SMALL batches of LARGE polynomials: one CTA per polynomial,
// its NW warps chained into one systolic array of 32 NW bulges.
//
// The unit kernel (rounds.cuh) keeps at most 32 bulges of a polynomial in flight, one warp per
// polynomial: 64 polynomials of degree 10 000 (BASELINE config 5, per GPU) occupy 64 of the 592 SMSPs
// and each advances at the latency of one dependent chase step (~1100 cycles).  Here lane 31 of warp w
// hands the rotator triple it has finished to lane 0 of warp w+1 through a double-buffered
// shared-memory slot, and the NW warps take their lock-steps together (one bar.sync per step, ~2% of a
// step): 32 NW bulges, three indices apart, per fat step, with 64 NW / reuse shifts from the dense trailing
// block.  Everything is uniform over the CTA (one polynomial), so there is no unit queue, no parking
// and no residency requirement: a CTA takes polynomials blockIdx.x, blockIdx.x + gridDim.x, ... to
// completion, alternating phase A (driver on thread 0, shifts on warp 0) and the chase.
//
// Same operations as the unit kernel; the -DAMRVW_STRICT build is bit-identical to the oracle's
// sequential multishift schedule with L = 32 NW (tests).
#pragma once
#include "rounds.cuh"

namespace amrvw {

template <int NW> struct ChainSmem {
    double2 ring[3 * AMRVW_RING];       // look-ahead ring of thread 0 (global -> shared, cp.async)
    double2 hand[NW][2][3];              // hand[w][parity]: triple leaving warp w at the end of a step of that parity
    unsigned long long nturn;
    int mode;
};

// Steps [t0, t1) during which every lane of the CTA is at a regular chase position: ingest -> 7 turnovers
// -> retire -> bar.sync (the chain twin of lean_chase, pipeline.cuh)
// BAR: named barrier of the NW chase warps (0 = the whole CTA, __syncthreads; stream.cuh has service warps beside them)
template <int NW, int BAR = 0>
AMRVW_DI int chain_lean(StepState& z, Rot<double>* gq, Rot<double>* gb, Rot<double>* gc, const unsigned stage_sa, const unsigned snap_sa, const unsigned hand_sa,
                        int* dstack, int sp, const int start, const int t0, const int t1, const int warp, const int lane32) {
    constexpr int LE = 32 * NW;
    [[maybe_unused]] constexpr unsigned SS = LE * 16;     // snapshot slot stride in bytes
    Rot<double> q0 = z.q0, q1 = z.q1, q2 = z.q2, b0 = z.b0, b1 = z.b1, b2 = z.b2, c0 = z.c0, c1 = z.c1, c2 = z.c2, U = z.U, V = z.V, W = z.W;
    const bool feeds_g = warp == 0 && lane32 == 0, feeds_h = warp > 0 && lane32 == 0;
    const bool ret_g = warp == NW - 1 && lane32 == 31, ret_h = warp < NW - 1 && lane32 == 31;
    gq += start; gb += start; gc += start;
    const int kofs = feeds_g ? AMRVW_RING - 1 : -(3 * (LE - 1) + 2);
    gq += t0 + kofs; gb += t0 + kofs; gc += t0 + kofs;
    int kidx = start + t0 + kofs;
    const unsigned hin = hand_sa + (unsigned)((warp > 0 ? warp - 1 : 0) * 2) * 48u, hout = hand_sa + (unsigned)(warp * 2) * 48u;
#pragma unroll 1
    for (int t = t0; t < t1; ++t) {
        Rot<double> iq = shfl_up_rot(q0, 32), ib = shfl_up_rot(b0, 32), ic = shfl_up_rot(c0, 32);
        if (feeds_g) {
            cp_async_wait_ring();
            const unsigned sg = stage_sa + (t & (AMRVW_RING - 1)) * 48, sn2 = stage_sa + ((t + AMRVW_RING - 1) & (AMRVW_RING - 1)) * 48;
            iq = lds_rot(sg); ib = lds_rot(sg + 16); ic = lds_rot(sg + 32);
            cp_async16_sa(sn2, gq); cp_async16_sa(sn2 + 16, gb); cp_async16_sa(sn2 + 32, gc);
            cp_async_commit();
        } else if (feeds_h) {
            const unsigned sg = hin + (unsigned)((t + 1) & 1) * 48u;
            iq = lds_rot(sg); ib = lds_rot(sg + 16); ic = lds_rot(sg + 32);
        }
        q0 = q1; q1 = q2; q2 = iq;
        b0 = b1; b1 = b2; b2 = ib;
        c0 = c1; c1 = c2; c2 = ic;
#ifndef AMRVW_STRICT
        sts_rot(snap_sa + 0 * SS, q0); sts_rot(snap_sa + 1 * SS, q1); sts_rot(snap_sa + 2 * SS, q2);
        sts_rot(snap_sa + 3 * SS, b0); sts_rot(snap_sa + 4 * SS, b1); sts_rot(snap_sa + 5 * SS, b2);
        sts_rot(snap_sa + 6 * SS, c0); sts_rot(snap_sa + 7 * SS, c1); sts_rot(snap_sa + 8 * SS, c2);
        sts_rot(snap_sa + 9 * SS, U); sts_rot(snap_sa + 10 * SS, V); sts_rot(snap_sa + 11 * SS, W);
        bool ok = true;
        {
            Rot<double> t1_, t2_, t3_;
            ok &= turnover_fast_try(b1, b2, V, t1_, t2_, t3_); V = t1_; b1 = t2_; b2 = t3_;
            ok &= turnover_fast_try(c2, c1, V, t1_, t2_, t3_); V = t1_; c2 = t2_; c1 = t3_;
            ok &= turnover_fast_try(b0, b1, U, t1_, t2_, t3_); U = t1_; b0 = t2_; b1 = t3_;
            ok &= turnover_fast_try(c1, c0, U, t1_, t2_, t3_); U = t1_; c1 = t2_; c0 = t3_;
            ok &= turnover_fast_try(q1, q2, V, t1_, t2_, t3_); V = t1_; q1 = t2_; q2 = t3_;
            ok &= turnover_fast_try(q0, q1, U, t1_, t2_, t3_); U = t1_; q0 = t2_; q1 = t3_;
            ok &= turnover_fast_try(W, V, U, t1_, t2_, t3_); V = t1_; U = t2_; W = t3_;
        }
        if (!ok) {   // a degenerate turnover (essentially never): redo the step on the general path
            StepState y;
            y.q0 = lds_rot(snap_sa + 0 * SS); y.q1 = lds_rot(snap_sa + 1 * SS); y.q2 = lds_rot(snap_sa + 2 * SS);
            y.b0 = lds_rot(snap_sa + 3 * SS); y.b1 = lds_rot(snap_sa + 4 * SS); y.b2 = lds_rot(snap_sa + 5 * SS);
            y.c0 = lds_rot(snap_sa + 6 * SS); y.c1 = lds_rot(snap_sa + 7 * SS); y.c2 = lds_rot(snap_sa + 8 * SS);
            y.U = lds_rot(snap_sa + 9 * SS); y.V = lds_rot(snap_sa + 10 * SS); y.W = lds_rot(snap_sa + 11 * SS);
            chase_step_general(y);
            q0 = y.q0; q1 = y.q1; q2 = y.q2; b0 = y.b0; b1 = y.b1; b2 = y.b2; c0 = y.c0; c1 = y.c1; c2 = y.c2; U = y.U; V = y.V; W = y.W;
        }
#else
        {
            StepState y = {q0, q1, q2, b0, b1, b2, c0, c1, c2, U, V, W};
            chase_step_general(y);
            q0 = y.q0; q1 = y.q1; q2 = y.q2; b0 = y.b0; b1 = y.b1; b2 = y.b2; c0 = y.c0; c1 = y.c1; c2 = y.c2; U = y.U; V = y.V; W = y.W;
        }
#endif
        if (ret_g) {
            stg_rot(gq, q0); stg_rot(gb, b0); stg_rot(gc, c0);
            if (fabs(q0.s) <= AMRVW_EPS) dstack[sp++] = kidx;
        } else if (ret_h) {
            const unsigned so = hout + (unsigned)(t & 1) * 48u;
            sts_rot(so, q0); sts_rot(so + 16, b0); sts_rot(so + 32, c0);
        }
        gq += 1; gb += 1; gc += 1; kidx += 1;
        if constexpr (BAR == 0) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(32 * NW) : "memory");
    }
    z.q0 = q0; z.q1 = q1; z.q2 = q2; z.b0 = b0; z.b1 = b1; z.b2 = b2; z.c0 = c0; z.c1 = c1; z.c2 = c2; z.U = U; z.V = V; z.W = W;
    return sp;
}

template <int NW, int MH>
__global__ void __launch_bounds__(32 * NW) chain_kernel(Problem<double> pr, PipeParams pp) {
    constexpr int LE = 32 * NW;
    [[maybe_unused]] constexpr int NT = LE;
    extern __shared__ __align__(16) unsigned char raw[];   // chain_raw_bytes<NW, MH>(): shift scratch (phase A) and step snapshots (chase) are never live together
    __shared__ __align__(16) ChainSmem<NW> sm;
    ShiftScratch<MH>& shift_scratch = *reinterpret_cast<ShiftScratch<MH>*>(raw);
    double2* const snap_t = reinterpret_cast<double2*>(raw) + threadIdx.x;                 // [12][NT] step snapshots
    const int gl = threadIdx.x, warp = gl >> 5, lane32 = gl & 31;
    const unsigned snap_sa = (unsigned)__cvta_generic_to_shared(snap_t);
    const unsigned stage_sa = (unsigned)__cvta_generic_to_shared(sm.ring);
    const unsigned hand_sa = (unsigned)__cvta_generic_to_shared(&sm.hand[0][0][0]);

    long long pc_all = 0, pc_a = 0, pc_lean = 0;      // warp-cycle accounting of warp 0 (amrvw_last_profile)
    long long pn_lean = 0, pn_all = 0;
    for (long long slot = blockIdx.x; slot < pr.nbatch; slot += gridDim.x) {
        const long long p = unit_poly(pr, slot);      // largest polynomials first
        if (!pr.rec[p].iterating) continue;           // written by the set-up kernel before this launch
        for (;;) {
            // ---------------- phase A: driver on thread 0, shifts on warp 0
            const long long pc_f0 = clock64();
            if (gl == 0) { sm.mode = unit_driver(pr, pp, LE, p); sm.nturn = 0ull; }
            if (gl < NW * 2 * 3) (&sm.hand[0][0][0])[gl] = make_double2(1.0, 0.0);
            __syncthreads();
            const int m0 = sm.mode;
            if (m0 == 0) break;
            if (m0 == 3 && warp == 0) unit_shifts<MH>(pr, pp, 32, p, shift_scratch, lane32, 0xffffffffu, LE);
            __syncthreads();
            pc_a += clock64() - pc_f0;

            // ---------------- the fat step: every thread reads the same record
            const PolyRec* r = pr.rec + p;
            const int mode = __ldcg(&r->mode), nb = __ldcg(&r->nb), start = __ldcg(&r->ls.ctr.start_index), stop = __ldcg(&r->ls.ctr.stop_index);
            const int ord0 = __ldcg(&r->ls.ord), N = __ldcg(&r->N);
            int sp = __ldcg(&r->ls.sp);
            int nturn = 0;
            const int T = stop - start + 3 * LE + 1;
            Fact<double, false> F = make_fact<double, false>(pr, p, N);
            int* const dstack = pr.dstack + (pr.offsets[p] + (long long)pr.extra * p);
            Rot<double>* const FQ = F.Q;
            Rot<double>* const FB = F.B;
            Rot<double>* const FC = F.Ct;
            Rot<double> q0, q1, q2, b0, b1, b2, c0, c1, c2, U, V, W;
            q0 = q1 = q2 = b0 = b1 = b2 = c0 = c1 = c2 = U = V = W = make_rot<double>(1.0, 0.0);
            const bool has_bulge = gl < nb;
            if (gl == 0) {
                for (int a = 0; a < AMRVW_RING - 1; ++a) {
                    const int nidx = min(start + a, stop + 1);
                    const unsigned sg = stage_sa + (a & (AMRVW_RING - 1)) * 48;
                    cp_async16_sa(sg, FQ + nidx); cp_async16_sa(sg + 16, FB + nidx); cp_async16_sa(sg + 32, FC + nidx);
                    cp_async_commit();
                }
            }
            for (int t = 0; t < T; ++t) {
                if (nb == LE && t >= 3 * LE && t <= stop - start) {   // every lane at a regular position: the lean loop
                    const int te = stop - start + 1;
                    const long long pc_l0 = clock64();
                    StepState z = {q0, q1, q2, b0, b1, b2, c0, c1, c2, U, V, W};
                    sp = chain_lean<NW>(z, FQ, FB, FC, stage_sa, snap_sa, hand_sa, dstack, sp, start, t, te, warp, lane32);
                    q0 = z.q0; q1 = z.q1; q2 = z.q2; b0 = z.b0; b1 = z.b1; b2 = z.b2; c0 = z.c0; c1 = z.c1; c2 = z.c2; U = z.U; V = z.V; W = z.W;
                    nturn += 7 * (te - t);
                    pc_lean += clock64() - pc_l0; pn_lean += te - t;
                    t = te - 1;
                    continue;
                }
                const int tp = t - 3 * gl;
                // ingest: what the previous lane finished last step (thread 0: from memory; lane 0 of a later warp: from the hand-off slot)
                Rot<double> iq = shfl_up_rot(q0, 32), ib = shfl_up_rot(b0, 32), ic = shfl_up_rot(c0, 32);
                if (gl == 0) {
                    cp_async_wait_ring();
                    const unsigned sg = stage_sa + (t & (AMRVW_RING - 1)) * 48;
                    iq = lds_rot(sg); ib = lds_rot(sg + 16); ic = lds_rot(sg + 32);
                    const int nidx = min(start + t + AMRVW_RING - 1, stop + 1);
                    const unsigned sn2 = stage_sa + ((t + AMRVW_RING - 1) & (AMRVW_RING - 1)) * 48;
                    cp_async16_sa(sn2, FQ + nidx); cp_async16_sa(sn2 + 16, FB + nidx); cp_async16_sa(sn2 + 32, FC + nidx);
                    cp_async_commit();
                } else if (lane32 == 0) {
                    const unsigned sg = hand_sa + (unsigned)((warp - 1) * 2 + ((t + 1) & 1)) * 48u;
                    iq = lds_rot(sg); ib = lds_rot(sg + 16); ic = lds_rot(sg + 32);
                }
                q0 = q1; q1 = q2; q2 = iq;
                b0 = b1; b1 = b2; b2 = ib;
                c0 = c1; c1 = c2; c2 = ic;
                const int i = start + tp - 2;   // position of this lane's bulge; window = indices i..i+2
                if (has_bulge && tp >= 2 && i <= stop - 1) {
                    if (tp == 2) {
                        StepState z = {q0, q1, q2, b0, b1, b2, c0, c1, c2, U, V, W};
                        create_and_absorb_at(z, F, start, stop, mode, pr.shifts + p * LE, gl, pr.seed, (uint64_t)p, ord0 + gl);
                        q0 = z.q0; q1 = z.q1; U = z.U; V = z.V; W = z.W;
                        nturn += 1;
                    }
                    if (i <= stop - 2) {
#ifndef AMRVW_STRICT
                        AMRVW_CHASE_REGULAR(snap_t, NT);
#else
                        AMRVW_CHASE_GENERAL();
#endif
                    } else {
                        StepState z = {q0, q1, q2, b0, b1, b2, c0, c1, c2, U, V, W};
                        chase_step_final_at(z, F, stop);
                        q0 = z.q0; q1 = z.q1; q2 = z.q2; b0 = z.b0; b1 = z.b1; b2 = z.b2; c0 = z.c0; c1 = z.c1; c2 = z.c2; U = z.U; V = z.V; W = z.W;
                    }
                    nturn += 7;
                }
                // retire: the last lane of the chain writes index i back, the last lane of a warp hands it on
                if (gl == LE - 1) {
                    if (i >= start && i <= stop + 1) {
                        stg_rot(FQ + i, q0); stg_rot(FB + i, b0); stg_rot(FC + i, c0);
                        if (i <= stop && fabs(q0.s) <= AMRVW_EPS) dstack[sp++] = i;   // deflatable: record for phase A
                    }
                } else if (lane32 == 31) {
                    const unsigned so = hand_sa + (unsigned)(warp * 2 + (t & 1)) * 48u;
                    sts_rot(so, q0); sts_rot(so + 16, b0); sts_rot(so + 32, c0);
                }
                __syncthreads();
            }
            pn_all += T;
            if (gl == 0) cp_async_wait_all();   // drain look-ahead requests past the end of the window
            for (int d = 16; d >= 1; d >>= 1) nturn += __shfl_down_sync(0xffffffffu, nturn, d);
            if (lane32 == 0) atomicAdd(&sm.nturn, (unsigned long long)nturn);
            __syncthreads();
            if (gl == LE - 1) {   // back to the record (this thread owns the deflation stack pointer); phase A is due
                PolyRec* rw = pr.rec + p;
                PolyRec rec = load_rec(rw);
                rec.ls.sp = sp;
                rec.st.sweeps += nb;
                rec.chase_turnovers += (long long)sm.nturn;
                if (mode == 2) { rec.st.exceptional += nb; rec.ls.ord += nb; rec.ls.ctr.it_count = AMRVW_IT_COUNT - 1; }
                store_rec(rw, rec);
            }
            __syncthreads();
            pc_all += clock64() - pc_f0;
        }
        __syncthreads();
    }
    if (pr.prof && gl == 0) {
        atomicAdd(pr.prof + 0, (unsigned long long)pc_all);
        atomicAdd(pr.prof + 1, (unsigned long long)pc_a);
        atomicAdd(pr.prof + 2, (unsigned long long)pc_lean);
        atomicAdd(pr.prof + 3, (unsigned long long)(pc_all - pc_lean - pc_a));
        atomicAdd(pr.prof + 4, (unsigned long long)pn_lean);
        atomicAdd(pr.prof + 5, (unsigned long long)(pn_all - pn_lean));
    }
}

template <int NW, int MH> constexpr size_t chain_raw_bytes() {
    return sizeof(ShiftScratch<MH>) > 12 * 32 * NW * sizeof(double2) ? sizeof(ShiftScratch<MH>) : 12 * 32 * NW * sizeof(double2);
}
template <int NW, int MH>
static cudaError_t launch_chain(const Problem<double>& pr, const PipeParams& pp, int sm_count, cudaStream_t st) {
    int per_sm = 0;
    cudaError_t e;
    constexpr size_t smem = chain_raw_bytes<NW, MH>();
    if ((e = cudaFuncSetAttribute(chain_kernel<NW, MH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chain_kernel<NW, MH>, 32 * NW, smem)) != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const long long cap = (long long)sm_count * per_sm;
    const int blocks = (int)(pr.nbatch < cap ? pr.nbatch : cap);
    chain_kernel<NW, MH><<<blocks, 32 * NW, smem, st>>>(pr, pp);
    return cudaGetLastError();
}

}  // namespace amrvw
