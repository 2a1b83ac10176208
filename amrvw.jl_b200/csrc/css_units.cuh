// css_units.cuh -- complex single shift (CSS.jl) on the unit queue: pipelined multishift chase.
//
// Same architecture as the RDS path (rounds.cuh): work units (one warp = 32/L polynomials) circulate
// through a device-side FIFO; a visit runs phase A when the unit's fat step is through (driver on
// lane 0 of each group, then the L shifts of the next fat step) and then up to SEG lock-steps of the
// systolic chase of L bulges, one per lane.
//
// What is different for the complex single shift (reference: CSS.jl, bulge-step.jl:22-33):
//   * a bulge is ONE rotator U at position i and touches indices i, i+1 of Q, B, Ct and of the two
//     phase diagonals D_Q, D_R, so bulges TWO indices apart commute; a lane holds a 2-index window
//     (13 doubles per index) and per step does the two diagonal swaps (diagonal.jl:84-96) and the three
//     turnovers of passthrough_triu / passthrough_Q (CSS.jl:17-41, 64-88);
//   * absorb_Ut (CSS.jl:8-14) fuses the new rotator into Q[start] and pushes the resulting phase down
//     the WHOLE Q chain (diagonal.jl:199-225: every Q[k].c, k > start, is multiplied by conj(alpha) until
//     an exact zero sine, where the phase lands in D_Q) -- an O(window) sweep per bulge that would
//     serialise the pipeline.  Here the phase travels with its bulge: lane m multiplies every Q rotator
//     it ingests by its own conj(alpha_m), exactly when the sequential order would have (after bulge
//     m-1 is through with that index, before bulge m's turnover on it), and deposits it in D_Q at the
//     first zero sine / the end of the window.  The push-down costs one complex multiplication per
//     lock-step instead of a pass over memory;
//   * shifts: the L / reuse eigenvalues of the trailing block, each driving `reuse` bulges, from the dense
//     block (entry recurrences over the reference's formulas) and a single-shift complex Hessenberg QR shared
//     by the lanes of the group; the core-chasing iteration on a shared-memory copy (lane 0) is the fall-back
//     and AMRVW_SHIFTS_CHASE.  The chase uses the division-free complex turnover inline, with the lane's
//     state in registers.
//
// Roots agree with the reference schedule to backward-error level (tests: minimum-distance matching
// against the oracle).  The oracle holds a sequential twin of this schedule (amrvw_algorithm_ms_c: same fat
// steps, dense block + complex Hessenberg QR, the reference's phase push-down); hypot, complex division and
// complex sqrt come from different math libraries on the two sides, so the no-FMA build is held to it at a
// tight tolerance and to the same fat-step and turnover counts within 2 %, not bit for bit.
#pragma once
#include "rounds.cuh"

namespace amrvw {

// one index of the five arrays
struct CssIdx {
    Rot<cplx> q, b, c;
    cplx dq, dr;
};
struct CssLane {            // what a lane owns between lock-steps
    CssIdx w0, w1;          // indices i, i+1 (i = position of the lane's bulge)
    Rot<cplx> U;            // the bulge
    cplx ca;                // conj(alpha) of the bulge's creation, still travelling down Q
    int alive;              // 1 while ca has not been deposited in D_Q
    int pad[3];
};
static_assert(sizeof(CssLane) % 16 == 0, "CssLane is parked in 16-byte words");

AMRVW_DI Rot<cplx> ldcg_crot(const Rot<cplx>* p) {
    const double2 a = __ldcg(reinterpret_cast<const double2*>(p)), b = __ldcg(reinterpret_cast<const double2*>(p) + 1);
    Rot<cplx> r;
    r.c = cplx(a.x, a.y); r.s = b.x; r.pad = 0.0;
    return r;
}
AMRVW_DI cplx ldcg_cplx(const cplx* p) {
    const double2 a = __ldcg(reinterpret_cast<const double2*>(p));
    return cplx(a.x, a.y);
}
AMRVW_DI void css_store_index(const Fact<cplx, false>& F, const int k, const CssIdx& x) {
    F.Q[k] = x.q; F.B[k] = x.b; F.Ct[k] = x.c; F.DQ[k] = x.dq; F.DR[k] = x.dr;
}
AMRVW_DI CssIdx css_shfl_up(const CssIdx& x, const int width) {
    CssIdx o;
    const double* src = reinterpret_cast<const double*>(&x);
    double* dst = reinterpret_cast<double*>(&o);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(CssIdx) / 8); ++i) dst[i] = __shfl_up_sync(0xffffffffu, src[i], 1, width);
    return o;
}
// look-ahead ring of a group's feeding lane (the same scheme as the real path, pipeline.cuh): the five
// arrays of index k go global -> shared with eight 16-byte cp.async.cg, AMRVW_RING - 1 steps ahead of their
// use.  (With a plain load one step ahead ptxas kept the load next to its first use: one DMUL carried 18%
// of all stall samples, long scoreboard: profiles/.)
#define AMRVW_CSS_SLOT 128   // bytes per ring slot: Q, B, Ct (32 each), D_Q, D_R (16 each)
AMRVW_DI void css_prefetch_index(const unsigned slot_sa, const Fact<cplx, false>& F, const int k) {
    const char* q = reinterpret_cast<const char*>(F.Q + k);
    const char* b = reinterpret_cast<const char*>(F.B + k);
    const char* c = reinterpret_cast<const char*>(F.Ct + k);
    cp_async16_sa(slot_sa, q); cp_async16_sa(slot_sa + 16, q + 16);
    cp_async16_sa(slot_sa + 32, b); cp_async16_sa(slot_sa + 48, b + 16);
    cp_async16_sa(slot_sa + 64, c); cp_async16_sa(slot_sa + 80, c + 16);
    cp_async16_sa(slot_sa + 96, F.DQ + k); cp_async16_sa(slot_sa + 112, F.DR + k);
}
AMRVW_DI double2 lds_d2(const unsigned sa) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(sa) : "memory");
    return v;
}
AMRVW_DI CssIdx css_read_slot(const unsigned slot_sa) {
    CssIdx x;
    double2 a = lds_d2(slot_sa), b = lds_d2(slot_sa + 16);
    x.q.c = cplx(a.x, a.y); x.q.s = b.x; x.q.pad = 0.0;
    a = lds_d2(slot_sa + 32); b = lds_d2(slot_sa + 48);
    x.b.c = cplx(a.x, a.y); x.b.s = b.x; x.b.pad = 0.0;
    a = lds_d2(slot_sa + 64); b = lds_d2(slot_sa + 80);
    x.c.c = cplx(a.x, a.y); x.c.s = b.x; x.c.pad = 0.0;
    a = lds_d2(slot_sa + 96); b = lds_d2(slot_sa + 112);
    x.dq = cplx(a.x, a.y); x.dr = cplx(b.x, b.y);
    return x;
}
AMRVW_DI CssIdx css_identity_index() {
    CssIdx x;
    x.q = make_rot<cplx>(cplx(1.0, 0.0), 0.0); x.b = x.q; x.c = x.q; x.q.pad = x.b.pad = x.c.pad = 0.0;
    x.dq = cplx(1.0, 0.0); x.dr = cplx(1.0, 0.0);
    return x;
}

// diagonal swap (diagonal.jl:84-96): the bulge passes the two phases at i, i+1
AMRVW_DI void css_pass_diag(cplx& d0, cplx& d1, Rot<cplx>& U) {
    const cplx alpha = d0, beta = d1;
    d0 = beta; d1 = alpha;
    U.c = (alpha * conj_(beta)) * U.c;
}
// passthrough_triu (CSS.jl:17-41): U at i right -> left through R = Ct (B D_R)
AMRVW_DI void css_pass_R(CssIdx& w0, CssIdx& w1, Rot<cplx>& U) {
    css_pass_diag(w0.dr, w1.dr, U);
    Rot<cplx> r1, r2, r3;
    turnover(w0.b, w1.b, U, r1, r2, r3);      // pass_desc on B: U moves to i+1
    U = r1; w0.b = r2; w1.b = r3;
    turnover(w1.c, w0.c, U, r1, r2, r3);      // pass_asc on Ct: U back at i
    U = r1; w1.c = r2; w0.c = r3;
}
// the same two passes for the systolic lanes: inline division-free turnovers, state in registers
AMRVW_DI void css_lane_regular(CssIdx& w0, CssIdx& w1, Rot<cplx>& U) {
    css_pass_diag(w0.dr, w1.dr, U);
    Rot<cplx> r1, r2, r3;
    turnover_fast(w0.b, w1.b, U, r1, r2, r3);
    U = r1; w0.b = r2; w1.b = r3;
    turnover_fast(w1.c, w0.c, U, r1, r2, r3);
    U = r1; w1.c = r2; w0.c = r3;
    css_pass_diag(w0.dq, w1.dq, U);
    turnover_fast(w0.q, w1.q, U, r1, r2, r3);
    U = r1; w0.q = r2; w1.q = r3;
}
// regular position (CSS.jl:64-75): through R, then through D_Q and Q; the bulge leaves at i+1
AMRVW_NI void css_step_regular(CssLane& z) {
    css_pass_R(z.w0, z.w1, z.U);
    css_pass_diag(z.w0.dq, z.w1.dq, z.U);
    Rot<cplx> r1, r2, r3;
    turnover(z.w0.q, z.w1.q, z.U, r1, r2, r3);
    z.U = r1; z.w0.q = r2; z.w1.q = r3;
}
// last position i == stop (CSS.jl:76-88): through R, through D_Q, fuse into Q[stop]
AMRVW_NI void css_step_final(CssLane& z) {
    css_pass_R(z.w0, z.w1, z.U);
    css_pass_diag(z.w0.dq, z.w1.dq, z.U);
    cplx dph;
    z.w0.q = fuse(z.w0.q, z.U, dph);
    z.w0.dq = z.w0.dq * dph;
    z.w1.dq = z.w1.dq * conj_(dph);
}
// create_bulge (create-bulge.jl:80-118) + absorb_Ut (CSS.jl:8-14) for the lane whose window just became
// (start, start+1); wm = index start-1 (static during the fat step, read from memory)
AMRVW_NI void css_create(CssLane& z, const Fact<cplx, false>& F, const int start, const int stop, const cplx shift) {
    CssIdx wm = css_identity_index();
    if (start >= 1) {
        wm.q = ldcg_crot(F.Q + start - 1); wm.dq = ldcg_cplx(F.DQ + start - 1);       // Q[0], D_Q[0] are sentinels
        if (start > 1) { wm.b = ldcg_crot(F.B + start - 1); wm.c = ldcg_crot(F.Ct + start - 1); wm.dr = ldcg_cplx(F.DR + start - 1); }
    }
    cplx c; double sn;
    cplx dph;
#ifndef AMRVW_STRICT
    // Only A11 and A21 of the block at `start` enter the bulge (create-bulge.jl:105-108): written out from
    // the entry formulas (chains.jl:109-158, rfactorization.jl:139-170) with reciprocals instead of the IEEE
    // complex divisions of diagonal_block.  The bulge only chooses the similarity -- it is applied exactly,
    // by turnovers -- so its last bits steer convergence, not accuracy (same argument as create_bulge_fast).
    bool fast = false;
    {
        const Rot<cplx> qm = wm.q, q0 = z.w0.q;
        const cplx qji = (-qm.s) * wm.dq, qjj = conj_(qm.c) * q0.c * z.w0.dq, qkj = (-q0.s) * z.w0.dq;
        const double ig0 = rcp_(z.w0.c.s);
        const cplx wjj = (-z.w0.b.s) * z.w0.dr;                   // w(j,j) D_R[j]
        const cplx rjj = wjj * ig0;
        cplx rij = cplx(0.0, 0.0);
        if (start > 1) {
            const double igm = rcp_(wm.c.s);
            const cplx wij = conj_(wm.b.c) * z.w0.b.c;            // w(i,j), i = j - 1
            rij = (wij * z.w0.dr) * igm - (wm.c.c * conj_(z.w0.c.c)) * (igm * ig0) * wjj;
        }
        const cplx a = (qji * rij + qjj * rjj) - shift, b = qkj * rjj;
        // givensrot(a, b) (transformations.jl:28-31): s = |b| / h, c = b conj(a) / (|b| h), h^2 = |a|^2 + |b|^2
        const double f2 = abs2_(b), g2 = abs2_(a), h2 = f2 + g2;
        if (f2 >= 1e-280 && h2 <= 1e280 && f2 > fmax(g2, 1.0) * AMRVW_SAFMIN) {
            const double ih = rsqrt_nr(h2), ib = rsqrt_nr(f2);
            sn = (f2 * ib) * ih;
            c = (b * conj_(a)) * (ib * ih);
            // fuse(R, Q[start]) (transformations.jl:101-118) with 1/|v| by rsqrt
            const Rot<cplx> qq = z.w0.q;
            const cplx u = c * qq.c - sn * qq.s;
            const cplx v = c * qq.s + sn * conj_(qq.c);
            const double v2 = abs2_(v);
            if (v2 >= 1e-280 && v2 <= 1e280) {
                const double iv = rsqrt_nr(v2);
                const cplx alpha = v * iv;
                cplx cc = u * alpha;
                double ss = v2 * iv;
                polish(cc, ss);
                dph = conj_(alpha);
                z.U = make_rot<cplx>(conj_(c), -sn);
                z.w0.q = make_rot<cplx>(cc, ss);
                fast = true;
            }
        }
    }
    if (!fast)
#endif
    {
    Rot<cplx> lq[3] = {wm.q, z.w0.q, z.w1.q}, lb[3] = {wm.b, z.w0.b, z.w1.b}, lc[3] = {wm.c, z.w0.c, z.w1.c};
    cplx ldq[3] = {wm.dq, z.w0.dq, z.w1.dq}, ldr[3] = {wm.dr, z.w0.dr, z.w1.dr};
    Fact<cplx, false> Fl = F;
    Fl.Q = lq - (start - 1); Fl.B = lb - (start - 1); Fl.Ct = lc - (start - 1); Fl.DQ = ldq - (start - 1); Fl.DR = ldr - (start - 1);
    cplx A[4];
    diagonal_block(Fl, start, A);
    cplx nrm;
    givensrot(A[0] - shift, A[2], c, sn, nrm);
    const Rot<cplx> R = make_rot<cplx>(c, sn);
    z.U = adjoint(R);
    z.w0.q = fuse(R, z.w0.q, dph);
    }
    // push_phase (diagonal.jl:199-225), first two actions; the rest travels with the bulge
    z.w0.dq = z.w0.dq * dph;
    z.ca = conj_(dph);
    z.alive = 1;
    if (start + 1 > stop || z.w1.q.s == 0.0) { z.w1.dq = z.w1.dq * z.ca; z.alive = 0; }
    else z.w1.q.c = z.ca * z.w1.q.c;
}

// ---------------------------------------------------------------- phase A of a complex unit
AMRVW_DI void css_diagonal_block_cg(const Fact<cplx, false>& F, const int j, cplx A[4]) {
    Rot<cplx> lq[3], lb[3], lc[3];
    cplx ldq[3], ldr[3];
    for (int t = 0; t < 3; ++t) {
        const int k = j - 1 + t;
        lq[t] = ldcg_crot(F.Q + k); ldq[t] = ldcg_cplx(F.DQ + k);
        if (k >= 1) { lb[t] = ldcg_crot(F.B + k); lc[t] = ldcg_crot(F.Ct + k); ldr[t] = ldcg_cplx(F.DR + k); }
        else { lb[t] = make_rot<cplx>(cplx(1.0, 0.0), 0.0); lc[t] = lb[t]; ldr[t] = cplx(1.0, 0.0); }
    }
    Fact<cplx, false> Fl = F;
    Fl.Q = lq - (j - 1); Fl.B = lb - (j - 1); Fl.Ct = lc - (j - 1); Fl.DQ = ldq - (j - 1); Fl.DR = ldr - (j - 1);
    diagonal_block(Fl, j, A);
}
// deflate (CSS.jl:98-122) with coherent reads: Q[k] becomes the identity, its phase goes down Q
AMRVW_DI void css_deflate_cg(Fact<cplx, false>& F, const int k) {
    const cplx alpha = ldcg_crot(F.Q + k).c;
    F.Q[k] = make_rot<cplx>(cplx(1.0, 0.0), 0.0);
    const int M = F.N - 1;
    F.DQ[k] = ldcg_cplx(F.DQ + k) * alpha;
    const cplx ca = conj_(alpha);
    int i = k;
    while (i < M) {
        Rot<cplx> q = ldcg_crot(F.Q + i + 1);
        if (q.s == 0.0) break;
        q.c = ca * q.c;
        F.Q[i + 1] = q;
        ++i;
    }
    F.DQ[i + 1] = ldcg_cplx(F.DQ + i + 1) * ca;
}
AMRVW_DI void css_check_deflation_stack(Fact<cplx, false>& F, LeaderState& ls, const int* dstack) {
    if (ls.sp > 0) {
        const int k = __ldcg(dstack + ls.sp - 1);
        if (k >= ls.ctr.start_index && k <= ls.ctr.stop_index) {
            css_deflate_cg(F, k);
            ls.ctr.zero_index = k;
            ls.ctr.start_index = k + 1;
            ls.ctr.it_count = AMRVW_IT_COUNT;
        }
    }
}
// the driver (AMRVW_algorithm.jl:10-138) up to the next fat step: 0 finished, 2 exceptional fat step,
// 3 fat step that needs nb + 1 shifts (nb = rotators of the trailing block)
AMRVW_NI int css_unit_driver(const Problem<cplx>& pr, const PipeParams pp, const int L, const long long p) {
    PolyRec rec = load_rec(pr.rec + p);
    if (!rec.iterating) return 0;
    Fact<cplx, false> F = make_fact<cplx, false>(pr, p, rec.N);
    cplx* E = pr.roots + pr.offsets[p] - 1;
    const int* dstack = pr.dstack + (pr.offsets[p] + (long long)pr.extra * p);
    SmallSink sink;
    sink.list = pr.swlist; sink.count = pr.counters + CNT_SMALL; sink.poly = (int)p;
    LeaderState& ls = rec.ls;
    Counter& ctr = ls.ctr;
    cplx A[4];
    int mode = 0, nb = 0;
    while (ls.kk <= ls.it_max) {
        if (ctr.stop_index < 1) break;
        css_check_deflation_stack(F, ls, dstack);
        ++ls.kk;
        const int k = ctr.stop_index;
        const int delta = ctr.stop_index - ctr.zero_index;
        if (delta == 1) {
            css_diagonal_block_cg(F, k, A);
            cplx e1, e2;
            eig2x2(A, e1, e2);
            E[k] = e1; E[k + 1] = e2;
            if (ctr.stop_index == 2) { css_diagonal_block_cg(F, 1, A); E[1] = A[0]; ctr.stop_index = 0; break; }
            pop_window(ls);
            ctr.stop_index = ctr.zero_index - 1; ctr.zero_index = 0; ctr.start_index = 1;
            css_check_deflation_stack(F, ls, dstack);
            if (ctr.stop_index < 1) break;
        } else if (delta <= 0) {
            css_diagonal_block_cg(F, k, A);
            if (ctr.stop_index == 1) { E[1] = A[0]; E[2] = A[3]; ctr.stop_index = 0; break; }
            E[k + 1] = A[3];
            pop_window(ls);
            ctr.stop_index = ctr.zero_index - 1;
            if (ctr.stop_index < 1) break;
            ctr.zero_index = 0; ctr.start_index = 1;
            css_check_deflation_stack(F, ls, dstack);
        } else if (delta <= pp.small) {
            const int lo2 = ctr.start_index, hi2 = ctr.stop_index;
            sink.push(lo2, hi2, ctr.it_count);
            if (lo2 == 2) { css_diagonal_block_cg(F, 1, A); E[1] = A[0]; }
            pop_window(ls);
            ctr.stop_index = lo2 - 2; ctr.zero_index = 0; ctr.start_index = 1; ctr.it_count = AMRVW_IT_COUNT;
            if (ctr.stop_index >= 1) css_check_deflation_stack(F, ls, dstack);
            ls.pn_small += 1;
        } else if (ctr.it_count == 0) {
            rec.st.fat_steps++;
            nb = L; mode = 2;
            break;
        } else {
            nb = min(L / pp.reuse - 1, delta - 1);       // rotators of the trailing block: nb + 1 shifts, each driving `reuse` bulges
            rec.st.fat_steps++;
            ctr.it_count -= 1;
            mode = 3;
            break;
        }
    }
    rec.mode = mode; rec.nb = nb;
    if (mode == 0) rec.iterating = 0;
    store_rec(pr.rec + p, rec);
    return mode;
}

// shared-memory scratch of one lane group for the shifts
template <int MH> struct __align__(16) CssShiftScratch {
    Rot<cplx> q[MH + 3], b[MH + 3], c[MH + 3];
    cplx dq[MH + 3], dr[MH + 3], e[MH + 3];
    cplx H[MH * (MH + 1)];          // dense trailing block, row major, leading dimension MH + 1
    cplx gs[2 * MH];                // the Givens rotations of one QR sweep
    double isg[MH + 2];
};

// ---------------------------------------------------------------- shifts from the dense trailing block (complex)
// The complex twin of shifts.cuh.  Entries of the trailing M x M block of (Q D_Q) Ct (B D_R), rotator Q[k0]
// diagonal, column by column (one column per lane), bottom-up, O(1) per entry:
//   w(j,k) = -s^B_k (j = k), conj(c^B_j) (prod_{j<l<k} s^B_l) c^B_k          (rfactorization.jl:83-97)
//   R[j,k] = D_R[k] (w(j,k) + c^Ct_j T_j) / s^Ct_j,  T_k = 0,  T_{j-1} = (-conj(c^Ct_j) w(j,k) - T_j) / s^Ct_j   (:139-170)
//   rows of D_Q R pushed up through the Q rotators [c s; -s conj(c)]          (chains.jl:109-158, qfactorization.jl:67-70)
// q/b/c/dq/dr: staged with GLOBAL indices (slot t <-> index k0 + t), q[k0] = (phase, 0).
AMRVW_DI void css_dense_trailing_block_warp(const Rot<cplx>* q, const Rot<cplx>* b, const Rot<cplx>* c, const cplx* dq, const cplx* dr, const int k0, const int stop,
                                            cplx* H, const int LD, double* isg, const int M, const int lane, const unsigned gmask, const int GL) {
    const int j0 = k0 + 1;
    for (int i = lane; i < M * LD; i += GL) H[i] = cplx(0.0, 0.0);
    for (int j = j0 + lane; j <= stop + 1; j += GL) isg[j - j0] = rcp_(c[j].s);
    __syncwarp(gmask);
    for (int k = j0 + lane; k <= stop + 1; k += GL) {
        double P = 1.0;
        cplx T = cplx(0.0, 0.0), Vacc = cplx(0.0, 0.0);
        const cplx ck = b[k].c, dk = dr[k];
        const double sk = b[k].s;
        for (int j = k; j >= j0; --j) {
            const cplx w = (j == k) ? cplx(-sk, 0.0) : (conj_(b[j].c) * P) * ck;
            const cplx ctc = c[j].c;
            const double ig = isg[j - j0];
            const cplx X = dq[j] * (((w + ctc * T) * ig) * dk);
            const Rot<cplx> qj = q[j];
            if (j + 1 <= stop + 1) H[(j + 1 - j0) * LD + (k - j0)] = conj_(qj.c) * Vacc - qj.s * X;
            Vacc = qj.c * X + qj.s * Vacc;
            T = (-(conj_(ctc) * w) - T) * ig;
            if (j < k) P = P * b[j].s;
        }
        H[k - j0] = conj_(q[k0].c) * Vacc;
    }
    __syncwarp(gmask);
}
// Single-shift QR on the complex upper Hessenberg matrix H (destroyed), eigenvalues only, the lanes of the
// group sharing the row and column updates: Wilkinson shift (exceptional shifts at iterations 10 and 20 as in
// EISPACK comqr), one sweep = Givens rotations G_k = [conj(c) conj(s); -s c] zeroing H[k+1][k] applied to the
// rows, then their adjoints to the columns.  false: no convergence in 30 M sweeps or a non-finite value.
AMRVW_DI bool chqr_eigenvalues_warp(cplx* H, const int LD, const int M, cplx* E, cplx* gs, const int lane, const unsigned gmask, const int GL) {
    const int gbase = __ffs(gmask) - 1;
#define AMRVW_HC(i, j) H[(i) * LD + (j)]
    int en = M - 1, itn = 30 * M, its = 0;
    cplx t = cplx(0.0, 0.0);
    while (en >= 0) {
        __syncwarp(gmask);
        // highest l in 1..en with a negligible subdiagonal entry, 0 if none
        int l = 0;
        for (int base = (en / GL) * GL; base >= 0 && l == 0; base -= GL) {
            const int lc = base + lane;
            bool small = false;
            if (lc >= 1 && lc <= en) {
                const cplx d0 = AMRVW_HC(lc - 1, lc - 1), d1 = AMRVW_HC(lc, lc), sb = AMRVW_HC(lc, lc - 1);
                const double sd = fabs(d0.re) + fabs(d0.im) + fabs(d1.re) + fabs(d1.im);
                small = (sd + (fabs(sb.re) + fabs(sb.im)) == sd);
            }
            const unsigned mask = __ballot_sync(gmask, small) & gmask;
            if (mask) l = base + (31 - __clz(mask)) - gbase;
        }
        if (l == en) {
            if (lane == 0) E[en] = AMRVW_HC(en, en) + t;
            en -= 1; its = 0;
            continue;
        }
        if (itn == 0) return false;
        cplx sft;
        if (its == 10 || its == 20) {
            sft = cplx(fabs(AMRVW_HC(en, en - 1).re) + (en >= 2 ? fabs(AMRVW_HC(en - 1, en - 2).re) : 0.0), 0.0);
        } else {
            sft = AMRVW_HC(en, en);
            const cplx x = AMRVW_HC(en - 1, en) * AMRVW_HC(en, en - 1);
            if (!iszero_(x)) {
                const cplx y = (AMRVW_HC(en - 1, en - 1) - sft) * 0.5;
                cplx z = csqrt_(y * y + x);
                if (y.re * z.re + y.im * z.im < 0.0) z = -z;
                const cplx den = y + z;
                if (!iszero_(den)) sft = sft - x / den;
            }
        }
        ++its; --itn;
        __syncwarp(gmask);
        for (int i = lane; i <= en; i += GL) AMRVW_HC(i, i) = AMRVW_HC(i, i) - sft;
        t = t + sft;
        for (int k = l; k < en; ++k) {   // rows
            __syncwarp(gmask);
            const cplx a = AMRVW_HC(k, k), bb = AMRVW_HC(k + 1, k);
            const double r2 = abs2_(a) + abs2_(bb);
            cplx gc = cplx(1.0, 0.0), gsn = cplx(0.0, 0.0);
            if (r2 > 0.0) { const double ir = (r2 >= 1e-280 && r2 <= 1e280) ? rsqrt_nr(r2) : 1.0 / sqrt(r2); gc = a * ir; gsn = bb * ir; }
            __syncwarp(gmask);            // every lane has read column k before it is overwritten
            if (lane == 0) { gs[2 * k] = gc; gs[2 * k + 1] = gsn; }
            for (int j = k + lane; j <= en; j += GL) {
                const cplx u = AMRVW_HC(k, j), v = AMRVW_HC(k + 1, j);
                AMRVW_HC(k, j) = conj_(gc) * u + conj_(gsn) * v;
                AMRVW_HC(k + 1, j) = gc * v - gsn * u;
            }
        }
        for (int k = l; k < en; ++k) {   // columns
            __syncwarp(gmask);
            const cplx gc = gs[2 * k], gsn = gs[2 * k + 1];
            const int ihi = min(k + 1, en);
            for (int i = l + lane; i <= ihi; i += GL) {
                const cplx u = AMRVW_HC(i, k), v = AMRVW_HC(i, k + 1);
                AMRVW_HC(i, k) = u * gc + v * gsn;
                AMRVW_HC(i, k + 1) = v * conj_(gc) - u * conj_(gsn);
            }
        }
    }
#undef AMRVW_HC
    __syncwarp(gmask);
    bool ok = true;
    for (int i = lane; i < M; i += GL)
        if (!(isfinite(E[i].re) && isfinite(E[i].im))) ok = false;
    return __all_sync(gmask, ok);
}
// the nb + 1 shifts of a pending fat step: eigenvalues of the trailing block, by the one-bulge iteration
// on a shared-memory copy whose top rotator is made diagonal (lane 0 of the group; the others stage)
template <int MH>
AMRVW_NI void css_unit_shifts(const Problem<cplx>& pr, const PipeParams pp, const int L, const long long p, CssShiftScratch<MH>& ws, const int lane, const unsigned gmask) {
    PolyRec rec = load_rec(pr.rec + p);
    const int m = rec.nb, stop = rec.ls.ctr.stop_index, k0 = stop - m;
    Fact<cplx, false> F = make_fact<cplx, false>(pr, p, rec.N);
    __syncwarp(gmask);
    for (int t = lane; t <= m + 1; t += L) {
        const int k = k0 + t;
        ws.q[t] = ldcg_crot(F.Q + k); ws.dq[t] = ldcg_cplx(F.DQ + k);
        if (k >= 1) { ws.b[t] = ldcg_crot(F.B + k); ws.c[t] = ldcg_crot(F.Ct + k); ws.dr[t] = ldcg_cplx(F.DR + k); }
        else { ws.b[t] = make_rot<cplx>(cplx(1.0, 0.0), 0.0); ws.c[t] = ws.b[t]; ws.dr[t] = cplx(1.0, 0.0); }
    }
    __syncwarp(gmask);
    if (lane == 0) {   // make Q[k0] diagonal: keep its phase (as deflate does), drop its sine
        const cplx c0 = ws.q[0].c;
        const double a = abs_(c0);
        ws.q[0] = make_rot<cplx>(a == 0.0 ? cplx(1.0, 0.0) : c0 / a, 0.0);
    }
    __syncwarp(gmask);
    bool ok = false;
    if (pp.dense) {
        css_dense_trailing_block_warp(ws.q - k0, ws.b - k0, ws.c - k0, ws.dq - k0, ws.dr - k0, k0, stop, ws.H, MH + 1, ws.isg, m + 1, lane, gmask, L);
        ok = chqr_eigenvalues_warp(ws.H, MH + 1, m + 1, ws.e, ws.gs, lane, gmask, L);
    }
    __syncwarp(gmask);
    if (lane == 0) {
        if (!ok) {   // core-chasing iteration on the staged copy (also AMRVW_SHIFTS_CHASE)
            Fact<cplx, false> W = F;
            W.Q = ws.q - k0; W.B = ws.b - k0; W.Ct = ws.c - k0; W.DQ = ws.dq - k0; W.DR = ws.dr - k0; W.turnovers = 0;
            W.N = stop + 1;                       // the copy ends at the window: phase push-downs stop at Q[stop]
            deflate(W, k0);                       // the phase of Q[k0] goes down the copy's Q chain
            cplx* Es = ws.e - (k0 + 1);
            for (int t = 0; t <= m; ++t) ws.e[t] = cplx(0.0, 0.0);
            Counter c2 = {k0, k0 + 1, stop, AMRVW_IT_COUNT, -1};
            PolyStats sub = {0, 0, 0, 0, 0};
            drive_window(W, k0 + 1, stop, c2, Es, AMRVW_IT_MAX(m + 1), pr.seed, (uint64_t)p, rec.ls.ord, sub, ReferenceStep());
            rec.serial_turnovers += W.turnovers;
        }
        Shifts* shl = pr.shifts + p * L;
        for (int rep = 0; rep < pp.reuse; ++rep)
            for (int t = 0; t <= m; ++t) { shl[rep * (m + 1) + t].e1 = ws.e[t]; shl[rep * (m + 1) + t].e2 = cplx(0.0, 0.0); }
        rec.nb = (m + 1) * pp.reuse;
        rec.mode = 1;
        store_rec(pr.rec + p, rec);
    }
    __syncwarp(gmask);
}

// ---------------------------------------------------------------- the persistent kernel
#ifndef AMRVW_CSS_MINBLOCKS
#define AMRVW_CSS_MINBLOCKS 8
#endif
template <int L, int MH>
__global__ void __launch_bounds__(32, AMRVW_CSS_MINBLOCKS) css_unit_kernel(Problem<cplx> pr, PipeParams pp, RoundQueue* q, int* slots, unsigned mask, UnitPark* park, CssLane* parked, int seg_steps) {
    constexpr int G = 32 / L;
    __shared__ CssShiftScratch<MH> shift_scratch[G];
    __shared__ __align__(16) unsigned char css_ring[G * AMRVW_RING * AMRVW_CSS_SLOT];
    const int lane32 = threadIdx.x & 31;
    const int grp = lane32 / L, lane = lane32 % L;
    const unsigned stage_sa = (unsigned)__cvta_generic_to_shared(css_ring + grp * (AMRVW_RING * AMRVW_CSS_SLOT));

    long long pc_all = 0, pc_a = 0, pc_pop = 0, pc_lean = 0;      // warp-cycle accounting (amrvw_last_profile)
    long long pn_steps = 0, pn_lean = 0;
    int pn_units = 0;
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
        const long long p = unit_poly(pr, (long long)u * G + grp);
        int t0 = __ldcg(&up->t);

        // ---------------- phase A when due
        if (t0 < 0) {
            int m0 = 0;
            if (lane == 0 && p < pr.nbatch) m0 = css_unit_driver(pr, pp, L, p);
            __syncwarp();
            {
                const int mg = __shfl_sync(0xffffffffu, m0, grp * L);
                const unsigned gmask = (L == 32 ? 0xffffffffu : ((1u << L) - 1u)) << (grp * L);
                if (mg == 3) css_unit_shifts<MH>(pr, pp, L, p, shift_scratch[grp], lane, gmask);
                __syncwarp();
            }
            pc_a += clock64() - pc_u0;
            if (!__any_sync(0xffffffffu, m0 != 0)) {
                if (lane32 == 0) atomicSub(&q->remaining, 1u);
                pc_all += clock64() - pc_u0;
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

        // ---------------- the unit's fat steps
        int mode = 0, nb = 0, start = 1, stop = 0, ord0 = 0, N = 0;
        if (p < pr.nbatch) {
            const PolyRec* r = pr.rec + p;
            if (__ldcg(&r->iterating)) {
                mode = __ldcg(&r->mode); nb = __ldcg(&r->nb); start = __ldcg(&r->ls.ctr.start_index); stop = __ldcg(&r->ls.ctr.stop_index);
                ord0 = __ldcg(&r->ls.ord); N = __ldcg(&r->N);
            }
        }
        const bool act = mode != 0;
        const int T = act ? (stop - start + 2 * L + 1) : 0;
        const int Tmax = __reduce_max_sync(0xffffffffu, T);
        const int t1 = min(Tmax, t0 + seg_steps);
        int sp = __ldcg(&up->sp[grp]);
        int nturn = 0;
        Fact<cplx, false> F;
        F.Q = F.B = F.Ct = F.WB = F.WCt = nullptr; F.DQ = F.DR = F.WD = nullptr; F.N = N; F.turnovers = 0;
        int* dstack = nullptr;
        if (act) {
            F = make_fact<cplx, false>(pr, p, N);
            dstack = pr.dstack + (pr.offsets[p] + (long long)pr.extra * p);
        }
        // the lane's state lives in registers: only copies of it are handed to the out-of-line creation and
        // last-position steps and to the parking loops (a reference to z itself would pin it in local memory:
        // 8% of the instructions of the first version were LDL/STL, 21% of its stall samples long-scoreboard)
        CssLane z;
        if (t0 == 0) {
            z.w0 = css_identity_index(); z.w1 = z.w0;
            z.U = make_rot<cplx>(cplx(1.0, 0.0), 0.0); z.U.pad = 0.0;
            z.ca = cplx(1.0, 0.0); z.alive = 0; z.pad[0] = z.pad[1] = z.pad[2] = 0;
        } else {
            CssLane y;
            const double2* src = reinterpret_cast<const double2*>(parked + ((long long)u * 32 + lane32));
            double2* dst = reinterpret_cast<double2*>(&y);
#pragma unroll
            for (int i = 0; i < (int)(sizeof(CssLane) / 16); ++i) dst[i] = __ldcg(src + i);
            z = y;
        }
        const bool has_bulge = act && lane < nb;
        cplx shift = cplx(0.0, 0.0);
        if (has_bulge) {
            if (mode == 2) {   // exceptional: a point on the unit circle (create-bulge.jl:85-93)
                const double ang = uniform01(pr.seed ^ 0xA5A5A5A5DEADBEEFull, (uint64_t)p, (uint64_t)(ord0 + lane)) * 3.141592653589793;
                double sn, cs;
                sincos(ang, &sn, &cs);
                shift = cplx(cs, sn);
            } else {
                shift = ldcg_cplx(&pr.shifts[p * L + lane].e1);
            }
        }
        // lane 0 reads AMRVW_RING - 1 indices ahead, through shared memory
        if (act && lane == 0) {
            for (int a = 0; a < AMRVW_RING - 1; ++a) {
                css_prefetch_index(stage_sa + ((t0 + a) & (AMRVW_RING - 1)) * AMRVW_CSS_SLOT, F, min(start + t0 + a, stop + 1));
                cp_async_commit();
            }
        }
        // lock-steps during which every lane of the warp (both groups) sits at a regular chase position: no creation, no last
        // position, no window end -- the lean loop below (the complex twin of pipeline.cuh: lean_chase)
        const bool lean_ok = __all_sync(0xffffffffu, act && nb == L);
        const int lean_hi = lean_ok ? __reduce_min_sync(0xffffffffu, stop - start + 1) : 0;
        for (int t = t0; t < t1; ++t) {
            if (lean_ok && t >= 2 * L && t < lean_hi) {
                const int te = min(t1, lean_hi);
                const long long pc_l0 = clock64();
                CssIdx w0 = z.w0, w1 = z.w1;
                Rot<cplx> U = z.U;
                const cplx ca = z.ca;
                int alive = z.alive;
                int i = start + t - 2 * lane - 1;          // position of this lane's bulge
                const int tb = t;
#pragma unroll 1
                for (; t < te; ++t, ++i) {
                    CssIdx in = css_shfl_up(w0, L);
                    if (lane == 0) {
                        cp_async_wait_ring();
                        in = css_read_slot(stage_sa + (t & (AMRVW_RING - 1)) * AMRVW_CSS_SLOT);
                        css_prefetch_index(stage_sa + ((t + AMRVW_RING - 1) & (AMRVW_RING - 1)) * AMRVW_CSS_SLOT, F, min(start + t + AMRVW_RING - 1, stop + 1));
                        cp_async_commit();
                    }
                    w0 = w1; w1 = in;
                    if (alive) {   // the creation phase of this lane's bulge reaches index i + 1 (diagonal.jl:199-225)
                        if (w1.q.s == 0.0) { w1.dq = w1.dq * ca; alive = 0; }
                        else w1.q.c = ca * w1.q.c;
                    }
                    css_lane_regular(w0, w1, U);
                    if (lane == L - 1) {
                        css_store_index(F, i, w0);
                        if (fabs(w0.q.s) <= AMRVW_EPS) dstack[sp++] = i;
                    }
                }
                nturn += 3 * (te - tb);
                z.w0 = w0; z.w1 = w1; z.U = U; z.alive = alive;
                pc_lean += clock64() - pc_l0; pn_lean += te - tb;
                t = te - 1;
                continue;
            }
            const int tp = t - 2 * lane;
            // ingest: the index the previous lane finished last step (lane 0: from memory)
            CssIdx in = css_shfl_up(z.w0, L);
            if (lane == 0) {
                cp_async_wait_ring();
                in = css_read_slot(stage_sa + (t & (AMRVW_RING - 1)) * AMRVW_CSS_SLOT);
                if (act) css_prefetch_index(stage_sa + ((t + AMRVW_RING - 1) & (AMRVW_RING - 1)) * AMRVW_CSS_SLOT, F, min(start + t + AMRVW_RING - 1, stop + 1));
                cp_async_commit();
            }
            z.w0 = z.w1; z.w1 = in;
            const int knew = start + tp;          // index just ingested; the bulge sits at knew - 1
            if (has_bulge && z.alive && tp >= 2) {
                // the creation phase of this lane's bulge reaches index knew (diagonal.jl:199-225)
                if (knew > stop || z.w1.q.s == 0.0) { z.w1.dq = z.w1.dq * z.ca; z.alive = 0; }
                else z.w1.q.c = z.ca * z.w1.q.c;
            }
            const int i = knew - 1;
            if (has_bulge && tp >= 1 && i <= stop) {
                if (tp == 1) { CssLane y = z; css_create(y, F, start, stop, shift); z = y; }
                if (i < stop) css_lane_regular(z.w0, z.w1, z.U);
                else { CssLane y = z; css_step_final(y); z = y; }
                nturn += (i < stop) ? 3 : 2;
            }
            // retire: the last lane writes index i back
            if (lane == L - 1 && act && i >= start && i <= stop + 1) {
                css_store_index(F, i, z.w0);
                if (i <= stop && fabs(z.w0.q.s) <= AMRVW_EPS) dstack[sp++] = i;
            }
        }
        if (lane == 0) cp_async_wait_all();   // drain look-ahead requests past the end of the segment
        sp = __shfl_sync(0xffffffffu, sp, grp * L + L - 1);
        for (int d = L / 2; d >= 1; d >>= 1) nturn += __shfl_down_sync(0xffffffffu, nturn, d, L);
        if (t1 < Tmax) {
            const CssLane y = z;
            double2* dst = reinterpret_cast<double2*>(parked + ((long long)u * 32 + lane32));
            const double2* src = reinterpret_cast<const double2*>(&y);
#pragma unroll
            for (int i = 0; i < (int)(sizeof(CssLane) / 16); ++i) __stcg(dst + i, src[i]);
            if (lane == 0) { __stcg(&up->sp[grp], sp); __stcg(&up->nturn[grp], __ldcg(&up->nturn[grp]) + nturn); }
            if (lane32 == 0) __stcg(&up->t, t1);
        } else {
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
        __threadfence();
        __syncwarp();
        if (lane32 == 0) {
            const unsigned k = atomicAdd(&q->tail, 1u);
            atomicExch(slots + (k & mask), u);
        }
        pn_steps += t1 - t0; pn_units += 1;
        pc_all += clock64() - pc_u0;
    }
    if (pr.prof && lane32 == 0) {
        atomicAdd(pr.prof + 0, (unsigned long long)pc_all);
        atomicAdd(pr.prof + 1, (unsigned long long)pc_a);
        atomicAdd(pr.prof + 2, (unsigned long long)pc_lean);
        atomicAdd(pr.prof + 3, (unsigned long long)(pc_all - pc_a - pc_lean));
        atomicAdd(pr.prof + 4, (unsigned long long)pn_lean);
        atomicAdd(pr.prof + 5, (unsigned long long)(pn_steps - pn_lean));
        atomicAdd(pr.prof + 7, (unsigned long long)pn_units);
        atomicAdd(pr.prof + 13, (unsigned long long)pc_pop);
    }
}

// ---------------------------------------------------------------- set-up, small windows, finish (complex)
__global__ void __launch_bounds__(64) css_setup_kernel(Problem<cplx> pr) {
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
        const cplx* ps = pr.coeffs + off + first;
        const int ncoef = last - first + 1;
        rec.K = first;
        if (ncoef <= 3) {
            const int cnt = solve_simple(ps, ncoef, out);
            finish_poly(pr, p, out, cnt, rec.K, 0, rec.st);
        } else {
            const int N = ncoef - 1;
            Fact<cplx, false> F = make_fact<cplx, false>(pr, p, N);
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
// every deferred window with the reference's one-bulge schedule, in place (windows are disjoint and
// nothing else runs any more)
__global__ void __launch_bounds__(64) css_small_kernel(Problem<cplx> pr) {
    const unsigned w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= pr.counters[CNT_SMALL]) return;
    const SmallWin sw = reinterpret_cast<const SmallWin*>(pr.swlist)[w];
    const long long p = sw.poly;
    PolyRec* rw = pr.rec + p;
    Fact<cplx, false> F = make_fact<cplx, false>(pr, p, rw->N);
    cplx* E = pr.roots + pr.offsets[p] - 1;
    Counter c2 = {sw.lo - 1, sw.lo, sw.hi, sw.it_count, -1};
    PolyStats st = {0, 0, 0, 0, 0};
    int ord = 1000000 + sw.lo;
    const int left = drive_window(F, sw.lo, sw.hi, c2, E, AMRVW_IT_MAX(rw->N - 1), pr.seed, (uint64_t)p, ord, st, ReferenceStep());
    atomicAdd((unsigned long long*)&rw->serial_turnovers, (unsigned long long)F.turnovers);
    atomicAdd(&rw->st.sweeps, st.sweeps);
    if (st.exceptional) atomicAdd(&rw->st.exceptional, st.exceptional);
    if (left) atomicAdd(&rw->ls.unconv, left);
}

// ---------------------------------------------------------------- host side
template <int L, int MH>
static cudaError_t launch_css_units(const Problem<cplx>& pr, const PipeParams& pp, RoundQueue* q, int* slots, unsigned mask, UnitPark* park, CssLane* parked, int seg_steps,
                                    int sm_count, cudaStream_t st) {
    constexpr int G = 32 / L;
    const long long nunits = (pr.nbatch + G - 1) / G;
    unit_enqueue_kernel<L, cplx><<<(int)((nunits + 127) / 128), 128, 0, st>>>(pr, q, slots, mask, park);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, css_unit_kernel<L, MH>, 32, 0)) != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const long long resident = (long long)sm_count * per_sm;
    const int blocks = (int)(nunits < resident ? nunits : resident);
    css_unit_kernel<L, MH><<<blocks, 32, 0, st>>>(pr, pp, q, slots, mask, park, parked, seg_steps);
    return cudaGetLastError();
}

// lanes per polynomial and the rest of the schedule's parameters for a batch (also sizes the workspace: capi.cu)
static void choose_css_params(const amrvw_options& opt, long long nbatch, int max_len, int sm_count, int& L, PipeParams& pp) {
    const int degree = max_len - 1;
    L = opt.lanes_per_poly;
    if (L == 0) {
        L = degree < 64 ? 4 : (degree < 256 ? 8 : 16);
        // batches that leave warps free get 32 bulges per polynomial (gpurun_out/sweep80.log, sweep75.log: 64 x degree
        // 3000 944 -> 551 ms, 1000 x 1000 155 -> 121 ms; 2000 x 1000 181 ms with 16 lanes against 199 with 32)
        if (degree >= 256 && nbatch <= (long long)sm_count * 8) L = 32;
    }
    if (L > 32) L = 32;      // chained warps exist for the real rank-one path only
    // bulges per shift: the trailing block has order L / reuse.  With the core-chasing sub-solve on ONE lane the
    // shifts were 40-49% of the busy warp-cycles at reuse 1, L = 16 (config 3: reuse 1 530 ms, 2 336 ms, 4 316 ms:
    // gpurun_out/sweep62.log); with the dense block + lane-shared complex Hessenberg QR: reuse 1 201 ms (phase A
    // 26%), 2 181 ms (11%, 1.06x the turnovers), 4 237 ms (1.34x) (gpurun_out/sweep75.log)
    pp.reuse = opt.shift_reuse ? opt.shift_reuse : (L >= 8 ? 2 : 1);
    while (pp.reuse > 1 && L / pp.reuse < 2) pp.reuse /= 2;
    pp.small = opt.small_window ? opt.small_window : (L + 8 > 24 ? L + 8 : 24);
    if (pp.small < L) pp.small = L;
    if (pp.small > 96) pp.small = 96;
    pp.ms = pp.small + 4; pp.dense = opt.shift_solver != AMRVW_SHIFTS_CHASE; pp.mh = L;
}

static cudaError_t launch_css_pipelined(Problem<cplx> pr, const amrvw_options& opt, int max_len, long long total, int sm_count, cudaStream_t st, int* launches,
                                        RoundQueue* q, int* slots, unsigned mask, UnitPark* park, CssLane* parked, cudaEvent_t* ev = nullptr, RoundsTimes* times = nullptr,
                                        int* lanes_out = nullptr, unsigned long long* order_keys = nullptr) {
    int L; PipeParams pp;
    choose_css_params(opt, pr.load_hint > pr.nbatch ? pr.load_hint : pr.nbatch, max_len, sm_count, L, pp);
    if (lanes_out) *lanes_out = L;
    cudaError_t e;
    if ((e = cudaMemsetAsync(pr.counters, 0, CNT_WORDS * sizeof(unsigned), st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(q, 0, sizeof(RoundQueue), st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(slots, 0xFF, (size_t)(mask + 1) * sizeof(int), st)) != cudaSuccess) return e;
    if (times) cudaEventRecord(ev[0], st);
    css_setup_kernel<<<(int)((pr.nbatch + 63) / 64), 64, 0, st>>>(pr);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (pr.order) {
        if ((e = launch_order_by_degree(pr.rec, pr.nbatch, max_len, order_keys, const_cast<int*>(pr.order), st, launches)) != cudaSuccess) return e;
    }
    if (times) cudaEventRecord(ev[1], st);
    const int seg_steps = 256;
    switch (L) {
        case 4: e = launch_css_units<4, 8>(pr, pp, q, slots, mask, park, parked, seg_steps, sm_count, st); break;
        case 8: e = launch_css_units<8, 8>(pr, pp, q, slots, mask, park, parked, seg_steps, sm_count, st); break;
        case 16: e = launch_css_units<16, 16>(pr, pp, q, slots, mask, park, parked, seg_steps, sm_count, st); break;
        default: e = launch_css_units<32, 32>(pr, pp, q, slots, mask, park, parked, seg_steps, sm_count, st); break;
    }
    if (e != cudaSuccess) return e;
    if (times) cudaEventRecord(ev[2], st);
    {
        const long long cap = total / 3 + pr.nbatch + 16;
        css_small_kernel<<<(int)((cap + 63) / 64), 64, 0, st>>>(pr);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if (times) cudaEventRecord(ev[3], st);
    {
        const int smem_elems = max_len - 1 <= AMRVW_SORT_SMEM ? (max_len > 1 ? max_len - 1 : 1) : AMRVW_SORT_SMEM;
        const size_t smem = (size_t)smem_elems * sizeof(cplx);
        if ((e = cudaFuncSetAttribute(rounds_finish_kernel<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        rounds_finish_kernel<cplx><<<(int)pr.nbatch, 256, smem, st>>>(pr, smem_elems);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    *launches += 5;
    if (times) return read_round_times(ev, times, st);
    return cudaSuccess;
}

}  // namespace amrvw
