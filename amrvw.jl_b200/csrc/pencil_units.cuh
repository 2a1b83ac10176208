// pencil_units.cuh -- roots(vs, ws), Float64 companion pencil (pencil.jl, companion.jl:163-204,282-303),
// on the unit queue: pipelined multishift real double shift with R = V W^{-1}.
//
// Same architecture as the rank-one RDS path (rounds.cuh): work units (one warp = 32/L polynomials)
// circulate through a device-side FIFO; a visit runs phase A when the unit's fat step is through
// (driver on lane 0 of each group, then the shifts of the next fat step) and then up to SEG lock-steps
// of the systolic chase of L bulges, one per lane.
//
// What the pencil changes (reference: pencil.jl:63-80, bulge-step.jl:36-49, RDS.jl):
//   * the upper-triangular factor is V W^{-1} with V and W both rank-one factorizations: FIVE chains
//     (Q; V.B, V.Ct; W.B, W.Ct).  passthrough!(RFp, U) sends U' left -> right through W.Ct and W.B and
//     then U right -> left through V.B and V.Ct: 4 turnovers where the rank-one R needs 2, so a chase
//     step is 11 turnovers (V and U through R: 8, through Q: 2, with the limb: 1);
//   * a bulge at position i still touches only indices i..i+2 of every chain, so bulges three indices
//     apart commute exactly as in the rank-one case; a lane owns a window of 3 indices x 5 chains
//     (30 doubles) plus its bulge U, V, W (6 doubles);
//   * matrix entries for the shifts and the bulge creation are the pencil's (pencil.jl:29-45, r_entry<P>).
//
// Shifts: the eigenvalues of the trailing block of the window, from the dense block (V and W written out
// by the rank-one recurrences, X = V W^{-1} by back substitution, H = Q X) and the lane-shared Hessenberg
// QR of shifts.cuh; the core-chasing iteration on a shared-memory copy is the fall-back.
#pragma once
#include "rounds.cuh"

namespace amrvw {

struct PenIdx { Rot<double> q, b, c, wb, wc; };                  // one index of the five chains
struct PenLane { PenIdx w0, w1, w2; Rot<double> U, V, W; };      // what a lane owns between lock-steps
static_assert(sizeof(PenLane) == 288, "PenLane is parked in 16-byte words");

AMRVW_DI PenIdx pen_identity_index() {
    PenIdx x;
    x.q = x.b = x.c = x.wb = x.wc = make_rot<double>(1.0, 0.0);
    return x;
}
// look-ahead ring of the feeding lane: five 16-byte cp.async.cg per index, AMRVW_RING - 1 steps ahead
#define AMRVW_PEN_SLOT 80
AMRVW_DI void pen_prefetch_index(const unsigned slot_sa, const Fact<double, true>& F, const int k) {
    cp_async16_sa(slot_sa, F.Q + k); cp_async16_sa(slot_sa + 16, F.B + k); cp_async16_sa(slot_sa + 32, F.Ct + k);
    cp_async16_sa(slot_sa + 48, F.WB + k); cp_async16_sa(slot_sa + 64, F.WCt + k);
}
AMRVW_DI PenIdx pen_read_slot(const unsigned slot_sa) {
    PenIdx x;
    x.q = lds_rot(slot_sa); x.b = lds_rot(slot_sa + 16); x.c = lds_rot(slot_sa + 32); x.wb = lds_rot(slot_sa + 48); x.wc = lds_rot(slot_sa + 64);
    return x;
}
AMRVW_DI void pen_store_index(const Fact<double, true>& F, const int k, const PenIdx& x) {
    stg_rot(F.Q + k, x.q); stg_rot(F.B + k, x.b); stg_rot(F.Ct + k, x.c); stg_rot(F.WB + k, x.wb); stg_rot(F.WCt + k, x.wc);
}
AMRVW_DI PenIdx pen_shfl_up(const PenIdx& x, const int width) {
    PenIdx o;
    o.q = shfl_up_rot(x.q, width); o.b = shfl_up_rot(x.b, width); o.c = shfl_up_rot(x.c, width);
    o.wb = shfl_up_rot(x.wb, width); o.wc = shfl_up_rot(x.wc, width);
    return o;
}

// passthrough!(RFp, U) (pencil.jl:63-71) with U at index i; x0, x1 = indices i, i+1.  U' goes left -> right
// through W.Ct (chains.jl:391) and W.B (chains.jl:381), then U right -> left through V.B (chains.jl:360)
// and V.Ct (chains.jl:370); U comes back at index i.
AMRVW_DI void pen_pass_R(PenIdx& x0, PenIdx& x1, Rot<double>& U) {
    Rot<double> Ut = adjoint(U), r1, r2, r3;
    turnover_fast(Ut, x1.wc, x0.wc, r1, r2, r3); x1.wc = r1; x0.wc = r2; Ut = r3;
    turnover_fast(Ut, x0.wb, x1.wb, r1, r2, r3); x0.wb = r1; x1.wb = r2; Ut = r3;
    U = adjoint(Ut);
    turnover_fast(x0.b, x1.b, U, r1, r2, r3); U = r1; x0.b = r2; x1.b = r3;
    turnover_fast(x1.c, x0.c, U, r1, r2, r3); U = r1; x1.c = r2; x0.c = r3;
}
// regular position (RDS.jl:42-70, 97-108): V@i+1 and U@i through R, then through Q, then the limb.
// Inline: the lane's state stays in registers (only the creation and last-position steps and the parking
// loops work on copies in local memory)
AMRVW_DI void pen_step_regular(PenLane& z) {
    pen_pass_R(z.w1, z.w2, z.V);
    pen_pass_R(z.w0, z.w1, z.U);
    Rot<double> r1, r2, r3;
    turnover_fast(z.w1.q, z.w2.q, z.V, r1, r2, r3); z.V = r1; z.w1.q = r2; z.w2.q = r3;
    turnover_fast(z.w0.q, z.w1.q, z.U, r1, r2, r3); z.U = r1; z.w0.q = r2; z.w1.q = r3;
    turnover_fast(z.W, z.V, z.U, r1, r2, r3); z.V = r1; z.U = r2; z.W = r3;
}
// last position j == stop (RDS.jl:110-128): both through R, V fused into Q[j], U through Q, limb fused
// into U, U through R once more and fused into Q[j]
AMRVW_NI void pen_step_final(PenLane& z, const double pbot) {
    pen_pass_R(z.w1, z.w2, z.V);
    pen_pass_R(z.w0, z.w1, z.U);
    z.V.s = sign_(pbot) * z.V.s;
    z.w1.q = fuse(z.w1.q, z.V);
    Rot<double> r1, r2, r3;
    turnover_fast(z.w0.q, z.w1.q, z.U, r1, r2, r3); z.U = r1; z.w0.q = r2; z.w1.q = r3;
    z.U = fuse(z.W, z.U);
    pen_pass_R(z.w1, z.w2, z.U);
    z.U.s = sign_(pbot) * z.U.s;
    z.w1.q = fuse(z.w1.q, z.U);
}
// create_bulge (create-bulge.jl:20-77) on the lane's window [start .. start+2] plus index start-1 (static
// during a fat step, read from memory) + absorb_Ut (RDS.jl:14-36)
AMRVW_NI void pen_create(PenLane& z, const Fact<double, true>& F, const int start, const int stop, const int mode, const Shifts* shl, const int lane,
                         const uint64_t seed, const uint64_t poly, const int ordinal) {
    PenIdx wm = pen_identity_index();
    wm.q = ldcg_rot(F.Q + start - 1);                 // Q[0] is the (1,0) sentinel
    if (start > 1) { wm.b = ldcg_rot(F.B + start - 1); wm.c = ldcg_rot(F.Ct + start - 1); wm.wb = ldcg_rot(F.WB + start - 1); wm.wc = ldcg_rot(F.WCt + start - 1); }
    Rot<double> U, V;
    if (mode == 2) {
        const double ang = uniform01(seed ^ 0xA5A5A5A5DEADBEEFull, poly, (uint64_t)ordinal) * 3.141592653589793;
        double sn, cs;
        sincos(ang, &sn, &cs);
        U = make_rot<double>(sn, cs);
        V = make_rot<double>(sn, -cs);
    } else {
        Shifts sh;
        const double2 v1 = __ldcg(reinterpret_cast<const double2*>(shl + lane)), v2 = __ldcg(reinterpret_cast<const double2*>(shl + lane) + 1);
        sh.e1 = cplx(v1.x, v1.y); sh.e2 = cplx(v2.x, v2.y);
        Rot<double> lq[4] = {wm.q, z.w0.q, z.w1.q, z.w2.q}, lb[4] = {wm.b, z.w0.b, z.w1.b, z.w2.b}, lc[4] = {wm.c, z.w0.c, z.w1.c, z.w2.c};
        Rot<double> lwb[4] = {wm.wb, z.w0.wb, z.w1.wb, z.w2.wb}, lwc[4] = {wm.wc, z.w0.wc, z.w1.wc, z.w2.wc};
        Fact<double, true> Fl = F;
        Fl.Q = lq - (start - 1); Fl.B = lb - (start - 1); Fl.Ct = lc - (start - 1); Fl.WB = lwb - (start - 1); Fl.WCt = lwc - (start - 1);
        Counter cl = {start - 1, start, stop, AMRVW_IT_COUNT, -1};
        int dummy = 0;
        PolyStats ds = {0, 0, 0, 0, 0};
        create_bulge_rds(Fl, cl, U, V, &sh, seed, poly, dummy, ds);
    }
    const double ptop = sign_(wm.q.c);
    Rot<double> w = z.w0.q;
    w.s = sign_(ptop) * w.s;
    Rot<double> r1, r2, r3;
    turnover_fast(adjoint(U), adjoint(V), w, r1, r2, r3);
    z.W = r1;
    z.w1.q = fuse(r3, z.w1.q);
    r2.s = sign_(ptop) * r2.s;
    z.w0.q = r2;
    z.U = U; z.V = V;
}

// ---------------------------------------------------------------- phase A of a pencil unit
AMRVW_DI void pen_diagonal_block_cg(const Fact<double, true>& F, const int j, double A[4]) {
    Rot<double> lq[3], lb[3], lc[3], lwb[3], lwc[3];
    const Rot<double> one = make_rot<double>(1.0, 0.0);
    for (int t = 0; t < 3; ++t) {
        const int k = j - 1 + t;
        lq[t] = ldcg_rot(F.Q + k);                            // Q[0], Q[N] are the (1,0) sentinels
        if (k >= 1) { lb[t] = ldcg_rot(F.B + k); lc[t] = ldcg_rot(F.Ct + k); lwb[t] = ldcg_rot(F.WB + k); lwc[t] = ldcg_rot(F.WCt + k); }
        else { lb[t] = one; lc[t] = one; lwb[t] = one; lwc[t] = one; }
    }
    Fact<double, true> Fl = F;
    Fl.Q = lq - (j - 1); Fl.B = lb - (j - 1); Fl.Ct = lc - (j - 1); Fl.WB = lwb - (j - 1); Fl.WCt = lwc - (j - 1);
    diagonal_block(Fl, j, A);
}
AMRVW_DI void pen_check_deflation_stack(Fact<double, true>& F, LeaderState& ls, const int* dstack) {
    if (ls.sp > 0) {
        const int k = __ldcg(dstack + ls.sp - 1);
        if (k >= ls.ctr.start_index && k <= ls.ctr.stop_index) {
            F.Q[k] = make_rot<double>(sign_(ldcg_rot(F.Q + k).c), 0.0);   // deflate (RDS.jl:134-141)
            ls.ctr.zero_index = k;
            ls.ctr.start_index = k + 1;
            ls.ctr.it_count = AMRVW_IT_COUNT;
        }
    }
}
// the driver (AMRVW_algorithm.jl:10-138) up to the next fat step: 0 finished, 2 exceptional fat step,
// 3 fat step whose shifts are the eigenvalues of the trailing (nb + 1) x (nb + 1) block
AMRVW_NI int pen_unit_driver(const Problem<double>& pr, const PipeParams pp, const int L, const long long p) {
    PolyRec rec = load_rec(pr.rec + p);
    if (!rec.iterating) return 0;
    Fact<double, true> F = make_fact<double, true>(pr, p, rec.N);
    cplx* E = pr.roots + pr.offsets[p] - 1;
    const int* dstack = pr.dstack + (pr.offsets[p] + (long long)pr.extra * p);
    SmallSink sink;
    sink.list = pr.swlist; sink.count = pr.counters + CNT_SMALL; sink.poly = (int)p;
    LeaderState& ls = rec.ls;
    Counter& ctr = ls.ctr;
    double A[4];
    int mode = 0, nb = 0;
    while (ls.kk <= ls.it_max) {
        if (ctr.stop_index < 1) break;
        pen_check_deflation_stack(F, ls, dstack);
        ++ls.kk;
        const int k = ctr.stop_index;
        const int delta = ctr.stop_index - ctr.zero_index;
        if (delta == 1) {
            pen_diagonal_block_cg(F, k, A);
            cplx e1, e2;
            eig2x2(A, e1, e2);
            E[k] = e1; E[k + 1] = e2;
            if (ctr.stop_index == 2) { pen_diagonal_block_cg(F, 1, A); E[1] = to_c(A[0]); ctr.stop_index = 0; break; }
            pop_window(ls);
            ctr.stop_index = ctr.zero_index - 1; ctr.zero_index = 0; ctr.start_index = 1;
            pen_check_deflation_stack(F, ls, dstack);
            if (ctr.stop_index < 1) break;
        } else if (delta <= 0) {
            pen_diagonal_block_cg(F, k, A);
            if (ctr.stop_index == 1) { E[1] = to_c(A[0]); E[2] = to_c(A[3]); ctr.stop_index = 0; break; }
            E[k + 1] = to_c(A[3]);
            pop_window(ls);
            ctr.stop_index = ctr.zero_index - 1;
            if (ctr.stop_index < 1) break;
            ctr.zero_index = 0; ctr.start_index = 1;
            pen_check_deflation_stack(F, ls, dstack);
        } else if (delta <= pp.small) {
            // decoupled small window: queued for pen_small_kernel (nothing above it touches it again)
            const int lo2 = ctr.start_index, hi2 = ctr.stop_index;
            sink.push(lo2, hi2, ctr.it_count);
            if (lo2 == 2) { pen_diagonal_block_cg(F, 1, A); E[1] = to_c(A[0]); }
            pop_window(ls);
            ctr.stop_index = lo2 - 2; ctr.zero_index = 0; ctr.start_index = 1; ctr.it_count = AMRVW_IT_COUNT;
            if (ctr.stop_index >= 1) pen_check_deflation_stack(F, ls, dstack);
            ls.pn_small += 1;
        } else if (ctr.it_count == 0) {
            rec.st.fat_steps++;
            nb = L; mode = 2;
            break;
        } else {
            int m = min(2 * L / pp.reuse - 1, delta - 1);
            if (m % 2 == 0) m -= 1;
            nb = m;
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

// Dense trailing block of the pencil, lanes of a group sharing the work.  V and W (each a rank-one factor)
// are written out column by column with the recurrences of dense_trailing_block (shifts.cuh) without the
// Q multiplication; X = V W^{-1} row by row (x W = v, forward substitution over the columns of the upper
// triangular W); H = Q X with Q the unitary Hessenberg matrix of the staged Q rotators, q[k0] diagonal.
// RV, RW: M x LD row-major upper triangular scratch; H overwrites RV.  One column (RV, RW, H) or one row
// (X) per lane at a time.
AMRVW_DI void pen_dense_R_warp(const Rot<double>* b, const Rot<double>* c, const int j0, const int stop, double* R, const int LD, double* isg, const int M,
                               const int lane, const unsigned gmask, const int L) {
    for (int t = lane; t < M; t += L) isg[t] = rcp_(c[j0 + t].s);
    __syncwarp(gmask);
    for (int k = j0 + lane; k <= stop + 1; k += L) {
        double P = 1.0, T = 0.0;
        const double ck = b[k].c, sk = b[k].s;
        for (int j = k; j >= j0; --j) {
            const double w = (j == k) ? -sk : (b[j].c * P) * ck;
            const double ctc = c[j].c;
            const double ig = isg[j - j0];
            R[(j - j0) * LD + (k - j0)] = (w + ctc * T) * ig;
            T = (-(ctc * w) - T) * ig;
            if (j < k) P = P * b[j].s;
        }
        for (int j = k + 1; j <= stop + 1; ++j) R[(j - j0) * LD + (k - j0)] = 0.0;
    }
    __syncwarp(gmask);
}
template <int MH>
struct __align__(16) PenShiftScratch {
    Rot<double> q[MH + 3], b[MH + 3], c[MH + 3], wb[MH + 3], wc[MH + 3];
    cplx e[MH + 3];
    double RV[MH * (MH + 1)], RW[MH * (MH + 1)];   // RV: V, then X = V W^{-1}, then H = Q X, in place
    double isg[MH + 2];
};
// q/b/c/wb/wc: staged rotators with GLOBAL indices (q[k0 .. stop+1], the others [k0+1 .. stop+1] valid)
template <int MH>
AMRVW_DI void pen_dense_trailing_block_warp(PenShiftScratch<MH>& ws, const int k0, const int stop, const int M, const int lane, const unsigned gmask, const int L) {
    constexpr int LD = MH + 1;
    const int j0 = k0 + 1;
    const Rot<double>* q = ws.q - k0;
    pen_dense_R_warp(ws.b - k0, ws.c - k0, j0, stop, ws.RV, LD, ws.isg, M, lane, gmask, L);
    pen_dense_R_warp(ws.wb - k0, ws.wc - k0, j0, stop, ws.RW, LD, ws.isg, M, lane, gmask, L);
    // X = V W^{-1}: row r of X solves x W = v; overwrite RV row by row (entries left of the diagonal are 0)
    for (int r = lane; r < M; r += L) {
        for (int k = r; k < M; ++k) {
            double acc = ws.RV[r * LD + k];
            for (int j = r; j < k; ++j) acc -= ws.RV[r * LD + j] * ws.RW[j * LD + k];
            ws.RV[r * LD + k] = acc * rcp_(ws.RW[k * LD + k]);
        }
    }
    __syncwarp(gmask);
    // H = Q X column by column, bottom-up (the Vacc recurrence of dense_trailing_block), in place: row
    // jj + 1 of the column is written after it has been read, and everything below the subdiagonal is
    // still the zero fill of pen_dense_R_warp
    for (int kk = lane; kk < M; kk += L) {
        double Vacc = 0.0;
        for (int jj = kk; jj >= 0; --jj) {
            const double R = ws.RV[jj * LD + kk];
            const Rot<double> qj = q[j0 + jj];
            if (jj + 1 < M) ws.RV[(jj + 1) * LD + kk] = qj.c * Vacc - qj.s * R;
            Vacc = qj.c * R + qj.s * Vacc;
        }
        ws.RV[kk] = q[k0].c * Vacc;
    }
    __syncwarp(gmask);
}

// the shifts of a pending fat step (mode 3), each lane group on its own polynomial
template <int MH>
AMRVW_NI void pen_unit_shifts(const Problem<double>& pr, const PipeParams pp, const int L, const long long p, PenShiftScratch<MH>& ws, const int lane, const unsigned gmask) {
    constexpr int LDH = MH + 1;
    PolyRec rec = load_rec(pr.rec + p);
    const int m = rec.nb, stop = rec.ls.ctr.stop_index, k0 = stop - m, M = m + 1;
    Fact<double, true> F = make_fact<double, true>(pr, p, rec.N);
    __syncwarp(gmask);
    const Rot<double> one = make_rot<double>(1.0, 0.0);
    for (int t = lane; t <= m + 1; t += L) {
        const int k = k0 + t;
        ws.q[t] = ldcg_rot(F.Q + k);
        if (k >= 1) { ws.b[t] = ldcg_rot(F.B + k); ws.c[t] = ldcg_rot(F.Ct + k); ws.wb[t] = ldcg_rot(F.WB + k); ws.wc[t] = ldcg_rot(F.WCt + k); }
        else { ws.b[t] = one; ws.c[t] = one; ws.wb[t] = one; ws.wc[t] = one; }
    }
    __syncwarp(gmask);
    if (lane == 0) {
        const double sg = sign_(ws.q[0].c);
        ws.q[0] = make_rot<double>(sg == 0.0 ? 1.0 : sg, 0.0);
    }
    __syncwarp(gmask);
    bool ok = false;
    if (pp.dense) {
        pen_dense_trailing_block_warp<MH>(ws, k0, stop, M, lane, gmask, L);
        ok = hqr_eigenvalues_warp(ws.RV, LDH, M, ws.e, lane, gmask, L);
    }
    __syncwarp(gmask);
    if (lane == 0) {
        if (!ok) {   // core-chasing iteration on the staged copy (also AMRVW_SHIFTS_CHASE)
            Fact<double, true> W = F;
            W.Q = ws.q - k0; W.B = ws.b - k0; W.Ct = ws.c - k0; W.WB = ws.wb - k0; W.WCt = ws.wc - k0; W.turnovers = 0;
            cplx* Es = ws.e - (k0 + 1);
            for (int t = 0; t <= m; ++t) ws.e[t] = cplx(0.0, 0.0);
            Counter c2 = {k0, k0 + 1, stop, AMRVW_IT_COUNT, -1};
            PolyStats sub = {0, 0, 0, 0, 0};
            drive_window(W, k0 + 1, stop, c2, Es, AMRVW_IT_MAX(m + 1), pr.seed, (uint64_t)p, rec.ls.ord, sub, ReferenceStep());
            rec.serial_turnovers += W.turnovers;
        }
        rec.nb = pair_shifts(ws.e, m, pr.shifts + p * L, L, pp.reuse);
        rec.mode = 1;
        store_rec(pr.rec + p, rec);
    }
    __syncwarp(gmask);
}

// ---------------------------------------------------------------- the persistent kernel
template <int L, int MH>
__global__ void __launch_bounds__(32, 8) pen_unit_kernel(Problem<double> pr, PipeParams pp, RoundQueue* q, int* slots, unsigned mask, UnitPark* park, PenLane* parked, int seg_steps) {
    constexpr int G = 32 / L;
    __shared__ PenShiftScratch<MH> shift_scratch[G];
    __shared__ __align__(16) unsigned char pen_ring[G * AMRVW_RING * AMRVW_PEN_SLOT];
    const int lane32 = threadIdx.x & 31;
    const int grp = lane32 / L, lane = lane32 % L;
    const unsigned stage_sa = (unsigned)__cvta_generic_to_shared(pen_ring + grp * (AMRVW_RING * AMRVW_PEN_SLOT));

    long long pc_all = 0, pc_a = 0, pc_pop = 0;      // warp-cycle accounting (amrvw_last_profile)
    long long pn_steps = 0;
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
            if (lane == 0 && p < pr.nbatch) m0 = pen_unit_driver(pr, pp, L, p);
            __syncwarp();
            {
                const int mg = __shfl_sync(0xffffffffu, m0, grp * L);
                const unsigned gmask = (L == 32 ? 0xffffffffu : ((1u << L) - 1u)) << (grp * L);
                if (mg == 3) pen_unit_shifts<MH>(pr, pp, L, p, shift_scratch[grp], lane, gmask);
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
        const int T = act ? (stop - start + 3 * L + 1) : 0;
        const int Tmax = __reduce_max_sync(0xffffffffu, T);
        const int t1 = min(Tmax, t0 + seg_steps);
        int sp = __ldcg(&up->sp[grp]);
        int nturn = 0;
        Fact<double, true> F;
        F.Q = F.B = F.Ct = F.WB = F.WCt = nullptr; F.DQ = F.DR = F.WD = nullptr; F.N = N; F.turnovers = 0;
        int* dstack = nullptr;
        if (act) {
            F = make_fact<double, true>(pr, p, N);
            dstack = pr.dstack + (pr.offsets[p] + (long long)pr.extra * p);
        }
        PenLane z;
        if (t0 == 0) {
            z.w0 = pen_identity_index(); z.w1 = z.w0; z.w2 = z.w0;
            z.U = z.V = z.W = make_rot<double>(1.0, 0.0);
        } else {
            PenLane y;
            const double2* src = reinterpret_cast<const double2*>(parked + ((long long)u * 32 + lane32));
            double2* dst = reinterpret_cast<double2*>(&y);
#pragma unroll
            for (int i = 0; i < (int)(sizeof(PenLane) / 16); ++i) dst[i] = __ldcg(src + i);
            z = y;
        }
        const bool has_bulge = act && lane < nb;
        double pbot = 1.0;
        if (act) pbot = sign_(ldcg_rot(F.Q + stop + 1).c);      // static during the fat step; Q[N] is the (1,0) sentinel
        // lane 0 reads AMRVW_RING - 1 indices ahead, through shared memory
        if (act && lane == 0) {
            for (int a = 0; a < AMRVW_RING - 1; ++a) {
                pen_prefetch_index(stage_sa + ((t0 + a) & (AMRVW_RING - 1)) * AMRVW_PEN_SLOT, F, min(start + t0 + a, stop + 1));
                cp_async_commit();
            }
        }
        for (int t = t0; t < t1; ++t) {
            const int tp = t - 3 * lane;
            // ingest: the index the previous lane finished last step (lane 0: from memory)
            PenIdx in = pen_shfl_up(z.w0, L);
            if (lane == 0) {
                cp_async_wait_ring();
                in = pen_read_slot(stage_sa + (t & (AMRVW_RING - 1)) * AMRVW_PEN_SLOT);
                if (act) pen_prefetch_index(stage_sa + ((t + AMRVW_RING - 1) & (AMRVW_RING - 1)) * AMRVW_PEN_SLOT, F, min(start + t + AMRVW_RING - 1, stop + 1));
                cp_async_commit();
            }
            z.w0 = z.w1; z.w1 = z.w2; z.w2 = in;
            const int i = start + tp - 2;   // position of this lane's bulge; window = indices i..i+2
            if (has_bulge && tp >= 2 && i <= stop - 1) {
                if (tp == 2) { PenLane y = z; pen_create(y, F, start, stop, mode, pr.shifts + p * L, lane, pr.seed, (uint64_t)p, ord0 + lane); z = y; nturn += 1; }
                if (i <= stop - 2) pen_step_regular(z);
                else { PenLane y = z; pen_step_final(y, pbot); z = y; }
                nturn += (i <= stop - 2) ? 11 : 13;
            }
            // retire: the last lane writes index i back
            if (lane == L - 1 && act && i >= start && i <= stop + 1) {
                pen_store_index(F, i, z.w0);
                if (i <= stop && fabs(z.w0.q.s) <= AMRVW_EPS) dstack[sp++] = i;
            }
        }
        if (lane == 0) cp_async_wait_all();   // drain look-ahead requests past the end of the segment
        sp = __shfl_sync(0xffffffffu, sp, grp * L + L - 1);
        for (int d = L / 2; d >= 1; d >>= 1) nturn += __shfl_down_sync(0xffffffffu, nturn, d, L);
        if (t1 < Tmax) {
            const PenLane y = z;
            double2* dst = reinterpret_cast<double2*>(parked + ((long long)u * 32 + lane32));
            const double2* src = reinterpret_cast<const double2*>(&y);
#pragma unroll
            for (int i = 0; i < (int)(sizeof(PenLane) / 16); ++i) __stcg(dst + i, src[i]);
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
        atomicAdd(pr.prof + 3, (unsigned long long)(pc_all - pc_a));
        atomicAdd(pr.prof + 5, (unsigned long long)pn_steps);
        atomicAdd(pr.prof + 7, (unsigned long long)pn_units);
        atomicAdd(pr.prof + 13, (unsigned long long)pc_pop);
    }
}

// ---------------------------------------------------------------- set-up, small windows (pencil)
__global__ void __launch_bounds__(64) pen_setup_kernel(Problem<double> pr) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= pr.nbatch) return;
    PolyRec rec;
    rec.ls.kk = 0; rec.ls.it_max = 0; rec.ls.ord = 0; rec.ls.unconv = 0; rec.ls.sp = 0; rec.ls.pc_sub = 0; rec.ls.pc_small = 0; rec.ls.pn_small = 0;
    rec.ls.ctr = {0, 1, 0, AMRVW_IT_COUNT, -1};
    rec.st = {0, 0, 0, 0, 0};
    rec.serial_turnovers = 0; rec.chase_turnovers = 0;
    rec.N = 0; rec.K = 0; rec.mode = 0; rec.nb = 0; rec.iterating = 0; rec.pad = 0;
    cplx* out = pr.roots + pr.offsets[p];
    Fact<double, true> F;
    int Kz = 0;
    const int n = pencil_prepare(pr, p, out, rec.st, F, Kz);   // trimming, closed forms, factorization (kernels.cuh)
    if (n > 0) {
        rec.N = n; rec.K = Kz;
        rec.ls.ctr = {0, 1, n - 1, AMRVW_IT_COUNT, -1};
        rec.ls.it_max = AMRVW_IT_MAX(n - 1);
        rec.iterating = 1;
    }
    pr.rec[p] = rec;
}
// every deferred window with the reference's one-bulge schedule on a local-memory copy (windows are
// disjoint, nothing else runs any more, only the eigenvalues leave)
template <int SM>
__global__ void __launch_bounds__(64) pen_small_kernel(Problem<double> pr) {
    const unsigned w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= pr.counters[CNT_SMALL]) return;
    const SmallWin sw = reinterpret_cast<const SmallWin*>(pr.swlist)[w];
    const long long p = sw.poly;
    PolyRec* rw = pr.rec + p;
    Fact<double, true> F = make_fact<double, true>(pr, p, rw->N);
    cplx* E = pr.roots + pr.offsets[p] - 1;
    Rot<double> sq[SM + 3], sb[SM + 3], sc[SM + 3], swb[SM + 3], swc[SM + 3];
    const int k = sw.lo - 1, cnt = sw.hi + 1 - k + 1;
    const Rot<double> one = make_rot<double>(1.0, 0.0);
    for (int t = 0; t < cnt; ++t) {
        sq[t] = F.Q[k + t];
        if (k + t >= 1) { sb[t] = F.B[k + t]; sc[t] = F.Ct[k + t]; swb[t] = F.WB[k + t]; swc[t] = F.WCt[k + t]; }
        else { sb[t] = one; sc[t] = one; swb[t] = one; swc[t] = one; }
    }
    Fact<double, true> W = F;
    W.Q = sq - k; W.B = sb - k; W.Ct = sc - k; W.WB = swb - k; W.WCt = swc - k; W.turnovers = 0;
    Counter c2 = {sw.lo - 1, sw.lo, sw.hi, sw.it_count, -1};
    PolyStats st = {0, 0, 0, 0, 0};
    int ord = 1000000 + sw.lo;
    const int left = drive_window(W, sw.lo, sw.hi, c2, E, AMRVW_IT_MAX(rw->N - 1), pr.seed, (uint64_t)p, ord, st, ReferenceStep());
    atomicAdd((unsigned long long*)&rw->serial_turnovers, (unsigned long long)W.turnovers);
    atomicAdd(&rw->st.sweeps, st.sweeps);
    if (st.exceptional) atomicAdd(&rw->st.exceptional, st.exceptional);
    if (left) atomicAdd(&rw->ls.unconv, left);
}

// ---------------------------------------------------------------- host side
template <int L, int MH>
static cudaError_t launch_pen_units(const Problem<double>& pr, const PipeParams& pp, const QueueBufs& qb, PenLane* parked, int seg_steps, int sm_count, cudaStream_t st) {
    constexpr int G = 32 / L;
    const long long nunits = (pr.nbatch + G - 1) / G;
    unit_enqueue_kernel<L, double><<<(int)((nunits + 127) / 128), 128, 0, st>>>(pr, qb.q, qb.slots, qb.mask, qb.park);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pen_unit_kernel<L, MH>, 32, 0)) != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const long long resident = (long long)sm_count * per_sm;     // every warp of the grid must be resident: they wait on each other
    const int blocks = (int)(nunits < resident ? nunits : resident);
    pen_unit_kernel<L, MH><<<blocks, 32, 0, st>>>(pr, pp, qb.q, qb.slots, qb.mask, qb.park, parked, seg_steps);
    return cudaGetLastError();
}

// lanes per polynomial and the rest of the schedule's parameters for a batch (also sizes the workspace: capi.cu)
static void choose_pencil_params(const amrvw_options& opt, long long nbatch, int max_len, int sm_count, int& L, PipeParams& pp) {
    const int degree = max_len;          // vs has one entry per degree
    L = opt.lanes_per_poly;
    if (L == 0) {
        L = degree < 48 ? 4 : (degree < 200 ? 8 : 16);
        // small batches of large pencils: 32 bulges (100 x degree 2000: 376 -> 257 ms; 500 x 500 stays at 16: 49 vs 53 ms)
        if (degree >= 1000 && nbatch <= (long long)sm_count * 8) L = 32;
    }
    if (L > 32) L = 32;      // chained warps exist for the real rank-one path only
    pp.reuse = opt.shift_reuse ? opt.shift_reuse : 2;
    if (pp.reuse > L / 2) pp.reuse = L / 2 > 0 ? L / 2 : 1;
    while (2 * L / pp.reuse > 32) pp.reuse *= 2;     // the largest dense trailing block of the pencil kernels is 32 x 32
    const int mshift = 2 * L / pp.reuse;
    pp.small = opt.small_window ? opt.small_window : (mshift + 8 > 24 ? mshift + 8 : 24);
    if (pp.small < mshift) pp.small = mshift;
    if (pp.small > 96) pp.small = 96;
    pp.ms = pp.small + 4;
    pp.dense = opt.shift_solver != AMRVW_SHIFTS_CHASE;
    pp.mh = mshift;
}

static cudaError_t launch_pencil_pipelined(Problem<double> pr, const amrvw_options& opt, int max_len, long long total, int sm_count, cudaStream_t st, int* launches,
                                           const QueueBufs& qb, PenLane* parked, cudaEvent_t* ev = nullptr, RoundsTimes* times = nullptr, int* lanes_out = nullptr,
                                           unsigned long long* order_keys = nullptr) {
    int L; PipeParams pp;
    choose_pencil_params(opt, pr.load_hint > pr.nbatch ? pr.load_hint : pr.nbatch, max_len, sm_count, L, pp);
    if (lanes_out) *lanes_out = L;
    cudaError_t e;
    if ((e = cudaMemsetAsync(pr.counters, 0, CNT_WORDS * sizeof(unsigned), st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(qb.q, 0, sizeof(RoundQueue), st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(qb.slots, 0xFF, (size_t)(qb.mask + 1) * sizeof(int), st)) != cudaSuccess) return e;
    if (times) cudaEventRecord(ev[0], st);
    pen_setup_kernel<<<(int)((pr.nbatch + 63) / 64), 64, 0, st>>>(pr);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (pr.order) {
        if ((e = launch_order_by_degree(pr.rec, pr.nbatch, max_len, order_keys, const_cast<int*>(pr.order), st, launches)) != cudaSuccess) return e;
    }
    if (times) cudaEventRecord(ev[1], st);
    const int seg_steps = 256;
    e = cudaErrorInvalidValue;
    switch (L) {
        case 4: if (pp.mh <= 8) e = launch_pen_units<4, 8>(pr, pp, qb, parked, seg_steps, sm_count, st); break;
        case 8: if (pp.mh <= 16) e = launch_pen_units<8, 16>(pr, pp, qb, parked, seg_steps, sm_count, st); break;
        case 16:
            if (pp.mh <= 16) e = launch_pen_units<16, 16>(pr, pp, qb, parked, seg_steps, sm_count, st);
            else e = launch_pen_units<16, 32>(pr, pp, qb, parked, seg_steps, sm_count, st);
            break;
        default:
            if (pp.mh <= 16) e = launch_pen_units<32, 16>(pr, pp, qb, parked, seg_steps, sm_count, st);
            else if (pp.mh <= 32) e = launch_pen_units<32, 32>(pr, pp, qb, parked, seg_steps, sm_count, st);
            break;
    }
    if (e != cudaSuccess) return e;
    if (times) cudaEventRecord(ev[2], st);
    {
        const long long cap = total / 3 + pr.nbatch + 16;
        const int blocks = (int)((cap + 63) / 64);
        if (pp.small <= 32) pen_small_kernel<32><<<blocks, 64, 0, st>>>(pr);
        else pen_small_kernel<96><<<blocks, 64, 0, st>>>(pr);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if (times) cudaEventRecord(ev[3], st);
    {
        const int smem_elems = max_len <= AMRVW_SORT_SMEM ? (max_len > 1 ? max_len : 1) : AMRVW_SORT_SMEM;
        const size_t smem = (size_t)smem_elems * sizeof(cplx);
        if ((e = cudaFuncSetAttribute(rounds_finish_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        rounds_finish_kernel<double><<<(int)pr.nbatch, 256, smem, st>>>(pr, smem_elems);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    *launches += 5;
    if (times) return read_round_times(ev, times, st);
    return cudaSuccess;
}

}  // namespace amrvw
