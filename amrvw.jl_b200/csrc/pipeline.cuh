// pipeline.cuh -- pipelined (multishift) real-double-shift kernel: the north-star path.
//
// The reference chases ONE bulge at a time down the three rotator chains (bulge-step.jl:36-49); a
// bulge at position i only touches indices i..i+2 of Q, B and Ct, so a second bulge
// can follow three indices behind without ever seeing a half-updated rotator.  This kernel keeps
// L bulges in flight per polynomial, one per lane of a group of L lanes (G = 32/L polynomials per
// warp), as a register-resident systolic array:
//
//        HBM/L2 --(lane 0 loads index start+t)--> [lane 0: window of 3 indices + bulge 0]
//               --shfl_up--> [lane 1: window + bulge 1] --shfl_up--> ... --> [lane L-1] --store--> HBM
//
//   * every lane holds a sliding window of three consecutive rotators of each chain (18 doubles) and
//     its bulge U, V, W (6 doubles) in registers; per step it does the 7 turnovers of one chase
//     step, retires the finished rotator triple to the next lane with warp shuffles and receives the
//     next one.  Each rotator is read from memory once and written once per L bulges: 96 B of
//     traffic per step per group instead of per bulge, no shared-memory staging of the chains, no
//     bank conflicts.
//   * the L bulges of one "fat step" use 2L/reuse shifts = the eigenvalues of the trailing block of
//     the active window.  They come from running the one-bulge algorithm on a copy of the last
//     2L/reuse rotators whose top rotator is made diagonal (so the copy is exactly the trailing
//     principal block), staged in shared memory by the group's lane 0 ("phase A").
//   * windows of at most `small` rotators are decoupled from the rest of the matrix; lane 0 finishes
//     them completely in shared memory with the reference's one-bulge schedule.
//   * phases alternate warp-synchronously: phase A (lane 0 of every group: deflation scan, 2x2/1x1
//     extraction, small windows, shift sub-solve), phase B (all 32 lanes chase).
//
// The operations are the reference's (same turnover/fuse/create_bulge code as solver.cuh); only
// their order differs, so roots agree with the reference to backward-error level, not bit for bit.
// oracle/amrvw_oracle.cpp holds the same schedule executed sequentially (amrvw_algorithm_ms): bulges
// three apart commute exactly, so that is what this kernel is tested against.
#pragma once
#include "kernels.cuh"
#include "shifts.cuh"
#include "../../include/amrvw_b200.h"

namespace amrvw {

struct PipeParams {
    int reuse;   // bulges per shift pair
    int small;   // windows <= small are finished serially in shared memory
    int ms;      // scratch slots per group and chain (>= max(2L/reuse, small) + 3)
    int dense;   // shifts from the dense trailing block (shifts.cuh) instead of the core-chasing sub-solve
    int mh;      // dense: order of the largest trailing block (2L/reuse)
    int stream;  // the CTAs take the fat steps of all polynomials as one stream (stream.cuh)
    int stream_ctas;   // stream: resident CTAs per SM (0 = as many as fit, at most 2)
};

AMRVW_DI Rot<double> shfl_up_rot(const Rot<double> r, int width) {
    Rot<double> o;
    o.c = __shfl_up_sync(0xffffffffu, r.c, 1, width);
    o.s = __shfl_up_sync(0xffffffffu, r.s, 1, width);
    return o;
}
AMRVW_DI Rot<double> ldg_rot(const Rot<double>* p) {
    const double2 v = *reinterpret_cast<const double2*>(p);
    return make_rot<double>(v.x, v.y);
}
AMRVW_DI void stg_rot(Rot<double>* p, const Rot<double> r) { *reinterpret_cast<double2*>(p) = make_double2(r.c, r.s); }
// L2-coherent loads.  In the persistent kernel (rounds.cuh) a polynomial's chains, deflation stack and
// record are written by whichever warp held its unit last, possibly on another SM; L1 is not
// coherent across SMs, so everything that was written during the kernel is read with ld.global.cg
// (and cp.async.cg).  Stores are write-through and need nothing.
AMRVW_DI Rot<double> ldcg_rot(const Rot<double>* p) {
    const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
    return make_rot<double>(v.x, v.y);
}

// 16-byte asynchronous global->shared copy (LDGSTS): lane 0's look-ahead of the next chain index goes
// through shared memory so that it holds no registers across the chase step (with the value kept in
// registers ptxas spilled it, and the spill store waited for the load: 14% of all stall samples).
AMRVW_DI void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
AMRVW_DI void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
AMRVW_DI void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// look-ahead ring of a group's feeding lane: AMRVW_RING slots of three rotators; the triple of step
// t sits in slot t % AMRVW_RING and is requested AMRVW_RING - 1 steps ahead (one step of look-ahead
// left lane 0 waiting on far-L2 / HBM latency: a step is ~1100 cycles).
#define AMRVW_RING 4
AMRVW_DI void cp_async_wait_ring() { asm volatile("cp.async.wait_group %0;" ::"n"(AMRVW_RING - 2) : "memory"); }

#define AMRVW_TURN(a, b, c, o1, o2, o3)      \
    do {                                      \
        Rot<double> t1_, t2_, t3_;            \
        turnover_fast(a, b, c, t1_, t2_, t3_); \
        o1 = t1_; o2 = t2_; o3 = t3_;         \
    } while (0)

// Shared-memory scratch of one group: a window of up to ms rotators of each chain + eigenvalue slots.
struct GroupScratch {
    Rot<double>* q; Rot<double>* b; Rot<double>* c; cplx* e;
    double* H; double* isg;   // dense shifts: mh x mh block, mh + 1 reciprocals
};

// Copy chain indices k..hi+1 of the polynomial into scratch and return a view whose indices are the
// global ones (scratch slot 0 <-> index k).
AMRVW_DI Fact<double, false> stage_window(const Fact<double, false>& F, const GroupScratch& sc, int k, int hi) {
    const int cnt = hi + 1 - k + 1;
    for (int t = 0; t < cnt; ++t) {
        sc.q[t] = ldg_rot(F.Q + k + t);
        sc.b[t] = ldg_rot(F.B + k + t);
        sc.c[t] = ldg_rot(F.Ct + k + t);
    }
    Fact<double, false> W = F;
    W.Q = sc.q - k; W.B = sc.b - k; W.Ct = sc.c - k;
    W.turnovers = 0;
    return W;
}

// The 24 doubles a lane owns during a chase step: its window of Q, B, Ct and its bulge.
struct StepState {
    Rot<double> q0, q1, q2, b0, b1, b2, c0, c1, c2, U, V, W;
};

// General-path versions of the chase step (all special cases of the turnover), out of line: taken
// when the division-free step flags a degenerate turnover, and for the last position of a bulge.
AMRVW_NI void chase_step_general(StepState& z) {
    turnover(z.b1, z.b2, z.V, z.V, z.b1, z.b2);
    turnover(z.c2, z.c1, z.V, z.V, z.c2, z.c1);
    turnover(z.b0, z.b1, z.U, z.U, z.b0, z.b1);
    turnover(z.c1, z.c0, z.U, z.U, z.c1, z.c0);
    turnover(z.q1, z.q2, z.V, z.V, z.q1, z.q2);
    turnover(z.q0, z.q1, z.U, z.U, z.q0, z.q1);
    turnover(z.W, z.V, z.U, z.V, z.U, z.W);
}
// last position (j == stop): passthrough_triu, then fuse the bulge into Q (RDS.jl:110-128)
AMRVW_DI void chase_step_final(StepState& z, double pbot) {
    turnover_fast(z.b1, z.b2, z.V, z.V, z.b1, z.b2);
    turnover_fast(z.c2, z.c1, z.V, z.V, z.c2, z.c1);
    turnover_fast(z.b0, z.b1, z.U, z.U, z.b0, z.b1);
    turnover_fast(z.c1, z.c0, z.U, z.U, z.c1, z.c0);
    z.V.s = sign_(pbot) * z.V.s;
    z.q1 = fuse(z.q1, z.V);
    turnover_fast(z.q0, z.q1, z.U, z.U, z.q0, z.q1);
    z.U = fuse(z.W, z.U);
    turnover_fast(z.b1, z.b2, z.U, z.U, z.b1, z.b2);
    turnover_fast(z.c2, z.c1, z.U, z.U, z.c2, z.c1);
    z.U.s = sign_(pbot) * z.U.s;
    z.q1 = fuse(z.q1, z.U);
}

// One chase step at a regular position (RDS.jl:42-70, 97-108) on registers: the 7 turnovers as ONE
// straight-line block so the independent ones (T2|T3, T4|T5) overlap.  The inputs are parked in the
// thread's shared-memory snapshot slots first; a degenerate turnover (flagged, essentially never)
// redoes the step on the general path from that snapshot.  Used by the systolic lanes and by the
// serial window solver.
#define AMRVW_CHASE_REGULAR(sn, NT)                                                                                              \
    do {                                                                                                                          \
        (sn)[0 * (NT)] = make_double2(q0.c, q0.s); (sn)[1 * (NT)] = make_double2(q1.c, q1.s); (sn)[2 * (NT)] = make_double2(q2.c, q2.s); \
        (sn)[3 * (NT)] = make_double2(b0.c, b0.s); (sn)[4 * (NT)] = make_double2(b1.c, b1.s); (sn)[5 * (NT)] = make_double2(b2.c, b2.s); \
        (sn)[6 * (NT)] = make_double2(c0.c, c0.s); (sn)[7 * (NT)] = make_double2(c1.c, c1.s); (sn)[8 * (NT)] = make_double2(c2.c, c2.s); \
        (sn)[9 * (NT)] = make_double2(U.c, U.s); (sn)[10 * (NT)] = make_double2(V.c, V.s); (sn)[11 * (NT)] = make_double2(W.c, W.s);    \
        bool ok_ = true;                                                                                                          \
        Rot<double> t1_, t2_, t3_;                                                                                                \
        ok_ &= turnover_fast_try(b1, b2, V, t1_, t2_, t3_); V = t1_; b1 = t2_; b2 = t3_;                                          \
        ok_ &= turnover_fast_try(c2, c1, V, t1_, t2_, t3_); V = t1_; c2 = t2_; c1 = t3_;                                          \
        ok_ &= turnover_fast_try(b0, b1, U, t1_, t2_, t3_); U = t1_; b0 = t2_; b1 = t3_;                                          \
        ok_ &= turnover_fast_try(c1, c0, U, t1_, t2_, t3_); U = t1_; c1 = t2_; c0 = t3_;                                          \
        ok_ &= turnover_fast_try(q1, q2, V, t1_, t2_, t3_); V = t1_; q1 = t2_; q2 = t3_;                                          \
        ok_ &= turnover_fast_try(q0, q1, U, t1_, t2_, t3_); U = t1_; q0 = t2_; q1 = t3_;                                          \
        ok_ &= turnover_fast_try(W, V, U, t1_, t2_, t3_); V = t1_; U = t2_; W = t3_;                                              \
        if (!ok_) {                                                                                                               \
            StepState z_;                                                                                                         \
            double2 v_;                                                                                                           \
            v_ = (sn)[0 * (NT)]; z_.q0 = make_rot<double>(v_.x, v_.y); v_ = (sn)[1 * (NT)]; z_.q1 = make_rot<double>(v_.x, v_.y); \
            v_ = (sn)[2 * (NT)]; z_.q2 = make_rot<double>(v_.x, v_.y); v_ = (sn)[3 * (NT)]; z_.b0 = make_rot<double>(v_.x, v_.y); \
            v_ = (sn)[4 * (NT)]; z_.b1 = make_rot<double>(v_.x, v_.y); v_ = (sn)[5 * (NT)]; z_.b2 = make_rot<double>(v_.x, v_.y); \
            v_ = (sn)[6 * (NT)]; z_.c0 = make_rot<double>(v_.x, v_.y); v_ = (sn)[7 * (NT)]; z_.c1 = make_rot<double>(v_.x, v_.y); \
            v_ = (sn)[8 * (NT)]; z_.c2 = make_rot<double>(v_.x, v_.y); v_ = (sn)[9 * (NT)]; z_.U = make_rot<double>(v_.x, v_.y);  \
            v_ = (sn)[10 * (NT)]; z_.V = make_rot<double>(v_.x, v_.y); v_ = (sn)[11 * (NT)]; z_.W = make_rot<double>(v_.x, v_.y); \
            chase_step_general(z_);                                                                                               \
            q0 = z_.q0; q1 = z_.q1; q2 = z_.q2; b0 = z_.b0; b1 = z_.b1; b2 = z_.b2; c0 = z_.c0; c1 = z_.c1; c2 = z_.c2;           \
            U = z_.U; V = z_.V; W = z_.W;                                                                                         \
        }                                                                                                                         \
    } while (0)

#define AMRVW_CHASE_GENERAL()                                                     \
    do {                                                                          \
        StepState z_ = {q0, q1, q2, b0, b1, b2, c0, c1, c2, U, V, W};             \
        chase_step_general(z_);                                                   \
        q0 = z_.q0; q1 = z_.q1; q2 = z_.q2; b0 = z_.b0; b1 = z_.b1; b2 = z_.b2;   \
        c0 = z_.c0; c1 = z_.c1; c2 = z_.c2; U = z_.U; V = z_.V; W = z_.W;         \
    } while (0)

// One whole bulge (bulge-step.jl:11-19 for Float64, rank-one R, no B-only shortcut) chased serially
// through the window [start, stop] with the same register window as the systolic lanes: per position
// three rotators in, three out, instead of six memory round trips per turnover.  Same operations in
// the same order as solver.cuh's bulge_step.
AMRVW_NI void bulge_step_window(Fact<double, false>& F, Counter& ctr, const Shifts* given, uint64_t seed, uint64_t poly, int& ord, PolyStats& st,
                                double2* sn, int NT) {
    st.sweeps++;
    Rot<double> U, V, W;
    create_bulge_rds(F, ctr, U, V, given, seed, poly, ord, st);
    int i = ctr.start_index;
    const int stop = ctr.stop_index;
    const double ptop = sign_(F.Q[i - 1].c);       // Q[0] = (1,0) sentinel
    const double pbot = sign_(F.Q[stop + 1].c);    // Q[N] = (1,0) sentinel
    Rot<double> q0 = F.Q[i], q1 = F.Q[i + 1], b0 = F.B[i], b1 = F.B[i + 1], c0 = F.Ct[i], c1 = F.Ct[i + 1];
    Rot<double> q2 = q1, b2 = b1, c2 = c1;
    {   // absorb_Ut (RDS.jl:14-36)
        Rot<double> w = q0;
        w.s = sign_(ptop) * w.s;
        Rot<double> r1, r2, r3;
        turnover(adjoint(U), adjoint(V), w, r1, r2, r3);
        W = r1;
        q1 = fuse(r3, q1);
        r2.s = sign_(ptop) * r2.s;
        q0 = r2;
        F.turnovers++;
    }
    for (; i <= stop - 2; ++i) {
        q2 = F.Q[i + 2]; b2 = F.B[i + 2]; c2 = F.Ct[i + 2];
#ifndef AMRVW_STRICT
        AMRVW_CHASE_REGULAR(sn, NT);
#else
        AMRVW_CHASE_GENERAL();
#endif
        F.Q[i] = q0; F.B[i] = b0; F.Ct[i] = c0;
        q0 = q1; q1 = q2; b0 = b1; b1 = b2; c0 = c1; c1 = c2;
        F.turnovers += 7;
    }
    {   // last position: i == stop - 1
        b2 = F.B[stop + 1]; c2 = F.Ct[stop + 1];
        StepState z = {q0, q1, q2, b0, b1, b2, c0, c1, c2, U, V, W};
        chase_step_final(z, pbot);
        F.Q[i] = z.q0; F.Q[i + 1] = z.q1;
        F.B[i] = z.b0; F.B[i + 1] = z.b1; F.B[i + 2] = z.b2;
        F.Ct[i] = z.c0; F.Ct[i + 1] = z.c1; F.Ct[i + 2] = z.c2;
        F.turnovers += 7;
    }
}

// create_bulge (create-bulge.jl:20-77) on the lane's window [start-1 .. start+2] + absorb_Ut
// (RDS.jl:14-36), out of line: once per bulge per fat step, keeps the chase loop small.
// Bulge creation with reciprocals instead of divisions.  The bulge rotators only choose WHICH unitary
// similarity is applied (it is applied exactly, by turnovers), so their last bits steer convergence,
// not accuracy; the IEEE divisions of the entry formulas made one creation ~7.7 k cycles during which
// the other 15 lanes of the group wait (12% of the kernel: profiles/).  Same formulas as
// r1_entry / diagonal_block / create_bulge_rds (solver.cuh); the strict build calls those.
AMRVW_DI double r1_entry_fast(const Rot<double>* B, const Rot<double>* Ct, const int j, const int k) {
    const double cj = Ct[j].c;
    double s = Ct[j].s;
    double par = -1.0;
    double tot = w_entry(B, j, k) * rcp_(s);
    for (int i = j + 1; i <= k; ++i) {
        const double w = w_entry(B, i, k);
        s = s * Ct[i].s;
        tot = tot + par * cj * Ct[i].c * rcp_(s) * w;
        par = -par;
    }
    return tot;
}
AMRVW_DI void givensrot_fast(const double a, const double b, double& c, double& s, double& r) {
    const double x = a * a + b * b;
    if (x >= 1e-280 && x <= 1e280) {
        const double ri = rsqrt_nr(x);
        c = a * ri; s = b * ri; r = x * ri;
    } else {
        givensrot(a, b, c, s, r);
    }
}
// lq/lb/lc hold indices start-1 .. start+2 (slot 0 = start-1); shifts e1, e2
AMRVW_DI void create_bulge_fast(const Rot<double>* lq, const Rot<double>* lb, const Rot<double>* lc, const int start, const Shifts sh, Rot<double>& U, Rot<double>& V) {
    const Rot<double>* Q = lq - (start - 1);
    const Rot<double>* B = lb - (start - 1);
    const Rot<double>* Ct = lc - (start - 1);
    const int j = start, i = j - 1, k = j + 1;
    // diagonal block at j (diagonal-block.jl:12-36) and block(j+1).A21
    const Rot<double> qm = Q[j - 1], q0 = Q[j], qp = Q[j + 1];
    const double qji = -qm.s, qjj = qm.c * q0.c, qjk = qp.c * q0.s * qm.c, qkj = -q0.s, qkk = q0.c * qp.c;
    const double rjj = r1_entry_fast(B, Ct, j, j), rjk = r1_entry_fast(B, Ct, j, k), rkk = r1_entry_fast(B, Ct, k, k);
    double rij = 0.0, rik = 0.0;
    if (i >= 1) { rij = r1_entry_fast(B, Ct, i, j); rik = r1_entry_fast(B, Ct, i, k); }
    const double bk11 = qji * rij + qjj * rjj;
    const double bk12 = qji * rik + qjj * rjk + qjk * rkk;
    const double bk21 = qkj * rjj;
    const double bk22 = qkj * rjk + qkk * rkk;
    const double bk32 = (-Q[j + 1].s) * rkk;
    const cplx e1 = sh.e1, e2 = sh.e2;
    const double c1 = -e1.im * e2.im + e1.re * e2.re - e1.re * bk11 - e2.re * bk11 + bk11 * bk11 + bk12 * bk21;
    const double c2 = -e1.re * bk21 - e2.re * bk21 + bk11 * bk21 + bk21 * bk22;
    const double c3 = bk21 * bk32;
    double c, s, nrm, tmp;
    givensrot_fast(c2, c3, c, s, nrm);
    V = make_rot<double>(c, -s);
    givensrot_fast(c1, nrm, c, s, tmp);
    U = make_rot<double>(c, -s);
}

AMRVW_DI void create_and_absorb(StepState& z, const Rot<double> qm, const Rot<double> bm, const Rot<double> cm, const Fact<double, false>& F,
                                int start, int stop, int mode, const Shifts sh, uint64_t seed, uint64_t poly, int ordinal) {
    Rot<double> U, V;
    if (mode == 2) {
        const double ang = uniform01(seed ^ 0xA5A5A5A5DEADBEEFull, poly, (uint64_t)ordinal) * 3.141592653589793;
        double sn, cs;
        sincos(ang, &sn, &cs);
        U = make_rot<double>(sn, cs);
        V = make_rot<double>(sn, -cs);
    } else {
        Rot<double> lq[4] = {qm, z.q0, z.q1, z.q2}, lb[4] = {bm, z.b0, z.b1, z.b2}, lc[4] = {cm, z.c0, z.c1, z.c2};
#ifndef AMRVW_STRICT
        (void)F; (void)stop;
        create_bulge_fast(lq, lb, lc, start, sh, U, V);
#else
        Fact<double, false> Fl = F;
        Fl.Q = lq - (start - 1); Fl.B = lb - (start - 1); Fl.Ct = lc - (start - 1);
        Counter cl = {start - 1, start, stop, AMRVW_IT_COUNT, -1};
        int dummy = 0;
        PolyStats ds = {0, 0, 0, 0, 0};
        create_bulge_rds(Fl, cl, U, V, &sh, seed, poly, dummy, ds);
#endif
    }
    const double ptop = sign_(qm.c);
    Rot<double> w = z.q0;
    w.s = sign_(ptop) * w.s;
    Rot<double> r1, r2, r3;
    turnover_fast(adjoint(U), adjoint(V), w, r1, r2, r3);
    z.W = r1;
    z.q1 = fuse(r3, z.q1);
    r2.s = sign_(ptop) * r2.s;
    z.q0 = r2;
    z.U = U; z.V = V;
}

// create_bulge + absorb_Ut for the lane whose bulge enters the window this step; everything it needs
// beyond the lane's window is read here (the rotators at start-1 never change during a fat step)
AMRVW_NI void create_and_absorb_at(StepState& z, const Fact<double, false>& F, const int start, const int stop, const int mode, const Shifts* shl, const int lane,
                                   const uint64_t seed, const uint64_t poly, const int ordinal) {
    const Rot<double> one = make_rot<double>(1.0, 0.0);
    const Rot<double> qm = ldcg_rot(F.Q + start - 1);                // Q[0] is the (1,0) sentinel
    Rot<double> bm = one, cm = one;
    if (start > 1) { bm = ldcg_rot(F.B + start - 1); cm = ldcg_rot(F.Ct + start - 1); }
    Shifts sh;
    sh.e1 = cplx(0.0, 0.0); sh.e2 = cplx(0.0, 0.0);
    if (mode == 1) {
        const double2 v1 = __ldcg(reinterpret_cast<const double2*>(shl + lane)), v2 = __ldcg(reinterpret_cast<const double2*>(shl + lane) + 1);
        sh.e1 = cplx(v1.x, v1.y); sh.e2 = cplx(v2.x, v2.y);
    }
    create_and_absorb(z, qm, bm, cm, F, start, stop, mode, sh, seed, poly, ordinal);
}
AMRVW_NI void chase_step_final_at(StepState& z, const Fact<double, false>& F, const int stop) {
    chase_step_final(z, sign_(ldcg_rot(F.Q + stop + 1).c));            // Q[N] is the (1,0) sentinel
}

// drive_window's step for the pipelined kernel's serial phases
struct WindowStep {
    double2* sn; int NT;
    AMRVW_DI void operator()(Fact<double, false>& F, Counter& ctr, uint64_t seed, uint64_t poly, int& ord, PolyStats& st) const {
        bulge_step_window(F, ctr, (const Shifts*)nullptr, seed, poly, ord, st, sn, NT);
    }
};

struct LeaderState {
    Counter ctr;
    long long kk, it_max;
    int ord;
    int unconv;
    int sp;        // size of the deflation stack
    long long pc_sub, pc_small;   // profile: cycles in shift sub-solves / small-window solves
    int pn_small;
};

// check_deflation (AMRVW_algorithm.jl:181-202) without the O(window) scan.  The reference scans
// k = stop..start for the first |s_k| <= eps.  Here every deflatable rotator is recorded when the
// chase writes it back (ascending index order, so the top of the stack is the highest one), and
// rotators above the active window never change, so "highest deflatable index <= stop" is the top of
// the stack: O(1) instead of n^2/4 dependent loads per polynomial (20% of the kernel in the ncu
// source view before this).
// diagonal_block (diagonal-block.jl:12-36) at j through a coherent copy of indices j-1..j+1
AMRVW_DI void diagonal_block_cg(const Fact<double, false>& F, const int j, double A[4]) {
    Rot<double> lq[3], lb[3], lc[3];
    for (int t = 0; t < 3; ++t) {
        lq[t] = ldcg_rot(F.Q + j - 1 + t);                         // Q[0], Q[N] are the (1,0) sentinels
        lb[t] = (j - 1 + t >= 1) ? ldcg_rot(F.B + j - 1 + t) : make_rot<double>(1.0, 0.0);
        lc[t] = (j - 1 + t >= 1) ? ldcg_rot(F.Ct + j - 1 + t) : make_rot<double>(1.0, 0.0);
    }
    Fact<double, false> Fl = F;
    Fl.Q = lq - (j - 1); Fl.B = lb - (j - 1); Fl.Ct = lc - (j - 1);
    diagonal_block(Fl, j, A);
}
AMRVW_DI void check_deflation_stack(Fact<double, false>& F, LeaderState& ls, const int* dstack) {
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
// the window below rotator zero_index is finished: drop that boundary, expose the one above it
AMRVW_DI void pop_window(LeaderState& ls) {
    if (ls.ctr.zero_index > 0 && ls.sp > 0) ls.sp -= 1;
}

// Phase A: the driver (AMRVW_algorithm.jl:10-138) up to the next fat step.  Returns 0 when the
// polynomial is finished, 3 for a fat step whose shifts are still to be computed (nb = m, the order of
// the trailing block minus one), 2 for a fat step of nb exceptional (random) bulges.
// pair the m+1 shifts e[0..m] for the real double shift: conjugate pairs first (slot order), then the
// reals two by two; each pair drives `reuse` bulges.  Returns the number of bulges.
AMRVW_DI int pair_shifts(const cplx* e, const int m, Shifts* shl, const int L, const int reuse) {
    int np = 0;
    for (int t = 0; t <= m; ++t) {
        const cplx z = e[t];
        if (z.im != 0.0) {
            if (t + 1 <= m && e[t + 1].im == -z.im && e[t + 1].re == z.re) { shl[np].e1 = z; shl[np].e2 = e[t + 1]; ++np; ++t; }
            else { shl[np].e1 = z; shl[np].e2 = conj_(z); ++np; }
        }
    }
    bool have = false;
    cplx pend;
    for (int t = 0; t <= m; ++t) {
        const cplx z = e[t];
        if (z.im == 0.0) {
            if (have) { shl[np].e1 = pend; shl[np].e2 = z; ++np; have = false; }
            else { pend = z; have = true; }
        }
    }
    if (have) { shl[np].e1 = pend; shl[np].e2 = pend; ++np; }
    if (np * reuse > L) np = L / reuse;
    for (int rep = 1; rep < reuse; ++rep)
        for (int b = 0; b < np; ++b) shl[rep * np + b] = shl[b];
    return np * reuse;
}

// shifts by the core-chasing iteration on the staged copy W of the trailing block (window
// [k0+1, stop], Q[k0] already diagonal): the fall-back of the dense solver, and AMRVW_SHIFTS_CHASE
AMRVW_NI long long chase_shifts(Fact<double, false>& W, const int k0, const int stop, cplx* e, const uint64_t seed, const uint64_t poly, int& ord, double2* sn, const int NT) {
    const WindowStep wstep = {sn, NT};
    const int m = stop - k0;
    cplx* Es = e - (k0 + 1);
    for (int t = 0; t <= m; ++t) e[t] = cplx(0.0, 0.0);
    Counter c2 = {k0, k0 + 1, stop, AMRVW_IT_COUNT, -1};
    PolyStats sub = {0, 0, 0, 0, 0};
    drive_window(W, k0 + 1, stop, c2, Es, AMRVW_IT_MAX(m), seed, poly, ord, sub, wstep);
    return W.turnovers;
}

// Decoupled small windows are not solved on the spot: they are queued for the small-window kernel
// (rounds.cuh), one thread each.
struct SmallWinRec { int poly, lo, hi, it_count; };
struct SmallSink {
    void* list; unsigned* count; int poly;
    AMRVW_DI void push(int lo, int hi, int it_count) const {
        const unsigned k = atomicAdd(count, 1u);
        SmallWinRec r = {poly, lo, hi, it_count};
        reinterpret_cast<SmallWinRec*>(list)[k] = r;
    }
};

AMRVW_DI int phase_a(Fact<double, false>& F, LeaderState& ls, cplx* E, const GroupScratch& sc, Shifts* shl, const int L, const PipeParams pp,
                     int& nb, const uint64_t seed, const uint64_t poly, PolyStats& st, const int* dstack, double2* sn, const int NT, const SmallSink& sink) {
    (void)sc; (void)shl; (void)seed; (void)poly; (void)sn; (void)NT;
    Counter& ctr = ls.ctr;
    double A[4];
    while (ls.kk <= ls.it_max) {
        if (ctr.stop_index < 1) return 0;
        check_deflation_stack(F, ls, dstack);
        ++ls.kk;
        const int k = ctr.stop_index;
        const int delta = ctr.stop_index - ctr.zero_index;
        if (delta == 1) {
            diagonal_block_cg(F, k, A);
            cplx e1, e2;
            eig2x2(A, e1, e2);
            E[k] = e1; E[k + 1] = e2;
            if (ctr.stop_index == 2) { diagonal_block_cg(F, 1, A); E[1] = to_c(A[0]); ctr.stop_index = 0; return 0; }
            pop_window(ls);
            ctr.stop_index = ctr.zero_index - 1; ctr.zero_index = 0; ctr.start_index = 1;
            check_deflation_stack(F, ls, dstack);
            if (ctr.stop_index < 1) return 0;
        } else if (delta <= 0) {
            diagonal_block_cg(F, k, A);
            if (ctr.stop_index == 1) { E[1] = to_c(A[0]); E[2] = to_c(A[3]); ctr.stop_index = 0; return 0; }
            E[k + 1] = to_c(A[3]);
            pop_window(ls);
            ctr.stop_index = ctr.zero_index - 1;
            if (ctr.stop_index < 1) return 0;
            ctr.zero_index = 0; ctr.start_index = 1;
            check_deflation_stack(F, ls, dstack);
        } else if (delta <= pp.small) {
            // decoupled small window [lo2, hi2]: nothing above it ever touches its rotators again, so it
            // is queued and finished later with the reference's one-bulge schedule (rounds_small_kernel)
            const int lo2 = ctr.start_index, hi2 = ctr.stop_index;
            sink.push(lo2, hi2, ctr.it_count);
            if (lo2 == 2) { diagonal_block_cg(F, 1, A); E[1] = to_c(A[0]); }
            pop_window(ls);
            ctr.stop_index = lo2 - 2; ctr.zero_index = 0; ctr.start_index = 1; ctr.it_count = AMRVW_IT_COUNT;
            if (ctr.stop_index >= 1) check_deflation_stack(F, ls, dstack);
            ls.pn_small += 1;
        } else if (ctr.it_count == 0) {
            st.fat_steps++;
            nb = L;
            return 2;
        } else {
            // fat step with 2L/reuse shifts = eigenvalues of the trailing (m+1) x (m+1) block: computed by
            // the shift kernel (rounds.cuh), one warp per polynomial
            int m = min(2 * L / pp.reuse - 1, delta - 1);
            if (m % 2 == 0) m -= 1;
            nb = m;            // handed to the shift kernel through the record
            st.fat_steps++;
            ctr.it_count -= 1;
            return 3;
        }
    }
    return 0;
}


// ---------------------------------------------------------------- the lean systolic loop
// Steps [t0, t1) of phase B during which EVERY lane of the warp is at a regular chase position (or
// belongs to a group with nothing to do): no bulge is being created or fused, so the step is
// ingest -> 7 turnovers -> retire with no per-lane case analysis.  Out of line on purpose: inside
// the kernel body the same loop carried ~300 register moves and spill instructions per step for the
// cold state around it (649 instructions per step against ~330 here; tools/ubench.cu measures the
// loop in isolation).  Shared memory is addressed through 32-bit shared-window addresses so that
// the snapshot and the look-ahead ring compile to STS/LDS/LDGSTS without generic-address glue.
AMRVW_DI void sts_rot(unsigned sa, const Rot<double> r) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(sa), "d"(r.c), "d"(r.s) : "memory"); }
AMRVW_DI Rot<double> lds_rot(unsigned sa) {
    Rot<double> r;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.c), "=d"(r.s) : "r"(sa) : "memory");
    return r;
}
AMRVW_DI void cp_async16_sa(unsigned sa, const void* gsrc) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory"); }

// role: bit 0: the lane's group is chasing, bit 1: lane 0 of its group (feeds), bit 2: lane L-1 (retires)
template <int L, int NT>
AMRVW_DI int lean_chase(StepState& z, Rot<double>* gq, Rot<double>* gb, Rot<double>* gc, const unsigned stage_sa, const unsigned snap_sa, int* dstack, int sp,
                        const int start, const int t0, const int t1, const int role) {
    Rot<double> q0 = z.q0, q1 = z.q1, q2 = z.q2, b0 = z.b0, b1 = z.b1, b2 = z.b2, c0 = z.c0, c1 = z.c1, c2 = z.c2, U = z.U, V = z.V, W = z.W;
    const bool chasing = role & 1, feeds = (role & 3) == 3, retires = (role & 5) == 5;
    [[maybe_unused]] constexpr unsigned SS = NT * 16;     // snapshot slot stride in bytes
    // feeding lane: next index to prefetch; retiring lane: index leaving the pipeline
    gq += start; gb += start; gc += start;
    const int kofs = feeds ? AMRVW_RING - 1 : -(3 * (L - 1) + 2);
    gq += t0 + kofs; gb += t0 + kofs; gc += t0 + kofs;
    int kidx = start + t0 + kofs;
#pragma unroll 1
    for (int t = t0; t < t1; ++t) {
        Rot<double> iq = shfl_up_rot(q0, L), ib = shfl_up_rot(b0, L), ic = shfl_up_rot(c0, L);
        if (feeds) {
            cp_async_wait_ring();
            const unsigned sg = stage_sa + (t & (AMRVW_RING - 1)) * 48, sn2 = stage_sa + ((t + AMRVW_RING - 1) & (AMRVW_RING - 1)) * 48;
            iq = lds_rot(sg); ib = lds_rot(sg + 16); ic = lds_rot(sg + 32);
            cp_async16_sa(sn2, gq); cp_async16_sa(sn2 + 16, gb); cp_async16_sa(sn2 + 32, gc);
            cp_async_commit();
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
        if (!ok && chasing) {   // a degenerate turnover (essentially never): redo the step on the general path
            StepState y;
            y.q0 = lds_rot(snap_sa + 0 * SS); y.q1 = lds_rot(snap_sa + 1 * SS); y.q2 = lds_rot(snap_sa + 2 * SS);
            y.b0 = lds_rot(snap_sa + 3 * SS); y.b1 = lds_rot(snap_sa + 4 * SS); y.b2 = lds_rot(snap_sa + 5 * SS);
            y.c0 = lds_rot(snap_sa + 6 * SS); y.c1 = lds_rot(snap_sa + 7 * SS); y.c2 = lds_rot(snap_sa + 8 * SS);
            y.U = lds_rot(snap_sa + 9 * SS); y.V = lds_rot(snap_sa + 10 * SS); y.W = lds_rot(snap_sa + 11 * SS);
            chase_step_general(y);
            q0 = y.q0; q1 = y.q1; q2 = y.q2; b0 = y.b0; b1 = y.b1; b2 = y.b2; c0 = y.c0; c1 = y.c1; c2 = y.c2; U = y.U; V = y.V; W = y.W;
        }
#else
        if (chasing) AMRVW_CHASE_GENERAL();
#endif
        if (retires) {
            stg_rot(gq, q0); stg_rot(gb, b0); stg_rot(gc, c0);
            if (fabs(q0.s) <= AMRVW_EPS) dstack[sp++] = kidx;
        }
        gq += 1; gb += 1; gc += 1; kidx += 1;
    }
    z.q0 = q0; z.q1 = q1; z.q2 = q2; z.b0 = b0; z.b1 = b1; z.b2 = b2; z.c0 = c0; z.c1 = c1; z.c2 = c2; z.U = U; z.V = V; z.W = W;
    return sp;
}

}  // namespace amrvw
