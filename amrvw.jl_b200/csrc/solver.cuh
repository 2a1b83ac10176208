// solver.cuh -- one polynomial's Francis iteration on the rotator-factored companion matrix,
// as device functions: factorization set-up, matrix entries, bulge creation, the chase, deflation,
// 2x2 extraction, sort.  Used (a) by the one-bulge-at-a-time kernels (reference schedule) and
// (b) as the serial phases (shift sub-solves, small windows) of the pipelined kernel.
//
// Reference counterparts (file:line under /root/reference/src):
//   companion.jl:6-24,64-137,150-204   basic_decompose / q_factorization / r_factorization / amrvw
//   chains.jl:109-158, qfactorization.jl:67-70, rfactorization.jl:83-97,139-170, pencil.jl:29-45
//                                      matrix entries  -> closed forms below (q_block / r1_entry / r_entry)
//   diagonal-block.jl:12-36            diagonal_block!
//   create-bulge.jl:20-118             create_bulge (RDS / CSS)
//   RDS.jl, CSS.jl, bulge-step.jl      absorb_Ut / passthrough_triu / passthrough_Q / absorb_U / deflate
//   chains.jl:360-400, rfactorization.jl:198-216, pencil.jl:63-80, diagonal.jl:84-96,199-225
//   AMRVW_algorithm.jl:10-138,181-202  driver + check_deflation
//
// Data layout (HBM): per polynomial, every chain is a contiguous array of Rot<S> indexed by ROTATOR
// INDEX (slot k holds the rotator acting on rows k,k+1).  Q carries two sentinels, Q[0] = Q[N] =
// (1,0), so the Hessenberg entry formulas need no edge branches (c_0 = c_N = 1, parity(N) = +1).
#pragma once
#include "rotators.cuh"

namespace amrvw {

// The reference's compile-time constants, runtime options here (amrvw_options): tol = eps
// (AMRVW_algorithm.jl:184), IT_COUNT = 15 (create-bulge.jl:9), it_max = 20 n (AMRVW_algorithm.jl:25).
// They live in constant memory (a c[bank][offset] operand costs what an immediate costs); the host
// writes them on the call's stream before its first launch.
struct AlgoConst { double tol; int it_count; int it_max_factor; };
__constant__ AlgoConst g_algo = {2.220446049250313e-16, 15, 20};
#define AMRVW_EPS (amrvw::g_algo.tol)
#define AMRVW_IT_COUNT (amrvw::g_algo.it_count)
#define AMRVW_IT_MAX(n1) ((long long)amrvw::g_algo.it_max_factor * (long long)(n1))

struct PolyStats {            // per polynomial, optional
    long long turnovers;      // executed turnovers (all schedules)
    int sweeps;               // bulges chased
    int exceptional;          // random shifts taken
    int unconverged;          // eigenvalue slots left at 0 when the trip cap was hit
    int fat_steps;            // pipelined multi-bulge steps (multishift schedule)
};

template <class S, bool PENCIL> struct Fact {
    Rot<S>* Q;    // descending chain, slots 0..N (1..N-1 live)
    Rot<S>* B;    // R = V factor: descending chain B, slots 1..N
    Rot<S>* Ct;   //               ascending chain Ct, slots 1..N (stored by index, not reversed)
    Rot<S>* WB;   // pencil: W factor
    Rot<S>* WCt;
    S* DQ;        // complex only: slots 0..N (DQ[0] = 1 sentinel)
    S* DR;        // complex only: slots 1..N+1
    S* WD;        // complex pencil only
    int N;        // degree of the (trimmed) polynomial
    long long turnovers;
};

template <class S> AMRVW_DI S dget(const S* D, int k) {
    if constexpr (is_real<S>::value) return 1.0; else return D[k];
}

struct Counter { int zero_index, start_index, stop_index, it_count, tr; };

// ---------------------------------------------------------------- entries
// Q-side entries of rows j, j+1 (qfactorization.jl:67-70 over chains.jl:109-158), closed form with
// the sentinels: Q[j,j-1] = -s_{j-1}, Q[j,j] = conj(c_{j-1}) c_j, Q[j,j+1] = c_{j+1} s_j conj(c_{j-1}).
template <class S, bool P> AMRVW_DI void q_block(const Fact<S, P>& F, int j, S& qji, S& qjj, S& qjk, S& qkj, S& qkk) {
    const Rot<S> qm = F.Q[j - 1], q0 = F.Q[j], qp = F.Q[j + 1];
    const S dm = dget(F.DQ, j - 1), d0 = dget(F.DQ, j), dp = dget(F.DQ, j + 1);
    qji = S(-qm.s) * dm;
    qjj = conj_(qm.c) * q0.c * d0;
    qjk = qp.c * q0.s * conj_(qm.c) * dp;
    qkj = S(-q0.s) * d0;
    qkk = conj_(q0.c) * qp.c * dp;
}
// w(j,k) of rfactorization.jl:83-97 for k - j <= 2
template <class S> AMRVW_DI S w_entry(const Rot<S>* B, int j, int k) {
    if (j == k) return S(-B[k].s);
    double s = 1.0;
    for (int i = j + 1; i <= k - 1; ++i) s *= B[i].s;
    return conj_(B[j].c) * s * B[k].c;
}
// R[j,k] of the rank-one factor, rfactorization.jl:139-170 (j <= k <= j+2 on this path)
template <class S> AMRVW_DI S r1_entry(const Rot<S>* B, const Rot<S>* Ct, const S* D, int j, int k) {
    const S cj = Ct[j].c;
    double s = Ct[j].s;
    double par = -1.0;
    const S dk = dget(D, k);
    S tot = (w_entry(B, j, k) * dk) / s;
    for (int i = j + 1; i <= k; ++i) {
        const S w = w_entry(B, i, k);
        s = s * Ct[i].s;
        tot = tot + par * cj * conj_(Ct[i].c) / s * w * dk;
        par = -par;
    }
    return tot;
}
// R[l,k]: rank one, or V W^{-1} for the pencil (pencil.jl:29-45)
template <class S, bool P> AMRVW_DI S r_entry(const Fact<S, P>& F, int l, int k) {
    if constexpr (!P) {
        return r1_entry(F.B, F.Ct, F.DR, l, k);
    } else {
        const int d = k - l;
        if (d == 0) return r1_entry(F.B, F.Ct, F.DR, k, k) / r1_entry(F.WB, F.WCt, F.WD, k, k);
        if (d == 1) {
            const int j = l;
            const S wkk = r1_entry(F.WB, F.WCt, F.WD, k, k);
            return -r1_entry(F.B, F.Ct, F.DR, j, j) * r1_entry(F.WB, F.WCt, F.WD, j, k) / (r1_entry(F.WB, F.WCt, F.WD, j, j) * wkk) +
                   r1_entry(F.B, F.Ct, F.DR, j, k) / wkk;
        }
        const int i = l, j = l + 1;
        const S wii = r1_entry(F.WB, F.WCt, F.WD, i, i), wij = r1_entry(F.WB, F.WCt, F.WD, i, j), wik = r1_entry(F.WB, F.WCt, F.WD, i, k);
        const S wjj = r1_entry(F.WB, F.WCt, F.WD, j, j), wjk = r1_entry(F.WB, F.WCt, F.WD, j, k), wkk = r1_entry(F.WB, F.WCt, F.WD, k, k);
        return r1_entry(F.B, F.Ct, F.DR, i, i) * (wij * wjk - wik * wjj) / (wii * wjj * wkk) -
               (r1_entry(F.B, F.Ct, F.DR, i, j) * wjk) / (wjj * wkk) + r1_entry(F.B, F.Ct, F.DR, i, k) / wkk;
    }
}
// A = (Q R)[j:j+1, j:j+1]   (diagonal-block.jl:12-36)
template <class S, bool P> AMRVW_NI void diagonal_block(const Fact<S, P>& F, int j, S A[4]) {
    S qji, qjj, qjk, qkj, qkk;
    q_block(F, j, qji, qjj, qjk, qkj, qkk);
    const int i = j - 1, k = j + 1;
    const S rjj = r_entry(F, j, j), rjk = r_entry(F, j, k), rkk = r_entry(F, k, k);
    S rij = S(0.0), rik = S(0.0);
    if (i >= 1) { rij = r_entry(F, i, j); rik = r_entry(F, i, k); }
    A[0] = qji * rij + qjj * rjj;
    A[1] = qji * rik + qjj * rjk + qjk * rkk;
    A[2] = qkj * rjj;
    A[3] = qkj * rjk + qkk * rkk;
}

// ---------------------------------------------------------------- passthroughs (one turnover each)
template <class S, bool P> AMRVW_DI void pass_desc(Fact<S, P>& F, Rot<S>* X, Rot<S>& U, int& i) {  // chains.jl:360
    Rot<S> r1, r2, r3;
    turnover(X[i], X[i + 1], U, r1, r2, r3);
    U = r1; X[i] = r2; X[i + 1] = r3; i += 1; F.turnovers++;
}
template <class S, bool P> AMRVW_DI void pass_asc(Fact<S, P>& F, Rot<S>* X, Rot<S>& U, int& i) {  // chains.jl:370
    Rot<S> r1, r2, r3;
    turnover(X[i], X[i - 1], U, r1, r2, r3);
    U = r1; X[i] = r2; X[i - 1] = r3; i -= 1; F.turnovers++;
}
template <class S, bool P> AMRVW_DI void pass_desc_lr(Fact<S, P>& F, Rot<S>& U, Rot<S>* X, int& i) {  // chains.jl:381
    Rot<S> r1, r2, r3;
    turnover(U, X[i - 1], X[i], r1, r2, r3);
    X[i - 1] = r1; X[i] = r2; U = r3; i -= 1; F.turnovers++;
}
template <class S, bool P> AMRVW_DI void pass_asc_lr(Fact<S, P>& F, Rot<S>& U, Rot<S>* X, int& i) {  // chains.jl:391
    Rot<S> r1, r2, r3;
    turnover(U, X[i + 1], X[i], r1, r2, r3);
    X[i + 1] = r1; X[i] = r2; U = r3; i += 1; F.turnovers++;
}
template <class S> AMRVW_DI void pass_diag(S* D, Rot<S>& U, int i) {  // diagonal.jl:84-96; real: no-op
    if constexpr (!is_real<S>::value) {
        const S alpha = D[i], beta = D[i + 1];
        D[i] = beta; D[i + 1] = alpha;
        U.c = (alpha * conj_(beta)) * U.c;
    }
}
// passthrough!(RF, U): right -> left through the upper-triangular factor; U comes back at index i.
template <class S, bool P> AMRVW_DI void pass_R(Fact<S, P>& F, Rot<S>& U, int i) {
    if constexpr (P) {  // pencil.jl:63-71: U' left->right through W, then U through V
        Rot<S> Ut = adjoint(U);
        pass_asc_lr(F, Ut, F.WCt, i);
        pass_desc_lr(F, Ut, F.WB, i);
        pass_diag(F.WD, Ut, i);
        U = adjoint(Ut);
    }
    pass_diag(F.DR, U, i);     // rfactorization.jl:198-206
    pass_desc(F, F.B, U, i);
    pass_asc(F, F.Ct, U, i);
}
// passthrough!(QF, U)  (qfactorization.jl:86-91): U@i leaves at i+1
template <class S, bool P> AMRVW_DI void pass_Q(Fact<S, P>& F, Rot<S>& U, int& i) {
    pass_diag(F.DQ, U, i);
    pass_desc(F, F.Q, U, i);
}
// phase push-down (diagonal.jl:199-225): Diag(alpha)@i moves down Q into D_Q
template <bool P> AMRVW_NI void push_phase(Fact<cplx, P>& F, cplx alpha, int i) {
    const int M = F.N - 1;
    F.DQ[i] = F.DQ[i] * alpha;
    const cplx ca = conj_(alpha);
    while (i < M) {
        Rot<cplx> q = F.Q[i + 1];
        if (q.s == 0.0) break;
        q.c = ca * q.c;
        F.Q[i + 1] = q;
        ++i;
    }
    F.DQ[i + 1] = F.DQ[i + 1] * ca;
}

// ---------------------------------------------------------------- bulge step, real double shift
struct Shifts { cplx e1, e2; };

// create_bulge (create-bulge.jl:20-77).  given != nullptr: use these shifts instead of the
// eigenvalues of block(stop) (multishift schedule).
template <bool P>
AMRVW_NI void create_bulge_rds(Fact<double, P>& F, Counter& ctr, Rot<double>& U, Rot<double>& V, const Shifts* given,
                               uint64_t seed, uint64_t poly, int& rng_ordinal, PolyStats& st) {
    const int i = ctr.start_index;
    if (ctr.it_count == 0 && !given) {
        ctr.it_count = AMRVW_IT_COUNT;
        st.exceptional++;
        const double t = uniform01(seed ^ 0xA5A5A5A5DEADBEEFull, poly, (uint64_t)rng_ordinal++) * 3.141592653589793;
        double sn, cs;
        sincos(t, &sn, &cs);
        U = make_rot<double>(sn, cs);
        V = make_rot<double>(sn, -cs);
        return;
    }
    double A[4];
    cplx e1, e2;
    if (given) { e1 = given->e1; e2 = given->e2; }
    else { diagonal_block(F, ctr.stop_index, A); eig2x2(A, e1, e2); }
    diagonal_block(F, i, A);
    const double bk11 = A[0], bk12 = A[1], bk21 = A[2], bk22 = A[3];
    // block(i+1).A21 = Q[i+2,i+1] * R[i+1,i+1]
    const double bk32 = (-F.Q[i + 1].s) * r_entry(F, i + 1, i + 1);
    const double c1 = -e1.im * e2.im + e1.re * e2.re - e1.re * bk11 - e2.re * bk11 + bk11 * bk11 + bk12 * bk21;
    const double c2 = -e1.re * bk21 - e2.re * bk21 + bk11 * bk21 + bk21 * bk22;
    const double c3 = bk21 * bk32;
    double c, s, nrm, tmp;
    givensrot(c2, c3, c, s, nrm);
    V = make_rot<double>(c, -s);
    givensrot(c1, nrm, c, s, tmp);
    U = make_rot<double>(c, -s);
}

template <bool P>
AMRVW_NI void bulge_step(Fact<double, P>& F, Counter& ctr, const Shifts* given, uint64_t seed, uint64_t poly, int& rng_ordinal, PolyStats& st) {
    st.sweeps++;
    Rot<double> U, V, W;
    create_bulge_rds(F, ctr, U, V, given, seed, poly, rng_ordinal, st);
    int i = ctr.start_index;
    const int nq = F.N - 1;
    {   // absorb_Ut (RDS.jl:14-36)
        const double p = i == 1 ? 1.0 : sign_(F.Q[i - 1].c);
        Rot<double> w = F.Q[i];
        w.s = sign_(p) * w.s;
        Rot<double> r1, r2, r3;
        turnover(adjoint(U), adjoint(V), w, r1, r2, r3);
        F.turnovers++;
        W = r1;
        F.Q[i + 1] = fuse(r3, F.Q[i + 1]);
        r2.s = sign_(p) * r2.s;
        F.Q[i] = r2;
    }
    for (;;) {  // absorb_U (bulge-step.jl:36-49)
        const int j = i + 1;
        // passthrough_triu (RDS.jl:42-70); the B-only shortcut of RDS.jl:73-88 while j <= tr
        if (!P && j <= ctr.tr) {
            Rot<double> v = V, u = U;
            int jv = j, iu = i;
            pass_desc(F, F.B, v, jv);
            pass_desc(F, F.B, u, iu);
            for (int k = -1; k <= 1; ++k) F.Ct[j + k] = adjoint(F.B[j + k]);
        } else {
            pass_R(F, V, j);
            pass_R(F, U, i);
        }
        ctr.tr -= 2;
        // passthrough_Q (RDS.jl:97-131)
        if (j < ctr.stop_index) {
            int jv = j, iu = i;
            pass_Q(F, V, jv);
            pass_Q(F, U, iu);
            Rot<double> r1, r2, r3;
            turnover(W, V, U, r1, r2, r3);
            F.turnovers++;
            V = r1; U = r2; W = r3;
            i += 1;
        } else {
            const double p = (j + 1 > nq) ? 1.0 : sign_(F.Q[j + 1].c);
            V.s = sign_(p) * V.s;
            F.Q[j] = fuse(F.Q[j], V);
            int iu = i;
            pass_desc(F, F.Q, U, iu);      // U now at j
            U = fuse(W, U);
            pass_R(F, U, iu);
            U.s = sign_(p) * U.s;
            F.Q[iu] = fuse(F.Q[iu], U);
            break;
        }
    }
}
template <bool P> AMRVW_DI void deflate(Fact<double, P>& F, int k) { F.Q[k] = make_rot<double>(sign_(F.Q[k].c), 0.0); }  // RDS.jl:134-141

// ---------------------------------------------------------------- bulge step, complex single shift
template <bool P>
AMRVW_NI void bulge_step(Fact<cplx, P>& F, Counter& ctr, const Shifts* given, uint64_t seed, uint64_t poly, int& rng_ordinal, PolyStats& st) {
    st.sweeps++;
    cplx A[4];
    cplx shift;
    if (ctr.it_count == 0 && !given) {  // create-bulge.jl:85-93
        ctr.it_count = AMRVW_IT_COUNT;
        st.exceptional++;
        const double t = uniform01(seed ^ 0xA5A5A5A5DEADBEEFull, poly, (uint64_t)rng_ordinal++) * 3.141592653589793;
        double sn, cs;
        sincos(t, &sn, &cs);
        shift = cplx(cs, sn);
    } else if (given) {
        shift = given->e1;
    } else {  // Wilkinson shift (create-bulge.jl:95-103)
        diagonal_block(F, ctr.stop_index, A);
        cplx e1, e2;
        eig2x2(A, e1, e2);
        shift = abs_(A[3] - e1) < abs_(A[3] - e2) ? e1 : e2;
    }
    int i = ctr.start_index;
    diagonal_block(F, i, A);
    cplx c, nrm; double sn;
    givensrot(A[0] - shift, A[2], c, sn, nrm);
    const Rot<cplx> R = make_rot<cplx>(c, sn);
    Rot<cplx> U = adjoint(R);
    {   // absorb_Ut (CSS.jl:8-14): fuse on the left, then push the phase down Q
        cplx dph;
        F.Q[i] = fuse(R, F.Q[i], dph);
        push_phase(F, dph, i);
    }
    for (;;) {
        // passthrough_triu (CSS.jl:17-41) with the B-only shortcut (CSS.jl:45-61) while i < tr
        if (!P && i < ctr.tr) {
            Rot<cplx> u = U;
            int iu = i;
            pass_desc(F, F.B, u, iu);
            F.Ct[i] = adjoint(F.B[i]);
            F.Ct[i + 1] = adjoint(F.B[i + 1]);
        } else {
            pass_R(F, U, i);
        }
        ctr.tr -= 1;
        // passthrough_Q (CSS.jl:64-88)
        if (i < ctr.stop_index) {
            pass_Q(F, U, i);
        } else {
            pass_diag(F.DQ, U, i);
            cplx dph;
            F.Q[i] = fuse(F.Q[i], U, dph);
            F.DQ[i] = F.DQ[i] * dph;
            F.DQ[i + 1] = F.DQ[i + 1] * conj_(dph);
            break;
        }
    }
}
template <bool P> AMRVW_DI void deflate(Fact<cplx, P>& F, int k) {  // CSS.jl:98-122
    const cplx alpha = F.Q[k].c;
    F.Q[k] = make_rot<cplx>(cplx(1.0, 0.0), 0.0);
    push_phase(F, alpha, k);
}

// ---------------------------------------------------------------- driver

template <class S, bool P> AMRVW_NI void check_deflation(Fact<S, P>& F, Counter& ctr) {  // AMRVW_algorithm.jl:181-202
    for (int k = ctr.stop_index; k >= ctr.start_index; --k) {
        if (fabs(F.Q[k].s) <= AMRVW_EPS) {
            deflate(F, k);
            ctr.zero_index = k;
            ctr.start_index = k + 1;
            ctr.it_count = AMRVW_IT_COUNT;
            return;
        }
    }
}
AMRVW_DI cplx to_c(double x) { return cplx(x, 0.0); }
AMRVW_DI cplx to_c(cplx x) { return x; }

// What the driver does with the window once it is neither a 1x1 nor a 2x2 block.
struct ReferenceStep {  // one bulge, shifts from the trailing 2x2 block: the reference's schedule
    template <class S, bool P>
    AMRVW_DI void operator()(Fact<S, P>& F, Counter& ctr, uint64_t seed, uint64_t poly, int& ord, PolyStats& st) const {
        bulge_step(F, ctr, (const Shifts*)nullptr, seed, poly, ord, st);
    }
};

// AMRVW_algorithm.jl:10-138 on eigenvalue slots lo..hi+1 (Q[lo-1] diagonal or lo == 1); E is the
// 1-based eigenvalue array (E[k] = slot k).  lo = 1, hi = N-1 is the reference driver.  Returns
// the number of slots left unwritten (trip cap hit).
template <class S, bool P, class Step>
AMRVW_DI int drive_window(Fact<S, P>& F, int lo, int hi, Counter& ctr, cplx* E, long long it_max,
                          uint64_t seed, uint64_t poly, int& ord, PolyStats& st, const Step& step) {
    (void)hi;
    long long kk = 0;
    S A[4];
    check_deflation(F, ctr);
    while (kk <= it_max) {
        if (ctr.stop_index < lo) return 0;
        check_deflation(F, ctr);
        ++kk;
        const int k = ctr.stop_index;
        const int delta = ctr.stop_index - ctr.zero_index;
        if (delta == 1) {
            diagonal_block(F, k, A);
            cplx e1, e2;
            eig2x2(A, e1, e2);
            E[k] = e1; E[k + 1] = e2;
            if (ctr.stop_index == lo + 1) {
                diagonal_block(F, lo, A);
                E[lo] = to_c(A[0]);
                ctr.stop_index = lo - 1;
                return 0;
            }
            ctr.stop_index = ctr.zero_index - 1;
            ctr.zero_index = lo - 1;
            ctr.start_index = lo;
            check_deflation(F, ctr);
            if (ctr.stop_index < lo) return 0;
        } else if (delta <= 0) {
            diagonal_block(F, k, A);
            if (ctr.stop_index == lo) {
                E[lo] = to_c(A[0]); E[lo + 1] = to_c(A[3]);
                ctr.stop_index = lo - 1;
                return 0;
            }
            E[k + 1] = to_c(A[3]);
            ctr.stop_index = ctr.zero_index - 1;
            if (ctr.stop_index < lo) return 0;
            ctr.zero_index = lo - 1;
            ctr.start_index = lo;
            check_deflation(F, ctr);
        } else {
            step(F, ctr, seed, poly, ord, st);
            ctr.it_count -= 1;
        }
    }
    return ctr.stop_index < lo ? 0 : ctr.stop_index - lo + 2;
}

// ---------------------------------------------------------------- sort (LinearAlgebra.sorteig!: by (re, im), isless order)
AMRVW_DI bool isless_(double a, double b) {
    if (a != a) return false;
    if (b != b) return true;
    return a < b || (a == b && signbit(a) && !signbit(b));
}
AMRVW_DI bool isequal_(double a, double b) { return (a == b && signbit(a) == signbit(b)) || (a != a && b != b); }
AMRVW_DI bool eig_less(cplx a, cplx b) { return isless_(a.re, b.re) || (isequal_(a.re, b.re) && isless_(a.im, b.im)); }

// in-place heap sort of E[0..n): O(n log n) compares, no extra memory, 0.1% of the flops
AMRVW_NI void sort_roots(cplx* E, int n) {
    auto sift = [&](int start, int end) {
        int root = start;
        const cplx v = E[root];
        for (;;) {
            int child = 2 * root + 1;
            if (child > end) break;
            if (child + 1 <= end && eig_less(E[child], E[child + 1])) child++;
            if (eig_less(v, E[child])) { E[root] = E[child]; root = child; } else break;
        }
        E[root] = v;
    };
    for (int start = (n - 2) / 2; start >= 0; --start) sift(start, n - 1);
    for (int end = n - 1; end > 0; --end) {
        const cplx t = E[end]; E[end] = E[0]; E[0] = t;
        sift(0, end - 1);
    }
}

// ---------------------------------------------------------------- companion set-up (companion.jl)
// xs[i] of basic_decompose (companion.jl:6-24), 1-based i in 1..n+1, from trimmed coefficients ps[0..n]
template <class S> AMRVW_DI S decompose_entry(const S* ps, int n, int i, S pNi) {
    const double par = (n % 2 == 0) ? 1.0 : -1.0;
    if (i == n + 1) return S(-1.0);
    if (i == n) return par * ps[0] * pNi;
    const double sg = ((i + 1) % 2 == 0) ? 1.0 : -1.0;
    return (par * sg) * ps[i] * pNi;
}
AMRVW_DI double inv_(double x) { return 1.0 / x; }
AMRVW_DI cplx inv_(cplx x) { return cplx(1.0, 0.0) / x; }

// r_factorization (companion.jl:79-93 real, :97-120 complex) for xs given by a functor x(i), i in 1..N+1
template <class XS> AMRVW_DI void r_factorization(Rot<double>* B, Rot<double>* Ct, double* /*D*/, int N, const XS& x) {
    double c, s, tmp;
    givensrot(x(N), -1.0, c, s, tmp);
    Ct[N] = make_rot<double>(c, -s);
    B[N] = make_rot<double>(s, -c);
    for (int i = N - 1; i >= 1; --i) {
        givensrot(x(i), tmp, c, s, tmp);
        Ct[i] = make_rot<double>(c, -s);
        B[i] = make_rot<double>(c, s);
    }
}
template <class XS> AMRVW_DI void r_factorization(Rot<cplx>* B, Rot<cplx>* Ct, cplx* D, int N, const XS& x) {
    cplx c, tmp; double s;
    givensrot(x(N), cplx(-1.0, 0.0), c, s, tmp);
    const double nrm = abs_(c);
    const cplx alpha = c / nrm;
    Ct[N] = make_rot<cplx>(conj_(c), -s);
    B[N] = make_rot<cplx>(s * alpha, -nrm);
    for (int i = 1; i <= N + 1; ++i) D[i] = cplx(1.0, 0.0);
    D[N] = conj_(alpha);
    D[N + 1] = alpha;
    for (int i = N - 1; i >= 1; --i) {
        cplx t2;
        givensrot(x(i), tmp, c, s, t2);
        tmp = t2;
        Ct[i] = make_rot<cplx>(conj_(c), -s);
        B[i] = make_rot<cplx>(c, s);
    }
}
// q_factorization (companion.jl:123-137): Q[i] = (0, 1), plus the two sentinels and D_Q = 1
template <class S, bool P> AMRVW_DI void q_factorization(Fact<S, P>& F) {
    const int N = F.N;
    F.Q[0] = make_rot<S>(S(1.0), 0.0);
    for (int i = 1; i <= N - 1; ++i) F.Q[i] = make_rot<S>(S(0.0), 1.0);
    F.Q[N] = make_rot<S>(S(1.0), 0.0);
    if constexpr (!is_real<S>::value) for (int i = 0; i <= N; ++i) F.DQ[i] = cplx(1.0, 0.0);
}

// closed forms for degree <= 2 (utils.jl:63-140)
AMRVW_DI double discr_(double a, double b, double c) {
    const double d = b * b - a * c, e = b * b + a * c;
    if (3.0 * fabs(d) > e) return d;
    const double p = b * b, dp = fma(b, b, -p), q = a * c, dq = fma(a, c, -q);
    return (p - q) + (dp - dq);
}
AMRVW_DI int solve_simple(const double* ps, int ncoef, cplx* out) {
    if (ncoef <= 1) return 0;
    if (ncoef == 2) { out[0] = cplx(-ps[0] / ps[1], 0.0); return 1; }
    const double a = ps[2], b = -ps[1] / 2, c = ps[0];
    const double d = discr_(a, b, c);
    if (d <= 0.0) {
        const double r = b / a, s = sqrt(-d) / a;
        out[0] = cplx(r, s); out[1] = cplx(r, -s);
    } else {
        const double r = sqrt(d) * (sign_(b) + (b == 0.0 ? 1.0 : 0.0)) + b;
        out[0] = cplx(r / a, 0.0); out[1] = cplx(c / r, 0.0);
    }
    return 2;
}
AMRVW_DI int solve_simple(const cplx* ps, int ncoef, cplx* out) {
    if (ncoef <= 1) return 0;
    if (ncoef == 2) { out[0] = -ps[0] / ps[1]; return 1; }
    const cplx c = ps[0], b = ps[1], a = ps[2];
    const cplx d = csqrt_(b * b - 4.0 * a * c);
    out[0] = (-b + d) / (2.0 * a);
    out[1] = (-b - d) / (2.0 * a);
    return 2;
}

}  // namespace amrvw
