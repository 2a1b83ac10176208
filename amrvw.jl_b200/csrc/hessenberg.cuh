// hessenberg.cuh -- eigenvalues of upper Hessenberg matrices through qr_factorization:
// the same core chase as the companion path, with R held as a dense upper triangular matrix
// (RFactorizationUpperTriangular, rfactorization.jl:259-357) or, for a unitary H, as a unitary diagonal
// (RFactorizationUnitaryDiagonal, rfactorization.jl:230-248).
//
// One warp per matrix (hess_kernel), or one thread per matrix for large batches of small matrices (hess_thread_kernel):
// the same code, the number of cooperating lanes CW = 32 or 1 is a template parameter.  The scalar part of the
// algorithm (shifts, turnovers in Q, deflation) is executed by all cooperating lanes on the same values; the O(n)
// part -- passing a rotator through the dense R is a column rotation of columns i, i+1 and a row rotation of rows
// i, i+1 -- is split over the lanes.  Only the diagonal block of the
// active window is kept up to date: rows above it and columns right of it never feed an eigenvalue (a row
// rotation mixes rows i, i+1 only, a column rotation columns i, i+1 only), so each entry that is read sees exactly
// the arithmetic of the reference and the eigenvalues do not change.  R lives in shared memory when the matrix
// fits (leading dimension odd: rows and columns are both conflict free), else in a global workspace (L2).
// Q, D_Q and the unitary diagonal are shared by the warp: every update is "all lanes read, __syncwarp, lane 0 writes,
// __syncwarp" (HWRITE below; both barriers compile away for CW = 1), so no lane can read a value another lane has
// already replaced.
#pragma once
#include "solver.cuh"

namespace amrvw {

template <class S, int CW> struct HFact {
    Rot<S>* Q;     // descending chain, slots 0..N (1..N-1 live, 0 and N identity sentinels)
    S* DQ;         // complex only: slots 0..N+1
    S* R;          // dense mode: column major, entry (l,k) 1-based at R[(l-1) + (k-1)*ld]
    S* D;          // unitary mode: diagonal of R, slots 1..N
    int ld, N;
    bool unitary;
    int lane;      // position in the cooperating group
    // CW: lanes cooperating on this matrix: 32 (one warp per matrix) or 1 (one thread per matrix, small n)
    long long turnovers;
};

// lane 0 stores, between two warp barriers: the reads every lane did before the call are complete, the new values are visible after it
#define HSYNC(F) do { if constexpr (CW == 32) __syncwarp(); } while (0)
#define HWRITE(F, ...) do { HSYNC(F); if ((F).lane == 0) { __VA_ARGS__; } HSYNC(F); } while (0)

template <class S, int CW> AMRVW_DI S& hr(const HFact<S, CW>& F, int l, int k) { return F.R[(size_t)(k - 1) * F.ld + (l - 1)]; }
template <class S, int CW> AMRVW_DI S hr_entry(const HFact<S, CW>& F, int l, int k) {   // rfactorization.jl:268-274 / diagonal.jl:30-36
    if (F.unitary) return l == k ? F.D[l] : S(0.0);
    return hr(F, l, k);
}
// A = (Q R)[j:j+1, j:j+1]   (diagonal-block.jl:12-36); same closed form for the Q side as solver.cuh: q_block
template <class S, int CW> AMRVW_NI void diagonal_block(const HFact<S, CW>& F, int j, S A[4]) {
    const Rot<S> qm = F.Q[j - 1], q0 = F.Q[j], qp = F.Q[j + 1];
    const S dm = dget(F.DQ, j - 1), d0 = dget(F.DQ, j), dp = dget(F.DQ, j + 1);
    const S qji = S(-qm.s) * dm;
    const S qjj = conj_(qm.c) * q0.c * d0;
    const S qjk = qp.c * q0.s * conj_(qm.c) * dp;
    const S qkj = S(-q0.s) * d0;
    const S qkk = conj_(q0.c) * qp.c * dp;
    const int i = j - 1, k = j + 1;
    const S rjj = hr_entry(F, j, j), rjk = hr_entry(F, j, k), rkk = hr_entry(F, k, k);
    S rij = S(0.0), rik = S(0.0);
    if (i >= 1) { rij = hr_entry(F, i, j); rik = hr_entry(F, i, k); }
    A[0] = qji * rij + qjj * rjj;
    A[1] = qji * rik + qjj * rjk + qjk * rkk;
    A[2] = qkj * rjj;
    A[3] = qkj * rjk + qkk * rkk;
}

// passthrough!(RF::RFactorizationUpperTriangular, V) (rfactorization.jl:317-352): R V = U' R'.  Rows lo..i of the column
// rotation and columns i+1..hi of the row rotation (the active window) are split over the lanes; the three entries of the
// 2x2 diagonal block are carried in registers by every lane, so the Givens rotation in the middle needs no exchange.
template <class S, int CW> AMRVW_DI void pass_dense(HFact<S, CW>& F, Rot<S>& V, int i, int lo, int hi) {
    const int j = i + 1, lane = F.lane;
    S* Ri = F.R + (size_t)(i - 1) * F.ld;
    S* Rj = Ri + F.ld;
    const S c = V.c; const double sn = V.s;
    HSYNC(F);
    // the block and the first slice of both rotations are requested together (their addresses are disjoint from this position's stores)
    const S rii = Ri[i - 1], rij = Rj[i - 1], rjj = Rj[j - 1];
    const int r0 = lo - 1 + lane, m0 = j + 1 + lane;
    const bool has_r = r0 < i - 1, has_m = m0 <= hi;
    const int r0s = has_r ? r0 : i - 1;
    S* Rm0 = F.R + (size_t)((has_m ? m0 : j) - 1) * F.ld;
    const S a0 = Ri[r0s], b0 = Rj[r0s];
    const S x0 = Rm0[i - 1], y0 = Rm0[j - 1];
    if (has_r) {
        Ri[r0] = c * a0 - sn * b0;
        Rj[r0] = sn * a0 + conj_(c) * b0;
    }
    for (int k = r0 + CW; k < i - 1; k += CW) {
        const S rki = Ri[k], rkj = Rj[k];
        Ri[k] = c * rki - sn * rkj;
        Rj[k] = sn * rki + conj_(c) * rkj;
    }
    const S nii = c * rii - sn * rij;
    const S nij = sn * rii + conj_(c) * rij;
    const S rji = c * S(0.0) - sn * rjj;              // R[j,i] is a structural zero
    const S njj = sn * S(0.0) + conj_(c) * rjj;
    S c2, r; double s2;
    givensrot(nii, rji, c2, s2, r);
    if (has_m) {
        Rm0[i - 1] = c2 * x0 + s2 * y0;
        Rm0[j - 1] = -(s2 * x0) + conj_(c2) * y0;
    }
    for (int k = m0 + CW; k <= hi; k += CW) {
        S* Rk = F.R + (size_t)(k - 1) * F.ld;
        const S rik = Rk[i - 1], rjk = Rk[j - 1];
        Rk[i - 1] = c2 * rik + s2 * rjk;
        Rk[j - 1] = -(s2 * rik) + conj_(c2) * rjk;
    }
    HSYNC(F);                                     // every lane has read the block before lane 0 replaces it
    if (lane == 0) {
        Ri[i - 1] = c2 * nii + s2 * rji;
        Rj[i - 1] = c2 * nij + s2 * njj;
        Rj[j - 1] = -(s2 * nij) + conj_(c2) * njj;
    }
    HSYNC(F);
    V = adjoint(make_rot<S>(c2, s2));
}
// The two passes of a real double-shift chase position -- V at j = i + 1, then U at i (RDS.jl:42-70) -- in one sweep over R:
// the column rotations of columns (j, j+1) and (i, j) act on the same rows, the row rotations of rows (j, j+1) and (i, j) on the
// same columns, so every entry above / right of the 3 x 3 diagonal block is read and written once instead of twice, with half the
// barriers.  Per entry the operations and their order are those of two pass_dense calls (the strict build stays bit-identical).
template <int CW> AMRVW_DI void pass_dense2(HFact<double, CW>& F, Rot<double>& V, Rot<double>& U, int i, int lo, int hi) {
    const int j = i + 1, k = i + 2, lane = F.lane;
    double* Ci = F.R + (size_t)(i - 1) * F.ld;
    double* Cj = Ci + F.ld;
    double* Ck = Cj + F.ld;
    const double cV = V.c, sV = V.s, cU = U.c, sU = U.s;
    HSYNC(F);
    // every load of this position is independent of its stores (rows < i, the 3 x 3 block, columns > k are disjoint): the block and
    // the first slice of both rotations are requested together, so the position pays one round trip to R instead of three
    const double rii = Ci[i - 1], rij = Cj[i - 1], rik = Ck[i - 1], rjj = Cj[j - 1], rjk = Ck[j - 1], rkk = Ck[k - 1];
    const int r0 = lo - 1 + lane, m0 = k + 1 + lane;
    const bool has_r = r0 < i - 1, has_m = m0 <= hi;
    const int r0s = has_r ? r0 : i - 1;                              // a valid address when this lane has no row
    double* Cm0 = F.R + (size_t)((has_m ? m0 : k) - 1) * F.ld;
    const double a0 = Ci[r0s], b0 = Cj[r0s], c0 = Ck[r0s];
    const double x0 = Cm0[i - 1], y0 = Cm0[j - 1], z0 = Cm0[k - 1];
    if (has_r) {
        const double b1 = cV * b0 - sV * c0, c1 = sV * b0 + cV * c0;  // V on columns (j, k)
        Ci[r0] = cU * a0 - sU * b1;                                    // U on columns (i, j)
        Cj[r0] = sU * a0 + cU * b1;
        Ck[r0] = c1;
    }
    for (int r = r0 + CW; r < i - 1; r += CW) {
        const double a = Ci[r], b = Cj[r], c = Ck[r];
        const double b1 = cV * b - sV * c, c1 = sV * b + cV * c;
        Ci[r] = cU * a - sU * b1;
        Cj[r] = sU * a + cU * b1;
        Ck[r] = c1;
    }
    // V at j: row i is an ordinary row of its column rotation, rows j, k its diagonal block
    const double bi1 = cV * rij - sV * rik, ci1 = sV * rij + cV * rik;
    const double njj = cV * rjj - sV * rjk, njk = sV * rjj + cV * rjk;
    const double rkj = cV * 0.0 - sV * rkk, nkk = sV * 0.0 + cV * rkk;
    double c2V, s2V, tV;
    givensrot(njj, rkj, c2V, s2V, tV);
    const double rjj1 = c2V * njj + s2V * rkj;
    const double rjk1 = c2V * njk + s2V * nkk, rkk1 = -(s2V * njk) + c2V * nkk;
    // U at i: rows i, j are its diagonal block
    const double nii = cU * rii - sU * bi1, nij = sU * rii + cU * bi1;
    const double rji = cU * 0.0 - sU * rjj1, njj2 = sU * 0.0 + cU * rjj1;
    double c2U, s2U, tU;
    givensrot(nii, rji, c2U, s2U, tU);
    if (has_m) {
        const double y1 = c2V * y0 + s2V * z0;                         // V' on rows (j, k)
        Cm0[k - 1] = -(s2V * y0) + c2V * z0;
        Cm0[i - 1] = c2U * x0 + s2U * y1;                              // U' on rows (i, j)
        Cm0[j - 1] = -(s2U * x0) + c2U * y1;
    }
    for (int m = m0 + CW; m <= hi; m += CW) {
        double* Cm = F.R + (size_t)(m - 1) * F.ld;
        const double x = Cm[i - 1], y = Cm[j - 1], z = Cm[k - 1];
        const double y1 = c2V * y + s2V * z;
        Cm[k - 1] = -(s2V * y) + c2V * z;
        Cm[i - 1] = c2U * x + s2U * y1;
        Cm[j - 1] = -(s2U * x) + c2U * y1;
    }
    HSYNC(F);                                                      // every lane has read the block before lane 0 replaces it
    if (lane == 0) {
        Ci[i - 1] = c2U * nii + s2U * rji;
        Cj[i - 1] = c2U * nij + s2U * njj2;
        Cj[j - 1] = -(s2U * nij) + c2U * njj2;
        Ck[i - 1] = c2U * ci1 + s2U * rjk1;
        Ck[j - 1] = -(s2U * ci1) + c2U * rjk1;
        Ck[k - 1] = rkk1;
    }
    HSYNC(F);
    V = adjoint(make_rot<double>(c2V, s2V));
    U = adjoint(make_rot<double>(c2U, s2U));
}
// diagonal.jl:84-96 on a diagonal shared by the warp; real: no-op
template <class S, int CW> AMRVW_DI void hpass_diag(HFact<S, CW>& F, S* D, Rot<S>& U, int i) {
    if constexpr (!is_real<S>::value) {
        const S alpha = D[i], beta = D[i + 1];
        HWRITE(F, D[i] = beta; D[i + 1] = alpha);
        U.c = (alpha * conj_(beta)) * U.c;
    }
}
// passthrough!(RF, U) for the two R factorizations of qr_factorization
template <class S, int CW> AMRVW_DI void pass_R(HFact<S, CW>& F, Rot<S>& U, int i, int lo, int hi) {
    if (F.unitary) {   // diagonal.jl:84-96 (complex: the entries swap, the cosine takes the phase), :113-118 (real: the sine takes d_i d_{i+1})
        if constexpr (is_real<S>::value) U.s = U.s * F.D[i] * F.D[i + 1];
        else hpass_diag(F, F.D, U, i);
        return;
    }
    pass_dense(F, U, i, lo, hi);
}
template <class S, int CW> AMRVW_DI void pass_Qh(HFact<S, CW>& F, Rot<S>& U, int& i) {   // qfactorization.jl:86-91
    hpass_diag(F, F.DQ, U, i);
    Rot<S> r1, r2, r3;
    turnover(F.Q[i], F.Q[i + 1], U, r1, r2, r3);   // chains.jl:360
    F.turnovers++;
    U = r1;
    HWRITE(F, F.Q[i] = r2; F.Q[i + 1] = r3);
    i += 1;
}

template <int CW> AMRVW_NI void push_phase(HFact<cplx, CW>& F, cplx alpha, int i) {   // diagonal.jl:199-225; nothing comes back, lane 0 does it
    HSYNC(F);
    if (F.lane == 0) {
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
    HSYNC(F);
}

// real double shift (create-bulge.jl:20-77, RDS.jl, bulge-step.jl:36-49) -- solver.cuh: bulge_step without the rank-one shortcut
template <int CW> AMRVW_NI void bulge_step(HFact<double, CW>& F, Counter& ctr, uint64_t seed, uint64_t mat, int& ord, PolyStats& st) {
    st.sweeps++;
    Rot<double> U, V, W;
    int i = ctr.start_index;
    const int lo = ctr.start_index, hi = ctr.stop_index + 1;
    if (ctr.it_count == 0) {
        ctr.it_count = AMRVW_IT_COUNT;
        st.exceptional++;
        const double t = uniform01(seed ^ 0xA5A5A5A5DEADBEEFull, mat, (uint64_t)ord++) * 3.141592653589793;
        double sn, cs;
        sincos(t, &sn, &cs);
        U = make_rot<double>(sn, cs);
        V = make_rot<double>(sn, -cs);
    } else {
        double A[4];
        cplx e1, e2;
        diagonal_block(F, ctr.stop_index, A);
        eig2x2(A, e1, e2);
        diagonal_block(F, i, A);
        const double bk11 = A[0], bk12 = A[1], bk21 = A[2], bk22 = A[3];
        const double bk32 = (-F.Q[i + 1].s) * hr_entry(F, i + 1, i + 1);
        const double c1 = -e1.im * e2.im + e1.re * e2.re - e1.re * bk11 - e2.re * bk11 + bk11 * bk11 + bk12 * bk21;
        const double c2 = -e1.re * bk21 - e2.re * bk21 + bk11 * bk21 + bk21 * bk22;
        const double c3 = bk21 * bk32;
        double c, s, nrm, tmp;
        givensrot(c2, c3, c, s, nrm);
        V = make_rot<double>(c, -s);
        givensrot(c1, nrm, c, s, tmp);
        U = make_rot<double>(c, -s);
    }
    const int nq = F.N - 1;
    {   // absorb_Ut (RDS.jl:14-36)
        const double p = i == 1 ? 1.0 : sign_(F.Q[i - 1].c);
        Rot<double> w = F.Q[i];
        w.s = sign_(p) * w.s;
        Rot<double> r1, r2, r3;
        turnover(adjoint(U), adjoint(V), w, r1, r2, r3);
        F.turnovers++;
        W = r1;
        const Rot<double> f = fuse(r3, F.Q[i + 1]);
        r2.s = sign_(p) * r2.s;
        HWRITE(F, F.Q[i + 1] = f; F.Q[i] = r2);
    }
    for (;;) {
        const int j = i + 1;
        if (F.unitary) {              // passthrough_triu (RDS.jl:42-70)
            pass_R(F, V, j, lo, hi);
            pass_R(F, U, i, lo, hi);
        } else {
            pass_dense2(F, V, U, i, lo, hi);
        }
        if (j < ctr.stop_index) {     // passthrough_Q (RDS.jl:97-131)
            int jv = j, iu = i;
            pass_Qh(F, V, jv);
            pass_Qh(F, U, iu);
            Rot<double> r1, r2, r3;
            turnover(W, V, U, r1, r2, r3);
            F.turnovers++;
            V = r1; U = r2; W = r3;
            i += 1;
        } else {
            const double p = (j + 1 > nq) ? 1.0 : sign_(F.Q[j + 1].c);
            V.s = sign_(p) * V.s;
            const Rot<double> fj = fuse(F.Q[j], V);
            HWRITE(F, F.Q[j] = fj);
            int iu = i;
            pass_Qh(F, U, iu);        // real: no diagonal; U now at j
            U = fuse(W, U);
            pass_R(F, U, iu, lo, hi);
            U.s = sign_(p) * U.s;
            const Rot<double> fu = fuse(F.Q[iu], U);
            HWRITE(F, F.Q[iu] = fu);
            break;
        }
    }
}
template <int CW> AMRVW_DI void deflate(HFact<double, CW>& F, int k) {  // RDS.jl:134-141
    const Rot<double> q = make_rot<double>(sign_(F.Q[k].c), 0.0);
    HWRITE(F, F.Q[k] = q);
}

// complex single shift (create-bulge.jl:85-110, CSS.jl)
template <int CW> AMRVW_NI void bulge_step(HFact<cplx, CW>& F, Counter& ctr, uint64_t seed, uint64_t mat, int& ord, PolyStats& st) {
    st.sweeps++;
    cplx A[4];
    cplx shift;
    const int lo = ctr.start_index, hi = ctr.stop_index + 1;
    if (ctr.it_count == 0) {
        ctr.it_count = AMRVW_IT_COUNT;
        st.exceptional++;
        const double t = uniform01(seed ^ 0xA5A5A5A5DEADBEEFull, mat, (uint64_t)ord++) * 3.141592653589793;
        double sn, cs;
        sincos(t, &sn, &cs);
        shift = cplx(cs, sn);
    } else {
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
    {
        cplx dph;
        const Rot<cplx> f = fuse(R, F.Q[i], dph);
        HWRITE(F, F.Q[i] = f);
        push_phase(F, dph, i);
    }
    for (;;) {
        pass_R(F, U, i, lo, hi);
        if (i < ctr.stop_index) {
            pass_Qh(F, U, i);
        } else {
            hpass_diag(F, F.DQ, U, i);
            cplx dph;
            const Rot<cplx> f = fuse(F.Q[i], U, dph);
            const cplx d0 = F.DQ[i] * dph, d1 = F.DQ[i + 1] * conj_(dph);
            HWRITE(F, F.Q[i] = f; F.DQ[i] = d0; F.DQ[i + 1] = d1);
            break;
        }
    }
}
template <int CW> AMRVW_DI void deflate(HFact<cplx, CW>& F, int k) {  // CSS.jl:98-122
    const cplx alpha = F.Q[k].c;
    HWRITE(F, F.Q[k] = make_rot<cplx>(cplx(1.0, 0.0), 0.0));
    push_phase(F, alpha, k);
}

template <class S, int CW> AMRVW_NI void check_deflation(HFact<S, CW>& F, Counter& ctr) {  // AMRVW_algorithm.jl:181-202
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

// AMRVW_algorithm.jl:10-138 -- the driver of solver.cuh: drive_window for lo = 1
template <class S, int CW> AMRVW_DI int drive_hessenberg(HFact<S, CW>& F, cplx* E, uint64_t seed, uint64_t mat, PolyStats& st) {
    const int n = F.N;
    Counter ctr = {0, 1, n - 1, AMRVW_IT_COUNT, -1};
    const long long it_max = AMRVW_IT_MAX(n);
    const bool w = F.lane == 0;
    long long kk = 0;
    int ord = 0;
    S A[4];
    check_deflation(F, ctr);
    while (kk <= it_max) {
        if (ctr.stop_index <= 0) return 0;
        check_deflation(F, ctr);
        ++kk;
        const int k = ctr.stop_index;
        const int delta = ctr.stop_index - ctr.zero_index;
        if (delta == 1) {
            diagonal_block(F, k, A);
            cplx e1, e2;
            eig2x2(A, e1, e2);
            if (w) { E[k] = e1; E[k + 1] = e2; }
            if (ctr.stop_index == 2) {
                diagonal_block(F, 1, A);
                if (w) E[1] = to_c(A[0]);
                return 0;
            }
            ctr.stop_index = ctr.zero_index - 1;
            ctr.zero_index = 0;
            ctr.start_index = 1;
            check_deflation(F, ctr);
            if (ctr.stop_index <= 0) return 0;
        } else if (delta <= 0) {
            diagonal_block(F, k, A);
            if (ctr.stop_index == 1) {
                if (w) { E[1] = to_c(A[0]); E[2] = to_c(A[3]); }
                return 0;
            }
            if (w) E[k + 1] = to_c(A[3]);
            ctr.stop_index = ctr.zero_index - 1;
            if (ctr.stop_index <= 0) return 0;
            ctr.zero_index = 0;
            ctr.start_index = 1;
            check_deflation(F, ctr);
        } else {
            bulge_step(F, ctr, seed, mat, ord, st);
            ctr.it_count -= 1;
        }
    }
    return ctr.stop_index <= 0 ? 0 : ctr.stop_index + 1;
}

struct HessArgs {
    const void* H;              // matrices back to back, column major, matrix p is dims[p] x dims[p]
    const int* dims;
    const long long* moff;      // element offset of matrix p in H (nbatch + 1)
    const long long* eoff;      // eigenvalue slot offset of matrix p (nbatch + 1)
    cplx* eig;
    int* nunconv;
    PolyStats* stats;           // one per matrix
    unsigned int* next;         // work counter
    void* work;                 // global workspace when R does not fit in shared memory: per warp slot `work_stride` bytes
    size_t work_stride;
    long long nbatch;
    int unitary;
    int ld_max;                 // leading dimension used for every matrix of the call (odd)
    uint64_t seed;
};

template <class S> __host__ __device__ inline size_t hess_warp_bytes(int nmax, int ld, bool with_r) {
    size_t b = (size_t)(nmax + 2) * sizeof(Rot<S>) + 2 * (size_t)(nmax + 2) * sizeof(S);   // Q, DQ, D
    if (with_r) b += (size_t)nmax * ld * sizeof(S);
    return (b + 15) & ~(size_t)15;
}

// qr_factorization + eigvals of matrix p by the group F describes (a warp, or one thread); E, nunconv, stats written by its lane 0
template <class S, int CW> AMRVW_DI void hess_solve_one(HFact<S, CW>& F, const HessArgs& a, const unsigned int p) {
    const int lane = F.lane; constexpr int w = CW;
    const S* Hall = static_cast<const S*>(a.H);
    const int n = a.dims[p];
    cplx* E = a.eig + a.eoff[p] - 1;     // 1-based slots
    PolyStats st = {0, 0, 0, 0, 0};
    F.N = n; F.unitary = false; F.turnovers = 0;
    int unconv = 0;
    if (n == 1) {
        if (lane == 0) E[1] = to_c(Hall[a.moff[p]]);
    } else if (n >= 2) {
        const S* H = Hall + a.moff[p];
        // load the upper Hessenberg part (column major, coalesced over rows when a warp shares the matrix)
        for (int k = 1; k <= n; ++k) {
            const int top = min(n, k + 1);
            for (int l = 1 + lane; l <= top; l += w) hr(F, l, k) = H[(size_t)(k - 1) * n + (l - 1)];
        }
        HSYNC(F);
        // qr_factorization! (qrfactorization.jl:69-85): Givens on (R[i,i], R[i+1,i]), row rotation of columns i+1..n
        if (lane == 0) {
            F.Q[0] = make_rot<S>(S(1.0), 0.0);
            F.Q[n] = make_rot<S>(S(1.0), 0.0);
        }
        if constexpr (!is_real<S>::value) for (int k = lane; k <= n + 1; k += w) F.DQ[k] = cplx(1.0, 0.0);
        for (int i = 1; i <= n - 1; ++i) {
            const int j = i + 1;
            S c, r; double sn;
            givensrot(hr(F, i, i), hr(F, j, i), c, sn, r);
            if (lane == 0) F.Q[i] = adjoint(make_rot<S>(c, sn));
            HSYNC(F);
            for (int k = j + lane; k <= n; k += w) {
                const S rik = hr(F, i, k), rjk = hr(F, j, k);
                hr(F, i, k) = c * rik + sn * rjk;
                hr(F, j, k) = -(sn * rik) + conj_(c) * rjk;
            }
            if (lane == 0) { hr(F, i, i) = r; hr(F, j, i) = S(0.0); }
            HSYNC(F);
        }
        if (a.unitary) {   // R = sign.(diag(R))  (qrfactorization.jl:87-93)
            for (int k = 1 + lane; k <= n; k += w) {
                const S z = hr(F, k, k);
                if constexpr (is_real<S>::value) F.D[k] = sign_(z);
                else F.D[k] = iszero_(z) ? z : z / abs_(z);
            }
            F.unitary = true;
        }
        HSYNC(F);
        if (lane == 0) for (int k = 1; k <= n; ++k) E[k] = cplx(0.0, 0.0);
        unconv = drive_hessenberg(F, E, a.seed, (uint64_t)p, st);
        HSYNC(F);
        if (lane == 0) sort_roots(E + 1, n);
    }
    if (lane == 0) {
        a.nunconv[p] = unconv;
        if (a.stats) { st.turnovers = F.turnovers; st.unconverged = unconv; a.stats[p] = st; }
    }
    HSYNC(F);
}

template <class S, int CW> AMRVW_DI void hess_place(HFact<S, CW>& F, unsigned char* mine, const int nmax, const int ld) {
    F.Q = reinterpret_cast<Rot<S>*>(mine);
    F.DQ = reinterpret_cast<S*>(mine + (size_t)(nmax + 2) * sizeof(Rot<S>));
    F.D = F.DQ + (nmax + 2);
    F.R = F.D + (nmax + 2);
    F.ld = ld;
}

// One warp per matrix, matrices handed out by an atomic counter.  SMEM: R in shared memory, else in a.work.
template <class S, bool SMEM>
__global__ void hess_kernel(HessArgs a, int nmax) {
    extern __shared__ __align__(16) unsigned char hsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    HFact<S, 32> F;
    hess_place(F, hsm + (size_t)warp * hess_warp_bytes<S>(nmax, a.ld_max, SMEM), nmax, a.ld_max);
    if (!SMEM) F.R = reinterpret_cast<S*>(static_cast<unsigned char*>(a.work) + ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * a.work_stride);
    F.lane = lane;
    for (;;) {
        unsigned int p = 0;
        if (lane == 0) p = atomicAdd(a.next, 1u);
        p = __shfl_sync(0xffffffffu, p, 0);
        if ((long long)p >= a.nbatch) break;
        hess_solve_one(F, a, p);
    }
}

// One THREAD per matrix for batches of small matrices (n <= 16 or so: the reference's own test size is 15): the warp-per-matrix
// kernel is bound by the instructions of the scalar part, which one warp issues for one matrix; here a warp issues them for up
// to 32 matrices (as far as their control flow agrees).  Every thread owns `stride` bytes of shared memory (an odd multiple of
// 16 bytes: the 32 threads' 16-byte accesses fall into different banks); same code, cooperation width 1.
template <class S>
__global__ void hess_thread_kernel(HessArgs a, int nmax, unsigned stride) {
    extern __shared__ __align__(16) unsigned char hsm[];
    HFact<S, 1> F;
    hess_place(F, hsm + (size_t)threadIdx.x * stride, nmax, a.ld_max);
    F.lane = 0;
    for (;;) {
        const unsigned int p = atomicAdd(a.next, 1u);
        if ((long long)p >= a.nbatch) break;
        hess_solve_one(F, a, p);
    }
}

}  // namespace amrvw
