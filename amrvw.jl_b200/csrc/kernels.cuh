// kernels.cuh -- __global__ entry points of the one-bulge-at-a-time path (reference schedule):
// one polynomial per thread, everything on device: zero trimming, closed forms for degree <= 2,
// factorization set-up, the Francis iteration, sort, zero roots.
#pragma once
#include "solver.cuh"

namespace amrvw {

struct PolyRec;   // rounds.cuh

template <class S> struct Problem {
    const S* coeffs;          // non-pencil: a0..an per polynomial; pencil: vs
    const S* ws;              // pencil only
    const long long* offsets; // nbatch+1, CSR
    long long nbatch;
    cplx* roots;              // complex slots; polynomial p at offsets[p] (one slot per input entry: empty vectors cannot overlap anything)
    int* nroots;
    int* nunconv;
    PolyStats* stats;         // nullable
    // workspace: chains are arrays of `slots` rotators; polynomial p starts at slot offsets[p] + extra*p
    Rot<S>* Q; Rot<S>* B; Rot<S>* Ct; Rot<S>* WB; Rot<S>* WCt;
    S* DQ; S* DR; S* WD;
    int* dstack;              // pipelined kernel: per-polynomial stack of deflated Q indices (same slot layout)
    int extra;                // 1 (non-pencil: len+1 slots) or 2 (pencil: N+2 slots)
    unsigned long long* prof; // pipelined kernels: 16 warp-cycle counters (nullable), see amrvw_last_profile
    // rounds schedule (rounds.cuh): per-polynomial records, shifts of the pending fat steps, queue of
    // deferred small windows, round counters
    PolyRec* rec; Shifts* shifts; void* swlist; unsigned* counters;
    // unit schedules: polynomials in order of decreasing trimmed degree (nullable = identity).  The groups of
    // a warp take their lock-steps together, up to the longer window, so units are formed from neighbours
    // in this order, not in batch order (ragged batches: SURVEY 8 f2)
    const int* order;
    unsigned long long seed;
    // host only: the batch this one competes with on the device, in polynomials of ITS size (degree classes of a ragged
    // batch run side by side: capi.cu); 0 = it is alone.  Only the choice of the lane count looks at it
    long long load_hint;
};
// polynomial of slot `slot` of the unit schedule (slot = unit * G + group); nbatch = none
template <class S> AMRVW_DI long long unit_poly(const Problem<S>& pr, const long long slot) {
    if (slot >= pr.nbatch) return pr.nbatch;
    return pr.order ? (long long)pr.order[slot] : slot;
}

template <class S, bool P> AMRVW_DI Fact<S, P> make_fact(const Problem<S>& pr, long long p, int N) {
    const long long base = pr.offsets[p] + (long long)pr.extra * p;
    Fact<S, P> F;
    F.Q = pr.Q + base; F.B = pr.B + base; F.Ct = pr.Ct + base;
    F.WB = P ? pr.WB + base : nullptr; F.WCt = P ? pr.WCt + base : nullptr;
    if constexpr (is_real<S>::value) { F.DQ = nullptr; F.DR = nullptr; F.WD = nullptr; }
    else { F.DQ = pr.DQ + base; F.DR = pr.DR + base; F.WD = P ? pr.WD + base : nullptr; }
    F.N = N;
    F.turnovers = 0;
    return F;
}

// Trim exact zeros at both ends (utils.jl:14-22): first/last non-zero positions, -1 if none.
template <class S> AMRVW_DI void trim_zeros(const S* a, int len, int& first, int& last) {
    first = -1; last = -1;
    for (int i = 0; i < len; ++i)
        if (!iszero_(a[i])) { if (first < 0) first = i; last = i; }
}

template <class S> AMRVW_DI void finish_poly(const Problem<S>& pr, long long p, cplx* out, int cnt, int K, int unconv, const PolyStats& st) {
    for (int i = 0; i < K; ++i) out[cnt + i] = cplx(0.0, 0.0);   // companion.jl:266-268
    pr.nroots[p] = cnt + K;
    pr.nunconv[p] = unconv;
    if (pr.stats) pr.stats[p] = st;
}

// Factor (companion.jl:150-159) the trimmed polynomial ps[0..n] into the workspace.
template <class S> AMRVW_DI void factor_companion(Fact<S, false>& F, const S* ps) {
    const int n = F.N;
    const S pNi = inv_(ps[n]);
    q_factorization(F);
    r_factorization(F.B, F.Ct, F.DR, n, [&](int i) { return decompose_entry(ps, n, i, pNi); });
}

// ---------------------------------------------------------------- roots(ps), one bulge at a time
template <class S>
__global__ void __launch_bounds__(64) roots_reference_kernel(Problem<S> pr) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= pr.nbatch) return;
    const long long off = pr.offsets[p];
    const int len = (int)(pr.offsets[p + 1] - off);
    cplx* out = pr.roots + off;
    PolyStats st = {0, 0, 0, 0, 0};
    int first, last;
    trim_zeros(pr.coeffs + off, len, first, last);
    if (first < 0) { finish_poly(pr, p, out, 0, 0, 0, st); return; }
    const S* ps = pr.coeffs + off + first;
    const int ncoef = last - first + 1, K = first;
    if (ncoef <= 3) {
        const int cnt = solve_simple(ps, ncoef, out);
        finish_poly(pr, p, out, cnt, K, 0, st);
        return;
    }
    const int N = ncoef - 1;
    Fact<S, false> F = make_fact<S, false>(pr, p, N);
    factor_companion(F, ps);
    for (int i = 0; i < N; ++i) out[i] = cplx(0.0, 0.0);   // EIGS = zeros (AMRVW_algorithm.jl:23)
    Counter ctr = {0, 1, N - 1, AMRVW_IT_COUNT, N - 2};                 // AMRVW_algorithm.jl:18
    int ord = 0;
    const int unconv = drive_window(F, 1, N - 1, ctr, out - 1, AMRVW_IT_MAX(N - 1), pr.seed, (unsigned long long)p, ord, st, ReferenceStep());
    sort_roots(out, N);
    st.turnovers = F.turnovers;
    st.unconverged = unconv;
    finish_poly(pr, p, out, N, K, unconv, st);
}

// ---------------------------------------------------------------- roots(vs, ws): trimming, closed forms, factorization
// deflate_leading_zeros(vs, ws) (utils.jl:29-58), the degree <= 2 shortcut (companion.jl:288-293) and
// amrvw(vs, ws) (companion.jl:163-204) for polynomial p.  Returns the order n of the pencil to iterate
// on (chains factored into the workspace, root slots zeroed) or 0 when the polynomial is already
// finished (finish_poly done); Kz = zero roots to append.
template <class S>
AMRVW_DI int pencil_prepare(const Problem<S>& pr, const long long p, cplx* out, const PolyStats& st, Fact<S, true>& F, int& Kz) {
    const long long off = pr.offsets[p];
    const int N0 = (int)(pr.offsets[p + 1] - off);
    Kz = 0;
    if (N0 <= 0) { finish_poly(pr, p, out, 0, 0, 0, st); return 0; }
    const S* vs = pr.coeffs + off;
    const S* ws = pr.ws + off;
    // deflate_leading_zeros(vs, ws) without touching the inputs: the reference's two in-place edits become
    // overrides of one entry each.  Indices below are 1-based like the reference.
    int N = N0, wsN_idx = 0, vsK_idx = 0;
    S wsN_val = S(0.0), vsK_val = S(0.0);
    if (iszero_(ws[N0 - 1])) {
        N -= 1;
        while (N >= 1) {
            const S ai = vs[N] + ws[N - 1];
            if (!iszero_(ai)) { wsN_idx = N; wsN_val = ai; break; }
            N -= 1;
        }
    }
    int K = 1;
    if (iszero_(vs[0])) {
        K = 2;
        while (K <= N) {
            const S wprev = (K - 1 == wsN_idx) ? wsN_val : ws[K - 2];
            const S ai = vs[K - 1] + wprev;
            if (!iszero_(ai)) { vsK_idx = K; vsK_val = ai; break; }
            K += 1;
        }
    }
    auto V = [&](int i) { return i == vsK_idx ? vsK_val : vs[i - 1]; };  // 1-based
    auto W = [&](int i) { return i == wsN_idx ? wsN_val : ws[i - 1]; };
    const int n = N - K + 1;   // length of vs[K:N]
    Kz = K - 1;
    if (n <= 0) { finish_poly(pr, p, out, 0, 0, 0, st); return 0; }
    if (n <= 2) {  // companion.jl:288-293: rs = vcat(ps,0) + vcat(0,qs); roots(rs)
        S rs[3] = {S(0.0), S(0.0), S(0.0)};
        for (int i = 0; i < n; ++i) { rs[i] = rs[i] + V(K + i); rs[i + 1] = rs[i + 1] + W(K + i); }
        int f2, l2;
        trim_zeros(rs, n + 1, f2, l2);
        int cnt = 0;
        if (f2 >= 0) {
            cnt = solve_simple(rs + f2, l2 - f2 + 1, out);
            for (int i = 0; i < f2; ++i) out[cnt + i] = cplx(0.0, 0.0);
            cnt += f2;
        }
        finish_poly(pr, p, out, cnt, Kz, 0, st);
        return 0;
    }
    // amrvw(vs, ws) (companion.jl:163-204): ps = basic_decompose([vs; 1]); ws sign-adjusted; qs = [ws; -1]
    F = make_fact<S, true>(pr, p, n);
    q_factorization(F);
    {
        const double par = (n % 2 == 0) ? 1.0 : -1.0;
        auto xs_v = [&](int i) -> S {   // basic_decompose of [vs[K:N]; 1], monic so pNi = 1
            if (i == n + 1) return S(-1.0);
            if (i == n) return par * V(K) * S(1.0);
            const double sg = ((i + 1) % 2 == 0) ? 1.0 : -1.0;
            return (par * sg) * V(K + i) * S(1.0);
        };
        auto xs_w = [&](int i) -> S {   // adjust_pencil: entries n-1, n-3, ... change sign (companion.jl:168-170)
            if (i == n + 1) return S(-1.0);
            const S w = W(K + i - 1);
            return ((n - 1 - i) % 2 == 0 && i <= n - 1) ? -w : w;
        };
        r_factorization(F.B, F.Ct, F.DR, n, xs_v);
        r_factorization(F.WB, F.WCt, F.WD, n, xs_w);
    }
    for (int i = 0; i < n; ++i) out[i] = cplx(0.0, 0.0);
    return n;
}

// ---------------------------------------------------------------- roots(vs, ws), one bulge at a time
template <class S>
__global__ void __launch_bounds__(64) roots_pencil_reference_kernel(Problem<S> pr) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= pr.nbatch) return;
    cplx* out = pr.roots + pr.offsets[p];
    PolyStats st = {0, 0, 0, 0, 0};
    Fact<S, true> F;
    int Kz = 0;
    const int n = pencil_prepare(pr, p, out, st, F, Kz);
    if (n == 0) return;
    Counter ctr = {0, 1, n - 1, AMRVW_IT_COUNT, n - 2};
    int ord = 0;
    const int unconv = drive_window(F, 1, n - 1, ctr, out - 1, AMRVW_IT_MAX(n - 1), pr.seed, (unsigned long long)p, ord, st, ReferenceStep());
    sort_roots(out, n);
    st.turnovers = F.turnovers;
    st.unconverged = unconv;
    finish_poly(pr, p, out, n, Kz, unconv, st);
}

// ---------------------------------------------------------------- real_polynomial_roots counts
__global__ void count_real_roots_kernel(const cplx* roots, const long long* offsets, long long nbatch, int shift, const int* nroots, int* counts3) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nbatch) return;
    const cplx* r = roots + (offsets[p] - (long long)shift * p);
    int nre = 0, npos = 0, nneg = 0;
    for (int i = 0; i < nroots[p]; ++i) {
        const double im = r[i].im;
        if (im == 0.0) nre++; else if (im > 0.0) npos++; else nneg++;
    }
    counts3[3 * p] = nre; counts3[3 * p + 1] = npos; counts3[3 * p + 2] = nneg;
}

// ---------------------------------------------------------------- device self-test of the fast primitives
// out[0]: max relative error of rsqrt_nr against 1/sqrt over a log-uniform sweep of [1e-280, 4]
// out[1]: max |Q1 Q2 Q3 - R1 R2 R3| over random real turnovers on the division-free path
// out[2]: max | |c|^2 + s^2 - 1 | of the rotators it returns
// out[3]: number of turnovers that took the general (IEEE) path
// (mirrors test/test-amrvw-basics.jl:35-42: the turnover must preserve the product)
AMRVW_DI void rot3(double M[9], const Rot<double> r, int i) {   // M <- M * G(r)@i, 3x3, i in {0,1}
    for (int k = 0; k < 3; ++k) {
        const double a = M[3 * k + i], b = M[3 * k + i + 1];
        M[3 * k + i] = r.c * a - r.s * b;
        M[3 * k + i + 1] = r.s * a + r.c * b;
    }
}
__global__ void selftest_kernel(double* out, int n, unsigned long long seed) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    double e_rsqrt = 0.0, e_prod = 0.0, e_unit = 0.0, n_gen = 0.0;
    for (int k = tid; k < n; k += gridDim.x * blockDim.x) {
        const double u = uniform01(seed, 1, (uint64_t)k);
        const double x = exp(u * (log(4.0) - log(1e-280)) + log(1e-280));
        const double y = rsqrt_nr(x), yref = 1.0 / sqrt(x);
        e_rsqrt = fmax(e_rsqrt, fabs(y - yref) / yref);
        Rot<double> q[3];
        for (int j = 0; j < 3; ++j) {
            double sn, cs;
            sincos(6.283185307179586 * uniform01(seed, 2 + j, (uint64_t)k), &sn, &cs);
            // every 7th triple: a nearly deflated rotator, sines down to 1e-18
            if (k % 7 == j) { sn *= exp(-40.0 * uniform01(seed, 9, (uint64_t)k)); cs = copysign(sqrt(fmax(0.0, 1.0 - sn * sn)), cs); }
            q[j] = make_rot<double>(cs, sn);
        }
        Rot<double> r1, r2, r3;
        if (!turnover_fast_try(q[0], q[1], q[2], r1, r2, r3)) {
            n_gen += 1.0;
            turnover_ieee(q[0], q[1], q[2], r1, r2, r3);
        }
        double A[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Bm[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        rot3(A, q[0], 0); rot3(A, q[1], 1); rot3(A, q[2], 0);
        rot3(Bm, r1, 1); rot3(Bm, r2, 0); rot3(Bm, r3, 1);
        for (int j = 0; j < 9; ++j) e_prod = fmax(e_prod, fabs(A[j] - Bm[j]));
        e_unit = fmax(e_unit, fmax(fabs(r1.c * r1.c + r1.s * r1.s - 1.0), fmax(fabs(r2.c * r2.c + r2.s * r2.s - 1.0), fabs(r3.c * r3.c + r3.s * r3.s - 1.0))));
    }
    // doubles are non-negative: compare as integers for an atomic max
    atomicMax((unsigned long long*)&out[0], (unsigned long long)__double_as_longlong(e_rsqrt));
    atomicMax((unsigned long long*)&out[1], (unsigned long long)__double_as_longlong(e_prod));
    atomicMax((unsigned long long*)&out[2], (unsigned long long)__double_as_longlong(e_unit));
    atomicAdd(&out[3], n_gen);
}

// fall-back turnover on triples with vanishing sines (down to the denormal range): out[0] product error of
// turnover_robust_call, out[1] its deviation from unit norm, out[2] product error of the reference's
// formulas (turnover_ieee) on the same inputs, out[3] number of triples
__global__ void selftest_robust_kernel(double* out, int n, unsigned long long seed) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    double e_rob = 0.0, e_unit = 0.0, e_ref = 0.0, cnt = 0.0;
    for (int k = tid; k < n; k += gridDim.x * blockDim.x) {
        Rot<double> q[3];
        for (int j = 0; j < 3; ++j) {
            double sn, cs;
            sincos(6.283185307179586 * uniform01(seed, 2 + j, (uint64_t)k), &sn, &cs);
            // two of the three rotators (sometimes all three) get a sine between 1e-320 and 1e-100
            if ((k + j) % 3 != 0 || k % 5 == 0) { sn = copysign(exp(-(230.0 + 507.0 * uniform01(seed, 9 + j, (uint64_t)k))), sn); cs = copysign(1.0, cs); }
            q[j] = make_rot<double>(cs, sn);
        }
        const Rot3<double> o = turnover_robust_call(q[0], q[1], q[2]);
        Rot<double> i1, i2, i3;
        turnover_ieee(q[0], q[1], q[2], i1, i2, i3);
        double A[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Bm[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Cm[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        rot3(A, q[0], 0); rot3(A, q[1], 1); rot3(A, q[2], 0);
        rot3(Bm, o.r1, 1); rot3(Bm, o.r2, 0); rot3(Bm, o.r3, 1);
        rot3(Cm, i1, 1); rot3(Cm, i2, 0); rot3(Cm, i3, 1);
        for (int j = 0; j < 9; ++j) { e_rob = fmax(e_rob, fabs(A[j] - Bm[j])); const double d = fabs(A[j] - Cm[j]); e_ref = fmax(e_ref, d == d ? d : 1.0); }
        e_unit = fmax(e_unit, fmax(fabs(o.r1.c * o.r1.c + o.r1.s * o.r1.s - 1.0), fmax(fabs(o.r2.c * o.r2.c + o.r2.s * o.r2.s - 1.0), fabs(o.r3.c * o.r3.c + o.r3.s * o.r3.s - 1.0))));
        cnt += 1.0;
    }
    atomicMax((unsigned long long*)&out[0], (unsigned long long)__double_as_longlong(e_rob));
    atomicMax((unsigned long long*)&out[1], (unsigned long long)__double_as_longlong(e_unit));
    atomicMax((unsigned long long*)&out[2], (unsigned long long)__double_as_longlong(e_ref));
    atomicAdd(&out[3], cnt);
}

// complex twin: out[0] max |fast - ieee| over the nine output doubles, out[1] max |Q1 Q2 Q3 - R1 R2 R3|,
// out[2] max deviation from unit norm, out[3] turnovers sent to the IEEE path
AMRVW_DI void crot3(cplx M[9], const Rot<cplx> r, int i) {   // M <- M * G(r)@i, G = [c s; -s conj(c)]
    for (int k = 0; k < 3; ++k) {
        const cplx a = M[3 * k + i], b = M[3 * k + i + 1];
        M[3 * k + i] = a * r.c - b * r.s;
        M[3 * k + i + 1] = a * r.s + b * conj_(r.c);
    }
}
__global__ void selftest_complex_kernel(double* out, int n, unsigned long long seed) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    double e_diff = 0.0, e_prod = 0.0, e_unit = 0.0, n_gen = 0.0;
    for (int k = tid; k < n; k += gridDim.x * blockDim.x) {
        Rot<cplx> q[3];
        for (int j = 0; j < 3; ++j) {
            double sn, cs, ps, pc;
            sincos(6.283185307179586 * uniform01(seed, 2 + j, (uint64_t)k), &sn, &cs);
            sincos(6.283185307179586 * uniform01(seed, 12 + j, (uint64_t)k), &ps, &pc);
            if (k % 7 == j) { sn *= exp(-40.0 * uniform01(seed, 9, (uint64_t)k)); cs = copysign(sqrt(fmax(0.0, 1.0 - sn * sn)), cs); }
            q[j] = make_rot<cplx>(cplx(cs * pc, cs * ps), sn);
        }
        Rot<cplx> r1, r2, r3, i1, i2, i3;
        turnover_ieee(q[0], q[1], q[2], i1, i2, i3);
        if (!turnover_fast_try(q[0], q[1], q[2], r1, r2, r3)) { n_gen += 1.0; r1 = i1; r2 = i2; r3 = i3; }
        e_diff = fmax(e_diff, fmax(fmax(abs_(r1.c - i1.c), abs_(r2.c - i2.c)), abs_(r3.c - i3.c)));
        e_diff = fmax(e_diff, fmax(fmax(fabs(r1.s - i1.s), fabs(r2.s - i2.s)), fabs(r3.s - i3.s)));
        cplx A[9], Bm[9];
        for (int j = 0; j < 9; ++j) { A[j] = cplx(j % 4 == 0 ? 1.0 : 0.0, 0.0); Bm[j] = A[j]; }
        crot3(A, q[0], 0); crot3(A, q[1], 1); crot3(A, q[2], 0);
        crot3(Bm, r1, 1); crot3(Bm, r2, 0); crot3(Bm, r3, 1);
        for (int j = 0; j < 9; ++j) e_prod = fmax(e_prod, abs_(A[j] - Bm[j]));
        e_unit = fmax(e_unit, fmax(fabs(abs2_(r1.c) + r1.s * r1.s - 1.0), fmax(fabs(abs2_(r2.c) + r2.s * r2.s - 1.0), fabs(abs2_(r3.c) + r3.s * r3.s - 1.0))));
    }
    atomicMax((unsigned long long*)&out[0], (unsigned long long)__double_as_longlong(e_diff));
    atomicMax((unsigned long long*)&out[1], (unsigned long long)__double_as_longlong(e_prod));
    atomicMax((unsigned long long*)&out[2], (unsigned long long)__double_as_longlong(e_unit));
    atomicAdd(&out[3], n_gen);
}

// ---------------------------------------------------------------- FP64 peak probe
// 8 independent DFMA chains per thread, no memory traffic: measures the FP64 vector-pipe roofline.
__global__ void fp64_peak_kernel(double* out, int iters, double seed) {
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    const double m = 1.0000001, b = 1e-9;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
            a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
        }
    }
    out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace amrvw
