// shifts.cuh -- shifts of a multishift fat step from the DENSE trailing block.
//
// Not in the reference (it takes its two shifts from the trailing 2x2 block, create-bulge.jl:40-52).
// The pipelined kernel needs 2L/reuse shifts per fat step = the eigenvalues of the trailing
// (m+1) x (m+1) principal block of Q*R with rotator Q[k0] made diagonal.  Running the core-chasing
// iteration on a copy of that block costs ~40 serial sweeps of dependent turnovers (0.8 M cycles per
// fat step on one lane, 230 M cycles per degree-3000 polynomial: profiles/), so instead the block is
// written out entry by entry and handed to the classical double-shift Hessenberg QR:
//   * entries: column k, rows bottom-up, O(1) per entry by recurrences over the closed forms the
//     reference uses for single entries -- R[j,k] (rfactorization.jl:83-97,139-170) and the unitary
//     Hessenberg Q[j,l] = c_{j-1} s_j..s_{l-1} c_l, Q[j+1,j] = -s_j (chains.jl:109-158);
//   * eigenvalues: EISPACK hqr (eigenvalues only), O(m^3) plain flops with independent row/column
//     updates instead of dependent rotator chains.
// Shifts only steer convergence (the roots themselves always come from the rotator iteration), so
// the accuracy of the written-out entries (amplified by at most 1/prod(Ct sines) <= ||a||/|a_n|) is
// immaterial.  If hqr does not converge or produces a non-finite value the caller falls back to the
// core-chasing sub-solve.  oracle/amrvw_oracle.cpp (ms_shifts_dense) performs the same operations
// in the same order; the -DAMRVW_STRICT build is bit-identical to it.
#pragma once
#include "solver.cuh"

namespace amrvw {

// H: M x M row-major scratch; isg: M + 1 doubles scratch; q/b/c: staged rotators with GLOBAL indices
// (q[k0 .. stop+1], b/c[k0+1 .. stop+1] valid), q[k0] already made diagonal.
AMRVW_DI void dense_trailing_block(const Rot<double>* q, const Rot<double>* b, const Rot<double>* c, int k0, int stop, double* H, double* isg, int M) {
    const int j0 = k0 + 1;
    for (int i = 0; i < M * M; ++i) H[i] = 0.0;
    for (int j = j0; j <= stop + 1; ++j) isg[j - j0] = rcp_(c[j].s);
    for (int k = j0; k <= stop + 1; ++k) {
        double P = 1.0, T = 0.0, Vacc = 0.0;
        const double ck = b[k].c, sk = b[k].s;
        for (int j = k; j >= j0; --j) {
            const double w = (j == k) ? -sk : (b[j].c * P) * ck;
            const double ctc = c[j].c;
            const double ig = isg[j - j0];
            const double R = (w + ctc * T) * ig;
            const Rot<double> qj = q[j];
            if (j + 1 <= stop + 1) H[(j + 1 - j0) * M + (k - j0)] = qj.c * Vacc - qj.s * R;
            Vacc = qj.c * R + qj.s * Vacc;
            T = (-(ctc * w) - T) * ig;
            if (j < k) P = P * b[j].s;
        }
        H[k - j0] = q[k0].c * Vacc;
    }
}

// Double-shift Hessenberg QR after EISPACK hqr (eigenvalues only), 0-based, on the full M x M upper
// Hessenberg matrix H (row major, destroyed).  Differences from the textbook routine: every bulge
// starts at the top of the active block (no search for two consecutive small subdiagonals) and
// quotients by a common divisor are one reciprocal and multiplications.  E[0..M): conjugate pairs
// adjacent and exactly conjugate.  false: no convergence in 30 M iterations or a non-finite value.
AMRVW_DI bool hqr_eigenvalues(double* H, int M, cplx* E) {
#define AMRVW_H(i, j) H[(i) * M + (j)]
    double norm = 0.0;
    for (int i = 0; i < M; ++i)
        for (int j = (i > 0 ? i - 1 : 0); j < M; ++j) norm += fabs(AMRVW_H(i, j));
    int en = M - 1, itn = 30 * M;
    double t = 0.0;
    while (en >= 0) {
        int its = 0;
        const int na = en - 1, enm2 = na - 1;
        for (;;) {
            int l;
            for (l = en; l >= 1; --l) {
                double s = fabs(AMRVW_H(l - 1, l - 1)) + fabs(AMRVW_H(l, l));
                if (s == 0.0) s = norm;
                const double tst2 = s + fabs(AMRVW_H(l, l - 1));
                if (tst2 == s) break;
            }
            double x = AMRVW_H(en, en);
            if (l == en) { E[en] = cplx(x + t, 0.0); en = na; break; }
            double y = AMRVW_H(na, na);
            double w = AMRVW_H(en, na) * AMRVW_H(na, en);
            if (l == na) {
                const double p = (y - x) * 0.5;
                const double q = p * p + w;
                double zz = sqrt_(fabs(q));
                x = x + t;
                if (q >= 0.0) {
                    zz = p + (p >= 0.0 ? zz : -zz);
                    E[na] = cplx(x + zz, 0.0);
                    E[en] = E[na];
                    if (zz != 0.0) E[en] = cplx(x - w * rcp_(zz), 0.0);
                } else {
                    E[na] = cplx(x + p, zz);
                    E[en] = cplx(x + p, -zz);
                }
                en = enm2;
                break;
            }
            if (itn == 0) return false;
            if (its == 10 || its == 20) {
                t += x;
                for (int i = 0; i <= en; ++i) AMRVW_H(i, i) -= x;
                const double s = fabs(AMRVW_H(en, na)) + fabs(AMRVW_H(na, enm2));
                x = 0.75 * s; y = x; w = -0.4375 * s * s;
            }
            ++its; --itn;
            // first column of (H - a)(H - b) at the top of the active block, scaled
            double p, q, r;
            {
                const double zz = AMRVW_H(l, l);
                const double r0 = x - zz, s0 = y - zz;
                p = (r0 * s0 - w) * rcp_(AMRVW_H(l + 1, l)) + AMRVW_H(l, l + 1);
                q = AMRVW_H(l + 1, l + 1) - zz - r0 - s0;
                r = AMRVW_H(l + 2, l + 1);
                const double is = rcp_(fabs(p) + fabs(q) + fabs(r));
                p = p * is; q = q * is; r = r * is;
            }
            for (int i = l + 2; i <= en; ++i) { AMRVW_H(i, i - 2) = 0.0; if (i != l + 2) AMRVW_H(i, i - 3) = 0.0; }
            for (int k = l; k <= na; ++k) {
                const bool notlas = (k != na);
                double xs = 0.0;
                if (k != l) {
                    p = AMRVW_H(k, k - 1); q = AMRVW_H(k + 1, k - 1); r = notlas ? AMRVW_H(k + 2, k - 1) : 0.0;
                    xs = fabs(p) + fabs(q) + fabs(r);
                    if (xs == 0.0) continue;
                    const double ix = rcp_(xs);
                    p = p * ix; q = q * ix; r = r * ix;
                }
                double s = sqrt_(p * p + q * q + r * r);
                if (p < 0.0) s = -s;
                if (k != l) AMRVW_H(k, k - 1) = -s * xs;
                p = p + s;
                const double is = rcp_(s), ip = rcp_(p);
                const double hx = p * is, hy = q * is, hz = r * is;
                q = q * ip; r = r * ip;
                const int ihi = min(en, k + 3);
                if (notlas) {
#pragma unroll 4
                    for (int j = k; j <= en; ++j) {
                        const double a0 = AMRVW_H(k, j), a1 = AMRVW_H(k + 1, j), a2 = AMRVW_H(k + 2, j);
                        const double pp = a0 + q * a1 + r * a2;
                        AMRVW_H(k, j) = a0 - pp * hx; AMRVW_H(k + 1, j) = a1 - pp * hy; AMRVW_H(k + 2, j) = a2 - pp * hz;
                    }
#pragma unroll 4
                    for (int i = l; i <= ihi; ++i) {
                        const double a0 = AMRVW_H(i, k), a1 = AMRVW_H(i, k + 1), a2 = AMRVW_H(i, k + 2);
                        const double pp = hx * a0 + hy * a1 + hz * a2;
                        AMRVW_H(i, k) = a0 - pp; AMRVW_H(i, k + 1) = a1 - pp * q; AMRVW_H(i, k + 2) = a2 - pp * r;
                    }
                } else {
#pragma unroll 4
                    for (int j = k; j <= en; ++j) {
                        const double a0 = AMRVW_H(k, j), a1 = AMRVW_H(k + 1, j);
                        const double pp = a0 + q * a1;
                        AMRVW_H(k, j) = a0 - pp * hx; AMRVW_H(k + 1, j) = a1 - pp * hy;
                    }
#pragma unroll 4
                    for (int i = l; i <= ihi; ++i) {
                        const double a0 = AMRVW_H(i, k), a1 = AMRVW_H(i, k + 1);
                        const double pp = hx * a0 + hy * a1;
                        AMRVW_H(i, k) = a0 - pp; AMRVW_H(i, k + 1) = a1 - pp * q;
                    }
                }
            }
        }
    }
#undef AMRVW_H
    for (int i = 0; i < M; ++i)
        if (!(isfinite(E[i].re) && isfinite(E[i].im))) return false;
    return true;
}

// ---------------------------------------------------------------- one lane group per block (unit schedule)
// The same two routines with a group of GL lanes of a warp (GL = 4..32, mask gmask, lane = index in
// the group) sharing the work; the groups of a warp run their blocks side by side.  Within a group: every lane executes the scalar
// control flow on the same shared-memory values (no divergence, no communication), the entries are
// written / the rows and columns are updated one per lane.  Each entry sees exactly the arithmetic of
// the sequential routine above, so the strict build stays bit-identical to the oracle.
// H has leading dimension LD >= M (LD = M + 1 keeps column accesses off a single bank).
AMRVW_DI void dense_trailing_block_warp(const Rot<double>* q, const Rot<double>* b, const Rot<double>* c, int k0, int stop, double* H, int LD, double* isg, int M, int lane,
                                        const unsigned gmask, const int GL) {
    const int j0 = k0 + 1;
    for (int i = lane; i < M * LD; i += GL) H[i] = 0.0;
    for (int j = j0 + lane; j <= stop + 1; j += GL) isg[j - j0] = rcp_(c[j].s);
    __syncwarp(gmask);
    for (int k = j0 + lane; k <= stop + 1; k += GL) {
        double P = 1.0, T = 0.0, Vacc = 0.0;
        const double ck = b[k].c, sk = b[k].s;
        for (int j = k; j >= j0; --j) {
            const double w = (j == k) ? -sk : (b[j].c * P) * ck;
            const double ctc = c[j].c;
            const double ig = isg[j - j0];
            const double R = (w + ctc * T) * ig;
            const Rot<double> qj = q[j];
            if (j + 1 <= stop + 1) H[(j + 1 - j0) * LD + (k - j0)] = qj.c * Vacc - qj.s * R;
            Vacc = qj.c * R + qj.s * Vacc;
            T = (-(ctc * w) - T) * ig;
            if (j < k) P = P * b[j].s;
        }
        H[k - j0] = q[k0].c * Vacc;
    }
    __syncwarp(gmask);
}

AMRVW_DI bool hqr_eigenvalues_warp(double* H, const int LD, const int M, cplx* E, const int lane, const unsigned gmask, const int GL) {
    const int gbase = __ffs(gmask) - 1;   // first lane of the group
#define AMRVW_H(i, j) H[(i) * LD + (j)]
    double norm = 0.0;
    if (lane == 0) {   // sequential sum: the value must not depend on the number of lanes
        for (int i = 0; i < M; ++i)
            for (int j = (i > 0 ? i - 1 : 0); j < M; ++j) norm += fabs(AMRVW_H(i, j));
    }
    norm = __shfl_sync(gmask, norm, gbase);
    int en = M - 1, itn = 30 * M;
    double t = 0.0;
    while (en >= 0) {
        int its = 0;
        const int na = en - 1, enm2 = na - 1;
        for (;;) {
            __syncwarp(gmask);
            // highest l in 1..en with a negligible subdiagonal entry, 0 if none
            int l = 0;
            for (int base = (en / GL) * GL; base >= 0 && l == 0; base -= GL) {
                const int lc = base + lane;
                bool small = false;
                if (lc >= 1 && lc <= en) {
                    double s = fabs(AMRVW_H(lc - 1, lc - 1)) + fabs(AMRVW_H(lc, lc));
                    if (s == 0.0) s = norm;
                    const double tst2 = s + fabs(AMRVW_H(lc, lc - 1));
                    small = (tst2 == s);
                }
                const unsigned mask = __ballot_sync(gmask, small) & gmask;
                if (mask) l = base + (31 - __clz(mask)) - gbase;
            }
            double x = AMRVW_H(en, en);
            if (l == en) { if (lane == 0) E[en] = cplx(x + t, 0.0); en = na; break; }
            double y = AMRVW_H(na, na);
            double w = AMRVW_H(en, na) * AMRVW_H(na, en);
            if (l == na) {
                const double p = (y - x) * 0.5;
                const double q = p * p + w;
                double zz = sqrt_(fabs(q));
                x = x + t;
                if (lane == 0) {
                    if (q >= 0.0) {
                        zz = p + (p >= 0.0 ? zz : -zz);
                        E[na] = cplx(x + zz, 0.0);
                        E[en] = E[na];
                        if (zz != 0.0) E[en] = cplx(x - w * rcp_(zz), 0.0);
                    } else {
                        E[na] = cplx(x + p, zz);
                        E[en] = cplx(x + p, -zz);
                    }
                }
                en = enm2;
                break;
            }
            if (itn == 0) return false;
            if (its == 10 || its == 20) {
                t += x;
                __syncwarp(gmask);
                for (int i = lane; i <= en; i += GL) AMRVW_H(i, i) -= x;
                __syncwarp(gmask);
                const double s = fabs(AMRVW_H(en, na)) + fabs(AMRVW_H(na, enm2));
                x = 0.75 * s; y = x; w = -0.4375 * s * s;
            }
            ++its; --itn;
            double p, q, r;
            {
                const double zz = AMRVW_H(l, l);
                const double r0 = x - zz, s0 = y - zz;
                p = (r0 * s0 - w) * rcp_(AMRVW_H(l + 1, l)) + AMRVW_H(l, l + 1);
                q = AMRVW_H(l + 1, l + 1) - zz - r0 - s0;
                r = AMRVW_H(l + 2, l + 1);
                const double is = rcp_(fabs(p) + fabs(q) + fabs(r));
                p = p * is; q = q * is; r = r * is;
            }
            __syncwarp(gmask);
            for (int i = l + 2 + lane; i <= en; i += GL) { AMRVW_H(i, i - 2) = 0.0; if (i != l + 2) AMRVW_H(i, i - 3) = 0.0; }
            for (int k = l; k <= na; ++k) {
                __syncwarp(gmask);
                const bool notlas = (k != na);
                double xs = 0.0;
                if (k != l) {
                    p = AMRVW_H(k, k - 1); q = AMRVW_H(k + 1, k - 1); r = notlas ? AMRVW_H(k + 2, k - 1) : 0.0;
                    xs = fabs(p) + fabs(q) + fabs(r);
                    if (xs == 0.0) continue;
                    const double ix = rcp_(xs);
                    p = p * ix; q = q * ix; r = r * ix;
                }
                double s = sqrt_(p * p + q * q + r * r);
                if (p < 0.0) s = -s;
                __syncwarp(gmask);   // every lane has read column k-1 before it is overwritten
                if (k != l && lane == 0) AMRVW_H(k, k - 1) = -s * xs;
                p = p + s;
                const double is = rcp_(s), ip = rcp_(p);
                const double hx = p * is, hy = q * is, hz = r * is;
                q = q * ip; r = r * ip;
                const int ihi = min(en, k + 3);
                if (notlas) {
                    for (int j = k + lane; j <= en; j += GL) {
                        const double a0 = AMRVW_H(k, j), a1 = AMRVW_H(k + 1, j), a2 = AMRVW_H(k + 2, j);
                        const double pp = a0 + q * a1 + r * a2;
                        AMRVW_H(k, j) = a0 - pp * hx; AMRVW_H(k + 1, j) = a1 - pp * hy; AMRVW_H(k + 2, j) = a2 - pp * hz;
                    }
                    __syncwarp(gmask);
                    for (int i = l + lane; i <= ihi; i += GL) {
                        const double a0 = AMRVW_H(i, k), a1 = AMRVW_H(i, k + 1), a2 = AMRVW_H(i, k + 2);
                        const double pp = hx * a0 + hy * a1 + hz * a2;
                        AMRVW_H(i, k) = a0 - pp; AMRVW_H(i, k + 1) = a1 - pp * q; AMRVW_H(i, k + 2) = a2 - pp * r;
                    }
                } else {
                    for (int j = k + lane; j <= en; j += GL) {
                        const double a0 = AMRVW_H(k, j), a1 = AMRVW_H(k + 1, j);
                        const double pp = a0 + q * a1;
                        AMRVW_H(k, j) = a0 - pp * hx; AMRVW_H(k + 1, j) = a1 - pp * hy;
                    }
                    __syncwarp(gmask);
                    for (int i = l + lane; i <= ihi; i += GL) {
                        const double a0 = AMRVW_H(i, k), a1 = AMRVW_H(i, k + 1);
                        const double pp = hx * a0 + hy * a1;
                        AMRVW_H(i, k) = a0 - pp; AMRVW_H(i, k + 1) = a1 - pp * q;
                    }
                }
            }
        }
    }
#undef AMRVW_H
    __syncwarp(gmask);
    bool ok = true;
    for (int i = lane; i < M; i += GL)
        if (!(isfinite(E[i].re) && isfinite(E[i].im))) ok = false;
    return __all_sync(gmask, ok);
}

// one out-of-line copy: formation + hqr
AMRVW_NI bool dense_shifts(const Rot<double>* q, const Rot<double>* b, const Rot<double>* c, int k0, int stop, double* H, double* isg, cplx* E) {
    const int M = stop - k0 + 1;
    dense_trailing_block(q, b, c, k0, stop, H, isg, M);
    return hqr_eigenvalues(H, M, E);
}

}  // namespace amrvw
