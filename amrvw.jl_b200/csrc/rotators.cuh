// rotators.cuh -- scalar rotator kernels of the core-chasing path, register-resident FP64.
//
// Device-side counterparts of the reference's L0 layer:
//   Rotator{T,S}            src/rotators.jl:80-84      -> Rot<S> (index is implicit in the slot)
//   givensrot               src/transformations.jl:17-33 (+ Julia stdlib givensAlgorithm = LAPACK xLARTG)
//   approx_givensrot        src/transformations.jl:57-64
//   polish_givens           src/transformations.jl:67-73
//   fuse                    src/transformations.jl:80-118
//   _turnover / turnover    src/transformations.jl:158-229
// Everything here is plain FP64 on the vector pipe: a turnover is a dependent chain of ~40
// DFMA-class operations with one sqrt and three divisions; there is no contraction to map to
// tensor cores.  Functions are __host__ __device__ free, __forceinline__, and carry no memory
// traffic of their own.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace amrvw {

#define AMRVW_DI __device__ __forceinline__
#define AMRVW_NI __device__ __noinline__

// ---------------------------------------------------------------- complex scalar (ComplexF64)
struct cplx {
    double re, im;
    AMRVW_DI cplx() {}
    AMRVW_DI cplx(double r) : re(r), im(0.0) {}
    AMRVW_DI cplx(double r, double i) : re(r), im(i) {}
};
AMRVW_DI cplx operator+(cplx a, cplx b) { return cplx(a.re + b.re, a.im + b.im); }
AMRVW_DI cplx operator-(cplx a, cplx b) { return cplx(a.re - b.re, a.im - b.im); }
AMRVW_DI cplx operator-(cplx a) { return cplx(-a.re, -a.im); }
AMRVW_DI cplx operator*(cplx a, cplx b) { return cplx(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
AMRVW_DI cplx operator*(cplx a, double s) { return cplx(a.re * s, a.im * s); }
AMRVW_DI cplx operator*(double s, cplx a) { return cplx(s * a.re, s * a.im); }
AMRVW_DI cplx operator/(cplx a, double s) { return cplx(a.re / s, a.im / s); }
AMRVW_DI cplx operator-(cplx a, double s) { return cplx(a.re - s, a.im); }
AMRVW_DI cplx conj_(cplx a) { return cplx(a.re, -a.im); }
AMRVW_DI double conj_(double a) { return a; }
AMRVW_DI double real_(cplx a) { return a.re; }
AMRVW_DI double real_(double a) { return a; }
AMRVW_DI bool iszero_(cplx a) { return a.re == 0.0 && a.im == 0.0; }
AMRVW_DI bool iszero_(double a) { return a == 0.0; }
AMRVW_DI double abs_(cplx a) { return hypot(a.re, a.im); }
AMRVW_DI double abs2_(cplx a) { return a.re * a.re + a.im * a.im; }
AMRVW_DI double sign_(double x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : x); }

// complex / complex, Smith's algorithm with the Baudin-Smith corner cases (what Julia's Base does
// in the normal range; the reference divides complex numbers only in eigen_values and the pencil
// entries, diagonal-block.jl:127-142, pencil.jl:29-45).
AMRVW_DI double cdiv_part(double a, double b, double c, double d, double r, double t) {
    if (r != 0.0) {
        double br = b * r;
        return br != 0.0 ? (a + br) * t : a * t + (b * t) * r;
    }
    return (a + d * (b / c)) * t;
}
AMRVW_DI cplx operator/(cplx z, cplx w) {
    double p, q;
    if (fabs(w.im) <= fabs(w.re)) {
        double r = w.im / w.re, t = 1.0 / (w.re + w.im * r);
        p = cdiv_part(z.re, z.im, w.re, w.im, r, t);
        q = cdiv_part(z.im, -z.re, w.re, w.im, r, t);
    } else {
        double r = w.re / w.im, t = 1.0 / (w.im + w.re * r);
        p = cdiv_part(z.im, z.re, w.im, w.re, r, t);
        q = -cdiv_part(z.re, -z.im, w.im, w.re, r, t);
    }
    return cplx(p, q);
}
// principal square root
AMRVW_DI cplx csqrt_(cplx z) {
    double x = z.re, y = z.im;
    if (x == 0.0 && y == 0.0) return cplx(0.0, y);
    double rho = sqrt((hypot(x, y) + fabs(x)) / 2);
    double xi = rho, eta = y;
    if (rho != 0.0) {
        eta = (eta / rho) / 2;
        if (x < 0.0) {
            xi = fabs(eta);
            eta = copysign(rho, y);
        }
    }
    return cplx(xi, eta);
}

template <class S> struct is_real { static constexpr bool value = false; };
template <> struct is_real<double> { static constexpr bool value = true; };

// ---------------------------------------------------------------- rotator
// [c s; -s conj(c)] with s real.  Rot<double> is 16 bytes (one LDG.128), Rot<cplx> is padded to
// 32 bytes so it moves as two 16-byte transactions.
template <class S> struct Rot;
template <> struct __align__(16) Rot<double> {
    double c, s;
};
template <> struct __align__(16) Rot<cplx> {
    cplx c;
    double s, pad;
};
template <class S> AMRVW_DI Rot<S> make_rot(S c, double s) {
    Rot<S> r;
    r.c = c;
    r.s = s;
    return r;
}
template <class S> AMRVW_DI Rot<S> adjoint(Rot<S> u) { return make_rot<S>(conj_(u.c), -u.s); }

// ---------------------------------------------------------------- givens
// xLARTG scaling thresholds (floatmin2(Float64) in Julia's stdlib)
#define AMRVW_SAFMN2 0x1p-485
#define AMRVW_SAFMX2 0x1p+485
#define AMRVW_SAFMIN 2.2250738585072014e-308

// real: [c s; -s c] [a; b] = [r; 0], r >= 0   (transformations.jl:17-21 around dlartg)
AMRVW_DI void givensrot(double a, double b, double& c, double& s, double& r) {
    double cs, sn, rr;
    if (b == 0.0) {
        cs = 1.0; sn = 0.0; rr = a;
    } else if (a == 0.0) {
        cs = 0.0; sn = 1.0; rr = b;
    } else {
        double f1 = a, g1 = b;
        double scale = fmax(fabs(f1), fabs(g1));
        if (scale >= AMRVW_SAFMX2 || scale <= AMRVW_SAFMN2) {  // cold: far outside the working range
            int count = 0;
            const double down = scale >= AMRVW_SAFMX2 ? AMRVW_SAFMN2 : AMRVW_SAFMX2;
            const bool big = scale >= AMRVW_SAFMX2;
            do {
                ++count; f1 *= down; g1 *= down; scale = fmax(fabs(f1), fabs(g1));
            } while (big ? (scale >= AMRVW_SAFMX2 && count < 20) : (scale <= AMRVW_SAFMN2));
            rr = sqrt(f1 * f1 + g1 * g1); cs = f1 / rr; sn = g1 / rr;
            const double up = big ? AMRVW_SAFMX2 : AMRVW_SAFMN2;
            for (int i = 0; i < count; ++i) rr *= up;
        } else {
            rr = sqrt(f1 * f1 + g1 * g1); cs = f1 / rr; sn = g1 / rr;
        }
        if (fabs(a) > fabs(b) && cs < 0.0) { cs = -cs; sn = -sn; rr = -rr; }
    }
    const double sg = rr >= 0.0 ? 1.0 : -1.0;
    c = sg * cs; s = sg * sn; r = sg * rr;
}

// complex: G, r = givens(b, a); c = G.s, s = real(G.c)   (transformations.jl:28-31 around zlartg)
AMRVW_DI double abs1_(cplx z) { return fmax(fabs(z.re), fabs(z.im)); }
AMRVW_DI void givensrot(cplx a, cplx b, cplx& c, double& s, cplx& r) {
    // f = b, g = a in xLARTG's naming
    double scale = fmax(abs1_(b), abs1_(a));
    cplx fs = b, gs = a;
    int count = 0;
    if (scale >= AMRVW_SAFMX2) {
        do { ++count; fs = fs * AMRVW_SAFMN2; gs = gs * AMRVW_SAFMN2; scale *= AMRVW_SAFMN2; } while (scale >= AMRVW_SAFMX2 && count < 20);
    } else if (scale <= AMRVW_SAFMN2) {
        if (iszero_(a)) { s = 1.0; c = cplx(0.0, 0.0); r = b; return; }
        do { --count; fs = fs * AMRVW_SAFMX2; gs = gs * AMRVW_SAFMX2; scale *= AMRVW_SAFMX2; } while (scale <= AMRVW_SAFMN2);
    }
    const double f2 = abs2_(fs), g2 = abs2_(gs);
    if (f2 <= fmax(g2, 1.0) * AMRVW_SAFMIN) {  // f negligible
        if (iszero_(b)) {
            s = 0.0;
            r = cplx(hypot(a.re, a.im), 0.0);
            const double d = hypot(gs.re, gs.im);
            c = cplx(gs.re / d, -gs.im / d);
            return;
        }
        const double f2s = hypot(fs.re, fs.im);
        const double g2s = sqrt(g2);
        s = f2s / g2s;
        cplx ff;
        if (abs1_(b) > 1.0) {
            const double d = hypot(b.re, b.im);
            ff = cplx(b.re / d, b.im / d);
        } else {
            const double dr = AMRVW_SAFMX2 * b.re, di = AMRVW_SAFMX2 * b.im;
            const double d = hypot(dr, di);
            ff = cplx(dr / d, di / d);
        }
        c = ff * cplx(gs.re / g2s, -gs.im / g2s);
        r = s * b + c * a;
    } else {
        const double f2s = sqrt(1.0 + g2 / f2);
        r = cplx(f2s * fs.re, f2s * fs.im);
        s = 1.0 / f2s;
        const double d = f2 + g2;
        c = cplx(r.re / d, r.im / d) * conj_(gs);
        if (count > 0) for (int i = 0; i < count; ++i) r = r * AMRVW_SAFMX2;
        if (count < 0) for (int i = 0; i < -count; ++i) r = r * AMRVW_SAFMN2;
    }
}
AMRVW_DI void givensrot(cplx a, double b, cplx& c, double& s, cplx& r) { givensrot(a, cplx(b, 0.0), c, s, r); }

// first-order renormalisation, avoids a sqrt when |a|^2 + b^2 ~ 1  (transformations.jl:57-64)
AMRVW_DI void approx_givensrot(double a, double b, double& c, double& s) {
    const double delta = fabs(a * a) + b * b - 1.0;
    const double r = 1.0 - delta / 2.0;
    c = a * r; s = b * r;
}
AMRVW_DI void approx_givensrot(cplx a, double b, cplx& c, double& s) {
    const cplx aa = a * conj_(a);
    const double delta = hypot(aa.re, aa.im) + b * b - 1.0;
    const double r = 1.0 - delta / 2.0;
    c = conj_(a * r); s = b * r;
}
AMRVW_DI void polish(double& c, double& s) {  // transformations.jl:67-73
    const double delta = c * c + s * s - 1.0;
    const double r = 1.0 - delta / 2.0;
    c = r * c; s = r * s;
}
AMRVW_DI void polish(cplx& c, double& s) {
    const double delta = real_(conj_(c) * c) + s * s - 1.0;
    const double r = 1.0 - delta / 2.0;
    c = r * c; s = r * s;
}

// ---------------------------------------------------------------- fuse (transformations.jl:80-118)
AMRVW_DI Rot<double> fuse(Rot<double> a, Rot<double> b) {
    double u = a.c * b.c - a.s * b.s;
    double v = a.c * b.s + a.s * b.c;
    polish(u, v);
    return make_rot<double>(u, v);
}
// complex: a*b = (ab) * Diag(conj alpha) at the rotator's index; `dphase` = conj(alpha)
AMRVW_DI Rot<cplx> fuse(Rot<cplx> a, Rot<cplx> b, cplx& dphase) {
    const cplx u = a.c * b.c - a.s * b.s;
    const cplx v = a.c * b.s + a.s * conj_(b.c);
    double s = abs_(v);
    const cplx alpha = iszero_(v) ? cplx(1.0, 0.0) : v / s;
    cplx c = u * alpha;
    polish(c, s);
    dphase = conj_(alpha);
    return make_rot<cplx>(c, s);
}

// ---------------------------------------------------------------- turnover (transformations.jl:158-229)
// Q1@i Q2@(i+-1) Q3@i  ->  R1@(i+-1) R2@i R3@(i+-1), same product.  37 flop real / 94 complex.
//
// Three forms of the same operation:
//   turnover_ieee  the reference's formulas operation by operation (IEEE sqrt and divisions, all
//                  special cases): ~170 dependent FP64-pipe instructions once div/sqrt are expanded.
//   turnover_fast  real case, the common path without any division: 1/r by rsqrt, and the
//                  reference's b = s1 s2 / s5 rewritten with s5 = -r (1 - d/2), 1/(1 - d/2) = 1 + d/2
//                  + O(d^2), d = O(u).  ~45 instructions; anything unusual (zero or denormal norm)
//                  goes to the out-of-line IEEE copy.  Results differ from turnover_ieee in the last
//                  bit or two, and are renormalised by the same first-order correction.
//   turnover       what the serial code calls: ONE out-of-line copy per scalar type (the whole
//                  iteration inlined 30 times over was 560 KB of SASS and ran out of the I-cache).
// -DAMRVW_STRICT (the no-FMA build used by the bit-parity tests) maps everything to turnover_ieee.
template <class S> struct Rot3 { Rot<S> r1, r2, r3; };

template <class S>
AMRVW_DI void turnover_ieee(const Rot<S> q1, const Rot<S> q2, const Rot<S> q3, Rot<S>& r1, Rot<S>& r2, Rot<S>& r3) {
    const S c1 = q1.c, c2 = q2.c, c3 = q3.c;
    const double s1 = q1.s, s2 = q2.s, s3 = q3.s;
    const S u21 = (-c2) * s3 * conj_(c1) - c3 * s1;
    const double u31 = s2 * s3;
    if (iszero_(u21) && u31 == 0.0) {
        if (iszero_(s2 * conj_(c3))) {  // product is diagonal
            r1 = make_rot<S>(S(1.0), 0.0);
            r2 = make_rot<S>(c1 * c3 - c2 * s1 * s3, 0.0);
            r3 = make_rot<S>(c2, 0.0);
            return;
        }
    }
    S c4, r4; double s4;
    givensrot(u21, u31, c4, s4, r4);
    c4 = conj_(c4); s4 = -s4;
    const S u11 = c1 * c3 - c2 * s1 * s3;
    S c5; double s5;
    approx_givensrot(u11, real_(r4), c5, s5);
    c5 = conj_(c5); s5 = -s5;
    S c6; double s6;
    if (s5 != 0.0) {
        const S a = conj_(c1) * s2 * s4 + conj_(c2) * c4;
        const double b = s1 * s2 / s5;
        approx_givensrot(a, b, c6, s6);
    } else {
        S r6;
        givensrot(c2 * conj_(c1), -s2, c6, s6, r6);
    }
    r1 = make_rot<S>(c4, s4);
    r2 = make_rot<S>(c5, s5);
    r3 = make_rot<S>(c6, s6);
}

template <class S> AMRVW_NI Rot3<S> turnover_ieee_call(const Rot<S> q1, const Rot<S> q2, const Rot<S> q3) {
    Rot3<S> o;
    turnover_ieee(q1, q2, q3, o.r1, o.r2, o.r3);
    return o;
}

// 1/sqrt(x) for x in the normal range: MUFU.RSQ64H seed + two Newton steps (no special-case branches;
// callers guard the range)
AMRVW_DI double rsqrt_nr(double x) {
#ifdef AMRVW_STRICT
    return 1.0 / sqrt(x);
#else
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#ifndef AMRVW_NEWTON2
    // one third-order step: y <- y (1 + e/2 + 3 e^2/8), e = 1 - x y^2
    const double t = x * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double q = y * e;
    return fma(q, p, y);
#else
    double t = x * y;
    double e = fma(-t, y, 1.0);
    y = fma(0.5 * y, e, y);
    t = x * y;
    e = fma(-t, y, 1.0);
    y = fma(0.5 * y, e, y);
    return y;
#endif
#endif
}

// reciprocal and square root where the last bit does not matter (shifts, bulge creation): IEEE in
// the strict build (bit parity with the oracle), MUFU seed + Newton otherwise -- an FP64 division
// expands to ~35 dependent instructions, ~300 cycles next to warps that keep the FP64 pipe busy
AMRVW_DI double rcp_(double x) {
#ifdef AMRVW_STRICT
    return 1.0 / x;
#else
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
#endif
}
AMRVW_DI double sqrt_(double x) {
#ifdef AMRVW_STRICT
    return sqrt(x);
#else
    return x > 0.0 ? x * rsqrt_nr(x) : 0.0;
#endif
}

// division-free common path (real); returns false when the caller must take the general path.
// 31 FP64-pipe instructions + one MUFU: every product is contracted by hand (the reference's
// formulas, transformations.jl:162-191, with 1/r by rsqrt).  b = s1 s2 / s5 with s5 = -r (1 - h),
// h = O(u) the first-order norm correction of the second rotator: the factor 1/(1 - h) is dropped,
// a relative perturbation of b of a few ulp that the first-order renormalisation of (a, b) absorbs
// (amrvw_selftest measures the product error of exactly this code against the IEEE path).
AMRVW_DI bool turnover_fast_try(const Rot<double> q1, const Rot<double> q2, const Rot<double> q3, Rot<double>& r1, Rot<double>& r2, Rot<double>& r3) {
    const double c1 = q1.c, s1 = q1.s, c2 = q2.c, s2 = q2.s, c3 = q3.c, s3 = q3.s;
    const double t = c2 * s3;
    const double m1 = c3 * s1;
    const double u21 = fma(-t, c1, -m1);
    const double u31 = s2 * s3;
    const double m2 = c1 * c3;
    const double u11 = fma(-t, s1, m2);
    const double x = fma(u21, u21, u31 * u31);
    const double y = rsqrt_nr(x);
    const double c4 = u21 * y, s4 = -(u31 * y);
    const double r4 = x * y;
    const double d5 = fma(u11, u11, x - 1.0);
    const double rho5 = fma(-0.5, d5, 1.0);
    const double a = fma(c2, c4, (c1 * s2) * s4);
    const double b = -((s1 * s2) * y);
    const double d6 = fma(b, b, fma(a, a, -1.0));
    const double rho6 = fma(-0.5, d6, 1.0);
    r1 = make_rot<double>(c4, s4);
    r2 = make_rot<double>(u11 * rho5, -(r4 * rho5));
    r3 = make_rot<double>(a * rho6, b * rho6);
    // x in [2^-930, inf): tested on the high word in the integer pipe (a DSETP would take an FP64 issue
    // slot in a loop that is bound by that pipe); false for 0, denormals, inf and NaN
    return (unsigned)(__double2hiint(x) - 0x05d00000) < (0x7ff00000u - 0x05d00000u);
}
// Degenerate inputs (u21^2 + u31^2 below 2^-930: the bulge meets rotators whose sines have all but vanished).
// The reference's last step, b = s1 s2 / s5 (transformations.jl:188-191), is a quotient of two such
// quantities: once they reach the denormal range it has a few bits left, (a, b) is no longer a unit vector,
// the first-order renormalisation does not repair it and the chain stops being unitary (backward error 1e-2).
// The reference never gets there -- it deflates at |s| <= eps before the next bulge -- but a fat step passes
// up to 128 bulges through rotators that converged under the earlier ones.  Here the third rotator is read
// off the explicit product instead, M = Q1 Q2 Q3, R3 = (R1 R2)^T M: c6 = s4 M23 + c4 M33 (the reference's a),
// s6 = -(s4 M22 + c4 M32), no division, absolute accuracy O(u) whatever the sizes; the three rotators are
// normalised exactly (scaled givens).
AMRVW_NI Rot3<double> turnover_robust_call(const Rot<double> q1, const Rot<double> q2, const Rot<double> q3) {
    const double c1 = q1.c, s1 = q1.s, c2 = q2.c, s2 = q2.s, c3 = q3.c, s3 = q3.s;
    const double u21 = -(c2 * s3) * c1 - c3 * s1;
    const double u31 = s2 * s3;
    const double u11 = c1 * c3 - (c2 * s1) * s3;
    double c, s, r4, tmp;
    givensrot(u21, u31, c, s, r4);
    const double c4 = c, s4 = -s;
    givensrot(u11, r4, c, s, tmp);
    const double c5 = c, s5 = -s;
    const double m22 = (c1 * c2) * c3 - s1 * s3, m32 = -(s2 * c3), m23 = c1 * s2, m33 = c2;
    const double a = s4 * m23 + c4 * m33;
    const double b = -(s4 * m22 + c4 * m32);
    const double h = sqrt(a * a + b * b);        // a^2 + b^2 = 1 + O(u): no scaling needed
    Rot3<double> o;
    o.r1 = make_rot<double>(c4, s4);
    o.r2 = make_rot<double>(c5, s5);
    o.r3 = h > 0.0 ? make_rot<double>(a / h, b / h) : make_rot<double>(1.0, 0.0);
    return o;
}
AMRVW_DI void turnover_fast(const Rot<double> q1, const Rot<double> q2, const Rot<double> q3, Rot<double>& r1, Rot<double>& r2, Rot<double>& r3) {
#ifdef AMRVW_STRICT
    const Rot3<double> o = turnover_ieee_call<double>(q1, q2, q3);
    r1 = o.r1; r2 = o.r2; r3 = o.r3;
#else
    if (!turnover_fast_try(q1, q2, q3, r1, r2, r3)) {
        const Rot3<double> o = turnover_robust_call(q1, q2, q3);
        r1 = o.r1; r2 = o.r2; r3 = o.r3;
    }
#endif
}
AMRVW_NI Rot3<double> turnover_fast_call(const Rot<double> q1, const Rot<double> q2, const Rot<double> q3) {
    Rot3<double> o;
    turnover_fast(q1, q2, q3, o.r1, o.r2, o.r3);
    return o;
}
// the serial code's turnover: one out-of-line copy
AMRVW_DI void turnover(const Rot<double> q1, const Rot<double> q2, const Rot<double> q3, Rot<double>& r1, Rot<double>& r2, Rot<double>& r3) {
#ifdef AMRVW_STRICT
    const Rot3<double> o = turnover_ieee_call<double>(q1, q2, q3);
#else
    const Rot3<double> o = turnover_fast_call(q1, q2, q3);
#endif
    r1 = o.r1; r2 = o.r2; r3 = o.r3;
}
// division-free common path (complex), the same rewriting as the real one: 1/r by rsqrt, the first-order
// renormalisations as FMAs, b = s1 s2 / s5 without the O(u) factor.  The reference's complex givens
// (transformations.jl:28-31 around zlartg) returns a real s >= 0 and puts the sign of u31 into c:
// c4' = sign(u31) conj(u21)/r, s4' = |u31|/r, r4 = sign(u31) r.  ~62 FP64-pipe instructions against
// ~400 (hypot, complex division, three IEEE divisions and two square roots) on the IEEE path.
AMRVW_DI bool turnover_fast_try(const Rot<cplx> q1, const Rot<cplx> q2, const Rot<cplx> q3, Rot<cplx>& r1, Rot<cplx>& r2, Rot<cplx>& r3) {
    const cplx c1 = q1.c, c2 = q2.c, c3 = q3.c;
    const double s1 = q1.s, s2 = q2.s, s3 = q3.s;
    const double tre = c2.re * s3, tim = c2.im * s3;                                   // t = c2 s3
    const double u21re = -fma(tre, c1.re, fma(tim, c1.im, c3.re * s1));                // u21 = -t conj(c1) - c3 s1
    const double u21im = -fma(tim, c1.re, fma(-tre, c1.im, c3.im * s1));
    const double u31 = s2 * s3;
    const double u11re = fma(c1.re, c3.re, fma(-c1.im, c3.im, -(tre * s1)));           // u11 = c1 c3 - t s1
    const double u11im = fma(c1.re, c3.im, fma(c1.im, c3.re, -(tim * s1)));
    const double x = fma(u21re, u21re, fma(u21im, u21im, u31 * u31));
    const double y = rsqrt_nr(x);
    const double sy = u31 >= 0.0 ? y : -y;                                             // sign(u31) / r
    const cplx c4 = cplx(u21re * sy, u21im * sy);                                      // conj of the givens' c
    const double s4 = -(u31 * sy);                                                     // -|u31| / r
    const double r4 = x * sy;
    const double d5 = fma(u11re, u11re, fma(u11im, u11im, x - 1.0));
    const double rho5 = fma(-0.5, d5, 1.0);
    const double w = s2 * s4;
    const double are = fma(c1.re, w, fma(c2.re, c4.re, c2.im * c4.im));                // a = conj(c1) s2 s4 + conj(c2) c4
    const double aim = fma(-c1.im, w, fma(c2.re, c4.im, -(c2.im * c4.re)));
    const double b = -((s1 * s2) * sy);
    const double d6 = fma(b, b, fma(are, are, fma(aim, aim, -1.0)));
    const double rho6 = fma(-0.5, d6, 1.0);
    r1 = make_rot<cplx>(c4, s4);
    r2 = make_rot<cplx>(cplx(u11re * rho5, u11im * rho5), -(r4 * rho5));
    r3 = make_rot<cplx>(cplx(are * rho6, -(aim * rho6)), b * rho6);
    return (unsigned)(__double2hiint(x) - 0x05d00000) < (0x7ff00000u - 0x05d00000u);
}
// complex twin of turnover_robust_call: M = Q1 Q2 Q3 with G(c, s) = [c s; -s conj(c)],
// conj(c6) = (R1^H M)33 = s4 M23 + c4 M33, s6 = -Re (R1^H M)32 = -Re(s4 M22 + c4 M32) (real up to rounding by
// the phase the complex givens puts into c4); exact normalisations
AMRVW_DI void turnover_robust(const Rot<cplx> q1, const Rot<cplx> q2, const Rot<cplx> q3, Rot<cplx>& r1, Rot<cplx>& r2, Rot<cplx>& r3) {
    const cplx c1 = q1.c, c2 = q2.c, c3 = q3.c;
    const double s1 = q1.s, s2 = q2.s, s3 = q3.s;
    const cplx u21 = (-c2) * s3 * conj_(c1) - c3 * s1;
    const double u31 = s2 * s3;
    const cplx u11 = c1 * c3 - c2 * s1 * s3;
    cplx c, r4c; double s;
    givensrot(u21, u31, c, s, r4c);
    const cplx c4 = conj_(c);
    const double s4 = -s, r4 = r4c.re;
    double h = sqrt(abs2_(u11) + r4 * r4);
    const cplx c5 = h > 0.0 ? u11 / h : cplx(1.0, 0.0);
    const double s5 = h > 0.0 ? -r4 / h : 0.0;
    const cplx m22 = conj_(c1) * c2 * conj_(c3) - s1 * s3, m32 = -(s2 * conj_(c3)), m23 = conj_(c1) * s2, m33 = conj_(c2);
    const cplx a = s4 * m23 + c4 * m33;
    const double b = -(s4 * m22 + c4 * m32).re;
    h = sqrt(abs2_(a) + b * b);
    r1 = make_rot<cplx>(c4, s4);
    r2 = make_rot<cplx>(c5, s5);
    r3 = h > 0.0 ? make_rot<cplx>(conj_(a) / h, b / h) : make_rot<cplx>(cplx(1.0, 0.0), 0.0);
}
AMRVW_NI Rot3<cplx> turnover_robust_call(const Rot<cplx> q1, const Rot<cplx> q2, const Rot<cplx> q3) {
    Rot3<cplx> o;
    turnover_robust(q1, q2, q3, o.r1, o.r2, o.r3);
    return o;
}
AMRVW_NI Rot3<cplx> turnover_fast_call(const Rot<cplx> q1, const Rot<cplx> q2, const Rot<cplx> q3) {
    Rot3<cplx> o;
    if (!turnover_fast_try(q1, q2, q3, o.r1, o.r2, o.r3)) o = turnover_robust_call(q1, q2, q3);
    return o;
}
// inline fast path, out-of-line fall-back: for the systolic lanes (the serial code calls turnover())
AMRVW_DI void turnover_fast(const Rot<cplx> q1, const Rot<cplx> q2, const Rot<cplx> q3, Rot<cplx>& r1, Rot<cplx>& r2, Rot<cplx>& r3) {
#ifdef AMRVW_STRICT
    const Rot3<cplx> o = turnover_ieee_call<cplx>(q1, q2, q3);
    r1 = o.r1; r2 = o.r2; r3 = o.r3;
#else
    if (!turnover_fast_try(q1, q2, q3, r1, r2, r3)) {
        const Rot3<cplx> o = turnover_robust_call(q1, q2, q3);
        r1 = o.r1; r2 = o.r2; r3 = o.r3;
    }
#endif
}
AMRVW_DI void turnover(const Rot<cplx> q1, const Rot<cplx> q2, const Rot<cplx> q3, Rot<cplx>& r1, Rot<cplx>& r2, Rot<cplx>& r3) {
#ifdef AMRVW_STRICT
    const Rot3<cplx> o = turnover_ieee_call<cplx>(q1, q2, q3);
#else
    const Rot3<cplx> o = turnover_fast_call(q1, q2, q3);
#endif
    r1 = o.r1; r2 = o.r2; r3 = o.r3;
}

// ---------------------------------------------------------------- 2x2 eigenvalues
// real: Kahan-style qdrtc(b, c) (utils.jl:130-140 via diagonal-block.jl:117-124): either two reals
// with imaginary part exactly 0 or an exact conjugate pair.
AMRVW_DI void eig2x2(const double A[4], cplx& e1, cplx& e2) {
    const double b = (A[0] + A[3]) / 2;
    const double c = A[0] * A[3] - A[1] * A[2];
    const double d = b * b - c;
    if (d <= 0.0) {
        const double im = sqrt(-d);
        e1 = cplx(b, im); e2 = cplx(b, -im);
    } else {
        const double r = sqrt(d) * (sign_(b) + (b == 0.0 ? 1.0 : 0.0)) + b;
        e1 = cplx(r, 0.0); e2 = cplx(c / r, 0.0);
    }
}
// complex (diagonal-block.jl:127-142)
AMRVW_DI void eig2x2(const cplx A[4], cplx& e1, cplx& e2) {
    const cplx tr = A[0] + A[3];
    const cplx detm = A[0] * A[3] - A[2] * A[1];
    const cplx disc = csqrt_(tr * tr - 4.0 * detm);
    const cplx u = abs_(tr + disc) > abs_(tr - disc) ? tr + disc : tr - disc;
    if (iszero_(u)) { e1 = cplx(0.0, 0.0); e2 = cplx(0.0, 0.0); return; }
    e1 = u / 2.0;
    e2 = detm / e1;
}

// ---------------------------------------------------------------- counter-based RNG (exceptional shifts)
AMRVW_DI uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
AMRVW_DI double uniform01(uint64_t seed, uint64_t poly, uint64_t k) {
    const uint64_t h = splitmix64(seed ^ splitmix64(poly * 0x9E3779B97F4A7C15ull + k));
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

}  // namespace amrvw
