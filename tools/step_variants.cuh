// step_variants.cuh -- chase-step loop variants for tools/ubench.cu
#pragma once

AMRVW_DI bool turnover_v2(const Rot<double> q1, const Rot<double> q2, const Rot<double> q3, Rot<double>& r1, Rot<double>& r2, Rot<double>& r3) {
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
#ifdef TURN_KEEP_H5
    const double b = -((s1 * s2) * y) * (2.0 - rho5);
#else
    const double b = -((s1 * s2) * y);
#endif
    const double d6 = fma(b, b, fma(a, a, -1.0));
    const double rho6 = fma(-0.5, d6, 1.0);
    r1 = make_rot<double>(c4, s4);
    r2 = make_rot<double>(u11 * rho5, -(r4 * rho5));
    r3 = make_rot<double>(a * rho6, b * rho6);
    return x >= 1e-280;
}

template <int TURN> AMRVW_DI bool turn(const Rot<double> q1, const Rot<double> q2, const Rot<double> q3, Rot<double>& r1, Rot<double>& r2, Rot<double>& r3) {
    if (TURN == 0) return turnover_fast_try(q1, q2, q3, r1, r2, r3);
    return turnover_v2(q1, q2, q3, r1, r2, r3);
}

AMRVW_DI Rot<double> shfl_up_rot16(const Rot<double> r) {
    Rot<double> o;
    o.c = __shfl_up_sync(0xffffffffu, r.c, 1, 16);
    o.s = __shfl_up_sync(0xffffffffu, r.s, 1, 16);
    return o;
}

#define D2(r) make_double2((r).c, (r).s)
#define CHASE7(TURN, q0, q1, q2, b0, b1, b2, c0, c1, c2, ok)                                         \
    do {                                                                                              \
        Rot<double> t1_, t2_, t3_;                                                                    \
        ok &= turn<TURN>(b1, b2, V, t1_, t2_, t3_); V = t1_; b1 = t2_; b2 = t3_;                       \
        ok &= turn<TURN>(c2, c1, V, t1_, t2_, t3_); V = t1_; c2 = t2_; c1 = t3_;                       \
        ok &= turn<TURN>(b0, b1, U, t1_, t2_, t3_); U = t1_; b0 = t2_; b1 = t3_;                       \
        ok &= turn<TURN>(c1, c0, U, t1_, t2_, t3_); U = t1_; c1 = t2_; c0 = t3_;                       \
        ok &= turn<TURN>(q1, q2, V, t1_, t2_, t3_); V = t1_; q1 = t2_; q2 = t3_;                       \
        ok &= turn<TURN>(q0, q1, U, t1_, t2_, t3_); U = t1_; q0 = t2_; q1 = t3_;                       \
        ok &= turn<TURN>(W, V, U, t1_, t2_, t3_); V = t1_; U = t2_; W = t3_;                           \
    } while (0)

#define SNAPSHOT(sn, NT, q0, q1, q2, b0, b1, b2, c0, c1, c2)                                             \
    do {                                                                                                  \
        (sn)[0 * (NT)] = D2(q0); (sn)[1 * (NT)] = D2(q1); (sn)[2 * (NT)] = D2(q2);                         \
        (sn)[3 * (NT)] = D2(b0); (sn)[4 * (NT)] = D2(b1); (sn)[5 * (NT)] = D2(b2);                         \
        (sn)[6 * (NT)] = D2(c0); (sn)[7 * (NT)] = D2(c1); (sn)[8 * (NT)] = D2(c2);                         \
        (sn)[9 * (NT)] = D2(U); (sn)[10 * (NT)] = D2(V); (sn)[11 * (NT)] = D2(W);                          \
    } while (0)

// one full systolic step with role rotation: slot A (finished) is handed to the next lane and refilled
#define SYS_STEP(TURN, SNAP, A_q, B_q, C_q, A_b, B_b, C_b, A_c, B_c, C_c, tt)                           \
    do {                                                                                                  \
        Rot<double> iq = shfl_up_rot16(A_q), ib = shfl_up_rot16(A_b), ic = shfl_up_rot16(A_c);           \
        if (lane == L - 1) { outp[0] = D2(A_q); outp[1] = D2(A_b); outp[2] = D2(A_c); }                   \
        if (lane == 0) {                                                                                  \
            const double2* sg = feed + ((tt) & 63) * 3;                                                   \
            const double2 vq = sg[0], vb = sg[1], vc = sg[2];                                             \
            iq = make_rot<double>(vq.x, vq.y); ib = make_rot<double>(vb.x, vb.y); ic = make_rot<double>(vc.x, vc.y); \
        }                                                                                                 \
        A_q = iq; A_b = ib; A_c = ic;                                                                     \
        if (SNAP) SNAPSHOT(snap + threadIdx.x, NT, B_q, C_q, A_q, B_b, C_b, A_b, B_c, C_c, A_c);          \
        CHASE7(TURN, B_q, C_q, A_q, B_b, C_b, A_b, B_c, C_c, A_c, ok);                                    \
    } while (0)

template <int UNROLL3, int TURN, int SNAP, int NT>
__global__ void __launch_bounds__(NT, 1) step_kernel(double* out, long long* cyc, int nsteps, const double2* feed_g, double2* sink) {
    constexpr int L = 16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* feed = reinterpret_cast<double2*>(smem_raw);          // [64][3]
    double2* snap = feed + 64 * 3;                                    // [12][NT]
    for (int i = threadIdx.x; i < 64 * 3; i += NT) feed[i] = feed_g[i];
    __syncthreads();
    const int lane = threadIdx.x & (L - 1);
    double2* outp = sink + (size_t)(blockIdx.x * NT + threadIdx.x) * 4;
    Rot<double> q0, q1, q2, b0, b1, b2, c0, c1, c2, U, V, W;
    {
        const int k = threadIdx.x % 61;
        auto R = [&](int j) { const double2 v = feed[(k + j) % 192]; return make_rot<double>(v.x, v.y); };
        q0 = R(0); q1 = R(1); q2 = R(2); b0 = R(3); b1 = R(4); b2 = R(5); c0 = R(6); c1 = R(7); c2 = R(8); U = R(9); V = R(10); W = R(11);
    }
    bool ok = true;
    __syncthreads();
    const long long t0 = clock64();
    if (UNROLL3) {
        for (int t = 0; t < nsteps; t += 3) {
            SYS_STEP(TURN, SNAP, q0, q1, q2, b0, b1, b2, c0, c1, c2, t);
            SYS_STEP(TURN, SNAP, q1, q2, q0, b1, b2, b0, c1, c2, c0, t + 1);
            SYS_STEP(TURN, SNAP, q2, q0, q1, b2, b0, b1, c2, c0, c1, t + 2);
        }
    } else {
        for (int t = 0; t < nsteps; ++t) {
            Rot<double> iq = shfl_up_rot16(q0), ib = shfl_up_rot16(b0), ic = shfl_up_rot16(c0);
            if (lane == L - 1) { outp[0] = D2(q0); outp[1] = D2(b0); outp[2] = D2(c0); }
            if (lane == 0) {
                const double2* sg = feed + (t & 63) * 3;
                const double2 vq = sg[0], vb = sg[1], vc = sg[2];
                iq = make_rot<double>(vq.x, vq.y); ib = make_rot<double>(vb.x, vb.y); ic = make_rot<double>(vc.x, vc.y);
            }
            q0 = q1; q1 = q2; q2 = iq; b0 = b1; b1 = b2; b2 = ib; c0 = c1; c1 = c2; c2 = ic;
            if (SNAP) SNAPSHOT(snap + threadIdx.x, NT, q0, q1, q2, b0, b1, b2, c0, c1, c2);
            CHASE7(TURN, q0, q1, q2, b0, b1, b2, c0, c1, c2, ok);
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    out[blockIdx.x * NT + threadIdx.x] = q0.c + q1.s + q2.c + b0.s + b1.c + b2.s + c0.c + c1.s + c2.c + U.c + V.s + W.c + (ok ? 0.0 : 1e300);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int UNROLL3, int TURN, int SNAP, int NT>
static void run_one(int sms, double* out, long long* cyc, const double2* feed_g, double2* sink, const char* name) {
    const int nsteps = 3000;
    const size_t smem = 64 * 3 * sizeof(double2) + (size_t)12 * NT * sizeof(double2);
    auto k = step_kernel<UNROLL3, TURN, SNAP, NT>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k));
    std::vector<long long> h(sms);
    std::vector<double> ho(NT);
    for (int rep = 0; rep < 2; ++rep) { k<<<sms, NT, smem>>>(out, cyc, nsteps, feed_g, sink); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(h.data(), cyc, 8 * sms, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ho.data(), out, 8 * NT, cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
    double chk = 0; for (int i = 0; i < NT; ++i) chk += ho[i];
    const double cps = (double)mx / nsteps;
    printf("step %-28s warps/SMSP=%d regs=%3d local=%4zu: %8.1f cycles/step/warp, %6.1f cycles per step per SMSP-warp-slot, chk %.6f\n", name, NT / 128, fa.numRegs,
           (size_t)fa.localSizeBytes, cps, cps / (NT / 128), chk);
}

template <int UNROLL3, int TURN, int SNAP>
static void run_variant(int sms, double* out, long long* cyc, const double2* feed_g, double2* sink, const char* name) {
    run_one<UNROLL3, TURN, SNAP, 128>(sms, out, cyc, feed_g, sink, name);
    run_one<UNROLL3, TURN, SNAP, 256>(sms, out, cyc, feed_g, sink, name);
    run_one<UNROLL3, TURN, SNAP, 384>(sms, out, cyc, feed_g, sink, name);
    run_one<UNROLL3, TURN, SNAP, 512>(sms, out, cyc, feed_g, sink, name);
}

static void run_step_bench(int sms, double* out, long long* cyc) {
    std::vector<double2> feed(192);
    unsigned long long s = 88172645463325252ull;
    for (int i = 0; i < 192; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const double ang = (double)(s >> 11) * (1.0 / 9007199254740992.0) * 6.283185307179586;
        feed[i] = make_double2(cos(ang), sin(ang));
    }
    double2* feed_g; double2* sink;
    CK(cudaMalloc(&feed_g, 192 * sizeof(double2))); CK(cudaMemcpy(feed_g, feed.data(), 192 * sizeof(double2), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&sink, (size_t)sms * 512 * 4 * sizeof(double2)));
    run_variant<0, 0, 1>(sms, out, cyc, feed_g, sink, "current(move,fast_try,snap)");
    run_variant<1, 0, 1>(sms, out, cyc, feed_g, sink, "unroll3,fast_try,snap");
    run_variant<1, 1, 1>(sms, out, cyc, feed_g, sink, "unroll3,v2,snap");
    run_variant<1, 1, 0>(sms, out, cyc, feed_g, sink, "unroll3,v2,nosnap");
    run_variant<0, 1, 0>(sms, out, cyc, feed_g, sink, "move,v2,nosnap");
}
