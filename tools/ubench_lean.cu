// ubench_lean.cu -- times the product's lean_chase() (pipeline.cuh) in isolation: every group of 16
// lanes streams its own chain of random rotators from global memory, 1..4 warps per SMSP.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../amrvw.jl_b200/csrc/pipeline.cuh"
using namespace amrvw;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

template <int L, int NT>
__global__ void __launch_bounds__(NT, 1) lean_kernel(Rot<double>* chains, int len, int nsteps, long long* cyc, double* out, int* dstack) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* snap = reinterpret_cast<double2*>(smem_raw);            // [12][NT]
    double2* stage_all = snap + 12 * NT;                               // per group [RING][3]
    const int lane = threadIdx.x % L, grp = threadIdx.x / L;
    const int ggrp = blockIdx.x * (NT / L) + grp;
    Rot<double>* Q = chains + (size_t)(3 * ggrp + 0) * len;
    Rot<double>* B = chains + (size_t)(3 * ggrp + 1) * len;
    Rot<double>* Ct = chains + (size_t)(3 * ggrp + 2) * len;
    double2* stage = stage_all + grp * 3 * AMRVW_RING;
    StepState z;
    { Rot<double> r = Q[threadIdx.x % 97 + 1]; z.q0 = z.q1 = z.q2 = r; r = B[threadIdx.x % 89 + 1]; z.b0 = z.b1 = z.b2 = r; r = Ct[threadIdx.x % 83 + 1]; z.c0 = z.c1 = z.c2 = r;
      z.U = Q[5 + lane]; z.V = B[7 + lane]; z.W = Ct[9 + lane]; }
    const int start = 3 * L + 8, t0 = 3 * L;
    if (lane == 0) {
        for (int a = 0; a < AMRVW_RING - 1; ++a) {
            cp_async16(stage + 3 * a + 0, Q + start + t0 + a); cp_async16(stage + 3 * a + 1, B + start + t0 + a); cp_async16(stage + 3 * a + 2, Ct + start + t0 + a);
            cp_async_commit();
        }
    }
    __syncthreads();
    const int role = 1 | (lane == 0 ? 2 : 0) | (lane == L - 1 ? 4 : 0);
    const long long c0 = clock64();
    int sp = lean_chase<L, NT>(z, Q, B, Ct, (unsigned)__cvta_generic_to_shared(stage), (unsigned)__cvta_generic_to_shared(snap + threadIdx.x), dstack + (size_t)ggrp * len, 0, start, t0,
                               t0 + nsteps, role);
    const long long c1 = clock64();
    if (lane == 0) cp_async_wait_all();
    out[blockIdx.x * NT + threadIdx.x] = z.q0.c + z.b1.s + z.c2.c + z.U.c + z.V.s + z.W.c + sp;
    if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
    if (blockIdx.x == 3 && (threadIdx.x & 31) == 0) { unsigned wid, sm; asm("mov.u32 %0, %%warpid;" : "=r"(wid)); asm("mov.u32 %0, %%smid;" : "=r"(sm)); printf("  block 3 warp %d: smid %u warpid %u\n", threadIdx.x / 32, sm, wid); }
}

template <int L, int NT> static void run(int sms, Rot<double>* chains, int len, int nsteps, long long* cyc, double* out, int* dstack) {
    const size_t smem = (size_t)12 * NT * sizeof(double2) + (size_t)(NT / L) * 3 * AMRVW_RING * sizeof(double2);
    auto k = lean_kernel<L, NT>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k));
    std::vector<long long> h(sms);
    for (int rep = 0; rep < 2; ++rep) { k<<<sms, NT, smem>>>(chains, len, nsteps, cyc, out, dstack); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(h.data(), cyc, 8 * sms, cudaMemcpyDeviceToHost));
    long long mx = 0, sum = 0; for (int i = 0; i < sms; ++i) { mx = h[i] > mx ? h[i] : mx; sum += h[i]; }
    printf("lean_chase<L=%d> NT=%d regs=%d: mean %.1f max %.1f cycles/step/warp\n", L, NT, fa.numRegs, (double)sum / sms / nsteps, (double)mx / nsteps);
}

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount, len = 3100, nsteps = 2900;
    const int maxgroups = sms * (512 / 16);
    std::vector<Rot<double>> h((size_t)maxgroups * 3 * len);
    unsigned long long s = 88172645463325252ull;
    for (size_t i = 0; i < h.size(); ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const double ang = (double)(s >> 11) * (1.0 / 9007199254740992.0) * 6.283185307179586;
        h[i].c = cos(ang); h[i].s = sin(ang);
    }
    Rot<double>* chains; CK(cudaMalloc(&chains, h.size() * sizeof(Rot<double>)));
    CK(cudaMemcpy(chains, h.data(), h.size() * sizeof(Rot<double>), cudaMemcpyHostToDevice));
    long long* cyc; double* out; int* dstack;
    CK(cudaMalloc(&cyc, 8 * sms)); CK(cudaMalloc(&out, 8 * sms * 512)); CK(cudaMalloc(&dstack, (size_t)maxgroups * len * sizeof(int)));
    run<16, 64>(sms, chains, len, nsteps, cyc, out, dstack);
    run<16, 32>(sms, chains, len, nsteps, cyc, out, dstack);
    run<16, 128>(sms, chains, len, nsteps, cyc, out, dstack);
    run<16, 256>(sms, chains, len, nsteps, cyc, out, dstack);
    run<16, 384>(sms, chains, len, nsteps, cyc, out, dstack);
    run<16, 512>(sms, chains, len, nsteps, cyc, out, dstack);
    run<32, 128>(sms, chains, len, nsteps, cyc, out, dstack);
    run<32, 256>(sms, chains, len, nsteps, cyc, out, dstack);
    return 0;
}
