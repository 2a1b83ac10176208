// ubench.cu -- development micro-benchmarks for the FP64 pipe of sm_100a (not part of the product):
// dependent-issue latencies of DFMA/DMUL/DADD/MUFU.RSQ64H/SHFL, DFMA throughput against warps per
// SMSP and ILP, and the cycles one systolic chase step (7 division-free turnovers + 6 shuffles) takes
// at 1..4 warps per SMSP.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../amrvw.jl_b200/csrc/rotators.cuh"
using namespace amrvw;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

template <int OP> __global__ void lat_kernel(double* out, long long* cyc, int n, double a, double b) {
    double x = a + threadIdx.x * 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (OP == 0) x = fma(x, b, a);
            if (OP == 1) x = x * b;
            if (OP == 2) x = x + b;
            if (OP == 3) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y; }
            if (OP == 4) x = __shfl_up_sync(0xffffffffu, x, 1, 16);
            if (OP == 5) x = rsqrt_nr(x);
            if (OP == 6) { x = __longlong_as_double(__double_as_longlong(x) ^ 1ll); }   // int op on the pair
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// throughput: ILP independent DFMA chains per thread
template <int ILP> __global__ void thr_kernel(double* out, long long* cyc, int n, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) x[j] = a + threadIdx.x * 1e-9 + j;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], b, a);
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---------------------------------------------------------------------------------------------
// chase-step loop: every lane keeps a 3-index window of 3 chains + bulge, ingests by shfl_up.
#ifndef STEP_VARIANT
#define STEP_VARIANT 0
#endif
#include "step_variants.cuh"

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
    printf("device %s, %d SMs, clock %d kHz\n", pr.name, pr.multiProcessorCount, pr.clockRate);
    double* out; long long* cyc; CK(cudaMalloc(&out, 1 << 24)); CK(cudaMalloc(&cyc, 1 << 16));
    std::vector<long long> h(4096);
    const int n = 4096;
    const char* names[] = {"DFMA", "DMUL", "DADD", "MUFU.RSQ64H(+glue)", "SHFL.UP f64", "rsqrt_nr", "LOP pair"};
#define LAT(OP) do { lat_kernel<OP><<<1, 32>>>(out, cyc, n, 1.0000001, 0.9999999); CK(cudaDeviceSynchronize()); lat_kernel<OP><<<1, 32>>>(out, cyc, n, 1.0000001, 0.9999999); CK(cudaDeviceSynchronize()); \
        CK(cudaMemcpy(h.data(), cyc, 8, cudaMemcpyDeviceToHost)); printf("latency %-20s %.2f cycles\n", names[OP], (double)h[0] / (n * 16.0)); } while (0)
    LAT(0); LAT(1); LAT(2); LAT(3); LAT(4); LAT(5); LAT(6);
    // throughput per SM: warps per SM = 4 * wps
#define THR(ILP) do { for (int wps = 1; wps <= 8; wps *= 2) { int thr = 128 * wps; if (thr > 1024) break; \
        thr_kernel<ILP><<<pr.multiProcessorCount, thr>>>(out, cyc, 2048, 1.0000001, 0.9999999); CK(cudaDeviceSynchronize()); \
        CK(cudaMemcpy(h.data(), cyc, 8 * pr.multiProcessorCount, cudaMemcpyDeviceToHost)); long long mx = 0; for (int i = 0; i < pr.multiProcessorCount; ++i) mx = h[i] > mx ? h[i] : mx; \
        double inst_per_smsp = 2048.0 * 8 * ILP * wps; printf("DFMA thr ILP=%d warps/SMSP=%d: %.3f cycles per warp-inst per SMSP (%.1f%% of 2-cycle peak)\n", ILP, wps, mx / inst_per_smsp, 200.0 * inst_per_smsp / mx); } } while (0)
    THR(1); THR(2); THR(4); THR(8);
    run_step_bench(pr.multiProcessorCount, out, cyc);
    return 0;
}
