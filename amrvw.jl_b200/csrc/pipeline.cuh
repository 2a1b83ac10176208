// pipeline.cuh -- pipelined (multishift) real-double-shift kernel.  PLACEHOLDER while the
// one-bulge path is brought up on hardware: forwards to the reference-schedule kernel.
#pragma once
#include "kernels.cuh"
#include "../../include/amrvw_b200.h"

namespace amrvw {
static int launch_pipelined_rds(const Problem<double>& pr, const amrvw_options&, int, int, cudaStream_t st, int* launches) {
    const int threads = 64;
    const int blocks = (int)((pr.nbatch + threads - 1) / threads);
    roots_reference_kernel<double><<<blocks, threads, 0, st>>>(pr);
    *launches += 1;
    return (int)cudaGetLastError();
}
}  // namespace amrvw
