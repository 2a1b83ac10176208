/* amrvw_b200.h -- C ABI of the B200-native batched core-chasing root finder.
 *
 * Drop-in boundary for ONE path of jverzani/AMRVW.jl: `AMRVW.roots` applied to many independent
 * polynomials.  The reference has no FFI layer; its boundary is the Julia method table
 *     roots(ps::Vector{S})                       src/companion.jl:252-272
 *     roots(vs::Vector{S}, ws::Vector{S})        src/companion.jl:282-303
 * and "a batch" is a user comprehension over it (README.md:121).  A maintainer adds one batched
 * method roots(::AbstractVector{<:AbstractVector{T}}) that `ccall`s the entry points below
 * (INTEGRATION.md shows the Julia stub; amrvw.jl_b200/roots.py is the same binding via ctypes).
 *
 * Conventions
 *   - Polynomials are ragged, CSR style: polynomial p owns coeffs[offsets[p] .. offsets[p+1]),
 *     lowest degree first (a0 .. an), exactly the vector AMRVW.roots takes.  Complex data is
 *     interleaved (re, im), bit-compatible with Julia's Vector{ComplexF64}.
 *   - Roots are interleaved (re, im) doubles: polynomial p's roots start at complex slot offsets[p]
 *     (the root buffer has offsets[nbatch] complex slots, one per input entry: a polynomial with len
 *     coefficients has at most len-1 roots, a pencil of order N_p has N_p; empty vectors own nothing).
 *     nroots[p] says how many were written: like the reference, leading/trailing zero coefficients
 *     are trimmed (utils.jl:14-58), degree <= 2 goes through the closed forms (utils.jl:63-140),
 *     roots come back sorted by (real, imag) (AMRVW_algorithm.jl:134) followed by the K zero roots
 *     of the trimmed low end (companion.jl:266-268).
 *   - Non-convergence is not an error (AMRVW_algorithm.jl:23-30 leaves the missing roots at 0):
 *     n_unconverged[p] counts the slots left at zero.
 *   - Every function returns 0 on success, a negative amrvw_status otherwise; the message is
 *     available from amrvw_last_error.  No exceptions cross the boundary.  A context is bound to
 *     one CUDA device and is not thread-safe (one per host thread), like the reference's
 *     single-threaded state.
 *   - There is no CPU fallback: without a CUDA device amrvw_create fails.
 */
#ifndef AMRVW_B200_H
#define AMRVW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct amrvw_ctx amrvw_ctx;

enum amrvw_status {
    AMRVW_OK = 0,
    AMRVW_ERR_ARGUMENT = -1,
    AMRVW_ERR_CUDA = -2,
    AMRVW_ERR_NO_DEVICE = -3,
    AMRVW_ERR_QUEUE_TIMEOUT = -4   /* the device-side unit queue gave up waiting (~10 s): results incomplete */
};

enum amrvw_schedule {
    AMRVW_SCHEDULE_AUTO = 0,       /* pipelined multishift for Float64 RDS, one-bulge otherwise */
    AMRVW_SCHEDULE_REFERENCE = 1,  /* one bulge at a time, shifts from the trailing 2x2 block: the
                                      reference's own order of operations (bulge-step.jl:11-19)    */
    AMRVW_SCHEDULE_PIPELINED = 2,  /* several bulges in flight per polynomial, three indices apart */
    AMRVW_SCHEDULE_STREAMED = 3    /* rds_f64 with 64 or 128 lanes_per_poly (0 = 128): chained warps that take the fat steps of
                                      all polynomials as one stream -- no fill/drain between fat steps, shifts computed by
                                      service warps meanwhile; AUTO picks it for batches of a few hundred large polynomials */
};

/* Where the 2L/reuse shifts of a pipelined fat step come from (no reference counterpart: the
 * reference takes two shifts from the trailing 2x2 block, create-bulge.jl:40-52). */
enum amrvw_shift_solver {
    AMRVW_SHIFTS_AUTO = 0,
    AMRVW_SHIFTS_CHASE = 1,   /* core-chasing iteration on a copy of the trailing block              */
    AMRVW_SHIFTS_DENSE = 2    /* trailing block written out densely + double-shift Hessenberg QR     */
};

/* The schedule knobs of this library plus the reference's compile-time constants as run-time options:
 * IT_COUNT = 15 (create-bulge.jl:9: bulge steps on one window before an exceptional shift), it_max = 20 n
 * (AMRVW_algorithm.jl:25: cap on driver trips) and the deflation tolerance eps (AMRVW_algorithm.jl:184).
 * amrvw_default_options fills in the reference's values; 0 in any of the three also means "reference value".
 * The three constants live in device constant memory: contexts that share a device must agree on them. */
typedef struct amrvw_options {
    int32_t schedule;        /* amrvw_schedule */
    int32_t lanes_per_poly;  /* pipelined: bulges in flight per polynomial (4, 8, 16, 32; rds_f64 also 64, 128, 256 = 2, 4 or 8 chained warps); 0 = auto */
    int32_t shift_reuse;     /* pipelined: bulges driven by each shift pair (1, 2, 4 or 8); 0 = auto  */
    int32_t small_window;    /* pipelined: windows up to this size use the one-bulge step; 0 = auto   */
    int32_t shift_solver;    /* pipelined: amrvw_shift_solver; 0 = auto (dense)                        */
    int32_t it_count;        /* IT_COUNT, create-bulge.jl:9; default 15                                */
    uint64_t seed;           /* exceptional (random) shifts: counter-based, reproducible              */
    int32_t it_max_factor;   /* it_max = it_max_factor * n, AMRVW_algorithm.jl:25; default 20          */
    int32_t stream_ctas_per_sm; /* streamed schedule: resident CTAs per SM (1..3); 0 = auto                    */
    double tol;              /* |s| <= tol deflates, AMRVW_algorithm.jl:184; default eps(Float64)      */
} amrvw_options;

typedef struct amrvw_stats {
    int64_t turnovers;     /* executed calls of the turnover primitive (transformations.jl:208) */
    int64_t sweeps;        /* bulges chased */
    int64_t exceptional;   /* random shifts taken (create-bulge.jl:25-37, 85-93) */
    int64_t unconverged;   /* root slots left at zero */
    int64_t fat_steps;     /* pipelined multi-bulge steps */
    int64_t launches;      /* kernels launched by this call */
    double device_ms;      /* CUDA-event time of the kernels of this call */
    int64_t lanes;         /* bulges in flight per polynomial chosen for this call (0: one-bulge schedule) */
    int64_t schedule_used; /* amrvw_schedule that ran: 1 one-bulge, 2 pipelined (unit queue or chained warps), 3 streamed */
} amrvw_stats;

/* device < 0: current device. */
int amrvw_create(amrvw_ctx** ctx, int device, uint64_t seed);
/* One context over several GPUs of the box: the HOST-buffer entry points below split
 * the batch contiguously by polynomial, balanced by the sum of squared lengths (the work per polynomial
 * grows like n^2), run every slice on its own device and stream side by side, and copy each slice's
 * results straight into the caller's buffers -- one call roots the whole batch, as one
 * `roots.(pss)` comprehension does in the reference (README.md:121, companion.jl:252-272).  No
 * collective: polynomials never interact.  ndev = 0: every visible device.  Results are identical to
 * the single-device call (same kernels) except where an exceptional shift is drawn: its random stream is keyed by the
 * polynomial's position in the device's slice (create-bulge.jl:28,88 draw from the global RNG).  The *_dev entry points and
 * amrvw_set_stream need a single-device context. */
int amrvw_create_multi(amrvw_ctx** ctx, const int* devices, int ndev, uint64_t seed);
int amrvw_device_count(const amrvw_ctx* ctx);
/* The split a multi-device context applies (pure host code, needs no device): device d roots polynomials
 * cuts[d] .. cuts[d+1]-1; cuts has ndev+1 entries, cuts[0] = 0, cuts[ndev] = nbatch. */
int amrvw_partition_batch(const int64_t* offsets, int64_t nbatch, int ndev, int64_t* cuts);
int amrvw_destroy(amrvw_ctx* ctx);
const char* amrvw_last_error(const amrvw_ctx* ctx);
void amrvw_default_options(amrvw_options* opt);
int amrvw_set_options(amrvw_ctx* ctx, const amrvw_options* opt);
/* Run on the caller's CUDA stream (cudaStream_t passed as void*).  NULL is a valid handle: the
 * legacy default stream (what torch.cuda.current_stream().cuda_stream is by default).
 * amrvw_use_own_stream goes back to the context's private non-blocking stream (the default). */
int amrvw_set_stream(amrvw_ctx* ctx, void* cuda_stream);
int amrvw_use_own_stream(amrvw_ctx* ctx);

/* ---- host buffers in, host buffers out (H2D + kernels + D2H inside the call) ------------------ */

/* roots(ps) for Vector{Float64}: real double shift (src/RDS.jl).  companion.jl:252. */
int amrvw_roots_rds_f64(amrvw_ctx* ctx, const double* coeffs, const int64_t* offsets, int64_t nbatch,
                        double* roots_reim, int32_t* nroots, int32_t* n_unconverged, amrvw_stats* stats);
/* roots(ps) for Vector{ComplexF64}: complex single shift (src/CSS.jl).  companion.jl:252. */
int amrvw_roots_css_c64(amrvw_ctx* ctx, const double* coeffs_reim, const int64_t* offsets, int64_t nbatch,
                        double* roots_reim, int32_t* nroots, int32_t* n_unconverged, amrvw_stats* stats);
/* roots(vs, ws): companion pencil (src/pencil.jl), Float64 / ComplexF64.  companion.jl:282.
 * Inputs are const: the reference's in-place edits of vs/ws (utils.jl:38,50, companion.jl:168-170)
 * are not reproduced. */
int amrvw_roots_pencil_f64(amrvw_ctx* ctx, const double* vs, const double* ws, const int64_t* offsets, int64_t nbatch,
                           double* roots_reim, int32_t* nroots, int32_t* n_unconverged, amrvw_stats* stats);
int amrvw_roots_pencil_c64(amrvw_ctx* ctx, const double* vs_reim, const double* ws_reim, const int64_t* offsets, int64_t nbatch,
                           double* roots_reim, int32_t* nroots, int32_t* n_unconverged, amrvw_stats* stats);

/* ---- device-resident variants: every pointer except stats is a DEVICE pointer on the context's
 * device; total = offsets[nbatch] and max_len = longest polynomial are passed by value so the
 * call needs no device->host read.  Asynchronous on the context's stream unless stats != NULL
 * (stats needs the event time and the device-side counters, so it synchronises). */
int amrvw_roots_rds_f64_dev(amrvw_ctx* ctx, const double* d_coeffs, const int64_t* d_offsets, int64_t nbatch, int64_t total, int64_t max_len,
                            double* d_roots_reim, int32_t* d_nroots, int32_t* d_n_unconverged, amrvw_stats* stats);
int amrvw_roots_css_c64_dev(amrvw_ctx* ctx, const double* d_coeffs_reim, const int64_t* d_offsets, int64_t nbatch, int64_t total, int64_t max_len,
                            double* d_roots_reim, int32_t* d_nroots, int32_t* d_n_unconverged, amrvw_stats* stats);
int amrvw_roots_pencil_f64_dev(amrvw_ctx* ctx, const double* d_vs, const double* d_ws, const int64_t* d_offsets, int64_t nbatch, int64_t total, int64_t max_len,
                               double* d_roots_reim, int32_t* d_nroots, int32_t* d_n_unconverged, amrvw_stats* stats);
int amrvw_roots_pencil_c64_dev(amrvw_ctx* ctx, const double* d_vs_reim, const double* d_ws_reim, const int64_t* d_offsets, int64_t nbatch, int64_t total, int64_t max_len,
                               double* d_roots_reim, int32_t* d_nroots, int32_t* d_n_unconverged, amrvw_stats* stats);

/* eigvals(qr_factorization(H; unitary)) for a batch of upper Hessenberg matrices, HOST buffers:
 * replaces `eigvals(AMRVW.qr_factorization(M, unitary=...))` per matrix (src/qrfactorization.jl:63-103 with
 * src/AMRVW_algorithm.jl:207-216; R as RFactorizationUpperTriangular, src/rfactorization.jl:259-357, or -- unitary != 0, for a
 * unitary H -- as RFactorizationUnitaryDiagonal, :230-248).  H holds the matrices back to back, each dims[p] x dims[p] in
 * column major order (Julia's layout; _c64: interleaved re, im); entries below the first subdiagonal are not read.  The
 * eigenvalues of matrix p are written, sorted like LinearAlgebra.sorteig!, to complex slots sum(dims[0..p-1]) onwards of
 * eig_reim; n_unconverged[p] counts the slots the iteration cap left at zero.  Real matrices run the real double shift
 * chase, complex ones the single shift chase, one bulge at a time as the reference does.  A multi-device context splits
 * the batch contiguously by matrix, balanced by the sum of n^3, and gathers into the caller's buffers like the roots calls.
 * Limits: the rotators of a matrix live in shared memory, so n <= 7030 (Float64) / 3510 (ComplexF64) -- larger matrices
 * return AMRVW_ERR_ARGUMENT; amrvw_stats.lanes reports the kernel that ran (2: one thread per matrix, 1: one warp per matrix
 * with R in shared memory, 0: one warp per matrix with R in a global workspace); it_count, it_max_factor and tol apply. */
int amrvw_eigvals_hessenberg_f64(amrvw_ctx* ctx, const double* H, const int32_t* dims, int64_t nbatch, int unitary,
                                 double* eig_reim, int32_t* n_unconverged, amrvw_stats* stats);
int amrvw_eigvals_hessenberg_c64(amrvw_ctx* ctx, const double* H_reim, const int32_t* dims, int64_t nbatch, int unitary,
                                 double* eig_reim, int32_t* n_unconverged, amrvw_stats* stats);
/* The split a multi-device context applies to a batch of matrices (pure host code, needs no device): device d takes matrices
 * cuts[d] .. cuts[d+1]-1; cuts has ndev+1 entries. */
int amrvw_partition_hessenberg(const int32_t* dims, int64_t nbatch, int ndev, int64_t* cuts);
/* The same with DEVICE pointers, asynchronous on the context's stream (stats != NULL synchronizes); single-device contexts.
 * d_moff[p] / d_eoff[p]: element offset of matrix p in d_H and complex slot offset of its eigenvalues (nbatch + 1 entries
 * each, the running sums of dims^2 and dims); max_dim: the largest dims[p].  A slot of an eigenvalue the iteration cap
 * left unconverged holds zero. */
int amrvw_eigvals_hessenberg_f64_dev(amrvw_ctx* ctx, const double* d_H, const int32_t* d_dims, const int64_t* d_moff, const int64_t* d_eoff,
                                     int64_t nbatch, int32_t max_dim, int unitary, double* d_eig_reim, int32_t* d_n_unconverged, amrvw_stats* stats);
int amrvw_eigvals_hessenberg_c64_dev(amrvw_ctx* ctx, const double* d_H_reim, const int32_t* d_dims, const int64_t* d_moff, const int64_t* d_eoff,
                                     int64_t nbatch, int32_t max_dim, int unitary, double* d_eig_reim, int32_t* d_n_unconverged, amrvw_stats* stats);

/* real_polynomial_roots (src/complex_conjugate_root_theorem.jl:32-54) for a batch, HOST buffers: roots the
 * polynomials like amrvw_roots_rds_f64 and returns only, per polynomial, the number of roots with imag == 0
 * exactly, imag > 0 and imag < 0 (counts3[3p..3p+2]) -- 12 bytes per polynomial come back instead of 16 n
 * (README.md:121-127 needs exactly count(isreal)).  Works on single- and multi-device contexts. */
int amrvw_count_real_roots_rds_f64(amrvw_ctx* ctx, const double* coeffs, const int64_t* offsets, int64_t nbatch,
                                   int32_t* counts3, int32_t* n_unconverged, amrvw_stats* stats);

/* The same reduction on roots already on the device (src/complex_conjugate_root_theorem.jl:32-54): for each
 * polynomial the number of roots with imag == 0 exactly, imag > 0 and imag < 0 (counts[3*p..]).
 * Device pointers; the README statistics workload (README.md:121) needs only this. */
int amrvw_count_real_roots_dev(amrvw_ctx* ctx, const double* d_roots_reim, const int64_t* d_offsets, int64_t nbatch,
                               int32_t slot_shift_per_poly /* 0: roots at slot offsets[p] */, const int32_t* d_nroots, int32_t* d_counts3);

/* FP64 DFMA micro-benchmark: dependent-chain-free FMA throughput of the device, in TFLOP/s
 * (the FP64 roofline denominator; MEASURED_PEAKS.json only has HBM and bf16). */
int amrvw_measure_fp64_peak(amrvw_ctx* ctx, double* tflops);

/* Device self-test of the division-free primitives (mirrors test/test-amrvw-basics.jl:35-42, the
 * turnover must preserve the product): out4 = {max rel. error of the Newton rsqrt, max |Q1Q2Q3 -
 * R1R2R3| over n random turnovers, max deviation from unit norm, turnovers sent to the IEEE path}. */
int amrvw_selftest(amrvw_ctx* ctx, int64_t n, double* out4);
/* The complex twin: out4 = {max |fast - IEEE| over the outputs of n random complex turnovers, max
 * |Q1Q2Q3 - R1R2R3| of the fast path, max deviation from unit norm, turnovers sent to the IEEE path}. */
int amrvw_selftest_complex(amrvw_ctx* ctx, int64_t n, double* out4);
/* Device self-test of the fall-back turnover of the multishift schedules (no reference counterpart: the
 * reference deflates before a bulge can meet a vanished sine) on n random real triples whose sines span
 * 1e-320 .. 1: out4 = {max |Q1 Q2 Q3 - R1 R2 R3| of the division-free fall-back, max deviation of its
 * rotators from unit norm, the same product error for the reference's formulas (transformations.jl:158-202)
 * on the same inputs, number of triples}. */
int amrvw_selftest_robust(amrvw_ctx* ctx, int64_t n, double* out4);

/* Accounting of the last pipelined RDS call that asked for stats (development aid, no reference
 * counterpart).  out16 = {[0] sum over phase-B warps of their cycles, [2] of which in the lean
 * systolic loop, [3] in the fill/drain loop; [4] lean steps, [5] fill/drain steps, [6] lean calls,
 * [7] phase-B warp launches; kernel times in microseconds: [8] set-up, [9] phase A (all rounds),
 * [10] phase B (all rounds), [11] small windows, [12] finish; [13] rounds, [14] small windows}. */
int amrvw_last_profile(amrvw_ctx* ctx, int64_t* out16);

/* Development counters of the last call made with stats (48 slots whose meaning belongs to the kernel under study;
 * see the kernel source; [47] = polynomials that moved to another unit in the last unit-queue call: the compaction of
 * half-empty units, on when units outnumber the resident warps).  No reference counterpart. */
int amrvw_debug_counters(amrvw_ctx* ctx, int64_t* out48);

int amrvw_version(void);

#ifdef __cplusplus
}
#endif
#endif /* AMRVW_B200_H */
