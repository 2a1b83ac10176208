// capi.cu -- host side of the C ABI declared in include/amrvw_b200.h: context, workspace,
// launches, H2D/D2H staging.  No CPU compute path exists here: every entry point launches the
// sm_100a kernels of kernels.cuh / pipeline.cuh or fails.
#include "../../include/amrvw_b200.h"
#include "kernels.cuh"
#include "pipeline.cuh"
#include "rounds.cuh"
#include "chain.cuh"
#include "stream.cuh"
#include "css_units.cuh"
#include "pencil_units.cuh"
#include "hessenberg.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

using namespace amrvw;

struct DevBuf {
    void* ptr = nullptr;
    size_t cap = 0;
};

struct amrvw_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    amrvw_options opt;
    std::string err;
    // grow-only device buffers
    DevBuf chains[5], diags[3], coeffs, coeffs2, offsets, roots, nroots, nunconv, stats, shifts, state, prof, rec, fatshifts, swlist, counters, queue, qslots, park, parked, order, orderkeys, counts, squeues, sslots;
    bool queue_used = false;
    cudaEvent_t evpool[32] = {};
    RoundsTimes times = {};
    bool prof_valid = false;
    int sm_count = 148;
    // multi-device context (amrvw_create_multi): one single-device context per GPU; this object then owns no
    // device memory of its own and only partitions host batches over `subs`
    std::vector<amrvw_ctx*> subs;
    // degree classes of a ragged host batch: one sub-context per class on THIS device, created on first use,
    // each with its own stream and workspace so that the classes' kernels run side by side with their own lane counts
    std::vector<amrvw_ctx*> klass;
    DevBuf hwork, hmeta;            // Hessenberg eigenproblems: global R workspace, offsets
    DevBuf members;                 // class sub-context: batch indices of its polynomials
    cudaEvent_t ev_in = nullptr, ev_done = nullptr;
};

static int fail(amrvw_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    if (code == AMRVW_ERR_CUDA) (void)cudaGetLastError();   // reported: do not leave it behind for the next call's launch check
    return code;
}
#define CUDA_TRY(ctx, call)                                                                     \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return fail(ctx, AMRVW_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

static int reserve(amrvw_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap && b.ptr) return 0;
    if (b.ptr) CUDA_TRY(ctx, cudaFree(b.ptr));
    b.ptr = nullptr; b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    CUDA_TRY(ctx, cudaMalloc(&b.ptr, want));
    b.cap = want;
    return 0;
}

extern "C" {

int amrvw_version(void) { return 100; }

void amrvw_default_options(amrvw_options* opt) {
    if (!opt) return;
    opt->schedule = AMRVW_SCHEDULE_AUTO;
    opt->lanes_per_poly = 0;
    opt->shift_reuse = 0;
    opt->small_window = 0;
    opt->shift_solver = AMRVW_SHIFTS_AUTO;
    opt->it_count = 15;                          // create-bulge.jl:9
    opt->seed = 0;
    opt->it_max_factor = 20;                     // AMRVW_algorithm.jl:25
    opt->stream_ctas_per_sm = 0;
    opt->tol = 2.220446049250313e-16;            // AMRVW_algorithm.jl:184: eps(Float64)
}

int amrvw_create(amrvw_ctx** out, int device, uint64_t seed) {
    if (!out) return AMRVW_ERR_ARGUMENT;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return AMRVW_ERR_NO_DEVICE;  // no CPU fallback
    amrvw_ctx* ctx = new amrvw_ctx();
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) { delete ctx; return AMRVW_ERR_CUDA; } }
    if (device >= ndev) { delete ctx; return AMRVW_ERR_ARGUMENT; }
    ctx->device = device;
    amrvw_default_options(&ctx->opt);
    ctx->opt.seed = seed;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
        delete ctx;
        return AMRVW_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = ctx;
    return AMRVW_OK;
}

int amrvw_create_multi(amrvw_ctx** out, const int* devices, int ndev, uint64_t seed) {
    if (!out) return AMRVW_ERR_ARGUMENT;
    *out = nullptr;
    if (ndev < 0 || (ndev > 0 && !devices)) return AMRVW_ERR_ARGUMENT;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) return AMRVW_ERR_NO_DEVICE;  // no CPU fallback
    std::vector<int> devs;
    if (ndev == 0) for (int d = 0; d < have; ++d) devs.push_back(d);          // every visible device
    else devs.assign(devices, devices + ndev);
    for (size_t i = 0; i < devs.size(); ++i) {
        if (devs[i] < 0 || devs[i] >= have) return AMRVW_ERR_ARGUMENT;
        for (size_t j = 0; j < i; ++j) if (devs[j] == devs[i]) return AMRVW_ERR_ARGUMENT;   // each device once
    }
    if (devs.size() == 1) return amrvw_create(out, devs[0], seed);
    amrvw_ctx* ctx = new amrvw_ctx();
    amrvw_default_options(&ctx->opt);
    ctx->opt.seed = seed;
    ctx->device = devs[0];
    for (int d : devs) {
        amrvw_ctx* sub = nullptr;
        const int rc = amrvw_create(&sub, d, seed);
        if (rc != AMRVW_OK) {
            for (amrvw_ctx* s2 : ctx->subs) amrvw_destroy(s2);
            delete ctx;
            return rc;
        }
        ctx->subs.push_back(sub);
    }
    ctx->sm_count = ctx->subs[0]->sm_count;
    *out = ctx;
    return AMRVW_OK;
}

int amrvw_device_count(const amrvw_ctx* ctx) { return ctx ? (ctx->subs.empty() ? 1 : (int)ctx->subs.size()) : 0; }

int amrvw_destroy(amrvw_ctx* ctx) {
    if (!ctx) return AMRVW_OK;
    if (!ctx->subs.empty()) {
        for (amrvw_ctx* s2 : ctx->subs) amrvw_destroy(s2);
        delete ctx;
        return AMRVW_OK;
    }
    for (amrvw_ctx* k : ctx->klass) amrvw_destroy(k);
    ctx->klass.clear();
    cudaSetDevice(ctx->device);
    if (ctx->members.ptr) cudaFree(ctx->members.ptr);
    if (ctx->ev_in) cudaEventDestroy(ctx->ev_in);
    if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
    DevBuf* all[] = {&ctx->chains[0], &ctx->chains[1], &ctx->chains[2], &ctx->chains[3], &ctx->chains[4], &ctx->diags[0], &ctx->diags[1], &ctx->diags[2],
                     &ctx->coeffs, &ctx->coeffs2, &ctx->offsets, &ctx->roots, &ctx->nroots, &ctx->nunconv, &ctx->stats, &ctx->shifts, &ctx->state, &ctx->prof,
                     &ctx->rec, &ctx->fatshifts, &ctx->swlist, &ctx->counters, &ctx->queue, &ctx->qslots, &ctx->park, &ctx->parked, &ctx->order, &ctx->orderkeys, &ctx->counts, &ctx->squeues, &ctx->sslots, &ctx->hwork, &ctx->hmeta};
    for (DevBuf* b : all) if (b->ptr) cudaFree(b->ptr);
    for (cudaEvent_t e : ctx->evpool) if (e) cudaEventDestroy(e);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return AMRVW_OK;
}

const char* amrvw_last_error(const amrvw_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

}  // extern "C"
// the single-device context behind the calls that need device buffers of their own (self-tests, probes)
static amrvw_ctx* primary(amrvw_ctx* ctx) { return (ctx && !ctx->subs.empty()) ? ctx->subs[0] : ctx; }
extern "C" {


int amrvw_set_options(amrvw_ctx* ctx, const amrvw_options* opt) {
    if (!ctx || !opt) return AMRVW_ERR_ARGUMENT;
    if (opt->schedule < 0 || opt->schedule > 3) return fail(ctx, AMRVW_ERR_ARGUMENT, "schedule must be 0, 1, 2 or 3");
    const int L = opt->lanes_per_poly;
    if (!(L == 0 || L == 4 || L == 8 || L == 16 || L == 32 || L == 64 || L == 128 || L == 256)) return fail(ctx, AMRVW_ERR_ARGUMENT, "lanes_per_poly must be 0, 4, 8, 16, 32, 64, 128 or 256");
    if (opt->shift_reuse < 0 || opt->shift_reuse > 16) return fail(ctx, AMRVW_ERR_ARGUMENT, "shift_reuse must be 0 (auto) or 1..16");
    if (opt->small_window < 0) return fail(ctx, AMRVW_ERR_ARGUMENT, "small_window must be >= 0");
    if (opt->shift_solver < 0 || opt->shift_solver > 2) return fail(ctx, AMRVW_ERR_ARGUMENT, "shift_solver must be 0, 1 or 2");
    if (opt->it_count < 0 || opt->it_max_factor < 0) return fail(ctx, AMRVW_ERR_ARGUMENT, "it_count and it_max_factor must be >= 0 (0 = the reference's 15 and 20)");
    if (!(opt->tol >= 0.0) || opt->tol > 1e-4) return fail(ctx, AMRVW_ERR_ARGUMENT, "tol must be in [0, 1e-4] (0 = the reference's eps)");
    if (opt->stream_ctas_per_sm < 0 || opt->stream_ctas_per_sm > 3) return fail(ctx, AMRVW_ERR_ARGUMENT, "stream_ctas_per_sm must be 0..3");
    ctx->opt = *opt;
    if (ctx->opt.it_count == 0) ctx->opt.it_count = 15;
    if (ctx->opt.it_max_factor == 0) ctx->opt.it_max_factor = 20;
    if (ctx->opt.tol == 0.0) ctx->opt.tol = 2.220446049250313e-16;
    for (amrvw_ctx* sub : ctx->subs) {
        amrvw_options o = ctx->opt;
        const int rc = amrvw_set_options(sub, &o);
        if (rc != AMRVW_OK) return fail(ctx, rc, sub->err);
    }
    return AMRVW_OK;
}

int amrvw_set_stream(amrvw_ctx* ctx, void* s) {
    if (!ctx) return AMRVW_ERR_ARGUMENT;
    if (!ctx->subs.empty()) return fail(ctx, AMRVW_ERR_ARGUMENT, "a multi-device context runs on its own per-device streams");
    ctx->stream = (cudaStream_t)s;   // NULL = legacy default stream
    return AMRVW_OK;
}
int amrvw_use_own_stream(amrvw_ctx* ctx) {
    if (!ctx) return AMRVW_ERR_ARGUMENT;
    ctx->stream = ctx->own_stream;
    return AMRVW_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
enum Variant { V_RDS = 0, V_CSS = 1, V_PENCIL_R = 2, V_PENCIL_C = 3 };

struct StatsAccum { long long turnovers, sweeps, exceptional, unconverged, fat_steps; };

__global__ void reduce_stats_kernel(const PolyStats* st, long long n, StatsAccum* out) {
    long long t = 0, s = 0, e = 0, u = 0, f = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        t += st[i].turnovers; s += st[i].sweeps; e += st[i].exceptional; u += st[i].unconverged; f += st[i].fat_steps;
    }
    atomicAdd((unsigned long long*)&out->turnovers, (unsigned long long)t);
    atomicAdd((unsigned long long*)&out->sweeps, (unsigned long long)s);
    atomicAdd((unsigned long long*)&out->exceptional, (unsigned long long)e);
    atomicAdd((unsigned long long*)&out->unconverged, (unsigned long long)u);
    atomicAdd((unsigned long long*)&out->fat_steps, (unsigned long long)f);
}

// the unit queue's sticky abort flag (rounds.cuh: queue_pop); call with the stream idle
static int check_queue_abort(amrvw_ctx* ctx) {
    if (!ctx->queue_used || !ctx->counters.ptr) return AMRVW_OK;
    unsigned flag = 0;
    CUDA_TRY(ctx, cudaMemcpy(&flag, (const unsigned*)ctx->counters.ptr + CNT_ABORT, sizeof(flag), cudaMemcpyDeviceToHost));
    if (flag) return fail(ctx, AMRVW_ERR_QUEUE_TIMEOUT, "unit queue timed out: a warp waited ~10 s for work that never came; results are incomplete");
    return AMRVW_OK;
}

// workspace of the unit schedules: records, shifts, small-window list, counters; queue and parking sized by the
// number of units the chosen lane count really gives (L lanes per polynomial -> 32/L polynomials per unit)
static int reserve_unit_workspace(amrvw_ctx* ctx, int64_t nbatch, int64_t total, int L, size_t lane_bytes, bool streamed, size_t* qcap_out) {
    int rc;
    if ((rc = reserve(ctx, ctx->rec, (size_t)nbatch * sizeof(PolyRec)))) return rc;
    if ((rc = reserve(ctx, ctx->fatshifts, (size_t)nbatch * (size_t)(L < 4 ? 4 : L) * sizeof(Shifts)))) return rc;
    if ((rc = reserve(ctx, ctx->swlist, ((size_t)total / 3 + (size_t)nbatch + 16) * sizeof(SmallWin)))) return rc;
    if ((rc = reserve(ctx, ctx->counters, CNT_WORDS * sizeof(unsigned)))) return rc;
    *qcap_out = 0;
    if (L > 32 || streamed) return 0;     // chained warps and streams: no unit queue, nothing is parked
    const size_t G = (size_t)(32 / L);
    const size_t nunits = ((size_t)nbatch + G - 1) / G;
    size_t qcap = 1024;
    while (qcap < 2 * nunits + 4096) qcap <<= 1;
    if ((rc = reserve(ctx, ctx->queue, sizeof(RoundQueue)))) return rc;
    if ((rc = reserve(ctx, ctx->qslots, qcap * sizeof(int)))) return rc;
    if ((rc = reserve(ctx, ctx->park, nunits * sizeof(UnitPark)))) return rc;
    if ((rc = reserve(ctx, ctx->parked, nunits * 32 * lane_bytes))) return rc;
    *qcap_out = qcap;
    return 0;
}

template <class S>
static int run_dev(amrvw_ctx* ctx, Variant var, const S* d_a, const S* d_w, const int64_t* d_offsets, int64_t nbatch, int64_t total, int64_t max_len,
                   double* d_roots, int32_t* d_nroots, int32_t* d_nunconv, amrvw_stats* stats, int64_t load_hint = 0) {
    if (!ctx) return AMRVW_ERR_ARGUMENT;
    if (!ctx->subs.empty()) return fail(ctx, AMRVW_ERR_ARGUMENT, "device-pointer entry points need a single-device context (amrvw_create)");
    if (nbatch < 0 || total < 0) return fail(ctx, AMRVW_ERR_ARGUMENT, "negative batch or total");
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (nbatch == 0) return AMRVW_OK;
    if (!d_a || !d_offsets || !d_roots || !d_nroots || !d_nunconv) return fail(ctx, AMRVW_ERR_ARGUMENT, "null buffer");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    (void)cudaGetLastError();      // a stale error of someone else's earlier call must not fail this call's launch checks
    const bool pencil = (var == V_PENCIL_R || var == V_PENCIL_C);
    if (pencil && !d_w) return fail(ctx, AMRVW_ERR_ARGUMENT, "null ws buffer");
    const int extra = pencil ? 2 : 1;
    const size_t slots = (size_t)total + (size_t)extra * (size_t)nbatch + 4;
    const int nchains = pencil ? 5 : 3;
    for (int c = 0; c < nchains; ++c) { int rc = reserve(ctx, ctx->chains[c], slots * sizeof(Rot<S>)); if (rc) return rc; }
    if (!is_real<S>::value) {
        const int nd = pencil ? 3 : 2;
        for (int c = 0; c < nd; ++c) { int rc = reserve(ctx, ctx->diags[c], slots * sizeof(S)); if (rc) return rc; }
    }
    { int rc = reserve(ctx, ctx->stats, (size_t)nbatch * sizeof(PolyStats) + sizeof(StatsAccum)); if (rc) return rc; }
    if (var == V_RDS || var == V_CSS || var == V_PENCIL_R) { int rc = reserve(ctx, ctx->shifts, slots * sizeof(int)); if (rc) return rc; }

    Problem<S> pr;
    pr.coeffs = d_a; pr.ws = d_w; pr.offsets = (const long long*)d_offsets; pr.nbatch = nbatch;
    pr.roots = (cplx*)d_roots; pr.nroots = d_nroots; pr.nunconv = d_nunconv;
    pr.stats = (PolyStats*)ctx->stats.ptr;
    pr.Q = (Rot<S>*)ctx->chains[0].ptr; pr.B = (Rot<S>*)ctx->chains[1].ptr; pr.Ct = (Rot<S>*)ctx->chains[2].ptr;
    pr.WB = (Rot<S>*)ctx->chains[3].ptr; pr.WCt = (Rot<S>*)ctx->chains[4].ptr;
    pr.DQ = (S*)ctx->diags[0].ptr; pr.DR = (S*)ctx->diags[1].ptr; pr.WD = (S*)ctx->diags[2].ptr;
    pr.dstack = (int*)ctx->shifts.ptr;
    pr.prof = nullptr; pr.rec = nullptr; pr.shifts = nullptr; pr.swlist = nullptr; pr.counters = nullptr; pr.order = nullptr;
    cudaStream_t st = ctx->stream;
    if (stats && var != V_PENCIL_C) {   // warp-cycle counters ride along with the statistics
        int rc = reserve(ctx, ctx->prof, 64 * sizeof(unsigned long long)); if (rc) return rc;      // 16 documented counters + 48 development slots
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->prof.ptr, 0, 64 * sizeof(unsigned long long), st));
        pr.prof = (unsigned long long*)ctx->prof.ptr;
    }
    pr.extra = extra;
    pr.seed = ctx->opt.seed;
    pr.load_hint = load_hint;
    const int64_t nload = load_hint > nbatch ? load_hint : nbatch;
    {   // the reference's constants as options (solver.cuh: g_algo), written on the call's stream
        const AlgoConst ac = {ctx->opt.tol, ctx->opt.it_count, ctx->opt.it_max_factor};
        CUDA_TRY(ctx, cudaMemcpyToSymbolAsync(g_algo, &ac, sizeof(ac), 0, cudaMemcpyHostToDevice, st));
    }

    // unit schedules form their units from polynomials in order of decreasing degree
    const bool unit_sched = ctx->opt.schedule != AMRVW_SCHEDULE_REFERENCE && var != V_PENCIL_C;
    if (unit_sched && nbatch > 1) {
        int rc;
        if ((rc = reserve(ctx, ctx->order, (size_t)nbatch * sizeof(int)))) return rc;
        if ((rc = reserve(ctx, ctx->orderkeys, (size_t)nbatch * sizeof(unsigned long long)))) return rc;
        pr.order = (const int*)ctx->order.ptr;
    }
    int launches = 0, lanes = 0, sched_used = AMRVW_SCHEDULE_REFERENCE;
    ctx->times = RoundsTimes{0.f, 0.f, 0.f, 0.f};
    ctx->prof_valid = (var != V_PENCIL_C && stats != nullptr);
    ctx->queue_used = false;
    if (stats && !ctx->evpool[0]) for (cudaEvent_t& e : ctx->evpool) CUDA_TRY(ctx, cudaEventCreate(&e));
    if (stats) CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, st));

    bool done = false;
    if (unit_sched) {
        int L; PipeParams pp; size_t qcap = 0; int rc;
        size_t lane_bytes;
        if constexpr (is_real<S>::value) {
            if (var == V_RDS) { choose_pipe_params(ctx->opt, nload, (int)max_len, ctx->sm_count, L, pp, nbatch); lane_bytes = sizeof(StepState); }
            else { choose_pencil_params(ctx->opt, nload, (int)max_len, ctx->sm_count, L, pp); lane_bytes = sizeof(PenLane); }
        } else {
            choose_css_params(ctx->opt, nload, (int)max_len, ctx->sm_count, L, pp); lane_bytes = sizeof(CssLane);
        }
        if ((rc = reserve_unit_workspace(ctx, nbatch, total, L, lane_bytes, var == V_RDS && pp.stream, &qcap))) return rc;
        StreamBufs sbuf = {nullptr, nullptr, nullptr, nullptr, 0u, nullptr};
        if (var == V_RDS && pp.stream) {     // two device-wide queues of polynomial ids + the count of iterating polynomials
            size_t cap = 1024;
            while (cap < (size_t)nbatch + 64) cap <<= 1;
            if ((rc = reserve(ctx, ctx->squeues, 2 * sizeof(MpmcQueue) + 64))) return rc;     // + remaining, progress, t0
            if ((rc = reserve(ctx, ctx->sslots, 2 * cap * sizeof(MpmcSlot)))) return rc;
            sbuf.ready = (MpmcQueue*)ctx->squeues.ptr; sbuf.work = sbuf.ready + 1; sbuf.remaining = (unsigned*)(sbuf.work + 1);
            sbuf.ready_slots = (MpmcSlot*)ctx->sslots.ptr; sbuf.work_slots = sbuf.ready_slots + cap; sbuf.mask = (unsigned)(cap - 1);
        }
        QueueBufs qb;
        qb.q = (RoundQueue*)ctx->queue.ptr; qb.slots = (int*)ctx->qslots.ptr; qb.mask = qcap ? (unsigned)(qcap - 1) : 0u;
        qb.park = (UnitPark*)ctx->park.ptr; qb.parked = nullptr;
        pr.rec = (PolyRec*)ctx->rec.ptr; pr.shifts = (Shifts*)ctx->fatshifts.ptr; pr.swlist = ctx->swlist.ptr; pr.counters = (unsigned*)ctx->counters.ptr;
        ctx->queue_used = qcap != 0 || sbuf.ready != nullptr;
        sched_used = sbuf.ready ? AMRVW_SCHEDULE_STREAMED : AMRVW_SCHEDULE_PIPELINED;
        cudaError_t e = cudaSuccess;
        const char* what = "";
        if constexpr (is_real<S>::value) {
            if (var == V_RDS) {
                what = "unit schedule: ";
                qb.parked = (StepState*)ctx->parked.ptr;
                e = launch_rounds_rds(pr, ctx->opt, (int)max_len, (long long)total, ctx->sm_count, st, &launches, ctx->evpool, stats ? &ctx->times : nullptr, qb, &lanes, (unsigned long long*)ctx->orderkeys.ptr, sbuf.ready ? &sbuf : nullptr);
            } else {
                what = "pencil unit schedule: ";
                e = launch_pencil_pipelined(pr, ctx->opt, (int)max_len, (long long)total, ctx->sm_count, st, &launches, qb, (PenLane*)ctx->parked.ptr, ctx->evpool, stats ? &ctx->times : nullptr, &lanes, (unsigned long long*)ctx->orderkeys.ptr);
            }
        } else {
            what = "complex unit schedule: ";
            e = launch_css_pipelined(pr, ctx->opt, (int)max_len, (long long)total, ctx->sm_count, st, &launches, qb.q, qb.slots, qb.mask, qb.park, (CssLane*)ctx->parked.ptr,
                                     ctx->evpool, stats ? &ctx->times : nullptr, &lanes, (unsigned long long*)ctx->orderkeys.ptr);
        }
        if (e != cudaSuccess) return fail(ctx, AMRVW_ERR_CUDA, std::string(what) + cudaGetErrorString(e));
        done = true;
    }
    if (!done) {
        const int threads = 64;   // one polynomial per thread; small blocks spread the batch over all SMs
        const int blocks = (int)((nbatch + threads - 1) / threads);
        if (pencil) roots_pencil_reference_kernel<S><<<blocks, threads, 0, st>>>(pr);
        else roots_reference_kernel<S><<<blocks, threads, 0, st>>>(pr);
        launches += 1;
        CUDA_TRY(ctx, cudaGetLastError());
    }
    if (stats) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, st));
        StatsAccum* acc = (StatsAccum*)((char*)ctx->stats.ptr + (size_t)nbatch * sizeof(PolyStats));
        CUDA_TRY(ctx, cudaMemsetAsync(acc, 0, sizeof(StatsAccum), st));
        reduce_stats_kernel<<<64, 256, 0, st>>>(pr.stats, nbatch, acc);
        CUDA_TRY(ctx, cudaGetLastError());
        StatsAccum h;
        CUDA_TRY(ctx, cudaMemcpyAsync(&h, acc, sizeof(h), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        float ms = 0.f;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        stats->turnovers = h.turnovers; stats->sweeps = h.sweeps; stats->exceptional = h.exceptional;
        stats->unconverged = h.unconverged; stats->fat_steps = h.fat_steps;
        stats->launches = launches; stats->device_ms = ms; stats->lanes = lanes; stats->schedule_used = sched_used;
        const int rc = check_queue_abort(ctx);     // the stream is idle here: one 4-byte read
        if (rc) return rc;
    }
    return AMRVW_OK;
}

// offsets checks shared by the host entry points; max_len out
static int check_offsets(amrvw_ctx* ctx, const int64_t* offsets, int64_t nbatch, int64_t* max_len) {
    if (offsets[0] != 0) return fail(ctx, AMRVW_ERR_ARGUMENT, "offsets[0] must be 0");
    int64_t m = 0;
    for (int64_t p = 0; p < nbatch; ++p) {
        const int64_t len = offsets[p + 1] - offsets[p];
        if (len < 0) return fail(ctx, AMRVW_ERR_ARGUMENT, "offsets must be non-decreasing");
        if (len > 0x7fffff00) return fail(ctx, AMRVW_ERR_ARGUMENT, "polynomial too long");
        if (len > m) m = len;
    }
    *max_len = m;
    return AMRVW_OK;
}

// ---------------------------------------------------------------- degree classes of a ragged batch
// One lane count serves one degree range: 16 lanes on a degree-60 polynomial are mostly fill and drain (2.1x the
// reference's turnovers at degree 100 against 1.56x with 8), and the few large polynomials of a heavy-tailed batch want
// chained warps while thousands of small ones want the unit queue.  A host batch whose lengths span several classes is
// therefore split: each class is gathered on the device into its own contiguous CSR batch, rooted by its own sub-context
// (own stream, own workspace, lane count chosen from ITS degree and ITS size), and scattered back into the caller's root
// slots; the classes' kernels run side by side.  Per polynomial nothing changes (same kernels, seeds keyed by the
// polynomial's index in ITS class batch only for the exceptional shift).  Not used when the caller pins lanes_per_poly
// or asks for the reference schedule.
static const int64_t CLASS_BOUNDS[] = {64, 256, 1280};        // coefficient counts: class k = number of bounds <= len
enum { NCLASS = 4 };
template <class S>
__global__ void class_gather_kernel(const S* src, const long long* src_off, const int* members, const long long* cls_off, S* dst) {
    const long long i = blockIdx.x;
    const long long from = src_off[members[i]], to = cls_off[i];
    const int len = (int)(cls_off[i + 1] - to);
    for (int k = threadIdx.x; k < len; k += blockDim.x) dst[to + k] = src[from + k];
}
__global__ void class_scatter_kernel(const cplx* cls_roots, const long long* cls_off, const int* cls_nroots, const int* cls_nunconv, const int* members,
                                     const long long* dst_off, cplx* roots, int* nroots, int* nunconv) {
    const long long i = blockIdx.x;
    const int p = members[i];
    const int n = cls_nroots[i];
    const long long from = cls_off[i], to = dst_off[p];
    for (int k = threadIdx.x; k < n; k += blockDim.x) roots[to + k] = cls_roots[from + k];
    if (threadIdx.x == 0) { nroots[p] = n; nunconv[p] = cls_nunconv[i]; }
}

// roots every class of the batch already resident in the parent's coeffs / coeffs2 / offsets buffers; results land in
// the parent's roots / nroots / nunconv buffers.  The parent's stream has the H2D copies queued; on return it waits for
// every class.
template <class S>
static int run_classes(amrvw_ctx* ctx, Variant var, const int64_t* offsets, int64_t nbatch, const std::vector<std::vector<int>>& members, amrvw_stats* stats) {
    const bool pencil = (var == V_PENCIL_R || var == V_PENCIL_C);
    if (!ctx->ev_in) { CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_in, cudaEventDisableTiming)); }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_in, ctx->stream));
    while ((int)ctx->klass.size() < NCLASS) {
        amrvw_ctx* k = nullptr;
        const int rc = amrvw_create(&k, ctx->device, ctx->opt.seed);
        if (rc != AMRVW_OK) return fail(ctx, rc, "cannot create the sub-context of a degree class");
        ctx->klass.push_back(k);
    }
    std::vector<std::vector<int64_t>> cls_off((size_t)NCLASS);
    // the classes share the device: a lower class's lane count is chosen for the whole batch's work expressed in polynomials
    // of ITS size (work grows like n^2), not for its own head count
    double work_all = 0.0;
    for (int64_t p = 0; p < nbatch; ++p) { const double len = (double)(offsets[p + 1] - offsets[p]); work_all += len * len; }
    bool top = true;
    for (int c = NCLASS - 1; c >= 0; --c) {      // the largest polynomials first: theirs is the longest chain of dependent fat steps
        const std::vector<int>& mem = members[(size_t)c];
        if (mem.empty()) continue;
        amrvw_ctx* k = ctx->klass[(size_t)c];
        k->opt = ctx->opt;
        const int64_t nb = (int64_t)mem.size();
        std::vector<int64_t>& off = cls_off[(size_t)c];
        off.assign((size_t)nb + 1, 0);
        int64_t ml = 0;
        for (int64_t i = 0; i < nb; ++i) {
            const int64_t len = offsets[mem[(size_t)i] + 1] - offsets[mem[(size_t)i]];
            off[(size_t)i + 1] = off[(size_t)i] + len;
            ml = std::max(ml, len);
        }
        const int64_t tot = off[(size_t)nb];
        double work_cls = 0.0;
        for (int64_t i = 0; i < nb; ++i) { const double len = (double)(off[(size_t)i + 1] - off[(size_t)i]); work_cls += len * len; }
        // the top class sets the time of the whole call: it gets the lanes its own head count allows; the others fill the
        // machine around it and are sized for the whole batch's work
        const int64_t hint = (top || work_cls <= 0.0) ? nb : (int64_t)(work_all / (work_cls / (double)nb));
        top = false;
        int rc;
        if ((rc = reserve(k, k->coeffs, (size_t)tot * sizeof(S) + 16))) return fail(ctx, rc, k->err);
        if (pencil && (rc = reserve(k, k->coeffs2, (size_t)tot * sizeof(S) + 16))) return fail(ctx, rc, k->err);
        if ((rc = reserve(k, k->offsets, (size_t)(nb + 1) * sizeof(int64_t)))) return fail(ctx, rc, k->err);
        if ((rc = reserve(k, k->roots, ((size_t)tot + 1) * sizeof(cplx)))) return fail(ctx, rc, k->err);
        if ((rc = reserve(k, k->nroots, (size_t)nb * sizeof(int32_t)))) return fail(ctx, rc, k->err);
        if ((rc = reserve(k, k->nunconv, (size_t)nb * sizeof(int32_t)))) return fail(ctx, rc, k->err);
        if ((rc = reserve(k, k->members, (size_t)nb * sizeof(int)))) return fail(ctx, rc, k->err);
        cudaStream_t ks = k->stream;
        CUDA_TRY(ctx, cudaMemcpyAsync(k->members.ptr, mem.data(), (size_t)nb * sizeof(int), cudaMemcpyHostToDevice, ks));
        CUDA_TRY(ctx, cudaMemcpyAsync(k->offsets.ptr, off.data(), (size_t)(nb + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ks));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ks, ctx->ev_in, 0));
        if (tot > 0) {
            class_gather_kernel<S><<<(int)nb, 128, 0, ks>>>((const S*)ctx->coeffs.ptr, (const long long*)ctx->offsets.ptr, (const int*)k->members.ptr,
                                                             (const long long*)k->offsets.ptr, (S*)k->coeffs.ptr);
            if (pencil) class_gather_kernel<S><<<(int)nb, 128, 0, ks>>>((const S*)ctx->coeffs2.ptr, (const long long*)ctx->offsets.ptr, (const int*)k->members.ptr,
                                                                         (const long long*)k->offsets.ptr, (S*)k->coeffs2.ptr);
            CUDA_TRY(ctx, cudaGetLastError());
        }
        amrvw_stats kst;
        rc = run_dev<S>(k, var, (const S*)k->coeffs.ptr, pencil ? (const S*)k->coeffs2.ptr : nullptr, (const int64_t*)k->offsets.ptr, nb, tot, ml,
                        (double*)k->roots.ptr, (int32_t*)k->nroots.ptr, (int32_t*)k->nunconv.ptr, stats ? &kst : nullptr, hint);
        if (rc) return fail(ctx, rc, "degree class " + std::to_string(c) + ": " + k->err);
        class_scatter_kernel<<<(int)nb, 128, 0, ks>>>((const cplx*)k->roots.ptr, (const long long*)k->offsets.ptr, (const int*)k->nroots.ptr, (const int*)k->nunconv.ptr,
                                                      (const int*)k->members.ptr, (const long long*)ctx->offsets.ptr, (cplx*)ctx->roots.ptr, (int*)ctx->nroots.ptr,
                                                      (int*)ctx->nunconv.ptr);
        CUDA_TRY(ctx, cudaGetLastError());
        if (!k->ev_done) { CUDA_TRY(ctx, cudaEventCreateWithFlags(&k->ev_done, cudaEventDisableTiming)); }
        CUDA_TRY(ctx, cudaEventRecord(k->ev_done, ks));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, k->ev_done, 0));
        if (stats) {
            stats->turnovers += kst.turnovers; stats->sweeps += kst.sweeps; stats->exceptional += kst.exceptional; stats->unconverged += kst.unconverged;
            stats->fat_steps += kst.fat_steps; stats->launches += kst.launches + 2 + (pencil ? 1 : 0); stats->device_ms += kst.device_ms;
            if (kst.lanes > stats->lanes) stats->lanes = kst.lanes;
            if (kst.schedule_used > stats->schedule_used) stats->schedule_used = kst.schedule_used;
        }
    }
    // the host vectors above were pageable sources of asynchronous copies: those are staged before the call returns
    (void)nbatch;
    return AMRVW_OK;
}

// One device: H2D of the inputs, the kernels, D2H of the results (roots, or with counts3 != NULL only the
// imag == 0 / > 0 / < 0 counts of real_polynomial_roots), all on the context's stream; returns when the
// results are in the caller's buffers.
template <class S>
static int run_host_single(amrvw_ctx* ctx, Variant var, const S* a, const S* w, const int64_t* offsets, int64_t nbatch, int64_t max_len,
                           double* roots, int32_t* nroots, int32_t* nunconv, int32_t* counts3, amrvw_stats* stats) {
    const bool pencil = (var == V_PENCIL_R || var == V_PENCIL_C);
    const int64_t total = offsets[nbatch];
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = reserve(ctx, ctx->coeffs, (size_t)total * sizeof(S) + 16))) return rc;
    if (pencil && (rc = reserve(ctx, ctx->coeffs2, (size_t)total * sizeof(S) + 16))) return rc;
    if ((rc = reserve(ctx, ctx->offsets, (size_t)(nbatch + 1) * sizeof(int64_t)))) return rc;
    if ((rc = reserve(ctx, ctx->roots, ((size_t)total + 1) * sizeof(cplx)))) return rc;
    if ((rc = reserve(ctx, ctx->nroots, (size_t)nbatch * sizeof(int32_t)))) return rc;
    if ((rc = reserve(ctx, ctx->nunconv, (size_t)nbatch * sizeof(int32_t)))) return rc;
    if (counts3 && (rc = reserve(ctx, ctx->counts, (size_t)nbatch * 3 * sizeof(int32_t)))) return rc;
    cudaStream_t st = ctx->stream;
    if (total > 0) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->coeffs.ptr, a, (size_t)total * sizeof(S), cudaMemcpyHostToDevice, st));
    if (pencil && total > 0) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->coeffs2.ptr, w, (size_t)total * sizeof(S), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->offsets.ptr, offsets, (size_t)(nbatch + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    // a ragged batch: one sub-batch per degree class, each with its own lane count (run_classes)
    bool by_class = false;
    std::vector<std::vector<int>> members;
    if (ctx->opt.schedule != AMRVW_SCHEDULE_REFERENCE && ctx->opt.lanes_per_poly == 0 && var != V_PENCIL_C && nbatch >= 2 && nbatch < 0x7fffffff) {
        members.assign((size_t)NCLASS, std::vector<int>());
        for (int64_t p = 0; p < nbatch; ++p) {
            const int64_t len = offsets[p + 1] - offsets[p];
            int c = 0;
            while (c < NCLASS - 1 && len >= CLASS_BOUNDS[c]) ++c;
            members[(size_t)c].push_back((int)p);
        }
        int used = 0;
        for (const std::vector<int>& m : members) used += m.empty() ? 0 : 1;
        // ... when the batch is bound by the latency of its largest polynomials, not by throughput: only then do more lanes
        // for the top class pay for running several persistent kernels side by side (measured, tools/ragged_bench.py: 2000
        // polynomials of degree 200-3000 384 -> 297 ms, 5000 of degree 40-120 + 8 of degree 2500-3000 325 -> 162 ms; but 8000
        // of degree 100-1500 233 -> 283 ms and 6000 of degree 500-2000 325 -> 374 ms, which this test leaves on one launch).
        // Model (config 2 on one B200): the unit queue retires 2.9e11 turnovers/s, a polynomial of degree n costs 5.4 n^2 of
        // them and, on 16 lanes alone, 3.8e-8 n^2 seconds
        double sum_sq = 0.0;
        for (int64_t p = 0; p < nbatch; ++p) { const double len = (double)(offsets[p + 1] - offsets[p]); sum_sq += len * len; }
        const double t_throughput = 5.4 * sum_sq / 2.9e11, t_latency = 3.8e-8 * (double)max_len * (double)max_len;
        by_class = used >= 2 && t_latency > 1.5 * t_throughput;
    }
    if (by_class) {
        if (stats) std::memset(stats, 0, sizeof(*stats));
        rc = run_classes<S>(ctx, var, offsets, nbatch, members, stats);
    } else {
        rc = run_dev<S>(ctx, var, (const S*)ctx->coeffs.ptr, pencil ? (const S*)ctx->coeffs2.ptr : nullptr, (const int64_t*)ctx->offsets.ptr, nbatch, total, max_len,
                        (double*)ctx->roots.ptr, (int32_t*)ctx->nroots.ptr, (int32_t*)ctx->nunconv.ptr, stats);
    }
    if (rc) return rc;
    if (counts3) {
        const int threads = 128;
        count_real_roots_kernel<<<(int)((nbatch + threads - 1) / threads), threads, 0, st>>>((const cplx*)ctx->roots.ptr, (const long long*)ctx->offsets.ptr, nbatch, 0,
                                                                                           (const int*)ctx->nroots.ptr, (int*)ctx->counts.ptr);
        CUDA_TRY(ctx, cudaGetLastError());
        CUDA_TRY(ctx, cudaMemcpyAsync(counts3, ctx->counts.ptr, (size_t)nbatch * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    } else if (total > 0) {
        CUDA_TRY(ctx, cudaMemcpyAsync(roots, ctx->roots.ptr, (size_t)total * sizeof(cplx), cudaMemcpyDeviceToHost, st));
    }
    if (nroots) CUDA_TRY(ctx, cudaMemcpyAsync(nroots, ctx->nroots.ptr, (size_t)nbatch * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(nunconv, ctx->nunconv.ptr, (size_t)nbatch * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (by_class) {
        for (amrvw_ctx* k : ctx->klass) { const int r2 = check_queue_abort(k); if (r2) return fail(ctx, r2, k->err); }
        return AMRVW_OK;
    }
    return check_queue_abort(ctx);
}

// Several devices (amrvw_create_multi): contiguous split of the batch by polynomial, balanced by the sum of
// squared lengths (the work of a polynomial grows like n^2); one host thread per device drives that device's
// own context and stream (H2D slice -> kernels -> D2H slice straight into the caller's buffers, so the
// "gather" is the D2H copies themselves); per-device errors fold into one status.  SURVEY.md 8(e): no
// collective, polynomials never interact.
static void partition_batch(const int64_t* offsets, int64_t nbatch, int ndev, std::vector<int64_t>& cut) {
    cut.assign((size_t)ndev + 1, nbatch);
    cut[0] = 0;
    double tot = 0.0;
    for (int64_t p = 0; p < nbatch; ++p) { const double len = (double)(offsets[p + 1] - offsets[p]); tot += len * len + 64.0; }
    double acc = 0.0;
    int d = 1;
    for (int64_t p = 0; p < nbatch && d < ndev; ++p) {
        const double len = (double)(offsets[p + 1] - offsets[p]);
        acc += len * len + 64.0;
        while (d < ndev && acc >= tot * (double)d / (double)ndev) cut[(size_t)d++] = p + 1;
    }
}

template <class S>
static int run_host(amrvw_ctx* ctx, Variant var, const S* a, const S* w, const int64_t* offsets, int64_t nbatch,
                    double* roots, int32_t* nroots, int32_t* nunconv, int32_t* counts3, amrvw_stats* stats) {
    if (!ctx) return AMRVW_ERR_ARGUMENT;
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (nbatch < 0) return fail(ctx, AMRVW_ERR_ARGUMENT, "negative batch");
    if (nbatch == 0) return AMRVW_OK;
    if (!offsets || (!roots && !counts3) || (!nroots && !counts3) || !nunconv) return fail(ctx, AMRVW_ERR_ARGUMENT, "null buffer");
    const bool pencil = (var == V_PENCIL_R || var == V_PENCIL_C);
    int64_t max_len = 0;
    int rc = check_offsets(ctx, offsets, nbatch, &max_len);
    if (rc) return rc;
    const int64_t total = offsets[nbatch];
    if (total > 0 && (!a || (pencil && !w))) return fail(ctx, AMRVW_ERR_ARGUMENT, "null coefficient buffer");
    if (ctx->subs.empty()) return run_host_single<S>(ctx, var, a, w, offsets, nbatch, max_len, roots, nroots, nunconv, counts3, stats);

    const int ndev = (int)ctx->subs.size();
    std::vector<int64_t> cut;
    partition_batch(offsets, nbatch, ndev, cut);
    std::vector<int> rcs((size_t)ndev, AMRVW_OK);
    std::vector<amrvw_stats> sts((size_t)ndev);
    std::vector<std::thread> workers;
    for (int d = 0; d < ndev; ++d) {
        const int64_t p0 = cut[(size_t)d], p1 = cut[(size_t)d + 1];
        if (p1 <= p0) continue;
        workers.emplace_back([&, d, p0, p1]() {
            amrvw_ctx* sub = ctx->subs[(size_t)d];
            const int64_t base = offsets[p0], nb = p1 - p0;
            std::vector<int64_t> loc((size_t)nb + 1);
            int64_t ml = 0;
            for (int64_t i = 0; i <= nb; ++i) loc[(size_t)i] = offsets[p0 + i] - base;
            for (int64_t i = 0; i < nb; ++i) ml = std::max(ml, loc[(size_t)i + 1] - loc[(size_t)i]);
            rcs[(size_t)d] = run_host_single<S>(sub, var, a ? a + base : nullptr, (pencil && w) ? w + base : nullptr, loc.data(), nb, ml,
                                                roots ? roots + 2 * base : nullptr, nroots ? nroots + p0 : nullptr, nunconv + p0,
                                                counts3 ? counts3 + 3 * p0 : nullptr, stats ? &sts[(size_t)d] : nullptr);
        });
    }
    for (std::thread& t : workers) t.join();
    for (int d = 0; d < ndev; ++d)
        if (rcs[(size_t)d] != AMRVW_OK) return fail(ctx, rcs[(size_t)d], "device " + std::to_string(ctx->subs[(size_t)d]->device) + ": " + ctx->subs[(size_t)d]->err);
    if (stats) {
        for (int d = 0; d < ndev; ++d) {
            const amrvw_stats& x = sts[(size_t)d];
            stats->turnovers += x.turnovers; stats->sweeps += x.sweeps; stats->exceptional += x.exceptional; stats->unconverged += x.unconverged;
            stats->fat_steps += x.fat_steps; stats->launches += x.launches;
            if (x.device_ms > stats->device_ms) stats->device_ms = x.device_ms;    // the devices run side by side
            if (x.lanes > stats->lanes) stats->lanes = x.lanes;
            if (x.schedule_used > stats->schedule_used) stats->schedule_used = x.schedule_used;
        }
    }
    return AMRVW_OK;
}

extern "C" {

int amrvw_roots_rds_f64(amrvw_ctx* ctx, const double* coeffs, const int64_t* offsets, int64_t nbatch, double* roots, int32_t* nroots, int32_t* nunconv, amrvw_stats* stats) {
    return run_host<double>(ctx, V_RDS, coeffs, nullptr, offsets, nbatch, roots, nroots, nunconv, nullptr, stats);
}
int amrvw_roots_css_c64(amrvw_ctx* ctx, const double* coeffs, const int64_t* offsets, int64_t nbatch, double* roots, int32_t* nroots, int32_t* nunconv, amrvw_stats* stats) {
    return run_host<cplx>(ctx, V_CSS, (const cplx*)coeffs, nullptr, offsets, nbatch, roots, nroots, nunconv, nullptr, stats);
}
int amrvw_roots_pencil_f64(amrvw_ctx* ctx, const double* vs, const double* ws, const int64_t* offsets, int64_t nbatch, double* roots, int32_t* nroots, int32_t* nunconv, amrvw_stats* stats) {
    return run_host<double>(ctx, V_PENCIL_R, vs, ws, offsets, nbatch, roots, nroots, nunconv, nullptr, stats);
}
int amrvw_roots_pencil_c64(amrvw_ctx* ctx, const double* vs, const double* ws, const int64_t* offsets, int64_t nbatch, double* roots, int32_t* nroots, int32_t* nunconv, amrvw_stats* stats) {
    return run_host<cplx>(ctx, V_PENCIL_C, (const cplx*)vs, (const cplx*)ws, offsets, nbatch, roots, nroots, nunconv, nullptr, stats);
}

int amrvw_partition_batch(const int64_t* offsets, int64_t nbatch, int ndev, int64_t* cuts) {
    if (!offsets || !cuts || nbatch < 0 || ndev < 1) return AMRVW_ERR_ARGUMENT;
    for (int64_t p = 0; p < nbatch; ++p) if (offsets[p + 1] < offsets[p]) return AMRVW_ERR_ARGUMENT;
    std::vector<int64_t> cut;
    partition_batch(offsets, nbatch, ndev, cut);
    for (int d = 0; d <= ndev; ++d) cuts[d] = cut[(size_t)d];
    return AMRVW_OK;
}

int amrvw_count_real_roots_rds_f64(amrvw_ctx* ctx, const double* coeffs, const int64_t* offsets, int64_t nbatch, int32_t* counts3, int32_t* nunconv, amrvw_stats* stats) {
    if (!counts3) return fail(ctx, AMRVW_ERR_ARGUMENT, "null counts buffer");
    return run_host<double>(ctx, V_RDS, coeffs, nullptr, offsets, nbatch, nullptr, nullptr, nunconv, counts3, stats);
}

int amrvw_roots_rds_f64_dev(amrvw_ctx* ctx, const double* a, const int64_t* off, int64_t nb, int64_t total, int64_t max_len, double* r, int32_t* nr, int32_t* nu, amrvw_stats* stats) {
    return run_dev<double>(ctx, V_RDS, a, nullptr, off, nb, total, max_len, r, nr, nu, stats);
}
int amrvw_roots_css_c64_dev(amrvw_ctx* ctx, const double* a, const int64_t* off, int64_t nb, int64_t total, int64_t max_len, double* r, int32_t* nr, int32_t* nu, amrvw_stats* stats) {
    return run_dev<cplx>(ctx, V_CSS, (const cplx*)a, nullptr, off, nb, total, max_len, r, nr, nu, stats);
}
int amrvw_roots_pencil_f64_dev(amrvw_ctx* ctx, const double* v, const double* w, const int64_t* off, int64_t nb, int64_t total, int64_t max_len, double* r, int32_t* nr, int32_t* nu, amrvw_stats* stats) {
    return run_dev<double>(ctx, V_PENCIL_R, v, w, off, nb, total, max_len, r, nr, nu, stats);
}
int amrvw_roots_pencil_c64_dev(amrvw_ctx* ctx, const double* v, const double* w, const int64_t* off, int64_t nb, int64_t total, int64_t max_len, double* r, int32_t* nr, int32_t* nu, amrvw_stats* stats) {
    return run_dev<cplx>(ctx, V_PENCIL_C, (const cplx*)v, (const cplx*)w, off, nb, total, max_len, r, nr, nu, stats);
}

int amrvw_count_real_roots_dev(amrvw_ctx* ctx, const double* d_roots, const int64_t* d_offsets, int64_t nbatch, int32_t shift, const int32_t* d_nroots, int32_t* d_counts3) {
    ctx = primary(ctx);
    if (!ctx) return AMRVW_ERR_ARGUMENT;
    if (nbatch <= 0) return AMRVW_OK;
    if (!d_roots || !d_offsets || !d_nroots || !d_counts3) return fail(ctx, AMRVW_ERR_ARGUMENT, "null buffer");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int threads = 128;
    count_real_roots_kernel<<<(int)((nbatch + threads - 1) / threads), threads, 0, ctx->stream>>>((const cplx*)d_roots, (const long long*)d_offsets, nbatch, shift, d_nroots, d_counts3);
    CUDA_TRY(ctx, cudaGetLastError());
    return AMRVW_OK;
}

int amrvw_last_profile(amrvw_ctx* ctx, int64_t* out16) {
    ctx = primary(ctx);
    if (!ctx || !out16) return AMRVW_ERR_ARGUMENT;
    std::memset(out16, 0, 16 * sizeof(int64_t));
    if (ctx->prof.ptr && ctx->prof_valid) {
        CUDA_TRY(ctx, cudaSetDevice(ctx->device));
        CUDA_TRY(ctx, cudaMemcpyAsync(out16, ctx->prof.ptr, 16 * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    // kernel times of the unit schedule, microseconds
    out16[8] = (int64_t)(ctx->times.setup_ms * 1e3); out16[9] = (int64_t)(ctx->times.main_ms * 1e3);   // [10] driver cycles, [13] pop cycles stay as counted
    out16[11] = (int64_t)(ctx->times.small_ms * 1e3); out16[12] = (int64_t)(ctx->times.finish_ms * 1e3);
    return AMRVW_OK;
}

int amrvw_debug_counters(amrvw_ctx* ctx, int64_t* out48) {
    if (!ctx || !out48) return AMRVW_ERR_ARGUMENT;
    ctx = primary(ctx);
    std::memset(out48, 0, 48 * sizeof(int64_t));
    if (ctx->prof.ptr && ctx->prof_valid) {
        CUDA_TRY(ctx, cudaSetDevice(ctx->device));
        CUDA_TRY(ctx, cudaMemcpyAsync(out48, (const int64_t*)ctx->prof.ptr + 16, 48 * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (ctx->counters.ptr) {      // [47]: polynomials that changed units in the last unit-queue call (compaction of half-empty units, rounds.cuh)
        unsigned adopted = 0;
        CUDA_TRY(ctx, cudaSetDevice(ctx->device));
        CUDA_TRY(ctx, cudaMemcpyAsync(&adopted, (const unsigned*)ctx->counters.ptr + CNT_ADOPTED, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        out48[47] = (int64_t)adopted;
    }
    return AMRVW_OK;
}

int amrvw_selftest(amrvw_ctx* ctx, int64_t n, double* out4) {
    ctx = primary(ctx);
    if (!ctx || !out4 || n <= 0) return AMRVW_ERR_ARGUMENT;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = reserve(ctx, ctx->state, 4 * sizeof(double));
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->state.ptr, 0, 4 * sizeof(double), st));
    selftest_kernel<<<ctx->sm_count, 128, 0, st>>>((double*)ctx->state.ptr, (int)n, ctx->opt.seed + 12345);
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(out4, ctx->state.ptr, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return AMRVW_OK;
}

int amrvw_selftest_complex(amrvw_ctx* ctx, int64_t n, double* out4) {
    ctx = primary(ctx);
    if (!ctx || !out4 || n <= 0) return AMRVW_ERR_ARGUMENT;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = reserve(ctx, ctx->state, 4 * sizeof(double));
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->state.ptr, 0, 4 * sizeof(double), st));
    selftest_complex_kernel<<<ctx->sm_count, 128, 0, st>>>((double*)ctx->state.ptr, (int)n, ctx->opt.seed + 54321);
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(out4, ctx->state.ptr, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return AMRVW_OK;
}

int amrvw_selftest_robust(amrvw_ctx* ctx, int64_t n, double* out4) {
    ctx = primary(ctx);
    if (!ctx || !out4 || n <= 0) return AMRVW_ERR_ARGUMENT;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = reserve(ctx, ctx->state, 4 * sizeof(double));
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->state.ptr, 0, 4 * sizeof(double), st));
    selftest_robust_kernel<<<ctx->sm_count, 128, 0, st>>>((double*)ctx->state.ptr, (int)n, ctx->opt.seed + 777);
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(out4, ctx->state.ptr, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return AMRVW_OK;
}

int amrvw_measure_fp64_peak(amrvw_ctx* ctx, double* tflops) {
    ctx = primary(ctx);
    if (!ctx || !tflops) return AMRVW_ERR_ARGUMENT;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 4096;
    int rc = reserve(ctx, ctx->state, (size_t)blocks * threads * sizeof(double));
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, st));
        fp64_peak_kernel<<<blocks, threads, 0, st>>>((double*)ctx->state.ptr, iters, 1.0 + rep);
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        float ms = 0.f;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
    *tflops = flops / (best * 1e-3) / 1e12;
    return AMRVW_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- Hessenberg eigenproblems
// eigvals(qr_factorization(H; unitary)) for a batch of upper Hessenberg matrices from host buffers: one warp per matrix
// (hessenberg.cuh), R in shared memory when the largest matrix of the batch fits, else in a global workspace.
template <class S>
static int run_hessenberg_single(amrvw_ctx* ctx, const S* H, const int32_t* dims, int64_t nbatch, int unitary, double* eig, int32_t* nunconv, amrvw_stats* stats);

// contiguous split of a batch of matrices over ndev devices, balanced by the sum of n^3; cut[d] .. cut[d+1]-1 are device d's
// matrices, moff / eoff the element and eigenvalue offsets of its first one
static void partition_hessenberg(const int32_t* dims, int64_t nbatch, int ndev, std::vector<int64_t>& cut, std::vector<int64_t>& moff, std::vector<int64_t>& eoff) {
    double work = 0.0;
    for (int64_t p = 0; p < nbatch; ++p) { const double n = dims[p]; work += n * n * n; }
    cut.assign((size_t)ndev + 1, nbatch); moff.assign((size_t)ndev + 1, 0); eoff.assign((size_t)ndev + 1, 0);
    cut[0] = 0;
    double acc = 0.0; long long mo = 0, eo = 0; int d = 1;
    for (int64_t p = 0; p < nbatch; ++p) {
        while (d < ndev && acc >= work * d / ndev) { cut[(size_t)d] = p; moff[(size_t)d] = mo; eoff[(size_t)d] = eo; ++d; }
        const long long n = dims[p];
        acc += (double)n * n * n; mo += n * n; eo += n;
    }
    while (d <= ndev) { cut[(size_t)d] = nbatch; moff[(size_t)d] = mo; eoff[(size_t)d] = eo; ++d; }
}

// Several devices: contiguous split by matrix, balanced by the sum of n^3 (the work of the dense-R chase), one host thread per
// device, results copied straight into the caller's buffers -- the same scheme as run_host.
template <class S>
static int run_hessenberg(amrvw_ctx* ctx, const S* H, const int32_t* dims, int64_t nbatch, int unitary, double* eig, int32_t* nunconv, amrvw_stats* stats) {
    if (!ctx) return AMRVW_ERR_ARGUMENT;
    if (nbatch < 0) return fail(ctx, AMRVW_ERR_ARGUMENT, "negative batch size");
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (nbatch == 0) return AMRVW_OK;
    if (!H || !dims || !eig || !nunconv) return fail(ctx, AMRVW_ERR_ARGUMENT, "null buffer");
    if (nbatch > 0x7fffffffLL) return fail(ctx, AMRVW_ERR_ARGUMENT, "batch too large");
    for (int64_t p = 0; p < nbatch; ++p)
        if (dims[p] < 0 || dims[p] > 32768) return fail(ctx, AMRVW_ERR_ARGUMENT, "matrix dimension out of range (0..32768)");
    if (ctx->subs.empty()) return run_hessenberg_single<S>(ctx, H, dims, nbatch, unitary, eig, nunconv, stats);
    const int ndev = (int)ctx->subs.size();
    std::vector<int64_t> cut, moff, eoff;
    partition_hessenberg(dims, nbatch, ndev, cut, moff, eoff);
    std::vector<int> rcs((size_t)ndev, AMRVW_OK);
    std::vector<amrvw_stats> sts((size_t)ndev);
    std::vector<std::thread> workers;
    for (int d = 0; d < ndev; ++d) {
        const int64_t p0 = cut[(size_t)d], p1 = cut[(size_t)d + 1];
        if (p1 <= p0) continue;
        workers.emplace_back([&, d, p0, p1]() {
            rcs[(size_t)d] = run_hessenberg_single<S>(ctx->subs[(size_t)d], H + moff[(size_t)d], dims + p0, p1 - p0, unitary, eig + 2 * eoff[(size_t)d], nunconv + p0,
                                                      stats ? &sts[(size_t)d] : nullptr);
        });
    }
    for (std::thread& t : workers) t.join();
    for (int d = 0; d < ndev; ++d)
        if (rcs[(size_t)d] != AMRVW_OK) return fail(ctx, rcs[(size_t)d], "device " + std::to_string(ctx->subs[(size_t)d]->device) + ": " + ctx->subs[(size_t)d]->err);
    if (stats) {
        for (int d = 0; d < ndev; ++d) {
            const amrvw_stats& x = sts[(size_t)d];
            stats->turnovers += x.turnovers; stats->sweeps += x.sweeps; stats->exceptional += x.exceptional; stats->unconverged += x.unconverged;
            stats->launches += x.launches; stats->device_ms = std::max(stats->device_ms, x.device_ms); stats->lanes = std::max(stats->lanes, x.lanes);
        }
        stats->schedule_used = AMRVW_SCHEDULE_REFERENCE;
    }
    return AMRVW_OK;
}

// The launch itself, device pointers in and out, asynchronous on the context's stream unless stats are asked for.
// d_moff / d_eoff: element offset of matrix p in d_H and eigenvalue slot offset (nbatch + 1 each); nmax: largest dimension.
template <class S>
static int run_hessenberg_dev(amrvw_ctx* ctx, const S* d_H, const int32_t* d_dims, const int64_t* d_moff, const int64_t* d_eoff, int64_t nbatch, int nmax, int unitary,
                              double* d_eig, int32_t* d_nunconv, amrvw_stats* stats) {
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (nmax < 1) nmax = 1;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    (void)cudaGetLastError();
    cudaStream_t st = ctx->stream;
    const int ld = nmax | 1;                                    // odd: rows and columns of R both map to distinct banks
    const size_t smem_cap = 220 * 1024;
    const size_t full = hess_warp_bytes<S>(nmax, ld, true), lean = hess_warp_bytes<S>(nmax, ld, false);
    // the chase is latency bound: 16 resident warps per SM matter more than where R lives.  R stays in shared memory only while
    // 16 warps' worth fits (real n <= 39, complex n <= 27; measured: n = 32 +12 %), else in the global workspace, which the L2 holds
    // (n = 64: 1.4x, n = 96: 3.2x, n = 128: 4.9x over one-warp-per-SM shared memory; profiles/r02_hessenberg.json)
    const bool in_smem = full * 16 <= smem_cap;
    const size_t per_warp = in_smem ? full : lean;
    if (per_warp > smem_cap) return fail(ctx, AMRVW_ERR_ARGUMENT, "matrix too large for the per-warp rotator storage");
    // small matrices, many of them: one thread per matrix (hess_thread_kernel) when a warp's worth of matrices fits in an SM's
    // shared memory and the batch fills the device (lanes_per_poly = 32 keeps one warp per matrix: development / comparison)
    size_t tstride = (full + 15) / 16; if (tstride % 2 == 0) ++tstride; tstride *= 16;      // odd multiple of 16 bytes
    const bool by_thread = tstride <= 5000 && nbatch >= 64ll * ctx->sm_count && ctx->opt.lanes_per_poly != 32;   // real n <= 22, complex n <= 15 (measured: n = 8 3.5x, 15 2.0x / 1.6x, 24 1.0x)
    // resident warps per SM (measured, tools/hessenberg_bench.py): R in shared memory: up to 20 (96 registers per thread) in CTAs of
    // four (n = 24, 32: +8-10 % over 16 in CTAs of eight); R in the global workspace: 16 in CTAs of eight (20 gain nothing at n = 64..128;
    // CTAs of four lose 12 % at n = 256)
    const int fit = (int)std::min<size_t>(in_smem ? 20 : 16, smem_cap / per_warp);
    int wpc = std::min(in_smem ? 4 : 8, fit);
    if ((int64_t)wpc > nbatch) wpc = (int)nbatch;
    int per_sm = std::max(1, fit / wpc);
    int64_t ctas = (nbatch + wpc - 1) / wpc;
    if (ctas > (int64_t)ctx->sm_count * per_sm) ctas = (int64_t)ctx->sm_count * per_sm;
    if (by_thread) {      // CTAs of one warp, as many per SM as the shared memory holds (at most 8)
        per_sm = (int)std::min<size_t>(8, smem_cap / (tstride * 32));
        ctas = std::min<int64_t>((nbatch + 31) / 32, (int64_t)ctx->sm_count * per_sm);
    }
    const size_t work_stride = ((size_t)nmax * ld * sizeof(S) + 255) & ~(size_t)255;
    int rc;
    if ((rc = reserve(ctx, ctx->stats, (size_t)nbatch * sizeof(PolyStats) + sizeof(StatsAccum)))) return rc;
    if ((rc = reserve(ctx, ctx->counters, 64 * sizeof(unsigned int)))) return rc;
    if (!in_smem && !by_thread && (rc = reserve(ctx, ctx->hwork, work_stride * (size_t)ctas * wpc))) return rc;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->counters.ptr, 0, sizeof(unsigned int), st));
    {
        const AlgoConst ac = {ctx->opt.tol, ctx->opt.it_count, ctx->opt.it_max_factor};
        CUDA_TRY(ctx, cudaMemcpyToSymbolAsync(g_algo, &ac, sizeof(ac), 0, cudaMemcpyHostToDevice, st));
    }
    HessArgs a;
    a.H = d_H; a.dims = d_dims; a.moff = (const long long*)d_moff; a.eoff = (const long long*)d_eoff;
    a.eig = (cplx*)d_eig; a.nunconv = d_nunconv; a.stats = (PolyStats*)ctx->stats.ptr;
    a.next = (unsigned int*)ctx->counters.ptr; a.work = (in_smem || by_thread) ? nullptr : ctx->hwork.ptr; a.work_stride = work_stride;
    a.nbatch = nbatch; a.unitary = unitary ? 1 : 0; a.ld_max = ld; a.seed = ctx->opt.seed;
    if (stats) CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, st));
    if (by_thread) {
        const size_t smem = tstride * 32;
        CUDA_TRY(ctx, cudaFuncSetAttribute(hess_thread_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        hess_thread_kernel<S><<<(int)ctas, 32, smem, st>>>(a, nmax, (unsigned)tstride);
    } else {
        const size_t smem = per_warp * wpc;
        auto kern = in_smem ? hess_kernel<S, true> : hess_kernel<S, false>;
        CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(int)ctas, 32 * wpc, smem, st>>>(a, nmax);
    }
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->prof_valid = false;
    if (stats) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, st));
        StatsAccum* acc = (StatsAccum*)((char*)ctx->stats.ptr + (size_t)nbatch * sizeof(PolyStats));
        CUDA_TRY(ctx, cudaMemsetAsync(acc, 0, sizeof(StatsAccum), st));
        reduce_stats_kernel<<<64, 256, 0, st>>>((const PolyStats*)ctx->stats.ptr, nbatch, acc);
        CUDA_TRY(ctx, cudaGetLastError());
        StatsAccum h;
        CUDA_TRY(ctx, cudaMemcpyAsync(&h, acc, sizeof(h), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        float ms = 0.f;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        stats->turnovers = h.turnovers; stats->sweeps = h.sweeps; stats->exceptional = h.exceptional; stats->unconverged = h.unconverged;
        stats->device_ms = ms; stats->launches = 1;
        stats->lanes = by_thread ? 2 : (in_smem ? 1 : 0);   // 2 = one thread per matrix (shared memory), 1 = one warp per matrix with R in shared memory, 0 = R in the global workspace
        stats->schedule_used = AMRVW_SCHEDULE_REFERENCE;
    }
    return AMRVW_OK;
}

// One device, host buffers: H2D, the launch, D2H; returns when the results are in the caller's buffers.
template <class S>
static int run_hessenberg_single(amrvw_ctx* ctx, const S* H, const int32_t* dims, int64_t nbatch, int unitary, double* eig, int32_t* nunconv, amrvw_stats* stats) {
    if (stats) std::memset(stats, 0, sizeof(*stats));
    std::vector<long long> meta(2 * (size_t)(nbatch + 1));
    long long* moff = meta.data(); long long* eoff = moff + nbatch + 1;
    int nmax = 1;
    moff[0] = 0; eoff[0] = 0;
    for (int64_t p = 0; p < nbatch; ++p) {
        const long long n = dims[p];
        if (n < 0 || n > 32768) return fail(ctx, AMRVW_ERR_ARGUMENT, "matrix dimension out of range (0..32768)");
        moff[p + 1] = moff[p] + n * n; eoff[p + 1] = eoff[p] + n;
        if (n > nmax) nmax = (int)n;
    }
    const long long nelem = moff[nbatch], neig = eoff[nbatch];
    if (neig == 0) { std::fill(nunconv, nunconv + nbatch, 0); return AMRVW_OK; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int rc;
    if ((rc = reserve(ctx, ctx->coeffs, (size_t)nelem * sizeof(S)))) return rc;
    if ((rc = reserve(ctx, ctx->hmeta, meta.size() * sizeof(long long) + (size_t)nbatch * sizeof(int)))) return rc;
    if ((rc = reserve(ctx, ctx->roots, (size_t)neig * sizeof(cplx)))) return rc;
    if ((rc = reserve(ctx, ctx->nunconv, (size_t)nbatch * sizeof(int)))) return rc;
    long long* d_meta = (long long*)ctx->hmeta.ptr;
    int* d_dims = (int*)(d_meta + meta.size());
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->coeffs.ptr, H, (size_t)nelem * sizeof(S), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_meta, meta.data(), meta.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_dims, dims, (size_t)nbatch * sizeof(int), cudaMemcpyHostToDevice, st));
    if ((rc = run_hessenberg_dev<S>(ctx, (const S*)ctx->coeffs.ptr, d_dims, (const int64_t*)d_meta, (const int64_t*)(d_meta + nbatch + 1), nbatch, nmax, unitary,
                                    (double*)ctx->roots.ptr, (int32_t*)ctx->nunconv.ptr, stats))) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(eig, ctx->roots.ptr, (size_t)neig * sizeof(cplx), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(nunconv, ctx->nunconv.ptr, (size_t)nbatch * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return AMRVW_OK;
}

extern "C" {
int amrvw_eigvals_hessenberg_f64(amrvw_ctx* ctx, const double* H, const int32_t* dims, int64_t nbatch, int unitary, double* eig, int32_t* nunconv, amrvw_stats* stats) {
    return run_hessenberg<double>(ctx, H, dims, nbatch, unitary, eig, nunconv, stats);
}
int amrvw_eigvals_hessenberg_c64(amrvw_ctx* ctx, const double* H, const int32_t* dims, int64_t nbatch, int unitary, double* eig, int32_t* nunconv, amrvw_stats* stats) {
    return run_hessenberg<cplx>(ctx, (const cplx*)H, dims, nbatch, unitary, eig, nunconv, stats);
}
int amrvw_partition_hessenberg(const int32_t* dims, int64_t nbatch, int ndev, int64_t* cuts) {
    if (!dims || !cuts || nbatch < 0 || ndev < 1) return AMRVW_ERR_ARGUMENT;
    for (int64_t p = 0; p < nbatch; ++p) if (dims[p] < 0 || dims[p] > 32768) return AMRVW_ERR_ARGUMENT;
    std::vector<int64_t> cut, moff, eoff;
    partition_hessenberg(dims, nbatch, ndev, cut, moff, eoff);
    for (int d = 0; d <= ndev; ++d) cuts[d] = cut[(size_t)d];
    return AMRVW_OK;
}
static int hess_dev_args(amrvw_ctx*& ctx, const void* H, const void* dims, const void* moff, const void* eoff, int64_t nbatch, int32_t max_dim, const void* eig, const void* nunconv) {
    if (!ctx) return AMRVW_ERR_ARGUMENT;
    if (!ctx->subs.empty()) return fail(ctx, AMRVW_ERR_ARGUMENT, "device-pointer entry points need a single-device context");
    if (nbatch < 0 || max_dim < 0 || max_dim > 32768) return fail(ctx, AMRVW_ERR_ARGUMENT, "batch size or largest dimension out of range");
    if (nbatch > 0 && (!H || !dims || !moff || !eoff || !eig || !nunconv)) return fail(ctx, AMRVW_ERR_ARGUMENT, "null buffer");
    return AMRVW_OK;
}
int amrvw_eigvals_hessenberg_f64_dev(amrvw_ctx* ctx, const double* d_H, const int32_t* d_dims, const int64_t* d_moff, const int64_t* d_eoff, int64_t nbatch, int32_t max_dim,
                                     int unitary, double* d_eig, int32_t* d_nunconv, amrvw_stats* stats) {
    const int rc = hess_dev_args(ctx, d_H, d_dims, d_moff, d_eoff, nbatch, max_dim, d_eig, d_nunconv);
    if (rc || nbatch == 0) { if (!rc && stats) std::memset(stats, 0, sizeof(*stats)); return rc; }
    return run_hessenberg_dev<double>(ctx, d_H, d_dims, d_moff, d_eoff, nbatch, max_dim, unitary, d_eig, d_nunconv, stats);
}
int amrvw_eigvals_hessenberg_c64_dev(amrvw_ctx* ctx, const double* d_H, const int32_t* d_dims, const int64_t* d_moff, const int64_t* d_eoff, int64_t nbatch, int32_t max_dim,
                                     int unitary, double* d_eig, int32_t* d_nunconv, amrvw_stats* stats) {
    const int rc = hess_dev_args(ctx, d_H, d_dims, d_moff, d_eoff, nbatch, max_dim, d_eig, d_nunconv);
    if (rc || nbatch == 0) { if (!rc && stats) std::memset(stats, 0, sizeof(*stats)); return rc; }
    return run_hessenberg_dev<cplx>(ctx, (const cplx*)d_H, d_dims, d_moff, d_eoff, nbatch, max_dim, unitary, d_eig, d_nunconv, stats);
}
}
