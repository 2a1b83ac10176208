// stream.cuh -- chained warps fed by a STREAM of fat steps of different polynomials (no fill/drain bubbles).
//
// chain.cuh gives one polynomial a CTA of NW warps chained into a systolic array of LE = 32 NW bulges.  A fat
// step of window w costs w + 3 LE lock-steps there: 3 LE of them fill and drain the array (38% of all lock-steps
// at 375 polynomials of degree 3000, LE = 128), and between two fat steps of the polynomial the whole CTA waits
// for phase A (driver + shifts, 19%) on one warp.  Polynomials are independent, so here the array never drains:
//
//   * a JOB is one fat step of one polynomial (window, shifts, bulge count -- the record phase A leaves behind).
//     The feeding lane streams the rotator triples of job after job into the array; every triple carries a TAG
//     (job slot, chain index, FIRST / LAST of its window), and each lane decides from the tags in its own
//     three-index window what to do in a lock-step: create its bulge for the job that is entering (window =
//     [*, start, start+1] -- one step BEFORE the bulge's first chase position, where the lane is idle anyway),
//     chase (7 turnovers), fuse the bulge into Q at the end of the window (split over two lock-steps: positions
//     [stop-1, stop, stop+1] and [stop, stop+1, *]), or pass the triple on.  The bulges of the next job follow
//     the last bulges of the previous one three indices behind, as the bulges of one job follow each other.
//   * phase A runs on a SERVICE warp of every CTA (driver on one lane, dense trailing block + Hessenberg QR on
//     the warp: rounds.cuh), for whichever polynomial needs it, while the chase warps work on other jobs; a
//     MANAGER warp moves polynomials between two device-wide queues (needs phase A / ready to chase) and the
//     CTA's ring of job descriptors in shared memory.  Jobs of one polynomial migrate between CTAs: every CTA
//     sees the average speed (same argument as the unit queue of rounds.cuh), and SMs with two or three
//     resident CTAs finish together.
//
// Per polynomial the operations and their order are those of the sequential multishift schedule (bulges three
// indices apart commute exactly, creation one step early reads final values, the split last step does the same
// turnovers in the same order): the -DAMRVW_STRICT build is bit-identical to the oracle's amrvw_algorithm_ms
// (tests), whatever the interleaving of polynomials.
#pragma once
#include "chain.cuh"

namespace amrvw {

// ---------------------------------------------------------------- device-wide queues of polynomial ids
// Bounded multi-producer multi-consumer queue with per-slot sequence numbers (after D. Vyukov): push never
// blocks (capacity > polynomials alive), try_pop returns -1 when the queue is empty instead of waiting -- the
// manager warps poll, nobody holds a ticket for work that may never come.
struct MpmcSlot { unsigned seq; int val; };
struct MpmcQueue { unsigned head, tail, pad0, pad1; };
struct StreamQueues {
    MpmcQueue* ready; MpmcSlot* ready_slots;     // polynomials whose next fat step is prepared
    MpmcQueue* work; MpmcSlot* work_slots;       // polynomials that need phase A
    unsigned mask;                                // capacity - 1 (both)
    unsigned* remaining;                          // polynomials still iterating
    unsigned* progress;                           // fat steps completed, device-wide (watchdog)
    unsigned* abort_flag;                         // counters[CNT_ABORT]: sticky, turned into AMRVW_ERR_QUEUE_TIMEOUT by the host
};
AMRVW_DI unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
AMRVW_DI unsigned ld_volatile_u32(const unsigned* p) { return *reinterpret_cast<const volatile unsigned*>(p); }

AMRVW_DI void mpmc_push(MpmcQueue* q, MpmcSlot* slots, const unsigned mask, const int v) {
    const unsigned pos = atomicAdd(&q->tail, 1u);
    MpmcSlot* s = slots + (pos & mask);
    while (ld_volatile_u32(&s->seq) != pos) __nanosleep(100);   // the slot's previous round has been popped (always, with capacity > items alive)
    __stcg(&s->val, v);
    __threadfence();
    atomicExch(&s->seq, pos + 1u);
}
AMRVW_DI int mpmc_try_pop(MpmcQueue* q, MpmcSlot* slots, const unsigned mask) {
    for (;;) {
        const unsigned pos = ld_volatile_u32(&q->head);
        MpmcSlot* s = slots + (pos & mask);
        const unsigned seq = ld_volatile_u32(&s->seq);
        const int diff = (int)(seq - (pos + 1u));
        if (diff == 0) {
            if (atomicCAS(&q->head, pos, pos + 1u) == pos) {
                __threadfence();
                const int v = __ldcg(&s->val);
                atomicExch(&s->seq, pos + mask + 1u);
                return v;
            }
        } else if (diff < 0) {
            return -1;
        }
    }
}

__global__ void stream_init_queues_kernel(StreamQueues sq) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= sq.mask) {
        sq.ready_slots[i].seq = i; sq.ready_slots[i].val = -1;
        sq.work_slots[i].seq = i; sq.work_slots[i].val = -1;
    }
    if (i == 0) {
        sq.ready->head = sq.ready->tail = 0u; sq.work->head = sq.work->tail = 0u;
        *sq.remaining = 0u; *sq.progress = 0u;
    }
}
// every iterating polynomial starts in the phase A queue, largest first (pr.order)
__global__ void stream_enqueue_kernel(Problem<double> pr, StreamQueues sq) {
    const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= pr.nbatch) return;
    const long long p = unit_poly(pr, slot);
    if (!pr.rec[p].iterating) return;
    atomicAdd(sq.remaining, 1u);
    mpmc_push(sq.work, sq.work_slots, sq.mask, (int)p);
}

// ---------------------------------------------------------------- tags and job descriptors
// tag of a rotator triple in flight: valid | first | last | job slot (4 bits) | chain index (24 bits)
#define AMRVW_TAG_VALID 0x80000000u
#define AMRVW_TAG_FIRST 0x40000000u
#define AMRVW_TAG_LAST 0x20000000u
#define AMRVW_TAG_SLOT_SHIFT 24
#define AMRVW_TAG_INDEX_MASK 0x00ffffffu
#define AMRVW_STREAM_JOBS 16          // descriptors per CTA: jobs in the array (<= 3 LE / smallest window + 2) plus the prefetched one

struct StreamJob {
    long long base;       // chain slot of the polynomial: offsets[p] + extra * p
    int p, start, stop, nb;
    int mode, ord0, sp, pad;
};
template <int NW> struct StreamSmem {
    StreamJob jobs[AMRVW_STREAM_JOBS];
    double2 ring[3 * AMRVW_RING];          // look-ahead ring of the feeding thread
    unsigned ringtag[AMRVW_RING];
    double2 hand[NW][2][3];                // hand[w][parity]: triple leaving warp w at the end of a step of that parity
    unsigned handtag[NW][2];
    volatile int jobs_tail;                // manager: descriptors written so far
    volatile int jobs_head;                // feeding thread: jobs started
    volatile int jobs_done;                // retiring thread: jobs whose last triple has been written back
    volatile int feed_left;                // feeding thread: triples of the current job not yet requested
    volatile int no_more;                  // manager: nothing iterates any more
    volatile int quit;                     // feeding thread: the array is empty and nothing will come
    volatile int idle;                     // feeding thread: nothing in the array right now
};

struct Created { Rot<double> qa, qb, U, V, W; };
// create_bulge + absorb_Ut (create-bulge.jl:20-77, RDS.jl:14-36) for the bulge of chain lane `bl` of a job: the
// lane's window holds final values of indices start (qa, ba, ca) and start+1 (qb, bb, cb); index start-1 never
// changes during the job and is read here, like the lane's shifts.
AMRVW_NI Created stream_create(const Rot<double> qa, const Rot<double> qb, const Rot<double> ba, const Rot<double> bb, const Rot<double> ca, const Rot<double> cb,
                               const Rot<double>* FQ, const Rot<double>* FB, const Rot<double>* FC, const int N, const int start, const int stop, const int mode,
                               const Shifts* shl, const int bl, const uint64_t seed, const uint64_t poly, const int ordinal) {
    const Rot<double> one = make_rot<double>(1.0, 0.0);
    const Rot<double> qm = ldcg_rot(FQ + start - 1);                // Q[0] is the (1,0) sentinel
    Rot<double> bm = one, cm = one;
    if (start > 1) { bm = ldcg_rot(FB + start - 1); cm = ldcg_rot(FC + start - 1); }
    Shifts sh;
    sh.e1 = cplx(0.0, 0.0); sh.e2 = cplx(0.0, 0.0);
    if (mode == 1) {
        const double2 v1 = __ldcg(reinterpret_cast<const double2*>(shl + bl)), v2 = __ldcg(reinterpret_cast<const double2*>(shl + bl) + 1);
        sh.e1 = cplx(v1.x, v1.y); sh.e2 = cplx(v2.x, v2.y);
    }
    StepState z;
    z.q0 = qa; z.q1 = qb; z.q2 = one; z.b0 = ba; z.b1 = bb; z.b2 = one; z.c0 = ca; z.c1 = cb; z.c2 = one;   // index start+2 is not read
    z.U = one; z.V = one; z.W = one;
    Fact<double, false> F;
    F.Q = F.B = F.Ct = F.WB = F.WCt = nullptr; F.DQ = F.DR = F.WD = nullptr; F.N = N; F.turnovers = 0;
    create_and_absorb(z, qm, bm, cm, F, start, stop, mode, sh, seed, poly, ordinal);
    Created o;
    o.qa = z.q0; o.qb = z.q1; o.U = z.U; o.V = z.V; o.W = z.W;
    return o;
}

#define AMRVW_BAR_CHASE(n) asm volatile("bar.sync 1, %0;" ::"n"(n) : "memory")

// ---------------------------------------------------------------- the three roles of a CTA
// manager warp (one lane works): completed jobs go back to their records and into the phase A queue; the next
// ready polynomial is fetched into the descriptor ring when the feeding thread is about to need it
template <int NW>
AMRVW_DI void stream_manager(const Problem<double>& pr, const StreamQueues& sq, StreamSmem<NW>& sm, const int LE) {
    if ((threadIdx.x & 31) != 0) return;
    int mgr_done = 0;
    // watchdog: no fat step completed anywhere on the device for ~20 s although polynomials are still iterating means
    // something is broken (a lost polynomial would otherwise hang the device): raise the sticky abort flag, leave
    unsigned long long wd_time = global_ns();
    unsigned wd_progress = ld_volatile_u32(sq.progress), wd_polls = 0;
    for (;;) {
        bool worked = false;
        if ((++wd_polls & 1023u) == 0u) {
            const unsigned pg = ld_volatile_u32(sq.progress);
            const unsigned long long now = global_ns();
            if (pg != wd_progress) { wd_progress = pg; wd_time = now; }
            else if (now - wd_time > 20000000000ull && ld_volatile_u32(sq.remaining) != 0u) atomicExch(sq.abort_flag, 1u);
            if (ld_volatile_u32(sq.abort_flag) != 0u) sm.no_more = 2;
        }
        while (mgr_done < sm.jobs_done) {    // finished fat steps: record, then phase A is due
            __threadfence_block();
            const StreamJob j = sm.jobs[mgr_done % AMRVW_STREAM_JOBS];
            PolyRec* r = pr.rec + j.p;
            PolyRec rec = load_rec(r);
            rec.ls.sp = j.sp;
            rec.st.sweeps += j.nb;
            rec.chase_turnovers += (long long)j.nb * (7ll * (j.stop - j.start) + 1ll);
            if (j.mode == 2) { rec.st.exceptional += j.nb; rec.ls.ord += j.nb; rec.ls.ctr.it_count = AMRVW_IT_COUNT - 1; }
            store_rec(r, rec);
            __threadfence();                  // the chase warps' write-backs of this job are visible before the polynomial is
            mpmc_push(sq.work, sq.work_slots, sq.mask, j.p);
            atomicAdd(sq.progress, 1u);
            mgr_done += 1;
            worked = true;
        }
        const int tail = sm.jobs_tail;
        if (!sm.no_more && tail - sm.jobs_head == 0 && sm.feed_left < 3 * LE && tail - mgr_done < AMRVW_STREAM_JOBS - 1) {
            const int p = mpmc_try_pop(sq.ready, sq.ready_slots, sq.mask);
            if (p >= 0) {
                const PolyRec* r = pr.rec + p;
                StreamJob j;
                j.p = p; j.base = pr.offsets[p] + (long long)pr.extra * p;
                j.mode = __ldcg(&r->mode); j.nb = __ldcg(&r->nb); j.start = __ldcg(&r->ls.ctr.start_index); j.stop = __ldcg(&r->ls.ctr.stop_index);
                j.ord0 = __ldcg(&r->ls.ord); j.sp = __ldcg(&r->ls.sp); j.pad = 0;
                sm.jobs[tail % AMRVW_STREAM_JOBS] = j;
                __threadfence_block();
                sm.jobs_tail = tail + 1;
                worked = true;
            } else if (ld_volatile_u32(sq.remaining) == 0u) {
                sm.no_more = 1;
            }
        }
        if (sm.quit) return;
        if (!worked) __nanosleep(200);
    }
}

// service warp: phase A of whichever polynomial needs it
template <int MH>
AMRVW_DI void stream_service(const Problem<double>& pr, const PipeParams& pp, const StreamQueues& sq, ShiftScratch<MH>& scratch, const int LE) {
    const int lane32 = threadIdx.x & 31;
    for (;;) {
        int p = -1;
        if (lane32 == 0) {
            for (unsigned spins = 0;; ++spins) {
                p = mpmc_try_pop(sq.work, sq.work_slots, sq.mask);
                if (p >= 0) break;
                if (ld_volatile_u32(sq.remaining) == 0u || ld_volatile_u32(sq.abort_flag) != 0u) { p = -2; break; }
                __nanosleep(spins < 64 ? 200 : 1000);
            }
        }
        p = __shfl_sync(0xffffffffu, p, 0);
        if (p < 0) return;
        __threadfence();
        int m0 = 0;
        if (lane32 == 0) m0 = unit_driver(pr, pp, LE, p);
        m0 = __shfl_sync(0xffffffffu, m0, 0);
        if (m0 == 3) unit_shifts<MH>(pr, pp, 32, p, scratch, lane32, 0xffffffffu, LE);
        __syncwarp();
        if (lane32 == 0) {
            __threadfence();
            if (m0 == 0) atomicSub(sq.remaining, 1u);
            else mpmc_push(sq.ready, sq.ready_slots, sq.mask, p);
        }
    }
}

template <int NW, int MH>
__global__ void __launch_bounds__(32 * (NW + 2), 2) stream_kernel(Problem<double> pr, PipeParams pp, StreamQueues sq) {
    constexpr int LE = 32 * NW;
    [[maybe_unused]] constexpr int NT = LE;
    extern __shared__ __align__(16) unsigned char raw[];       // [12][NT] step snapshots, then the service warp's shift scratch
    __shared__ __align__(16) StreamSmem<NW> sm;
    const int gl = threadIdx.x, warp = gl >> 5, lane32 = gl & 31;
    if (gl == 0) {
        sm.jobs_tail = 0; sm.jobs_head = 0; sm.jobs_done = 0; sm.feed_left = 0; sm.no_more = 0; sm.quit = 0; sm.idle = 1;
    }
    if (gl < NW * 2 * 3) (&sm.hand[0][0][0])[gl] = make_double2(1.0, 0.0);
    if (gl < NW * 2) (&sm.handtag[0][0])[gl] = 0u;
    if (gl < AMRVW_RING) sm.ringtag[gl] = 0u;
    __syncthreads();
    if (warp == NW) { stream_manager<NW>(pr, sq, sm, LE); return; }
    if (warp == NW + 1) {
        ShiftScratch<MH>& scratch = *reinterpret_cast<ShiftScratch<MH>*>(raw + 12 * NT * sizeof(double2));
        stream_service<MH>(pr, pp, sq, scratch, LE);
        return;
    }

    // ---------------- chase warps
    [[maybe_unused]] double2* const snap_t = reinterpret_cast<double2*>(raw) + gl;
    [[maybe_unused]] const unsigned snap_sa = (unsigned)__cvta_generic_to_shared(snap_t);
    const unsigned stage_sa = (unsigned)__cvta_generic_to_shared(sm.ring);
    const unsigned hand_sa = (unsigned)__cvta_generic_to_shared(&sm.hand[0][0][0]);
    const bool feeds_g = gl == 0, feeds_h = warp > 0 && lane32 == 0;
    const bool ret_g = gl == LE - 1, ret_h = warp < NW - 1 && lane32 == 31;
    const unsigned hin = hand_sa + (unsigned)((warp > 0 ? warp - 1 : 0) * 2) * 48u, hout = hand_sa + (unsigned)(warp * 2) * 48u;

    Rot<double> q0, q1, q2, b0, b1, b2, c0, c1, c2, U, V, W;
    q0 = q1 = q2 = b0 = b1 = b2 = c0 = c1 = c2 = U = V = W = make_rot<double>(1.0, 0.0);
    unsigned tag0 = 0u, tag1 = 0u, tag2 = 0u;
    bool active = false, pend_b = false;

    // feeding thread: the job being streamed in
    int f_slot = -1, f_k = 0, f_stop1 = -1, f_start = 0;       // next index to request, last index (stop + 1) of the job
    const Rot<double>* f_q = nullptr; const Rot<double>* f_b = nullptr; const Rot<double>* f_c = nullptr;
    int f_started = 0;                                         // jobs started
    auto request = [&](const int tslot) {
        // the triple that enters the array AMRVW_RING - 1 steps from now: next index of the current job, else the first
        // index of the next prepared job, else nothing
        unsigned tag = 0u;
        if (f_slot >= 0 && f_k > f_stop1) f_slot = -1;
        if (f_slot < 0 && f_started < sm.jobs_tail) {
            __threadfence_block();
            const StreamJob& j = sm.jobs[f_started % AMRVW_STREAM_JOBS];
            f_slot = f_started % AMRVW_STREAM_JOBS; f_start = j.start; f_k = j.start; f_stop1 = j.stop + 1;
            f_q = pr.Q + j.base; f_b = pr.B + j.base; f_c = pr.Ct + j.base;
            f_started += 1;
            sm.jobs_head = f_started;
        }
        if (f_slot >= 0) {
            const unsigned sg = stage_sa + (unsigned)tslot * 48u;
            cp_async16_sa(sg, f_q + f_k); cp_async16_sa(sg + 16, f_b + f_k); cp_async16_sa(sg + 32, f_c + f_k);
            tag = AMRVW_TAG_VALID | ((unsigned)f_slot << AMRVW_TAG_SLOT_SHIFT) | (unsigned)f_k | (f_k == f_start ? AMRVW_TAG_FIRST : 0u) | (f_k == f_stop1 ? AMRVW_TAG_LAST : 0u);
            f_k += 1;
            sm.feed_left = f_stop1 - f_k + 1;
        } else {
            sm.feed_left = 0;
        }
        cp_async_commit();
        sm.ringtag[tslot] = tag;
    };
    if (feeds_g) for (int a = 0; a < AMRVW_RING - 1; ++a) request(a);

    int r_done = 0;      // retiring thread: jobs completed
    for (unsigned t = 0;; ++t) {
        // ---------------- ingest: what the previous lane finished last step
        Rot<double> iq = shfl_up_rot(q0, 32), ib = shfl_up_rot(b0, 32), ic = shfl_up_rot(c0, 32);
        unsigned itag = __shfl_up_sync(0xffffffffu, tag0, 1);
        if (feeds_g) {
            cp_async_wait_ring();
            const int rs = (int)(t & (AMRVW_RING - 1));
            const unsigned sg = stage_sa + (unsigned)rs * 48u;
            iq = lds_rot(sg); ib = lds_rot(sg + 16); ic = lds_rot(sg + 32);
            itag = sm.ringtag[rs];
            request((int)((t + AMRVW_RING - 1) & (AMRVW_RING - 1)));
        } else if (feeds_h) {
            const unsigned par = (t + 1u) & 1u;
            const unsigned sg = hin + par * 48u;
            iq = lds_rot(sg); ib = lds_rot(sg + 16); ic = lds_rot(sg + 32);
            itag = sm.handtag[warp - 1][par];
        }
        q0 = q1; q1 = q2; q2 = iq;
        b0 = b1; b1 = b2; b2 = ib;
        c0 = c1; c1 = c2; c2 = ic;
        tag0 = tag1; tag1 = tag2; tag2 = itag;

        // ---------------- what this lane does in this lock-step (the lanes of a warp that differ take their branches in turn)
        const bool cr = (tag1 & AMRVW_TAG_FIRST) != 0u;
        const bool fa = active && (tag2 & AMRVW_TAG_LAST) != 0u;
        const bool reg = active && !fa;
        if (cr) {              // window [*, start, start+1]: create the bulge one step before its first chase position
            const StreamJob& j = sm.jobs[(tag1 >> AMRVW_TAG_SLOT_SHIFT) & (AMRVW_STREAM_JOBS - 1)];
            if (gl < j.nb) {
                const Created o = stream_create(q1, q2, b1, b2, c1, c2, pr.Q + j.base, pr.B + j.base, pr.Ct + j.base, 0, j.start, j.stop, j.mode,
                                                pr.shifts + (long long)j.p * LE, gl, pr.seed, (uint64_t)j.p, j.ord0 + gl);
                q1 = o.qa; q2 = o.qb; U = o.U; V = o.V; W = o.W;
                active = true;
            }
        } else if (fa) {       // window [stop-1, stop, stop+1]: passthrough_triu, V into Q, U through Q, U <- W U  (RDS.jl:110-128, first part)
            turnover_fast(b1, b2, V, V, b1, b2);
            turnover_fast(c2, c1, V, V, c2, c1);
            turnover_fast(b0, b1, U, U, b0, b1);
            turnover_fast(c1, c0, U, U, c1, c0);
            V.s = sign_(sign_(q2.c)) * V.s;          // Q[stop+1]: the diagonal rotator below the window (or the (1,0) sentinel)
            q1 = fuse(q1, V);
            turnover_fast(q0, q1, U, U, q0, q1);
            U = fuse(W, U);
            active = false; pend_b = true;
        } else if (pend_b) {   // window [stop, stop+1, *]: U through R, then into Q (second part)
            turnover_fast(b0, b1, U, U, b0, b1);
            turnover_fast(c1, c0, U, U, c1, c0);
            U.s = sign_(sign_(q1.c)) * U.s;
            q0 = fuse(q0, U);
            pend_b = false;
        }
        if (reg) {             // regular position: the 7 turnovers of one chase step as one straight-line block
#ifndef AMRVW_STRICT
            AMRVW_CHASE_REGULAR(snap_t, NT);
#else
            AMRVW_CHASE_GENERAL();
#endif
        }

        // ---------------- retire: the last lane of the chain writes back, the last lane of a warp hands on
        if (ret_g) {
            if (tag0 & AMRVW_TAG_VALID) {
                StreamJob& j = sm.jobs[(tag0 >> AMRVW_TAG_SLOT_SHIFT) & (AMRVW_STREAM_JOBS - 1)];
                const int k = (int)(tag0 & AMRVW_TAG_INDEX_MASK);
                const long long at = j.base + k;
                stg_rot(pr.Q + at, q0); stg_rot(pr.B + at, b0); stg_rot(pr.Ct + at, c0);
                if (!(tag0 & AMRVW_TAG_LAST)) {
                    if (fabs(q0.s) <= AMRVW_EPS) { pr.dstack[j.base + j.sp] = k; j.sp += 1; }     // deflatable: record for phase A
                } else {                      // the job's last triple: its write-backs are visible before the manager hands the polynomial on
                    __threadfence();
                    r_done += 1;
                    sm.jobs_done = r_done;
                }
            }
        } else if (ret_h) {
            const unsigned par = t & 1u;
            const unsigned so = hout + par * 48u;
            sts_rot(so, q0); sts_rot(so + 16, b0); sts_rot(so + 32, c0);
            sm.handtag[warp][par] = tag0;
        }
        if (feeds_g) {
            // nothing in the array and nothing to come: leave; nothing in the array right now: do not spin at full speed
            const bool empty = f_slot < 0 && sm.jobs_done == f_started;
            if ((empty && sm.no_more && f_started == sm.jobs_tail) || sm.no_more == 2) sm.quit = 1;     // 2: aborted by the watchdog
            sm.idle = empty ? 1 : 0;
        }
        AMRVW_BAR_CHASE(32 * NW);
        if (sm.quit) break;
        if (sm.idle) __nanosleep(400);
    }
    if (feeds_g) cp_async_wait_all();
}

template <int NW, int MH> constexpr size_t stream_raw_bytes() { return 12 * 32 * NW * sizeof(double2) + sizeof(ShiftScratch<MH>); }

struct StreamBufs { MpmcQueue* ready; MpmcSlot* ready_slots; MpmcQueue* work; MpmcSlot* work_slots; unsigned mask; unsigned* remaining; };   // remaining[1] = progress

template <int NW, int MH>
static cudaError_t launch_stream(const Problem<double>& pr, const PipeParams& pp, const StreamBufs& sb, int sm_count, cudaStream_t st) {
    StreamQueues sq;
    sq.ready = sb.ready; sq.ready_slots = sb.ready_slots; sq.work = sb.work; sq.work_slots = sb.work_slots; sq.mask = sb.mask; sq.remaining = sb.remaining;
    sq.progress = sb.remaining + 1; sq.abort_flag = pr.counters + CNT_ABORT;
    stream_init_queues_kernel<<<(int)((sb.mask + 1 + 255) / 256), 256, 0, st>>>(sq);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    stream_enqueue_kernel<<<(int)((pr.nbatch + 127) / 128), 128, 0, st>>>(pr, sq);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    constexpr size_t smem = stream_raw_bytes<NW, MH>();
    int per_sm = 0;
    if ((e = cudaFuncSetAttribute(stream_kernel<NW, MH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stream_kernel<NW, MH>, 32 * (NW + 2), smem)) != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;
    // every CTA must be resident (they exchange polynomials through the queues); more CTAs than polynomials would only poll
    const long long cap = (long long)sm_count * per_sm;
    const int blocks = (int)(pr.nbatch < cap ? pr.nbatch : cap);
    stream_kernel<NW, MH><<<blocks, 32 * (NW + 2), smem, st>>>(pr, pp, sq);
    return cudaGetLastError();
}


// ================================================================ sub-warp pipelines (L <= 32 lanes per polynomial)
// The same stream for large batches (config 2 on one GPU: 3000 polynomials, 16 bulges each): every group of L lanes
// of a chase warp is an independent systolic array with its own stream of jobs -- no hand-off between warps, no CTA
// barrier, each group steps with its warp.  What it removes from the unit schedule of rounds.cuh: the 3 L fill/drain
// lock-steps of every fat step (lanes idle 9% of the lock-steps at 1.9x the cost of a lean one), the waiting of a
// unit's shorter window for its longer one, parking (the bulges in flight never leave their registers), and phase A
// inside the chase warps (service warps compute it while other jobs stream).  One manager warp per CTA, one lane
// per pipeline.
#define AMRVW_USTREAM_JOBS 8
struct UPipe {
    StreamJob jobs[AMRVW_USTREAM_JOBS];
    double2 ring[3 * AMRVW_RING];
    unsigned ringtag[AMRVW_RING];
    volatile int jobs_tail, jobs_head, jobs_done, feed_left, no_more, quit;
    int pad[2];
};

// manager warp: lane i serves pipeline i of the CTA
template <int NP>
AMRVW_DI void ustream_manager(const Problem<double>& pr, const StreamQueues& sq, UPipe* pipes, const int L) {
    const int lane32 = threadIdx.x & 31;
    if (lane32 >= NP) return;
    UPipe& sm = pipes[lane32];
    int mgr_done = 0;
    unsigned long long wd_time = global_ns();
    unsigned wd_progress = ld_volatile_u32(sq.progress), wd_polls = 0;
    for (;;) {
        bool worked = false;
        if ((++wd_polls & 1023u) == 0u) {
            const unsigned pg = ld_volatile_u32(sq.progress);
            const unsigned long long now = global_ns();
            if (pg != wd_progress) { wd_progress = pg; wd_time = now; }
            else if (now - wd_time > 20000000000ull && ld_volatile_u32(sq.remaining) != 0u) atomicExch(sq.abort_flag, 1u);
            if (ld_volatile_u32(sq.abort_flag) != 0u) sm.no_more = 2;
        }
        while (mgr_done < sm.jobs_done) {
            __threadfence_block();
            const StreamJob j = sm.jobs[mgr_done % AMRVW_USTREAM_JOBS];
            PolyRec* r = pr.rec + j.p;
            PolyRec rec = load_rec(r);
            rec.ls.sp = j.sp;
            rec.st.sweeps += j.nb;
            rec.chase_turnovers += (long long)j.nb * (7ll * (j.stop - j.start) + 1ll);
            if (j.mode == 2) { rec.st.exceptional += j.nb; rec.ls.ord += j.nb; rec.ls.ctr.it_count = AMRVW_IT_COUNT - 1; }
            store_rec(r, rec);
            __threadfence();
            mpmc_push(sq.work, sq.work_slots, sq.mask, j.p);
            atomicAdd(sq.progress, 1u);
            mgr_done += 1;
            worked = true;
        }
        const int tail = sm.jobs_tail;
        if (!sm.no_more && tail - sm.jobs_head == 0 && sm.feed_left < 6 * L + 64 && tail - mgr_done < AMRVW_USTREAM_JOBS - 1) {
            const int p = mpmc_try_pop(sq.ready, sq.ready_slots, sq.mask);
            if (p >= 0) {
                const PolyRec* r = pr.rec + p;
                StreamJob j;
                j.p = p; j.base = pr.offsets[p] + (long long)pr.extra * p;
                j.mode = __ldcg(&r->mode); j.nb = __ldcg(&r->nb); j.start = __ldcg(&r->ls.ctr.start_index); j.stop = __ldcg(&r->ls.ctr.stop_index);
                j.ord0 = __ldcg(&r->ls.ord); j.sp = __ldcg(&r->ls.sp); j.pad = 0;
                sm.jobs[tail % AMRVW_USTREAM_JOBS] = j;
                __threadfence_block();
                sm.jobs_tail = tail + 1;
                worked = true;
            } else if (ld_volatile_u32(sq.remaining) == 0u) {
                sm.no_more = 1;
            }
        }
        if (sm.quit) return;
        if (!worked) __nanosleep(300);
    }
}

// service warp of the sub-warp stream: the lane groups of the warp each take a polynomial (the unit kernel's
// side-by-side phase A, rounds.cuh)
template <int L, int MH>
AMRVW_DI void ustream_service(const Problem<double>& pr, const PipeParams& pp, const StreamQueues& sq, ShiftScratch<MH>* scratch) {
    constexpr int G = 32 / L;
    const int lane32 = threadIdx.x & 31, grp = lane32 / L, lane = lane32 % L;
    const unsigned gmask = (L == 32 ? 0xffffffffu : ((1u << L) - 1u)) << (grp * L);
    for (;;) {
        // up to G polynomials, one per group; a group without one idles through this round
        int p = -1;
        if (lane == 0) p = mpmc_try_pop(sq.work, sq.work_slots, sq.mask);
        if (!__any_sync(0xffffffffu, p >= 0)) {
            if (ld_volatile_u32(sq.remaining) == 0u || ld_volatile_u32(sq.abort_flag) != 0u) return;
            __nanosleep(500);
            continue;
        }
        __threadfence();
        int m0 = 0;
        if (lane == 0 && p >= 0) m0 = unit_driver(pr, pp, L, p);
        __syncwarp();
        const int mg = __shfl_sync(0xffffffffu, m0, grp * L);
        const int pg = __shfl_sync(0xffffffffu, p, grp * L);
        if (mg == 3) unit_shifts<MH>(pr, pp, L, pg, scratch[grp], lane, gmask);
        __syncwarp();
        if (lane == 0 && p >= 0) {
            __threadfence();
            if (m0 == 0) atomicSub(sq.remaining, 1u);
            else mpmc_push(sq.ready, sq.ready_slots, sq.mask, p);
        }
        (void)G;
    }
}

// NW chase warps (G = 32/L pipelines each), 1 manager warp, NS service warps per CTA
template <int L, int MH, int NW, int NS>
__global__ void __launch_bounds__(32 * (NW + 1 + NS), 2) ustream_kernel(Problem<double> pr, PipeParams pp, StreamQueues sq) {
    constexpr int G = 32 / L, NP = NW * G, NT = 32 * NW;
    static_assert(NP <= 32, "one manager lane per pipeline");
    extern __shared__ __align__(16) unsigned char raw[];       // [12][NT] step snapshots, then the service warps' shift scratch
    __shared__ __align__(16) UPipe pipes[NP];
    const int gl = threadIdx.x, warp = gl >> 5, lane32 = gl & 31;
    if (gl < NP) {
        UPipe& z = pipes[gl];
        z.jobs_tail = 0; z.jobs_head = 0; z.jobs_done = 0; z.feed_left = 0; z.no_more = 0; z.quit = 0;
        for (int a = 0; a < AMRVW_RING; ++a) z.ringtag[a] = 0u;
    }
    __syncthreads();
    if (warp == NW) { ustream_manager<NP>(pr, sq, pipes, L); return; }
    if (warp > NW) {
        ShiftScratch<MH>* scratch = reinterpret_cast<ShiftScratch<MH>*>(raw + 12 * NT * sizeof(double2)) + (warp - NW - 1) * G;
        ustream_service<L, MH>(pr, pp, sq, scratch);
        return;
    }

    // ---------------- chase warps
    const int grp = lane32 / L, lane = lane32 % L;
    UPipe& sm = pipes[warp * G + grp];
    [[maybe_unused]] double2* const snap_t = reinterpret_cast<double2*>(raw) + gl;
    const unsigned stage_sa = (unsigned)__cvta_generic_to_shared(sm.ring);
    const bool feeds = lane == 0, retires = lane == L - 1;

    Rot<double> q0, q1, q2, b0, b1, b2, c0, c1, c2, U, V, W;
    q0 = q1 = q2 = b0 = b1 = b2 = c0 = c1 = c2 = U = V = W = make_rot<double>(1.0, 0.0);
    unsigned tag0 = 0u, tag1 = 0u, tag2 = 0u;
    bool active = false, pend_b = false;

    int f_slot = -1, f_k = 0, f_stop1 = -1, f_start = 0;
    const Rot<double>* f_q = nullptr; const Rot<double>* f_b = nullptr; const Rot<double>* f_c = nullptr;
    int f_started = 0;
    auto request = [&](const int tslot) {
        unsigned tag = 0u;
        if (f_slot >= 0 && f_k > f_stop1) f_slot = -1;
        if (f_slot < 0 && f_started < sm.jobs_tail) {
            __threadfence_block();
            const StreamJob& j = sm.jobs[f_started % AMRVW_USTREAM_JOBS];
            f_slot = f_started % AMRVW_USTREAM_JOBS; f_start = j.start; f_k = j.start; f_stop1 = j.stop + 1;
            f_q = pr.Q + j.base; f_b = pr.B + j.base; f_c = pr.Ct + j.base;
            f_started += 1;
            sm.jobs_head = f_started;
        }
        if (f_slot >= 0) {
            const unsigned sg = stage_sa + (unsigned)tslot * 48u;
            cp_async16_sa(sg, f_q + f_k); cp_async16_sa(sg + 16, f_b + f_k); cp_async16_sa(sg + 32, f_c + f_k);
            tag = AMRVW_TAG_VALID | ((unsigned)f_slot << AMRVW_TAG_SLOT_SHIFT) | (unsigned)f_k | (f_k == f_start ? AMRVW_TAG_FIRST : 0u) | (f_k == f_stop1 ? AMRVW_TAG_LAST : 0u);
            f_k += 1;
            sm.feed_left = f_stop1 - f_k + 1;
        } else {
            sm.feed_left = 0;
        }
        cp_async_commit();
        sm.ringtag[tslot] = tag;
    };
    if (feeds) for (int a = 0; a < AMRVW_RING - 1; ++a) request(a);

    int r_done = 0;
    bool my_quit = false;      // this lane's pipeline is through (feeding lane decides, the group follows)
    for (unsigned t = 0;; ++t) {
        Rot<double> iq = shfl_up_rot(q0, L), ib = shfl_up_rot(b0, L), ic = shfl_up_rot(c0, L);
        unsigned itag = __shfl_up_sync(0xffffffffu, tag0, 1, L);
        if (feeds) {
            cp_async_wait_ring();
            const int rs = (int)(t & (AMRVW_RING - 1));
            const unsigned sg = stage_sa + (unsigned)rs * 48u;
            iq = lds_rot(sg); ib = lds_rot(sg + 16); ic = lds_rot(sg + 32);
            itag = sm.ringtag[rs];
            request((int)((t + AMRVW_RING - 1) & (AMRVW_RING - 1)));
        }
        q0 = q1; q1 = q2; q2 = iq;
        b0 = b1; b1 = b2; b2 = ib;
        c0 = c1; c1 = c2; c2 = ic;
        tag0 = tag1; tag1 = tag2; tag2 = itag;

        const bool cr = (tag1 & AMRVW_TAG_FIRST) != 0u;
        const bool fa = active && (tag2 & AMRVW_TAG_LAST) != 0u;
        const bool reg = active && !fa;
        if (cr) {
            const StreamJob& j = sm.jobs[(tag1 >> AMRVW_TAG_SLOT_SHIFT) & (AMRVW_USTREAM_JOBS - 1)];
            if (lane < j.nb) {
                const Created o = stream_create(q1, q2, b1, b2, c1, c2, pr.Q + j.base, pr.B + j.base, pr.Ct + j.base, 0, j.start, j.stop, j.mode,
                                                pr.shifts + (long long)j.p * L, lane, pr.seed, (uint64_t)j.p, j.ord0 + lane);
                q1 = o.qa; q2 = o.qb; U = o.U; V = o.V; W = o.W;
                active = true;
            }
        } else if (fa) {
            turnover_fast(b1, b2, V, V, b1, b2);
            turnover_fast(c2, c1, V, V, c2, c1);
            turnover_fast(b0, b1, U, U, b0, b1);
            turnover_fast(c1, c0, U, U, c1, c0);
            V.s = sign_(sign_(q2.c)) * V.s;
            q1 = fuse(q1, V);
            turnover_fast(q0, q1, U, U, q0, q1);
            U = fuse(W, U);
            active = false; pend_b = true;
        } else if (pend_b) {
            turnover_fast(b0, b1, U, U, b0, b1);
            turnover_fast(c1, c0, U, U, c1, c0);
            U.s = sign_(sign_(q1.c)) * U.s;
            q0 = fuse(q0, U);
            pend_b = false;
        }
        if (reg) {
#ifndef AMRVW_STRICT
            AMRVW_CHASE_REGULAR(snap_t, NT);
#else
            AMRVW_CHASE_GENERAL();
#endif
        }

        if (retires && (tag0 & AMRVW_TAG_VALID)) {
            StreamJob& j = sm.jobs[(tag0 >> AMRVW_TAG_SLOT_SHIFT) & (AMRVW_USTREAM_JOBS - 1)];
            const int k = (int)(tag0 & AMRVW_TAG_INDEX_MASK);
            const long long at = j.base + k;
            stg_rot(pr.Q + at, q0); stg_rot(pr.B + at, b0); stg_rot(pr.Ct + at, c0);
            if (!(tag0 & AMRVW_TAG_LAST)) {
                if (fabs(q0.s) <= AMRVW_EPS) { pr.dstack[j.base + j.sp] = k; j.sp += 1; }
            } else {
                __threadfence();
                r_done += 1;
                sm.jobs_done = r_done;
            }
        }
        bool empty = false;
        if (feeds) {
            empty = f_slot < 0 && sm.jobs_done == f_started;
            if (!my_quit && ((empty && sm.no_more && f_started == sm.jobs_tail) || sm.no_more == 2)) { my_quit = true; sm.quit = 1; }
        }
        __syncwarp();          // the retiring lane's writes to the job ring before the next step's reads
        // leave when every pipeline of the warp is through; sleep when none has anything in flight
        const unsigned feeders = __ballot_sync(0xffffffffu, feeds);
        if ((__ballot_sync(0xffffffffu, feeds && my_quit) & feeders) == feeders) break;
        if ((__ballot_sync(0xffffffffu, feeds && empty) & feeders) == feeders) __nanosleep(400);
    }
    if (feeds) cp_async_wait_all();
}

template <int L, int MH, int NW, int NS> constexpr size_t ustream_raw_bytes() { return 12 * 32 * NW * sizeof(double2) + (size_t)NS * (32 / L) * sizeof(ShiftScratch<MH>); }

template <int L, int MH, int NW, int NS>
static cudaError_t launch_ustream(const Problem<double>& pr, const PipeParams& pp, const StreamBufs& sb, int sm_count, cudaStream_t st) {
    StreamQueues sq;
    sq.ready = sb.ready; sq.ready_slots = sb.ready_slots; sq.work = sb.work; sq.work_slots = sb.work_slots; sq.mask = sb.mask; sq.remaining = sb.remaining;
    sq.progress = sb.remaining + 1; sq.abort_flag = pr.counters + CNT_ABORT;
    stream_init_queues_kernel<<<(int)((sb.mask + 1 + 255) / 256), 256, 0, st>>>(sq);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    stream_enqueue_kernel<<<(int)((pr.nbatch + 127) / 128), 128, 0, st>>>(pr, sq);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    constexpr size_t smem = ustream_raw_bytes<L, MH, NW, NS>();
    int per_sm = 0;
    if ((e = cudaFuncSetAttribute(ustream_kernel<L, MH, NW, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ustream_kernel<L, MH, NW, NS>, 32 * (NW + 1 + NS), smem)) != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;
    constexpr int NP = NW * (32 / L);
    const long long cap = (long long)sm_count * per_sm;
    const long long want = (pr.nbatch + NP - 1) / NP;       // more pipelines than polynomials would only poll
    const int blocks = (int)(want < cap ? want : cap);
    ustream_kernel<L, MH, NW, NS><<<blocks, 32 * (NW + 1 + NS), smem, st>>>(pr, pp, sq);
    return cudaGetLastError();
}

}  // namespace amrvw
