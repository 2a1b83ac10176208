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
    unsigned long long* t0;                       // globaltimer at queue initialisation (accounting only)
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
        *sq.remaining = 0u; *sq.progress = 0u; *sq.t0 = global_ns();
    }
}
// every iterating polynomial starts in the phase A queue, largest first (pr.order)
__global__ void stream_enqueue_kernel(Problem<double> pr, StreamQueues sq) {
    const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= pr.nbatch) return;
    const long long p = unit_poly(pr, slot);
    if (!pr.rec[p].iterating) return;
    atomicAdd(sq.remaining, 1u);
    pr.rec[p].ls.pc_sub = (long long)global_ns();      // time stamps of the queue hand-overs (accounting only)
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
    double2 qm, bm, cm;   // rotators of index start - 1: constant during the job, needed by every bulge creation
};
// descriptor of the next fat step of polynomial p from its record (manager lanes)
AMRVW_DI StreamJob stream_job_from_record(const Problem<double>& pr, const int p) {
    const PolyRec* r = pr.rec + p;
    StreamJob j;
    j.p = p; j.base = pr.offsets[p] + (long long)pr.extra * p;
    j.mode = __ldcg(&r->mode); j.nb = __ldcg(&r->nb); j.start = __ldcg(&r->ls.ctr.start_index); j.stop = __ldcg(&r->ls.ctr.stop_index);
    j.ord0 = __ldcg(&r->ls.ord); j.sp = __ldcg(&r->ls.sp); j.pad = 0;
    j.qm = __ldcg(reinterpret_cast<const double2*>(pr.Q + j.base + j.start - 1));       // Q[0] is the (1,0) sentinel
    j.bm = make_double2(1.0, 0.0); j.cm = make_double2(1.0, 0.0);
    if (j.start > 1) {
        j.bm = __ldcg(reinterpret_cast<const double2*>(pr.B + j.base + j.start - 1));
        j.cm = __ldcg(reinterpret_cast<const double2*>(pr.Ct + j.base + j.start - 1));
    }
    return j;
}
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
    volatile int ctl;                      // feeding thread, every lock-step: > 0 lock-steps the array may run lean, 0 go on, -1 nap, -2 leave
    volatile int quit;                     // chase warps have left (manager's exit)
    volatile int sptype;                   // development: special operation of this lock-step (1 create, 2 last step part one, 3 part two)
};

struct Created { Rot<double> qa, qb, U, V, W; };
// create_bulge + absorb_Ut (create-bulge.jl:20-77, RDS.jl:14-36) for one bulge of a job: the lane's window holds final
// values of indices start (qa, ba, ca) and start+1 (qb, bb, cb); index start-1 (qm, bm, cm: constant during the job)
// comes from the job descriptor, the lane's shift pair (e1, e2) was loaded one lock-step earlier.  Everything by
// value: no memory access on the path.
AMRVW_NI Created stream_create(const Rot<double> qa, const Rot<double> qb, const Rot<double> ba, const Rot<double> bb, const Rot<double> ca, const Rot<double> cb,
                               const double2 qm2, const double2 bm2, const double2 cm2, const double2 e1, const double2 e2,
                               const int start, const int stop, const int mode, const uint64_t seed, const uint64_t poly, const int ordinal) {
    const Rot<double> one = make_rot<double>(1.0, 0.0);
    const Rot<double> qm = make_rot<double>(qm2.x, qm2.y), bm = make_rot<double>(bm2.x, bm2.y), cm = make_rot<double>(cm2.x, cm2.y);
    Shifts sh;
    sh.e1 = cplx(e1.x, e1.y); sh.e2 = cplx(e2.x, e2.y);
    StepState z;
    z.q0 = qa; z.q1 = qb; z.q2 = one; z.b0 = ba; z.b1 = bb; z.b2 = one; z.c0 = ca; z.c1 = cb; z.c2 = one;   // index start+2 is not read
    z.U = one; z.V = one; z.W = one;
    Fact<double, false> F;
    F.Q = F.B = F.Ct = F.WB = F.WCt = nullptr; F.DQ = F.DR = F.WD = nullptr; F.N = 0; F.turnovers = 0;
    create_and_absorb(z, qm, bm, cm, F, start, stop, mode, sh, seed, poly, ordinal);
    Created o;
    o.qa = z.q0; o.qb = z.q1; o.U = z.U; o.V = z.V; o.W = z.W;
    return o;
}

// ---------------------------------------------------------------- one lock-step of a lane that holds a bulge
// A lane is at a regular position (reg: 7 turnovers, RDS.jl:42-70, 97-108), at the first half of its bulge's last position
// (fa: window [stop-1, stop, stop+1] -- the four passthrough_triu turnovers, V fused into Q, U through Q, U <- W U;
// RDS.jl:110-128) or at the second half (fb: window [stop, stop+1, *] -- U through R, then fused into Q).  The three
// share their operand patterns turnover by turnover, so in the product build they are ONE straight-line instruction
// stream with the results committed under per-lane predicates: the lanes of a warp that sit at the end of a window cost
// nothing extra (taken as divergent branches the same warp paid regular + last-a + last-b in turn, 1.5x - 1.9x a regular
// step while a window boundary crosses it).  A degenerate turnover (flagged, essentially never) redoes the lane's step on
// the general path from the shared-memory snapshot.  The strict build keeps the branches (it exists to be bit-identical
// to the oracle, which needs the IEEE turnover).
AMRVW_DI Rot<double> rsel(const bool p, const Rot<double> a, const Rot<double> b) { return make_rot<double>(p ? a.c : b.c, p ? a.s : b.s); }
AMRVW_NI void stream_last_a_general(StepState& z) {
    turnover_fast(z.b1, z.b2, z.V, z.V, z.b1, z.b2);
    turnover_fast(z.c2, z.c1, z.V, z.V, z.c2, z.c1);
    turnover_fast(z.b0, z.b1, z.U, z.U, z.b0, z.b1);
    turnover_fast(z.c1, z.c0, z.U, z.U, z.c1, z.c0);
    z.V.s = sign_(sign_(z.q2.c)) * z.V.s;          // Q[stop+1]: the diagonal rotator below the window (or the (1,0) sentinel)
    z.q1 = fuse(z.q1, z.V);
    turnover_fast(z.q0, z.q1, z.U, z.U, z.q0, z.q1);
    z.U = fuse(z.W, z.U);
}
AMRVW_NI void stream_last_b_general(StepState& z) {
    turnover_fast(z.b0, z.b1, z.U, z.U, z.b0, z.b1);
    turnover_fast(z.c1, z.c0, z.U, z.U, z.c1, z.c0);
    z.U.s = sign_(sign_(z.q1.c)) * z.U.s;
    z.q0 = fuse(z.q0, z.U);
}
#define AMRVW_STREAM_COMPUTE(sn, NT)                                                                                                  \
    do {                                                                                                                              \
        if (reg | fa | fb) {                                                                                                          \
            (sn)[0 * (NT)] = make_double2(q0.c, q0.s); (sn)[1 * (NT)] = make_double2(q1.c, q1.s); (sn)[2 * (NT)] = make_double2(q2.c, q2.s); \
            (sn)[3 * (NT)] = make_double2(b0.c, b0.s); (sn)[4 * (NT)] = make_double2(b1.c, b1.s); (sn)[5 * (NT)] = make_double2(b2.c, b2.s); \
            (sn)[6 * (NT)] = make_double2(c0.c, c0.s); (sn)[7 * (NT)] = make_double2(c1.c, c1.s); (sn)[8 * (NT)] = make_double2(c2.c, c2.s); \
            (sn)[9 * (NT)] = make_double2(U.c, U.s); (sn)[10 * (NT)] = make_double2(V.c, V.s); (sn)[11 * (NT)] = make_double2(W.c, W.s);    \
            const bool m12_ = reg | fa, m34_ = reg | fa | fb;                                                                         \
            bool ok_ = true, k_;                                                                                                      \
            Rot<double> t1_, t2_, t3_;                                                                                                \
            k_ = turnover_fast_try(b1, b2, V, t1_, t2_, t3_); ok_ &= k_ | !m12_; V = rsel(m12_, t1_, V); b1 = rsel(m12_, t2_, b1); b2 = rsel(m12_, t3_, b2); \
            k_ = turnover_fast_try(c2, c1, V, t1_, t2_, t3_); ok_ &= k_ | !m12_; V = rsel(m12_, t1_, V); c2 = rsel(m12_, t2_, c2); c1 = rsel(m12_, t3_, c1); \
            k_ = turnover_fast_try(b0, b1, U, t1_, t2_, t3_); ok_ &= k_ | !m34_; U = rsel(m34_, t1_, U); b0 = rsel(m34_, t2_, b0); b1 = rsel(m34_, t3_, b1); \
            k_ = turnover_fast_try(c1, c0, U, t1_, t2_, t3_); ok_ &= k_ | !m34_; U = rsel(m34_, t1_, U); c1 = rsel(m34_, t2_, c1); c0 = rsel(m34_, t3_, c0); \
            {   /* regular: V through Q; last-a: V (with the sign of the rotator below the window) fused into Q[stop] */              \
                Rot<double> vf_ = V; vf_.s = sign_(sign_(q2.c)) * V.s;                                                                \
                const Rot<double> qf_ = fuse(q1, vf_);                                                                                \
                k_ = turnover_fast_try(q1, q2, V, t1_, t2_, t3_); ok_ &= k_ | !reg;                                                   \
                V = rsel(reg, t1_, V); q1 = rsel(reg, t2_, rsel(fa, qf_, q1)); q2 = rsel(reg, t3_, q2);                               \
            }                                                                                                                         \
            {   /* regular, last-a: U through Q; last-b: U fused into Q[stop] */                                                      \
                Rot<double> uf_ = U; uf_.s = sign_(sign_(q1.c)) * U.s;                                                                \
                const Rot<double> qg_ = fuse(q0, uf_);                                                                                \
                k_ = turnover_fast_try(q0, q1, U, t1_, t2_, t3_); ok_ &= k_ | !m12_;                                                  \
                U = rsel(m12_, t1_, U); q0 = rsel(m12_, t2_, rsel(fb, qg_, q0)); q1 = rsel(m12_, t3_, q1);                            \
            }                                                                                                                         \
            {   /* regular: the limb turnover; last-a: U <- W U */                                                                    \
                const Rot<double> uw_ = fuse(W, U);                                                                                   \
                k_ = turnover_fast_try(W, V, U, t1_, t2_, t3_); ok_ &= k_ | !reg;                                                     \
                V = rsel(reg, t1_, V); U = rsel(reg, t2_, rsel(fa, uw_, U)); W = rsel(reg, t3_, W);                                   \
            }                                                                                                                         \
            if (!ok_) {                                                                                                               \
                StepState z_;                                                                                                         \
                double2 v_;                                                                                                           \
                v_ = (sn)[0 * (NT)]; z_.q0 = make_rot<double>(v_.x, v_.y); v_ = (sn)[1 * (NT)]; z_.q1 = make_rot<double>(v_.x, v_.y); \
                v_ = (sn)[2 * (NT)]; z_.q2 = make_rot<double>(v_.x, v_.y); v_ = (sn)[3 * (NT)]; z_.b0 = make_rot<double>(v_.x, v_.y); \
                v_ = (sn)[4 * (NT)]; z_.b1 = make_rot<double>(v_.x, v_.y); v_ = (sn)[5 * (NT)]; z_.b2 = make_rot<double>(v_.x, v_.y); \
                v_ = (sn)[6 * (NT)]; z_.c0 = make_rot<double>(v_.x, v_.y); v_ = (sn)[7 * (NT)]; z_.c1 = make_rot<double>(v_.x, v_.y); \
                v_ = (sn)[8 * (NT)]; z_.c2 = make_rot<double>(v_.x, v_.y); v_ = (sn)[9 * (NT)]; z_.U = make_rot<double>(v_.x, v_.y);  \
                v_ = (sn)[10 * (NT)]; z_.V = make_rot<double>(v_.x, v_.y); v_ = (sn)[11 * (NT)]; z_.W = make_rot<double>(v_.x, v_.y); \
                if (reg) chase_step_general(z_); else if (fa) stream_last_a_general(z_); else stream_last_b_general(z_);              \
                q0 = z_.q0; q1 = z_.q1; q2 = z_.q2; b0 = z_.b0; b1 = z_.b1; b2 = z_.b2; c0 = z_.c0; c1 = z_.c1; c2 = z_.c2;           \
                U = z_.U; V = z_.V; W = z_.W;                                                                                         \
            }                                                                                                                         \
        }                                                                                                                             \
    } while (0)

#define AMRVW_BAR_CHASE(n) asm volatile("bar.sync 1, %0;" ::"n"(n) : "memory")
// the same barrier with an AND reduction of a predicate over the N chase threads
template <int N> AMRVW_DI bool bar_chase_and(const bool pred) {
    int r;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.s32 q, %1, 0;\n\tbar.red.and.pred p, 1, %2, q;\n\tselp.s32 %0, 1, 0, p;\n\t}" : "=r"(r) : "r"((int)pred), "n"(N) : "memory");
    return r != 0;
}

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
        if ((++wd_polls & 127u) == 0u) {
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
            rec.ls.pc_sub = (long long)global_ns();
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
                const StreamJob j = stream_job_from_record(pr, p);
                sm.jobs[tail % AMRVW_STREAM_JOBS] = j;
                __threadfence_block();
                sm.jobs_tail = tail + 1;
                worked = true;
            } else if (ld_volatile_u32(sq.remaining) == 0u) {
                sm.no_more = 1;
            }
        }
        if (sm.quit) return;
        // nap between polls: a manager that spins takes issue slots from the chase warp that shares its scheduler (ncu, round 2:
        // the 296 managers of config 2's 375-share executed a third of the kernel's warp-instructions at 200 ns).  Its events are
        // a job completed and a job about to end, both hundreds of lock-steps (~microseconds each) apart
        if (!worked) __nanosleep(2000);
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
                __nanosleep(spins < 8 ? 500 : 2000);
            }
        }
        p = __shfl_sync(0xffffffffu, p, 0);
        if (p < 0) return;
        const long long pc_s0 = clock64();
        __threadfence();
        int m0 = 0;
        if (lane32 == 0) m0 = unit_driver(pr, pp, LE, p);
        m0 = __shfl_sync(0xffffffffu, m0, 0);
        if (m0 == 3) unit_shifts<MH>(pr, pp, 32, p, scratch, lane32, 0xffffffffu, LE);
        __syncwarp();
        if (lane32 == 0) {
            __stcg(&pr.rec[p].ls.pc_small, (long long)global_ns());
            __threadfence();
            if (m0 == 0) atomicSub(sq.remaining, 1u);
            else mpmc_push(sq.ready, sq.ready_slots, sq.mask, p);
            if (pr.prof) { atomicAdd(pr.prof + 1, (unsigned long long)(clock64() - pc_s0)); atomicAdd(pr.prof + 10, 1ull); }   // phase A: cycles, count

        }
    }
}

template <int NW, int MH>
__global__ void __launch_bounds__(32 * (NW + 2), NW <= 2 ? 3 : 2) stream_kernel(Problem<double> pr, PipeParams pp, StreamQueues sq) {
    constexpr int LE = 32 * NW;
    [[maybe_unused]] constexpr int NT = LE;
    extern __shared__ __align__(16) unsigned char raw[];       // [12][NT] step snapshots, then the service warp's shift scratch
    __shared__ __align__(16) StreamSmem<NW> sm;
    const int gl = threadIdx.x, warp = gl >> 5, lane32 = gl & 31;
    if (gl == 0) {
        sm.jobs_tail = 0; sm.jobs_head = 0; sm.jobs_done = 0; sm.feed_left = 0; sm.no_more = 0; sm.quit = 0; sm.ctl = -1; sm.sptype = 0;
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
    double2 she1 = make_double2(0.0, 0.0), she2 = make_double2(0.0, 0.0);      // shift pair of the bulge this lane creates next

    // feeding thread: the job being streamed in
    int f_slot = -1, f_k = 0, f_stop1 = -1, f_start = 0;       // next index to request, last index (stop + 1) of the job
    long long f_base = 0;                                      // chain slot of the job being fed
    int f_started = 0;                                         // jobs started
    auto request = [&](const int tslot) {
        // the triple that enters the array AMRVW_RING - 1 steps from now: next index of the current job, else the first
        // index of the next prepared job, else nothing
        unsigned tag = 0u;
        if (f_slot >= 0 && f_k > f_stop1) { f_slot = -1; sm.feed_left = 0; }
        if (f_slot < 0 && f_started < sm.jobs_tail) {
            __threadfence_block();
            const StreamJob& j = sm.jobs[f_started % AMRVW_STREAM_JOBS];
            f_slot = f_started % AMRVW_STREAM_JOBS; f_start = j.start; f_k = j.start; f_stop1 = j.stop + 1;
            f_base = j.base;
            f_started += 1;
            sm.jobs_head = f_started;
            sm.feed_left = f_stop1 - f_k + 1;
        }
        if (f_slot >= 0) {
            const unsigned sg = stage_sa + (unsigned)tslot * 48u;
            const long long at = f_base + f_k;
            cp_async16_sa(sg, pr.Q + at); cp_async16_sa(sg + 16, pr.B + at); cp_async16_sa(sg + 32, pr.Ct + at);
            tag = AMRVW_TAG_VALID | ((unsigned)f_slot << AMRVW_TAG_SLOT_SHIFT) | (unsigned)f_k | (f_k == f_start ? AMRVW_TAG_FIRST : 0u) | (f_k == f_stop1 ? AMRVW_TAG_LAST : 0u);
            f_k += 1;
            if (f_stop1 - f_k < 3 * LE) sm.feed_left = f_stop1 - f_k + 1;      // the manager only needs it near the end of the job
        }
        cp_async_commit();
        sm.ringtag[tslot] = tag;
    };
    if (feeds_g) for (int a = 0; a < AMRVW_RING - 1; ++a) request(a);

    int r_done = 0;      // retiring thread: jobs completed
    // accounting of the feeding thread (amrvw_last_profile): lock-steps, of which in the lean loop / with an empty array / with
    // bulges in flight but nothing to feed
    long long pn_steps = 0, pn_fast = 0, pn_empty = 0, pn_starved = 0, pc_fast = 0, pc_nap = 0;
#ifdef AMRVW_STREAM_DEV
    long long dv_c[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dv_n[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dv_last = clock64();      // development: lock-step time by special operation
#endif
    unsigned long long ns_first = 0ull, ns_last = 0ull;      // first / last lock-step with something in the array
    const long long pc_t0 = clock64();
#ifdef AMRVW_STREAM_DEV
    long long dp[4] = {0, 0, 0, 0}, dpn = 0;      // development: ingest / compute / retire+control / barrier wait of threads 0, 40, LE-1 in steps without a special
    const int dslot = gl == 0 ? 0 : (gl == 40 ? 1 : (gl == LE - 1 ? 2 : -1));
#endif
    for (unsigned t = 0;; ++t) {
#ifdef AMRVW_STREAM_DEV
        const long long dta = clock64();
#endif
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
        if (tag2 & AMRVW_TAG_FIRST) {      // a window enters: this lane creates its bulge NEXT step -- its shift pair is on its way meanwhile
            const StreamJob& j = sm.jobs[(tag2 >> AMRVW_TAG_SLOT_SHIFT) & (AMRVW_STREAM_JOBS - 1)];
            if (gl < j.nb && j.mode == 1) {
                const double2* sp2 = reinterpret_cast<const double2*>(pr.shifts + (long long)j.p * LE + gl);
                she1 = __ldcg(sp2); she2 = __ldcg(sp2 + 1);
            }
        }

        // ---------------- what this lane does in this lock-step (the lanes of a warp that differ take their branches in turn)
#ifdef AMRVW_STREAM_DEV
        const long long dtb = clock64();
#endif
        const bool cr = (tag1 & AMRVW_TAG_FIRST) != 0u;
        const bool fa = active && (tag2 & AMRVW_TAG_LAST) != 0u;
        const bool fb = pend_b;
        const bool reg = active && !fa;
#ifdef AMRVW_STREAM_DEV
        if (pr.prof && (fa || fb || (cr && active == false))) atomicOr((int*)&sm.sptype, (cr ? 1 : 0) | (fa ? 2 : 0) | (fb ? 4 : 0));
#endif
        if (cr) {              // window [*, start, start+1]: create the bulge one step before its first chase position
            const StreamJob& j = sm.jobs[(tag1 >> AMRVW_TAG_SLOT_SHIFT) & (AMRVW_STREAM_JOBS - 1)];
            if (gl < j.nb) {
                const Created o = stream_create(q1, q2, b1, b2, c1, c2, j.qm, j.bm, j.cm, she1, she2, j.start, j.stop, j.mode, pr.seed, (uint64_t)j.p, j.ord0 + gl);
                q1 = o.qa; q2 = o.qb; U = o.U; V = o.V; W = o.W;
                active = true;
            }
        }
#ifndef AMRVW_STRICT
        AMRVW_STREAM_COMPUTE(snap_t, NT);       // regular / last-a / last-b lanes in one predicated instruction stream
#else
        if (reg) {
            AMRVW_CHASE_GENERAL();
        } else if (fa) {
            StepState z_ = {q0, q1, q2, b0, b1, b2, c0, c1, c2, U, V, W};
            stream_last_a_general(z_);
            q0 = z_.q0; q1 = z_.q1; q2 = z_.q2; b0 = z_.b0; b1 = z_.b1; b2 = z_.b2; c0 = z_.c0; c1 = z_.c1; c2 = z_.c2; U = z_.U; V = z_.V; W = z_.W;
        } else if (fb) {
            StepState z_ = {q0, q1, q2, b0, b1, b2, c0, c1, c2, U, V, W};
            stream_last_b_general(z_);
            q0 = z_.q0; q1 = z_.q1; q2 = z_.q2; b0 = z_.b0; b1 = z_.b1; b2 = z_.b2; c0 = z_.c0; c1 = z_.c1; c2 = z_.c2; U = z_.U; V = z_.V; W = z_.W;
        }
#endif
        if (fa) { active = false; pend_b = true; }
        else if (fb) pend_b = false;

#ifdef AMRVW_STREAM_DEV
        const long long dtc = clock64();
#endif
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

        // ---------------- can the whole array run lean next?  Every lane chasing inside ONE window (no first / last triple of
        // a window anywhere in flight) and the rest of that window still to be requested: nothing but regular steps until then
        const bool clean = active && (tag0 & tag1 & tag2 & AMRVW_TAG_VALID) != 0u && ((tag0 | tag1 | tag2) & (AMRVW_TAG_FIRST | AMRVW_TAG_LAST)) == 0u;
        if (feeds_g) {
            int code = 0;      // > 0: lock-steps the array may run lean if every lane is clean; -1: nothing in the array (nap); -2: leave
            if (f_slot >= 0) {
                const int left = f_stop1 - f_k + 1;               // triples of the window not yet requested
                if (left >= 16 && f_k - f_start >= 4) code = left;
            } else {
                // nothing to feed.  Nothing in the array either: do not spin at full speed; and nothing to come: leave
                const bool empty = sm.jobs_done == f_started;
                if (empty) code = -1;
                if ((empty && sm.no_more && f_started == sm.jobs_tail) || sm.no_more == 2) code = -2;     // no_more == 2: aborted by the watchdog
                pn_empty += empty ? 1 : 0; pn_starved += empty ? 0 : 1;
            }
            sm.ctl = code;
            pn_steps += 1;
            if (pr.prof && code >= 0 && (t & 63u) == 0u) { ns_last = global_ns(); if (ns_first == 0ull) ns_first = ns_last; }
        }
#ifdef AMRVW_STREAM_DEV
        const long long dtd = clock64();
#endif
        const bool all_clean = bar_chase_and<32 * NW>(clean);      // the lock-step's barrier, with the AND of `clean` over the array
#ifdef AMRVW_STREAM_DEV
        if (pr.prof && dslot >= 0 && (sm.sptype & 7) == 0 && sm.ctl >= 0) {
            const long long dte = clock64();
            dp[0] += dtb - dta; dp[1] += dtc - dtb; dp[2] += dtd - dtc; dp[3] += dte - dtd; dpn += 1;
        }
        AMRVW_BAR_CHASE(32 * NW);      // (development build only: every thread has read sptype before the feeding thread clears it)
        if (pr.prof && feeds_g) {
            const long long now = clock64();
            const int ty = sm.sptype & 7;
            if (sm.ctl >= 0) { dv_c[ty] += now - dv_last; dv_n[ty] += 1; }
            dv_last = now; sm.sptype = 0;
        }
#endif
        const int code = sm.ctl;
        if (code == -2) break;
        if (code == -1) { const long long c0_ = clock64(); __nanosleep(400); pc_nap += clock64() - c0_; }
        const int fk = all_clean ? code : 0;
        if (fk > 0) {
            // ---------------- the lean loop of chain.cuh for fk lock-steps: same operations, no tags, no case analysis
            StreamJob& j = sm.jobs[(tag0 >> AMRVW_TAG_SLOT_SHIFT) & (AMRVW_STREAM_JOBS - 1)];
            const int idx_next = (int)(tag2 & AMRVW_TAG_INDEX_MASK) + 3 * gl + 1;     // index the feeding thread ingests next
            const int t0 = (int)t + 1, t1 = t0 + fk;
            int sp = ret_g ? j.sp : 0;
            const long long pc_f0 = clock64();
            StepState z = {q0, q1, q2, b0, b1, b2, c0, c1, c2, U, V, W};
            sp = chain_lean<NW, 1>(z, pr.Q + j.base, pr.B + j.base, pr.Ct + j.base, stage_sa, snap_sa, hand_sa, pr.dstack + j.base, sp, idx_next - t0, t0, t1, warp, lane32);
            q0 = z.q0; q1 = z.q1; q2 = z.q2; b0 = z.b0; b1 = z.b1; b2 = z.b2; c0 = z.c0; c1 = z.c1; c2 = z.c2; U = z.U; V = z.V; W = z.W;
            pc_fast += clock64() - pc_f0;
            tag0 += (unsigned)fk; tag1 += (unsigned)fk; tag2 += (unsigned)fk;
            if (ret_g) j.sp = sp;
            if (ret_h) sm.handtag[warp][(unsigned)(t1 - 1) & 1u] = tag0;
            if (feeds_g) {
                f_k += fk;                                     // the window's remaining triples have all been requested
                sm.feed_left = f_stop1 - f_k + 1;
                const unsigned base_tag = AMRVW_TAG_VALID | ((unsigned)f_slot << AMRVW_TAG_SLOT_SHIFT);
                for (int a = 0; a < AMRVW_RING - 1; ++a) {
                    const int idx = f_k - (AMRVW_RING - 1) + a;
                    sm.ringtag[(unsigned)(t1 + a) & (AMRVW_RING - 1)] = base_tag | (unsigned)idx | (idx == f_stop1 ? AMRVW_TAG_LAST : 0u);
                }
                pn_steps += fk; pn_fast += fk;
            }
            t = (unsigned)(t1 - 1);
            AMRVW_BAR_CHASE(32 * NW);      // the tags written above, before the next step reads them
#ifdef AMRVW_STREAM_DEV
            if (pr.prof && feeds_g) dv_last = clock64();
#endif
        }
    }
    if (feeds_g) { cp_async_wait_all(); sm.quit = 1; }
#ifdef AMRVW_STREAM_DEV
    if (pr.prof && dslot >= 0) {
        for (int a = 0; a < 4; ++a) atomicAdd(pr.prof + 32 + 5 * dslot + a, (unsigned long long)dp[a]);
        atomicAdd(pr.prof + 32 + 5 * dslot + 4, (unsigned long long)dpn);
    }
#endif
    if (pr.prof && feeds_g) {
        atomicAdd(pr.prof + 0, (unsigned long long)(clock64() - pc_t0));   // chase loop, feeding thread
        atomicAdd(pr.prof + 2, (unsigned long long)pn_fast);               // lock-steps in the lean loop
        atomicAdd(pr.prof + 3, (unsigned long long)pc_fast);               // cycles in the lean loop
        atomicAdd(pr.prof + 13, (unsigned long long)pc_nap);               // cycles napping with an empty array
        atomicAdd(pr.prof + 4, (unsigned long long)pn_steps);
        atomicAdd(pr.prof + 5, (unsigned long long)pn_empty);              // lock-steps with nothing in the array
        atomicAdd(pr.prof + 6, (unsigned long long)f_started);             // jobs
        atomicAdd(pr.prof + 7, (unsigned long long)pn_starved);            // lock-steps with bulges in flight but no job to feed
        atomicAdd(pr.prof + 14, ns_last ? (ns_last - *sq.t0) / 1000ull : 0ull);     // sum over CTAs of the time of their last busy step (us)
        atomicAdd(pr.prof + 15, ns_first ? (ns_first - *sq.t0) / 1000ull : 0ull);   // and of their first
#ifdef AMRVW_STREAM_DEV
        for (int a = 0; a < 8; ++a) { atomicAdd(pr.prof + 16 + a, (unsigned long long)dv_c[a]); atomicAdd(pr.prof + 24 + a, (unsigned long long)dv_n[a]); }
#endif
    }
}

template <int NW, int MH> constexpr size_t stream_raw_bytes() { return 12 * 32 * NW * sizeof(double2) + sizeof(ShiftScratch<MH>); }

struct StreamBufs { MpmcQueue* ready; MpmcSlot* ready_slots; MpmcQueue* work; MpmcSlot* work_slots; unsigned mask; unsigned* remaining; };   // remaining[1] = progress

template <int NW, int MH>
static cudaError_t launch_stream(const Problem<double>& pr, const PipeParams& pp, const StreamBufs& sb, int sm_count, cudaStream_t st) {
    StreamQueues sq;
    sq.ready = sb.ready; sq.ready_slots = sb.ready_slots; sq.work = sb.work; sq.work_slots = sb.work_slots; sq.mask = sb.mask; sq.remaining = sb.remaining;
    sq.progress = sb.remaining + 1; sq.abort_flag = pr.counters + CNT_ABORT; sq.t0 = reinterpret_cast<unsigned long long*>(sb.remaining + 2);
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
    if (per_sm > (NW <= 2 ? 3 : 2)) per_sm = NW <= 2 ? 3 : 2;
    if (pp.stream_ctas > 0 && pp.stream_ctas < per_sm) per_sm = pp.stream_ctas;
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
        if ((++wd_polls & 127u) == 0u) {
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
                const StreamJob j = stream_job_from_record(pr, p);
                sm.jobs[tail % AMRVW_USTREAM_JOBS] = j;
                __threadfence_block();
                sm.jobs_tail = tail + 1;
                worked = true;
            } else if (ld_volatile_u32(sq.remaining) == 0u) {
                sm.no_more = 1;
            }
        }
        if (sm.quit) return;
        if (!worked) __nanosleep(2000);      // see stream_manager
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
            __nanosleep(2000);
            continue;
        }
        const long long pc_s0 = clock64();
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
            if (pr.prof) { atomicAdd(pr.prof + 1, (unsigned long long)(clock64() - pc_s0)); atomicAdd(pr.prof + 10, 1ull); }
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
    double2 she1 = make_double2(0.0, 0.0), she2 = make_double2(0.0, 0.0);      // shift pair of the bulge this lane creates next

    int f_slot = -1, f_k = 0, f_stop1 = -1, f_start = 0;
    long long f_base = 0;                                      // chain slot of the job being fed
    int f_started = 0;
    auto request = [&](const int tslot) {
        unsigned tag = 0u;
        if (f_slot >= 0 && f_k > f_stop1) f_slot = -1;
        if (f_slot < 0 && f_started < sm.jobs_tail) {
            __threadfence_block();
            const StreamJob& j = sm.jobs[f_started % AMRVW_USTREAM_JOBS];
            f_slot = f_started % AMRVW_USTREAM_JOBS; f_start = j.start; f_k = j.start; f_stop1 = j.stop + 1;
            f_base = j.base;
            f_started += 1;
            sm.jobs_head = f_started;
        }
        if (f_slot >= 0) {
            const unsigned sg = stage_sa + (unsigned)tslot * 48u;
            const long long at = f_base + f_k;
            cp_async16_sa(sg, pr.Q + at); cp_async16_sa(sg + 16, pr.B + at); cp_async16_sa(sg + 32, pr.Ct + at);
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
    long long pn_steps = 0, pn_empty = 0, pn_starved = 0, pn_fast = 0;
    const long long pc_t0 = clock64();
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
        if (tag2 & AMRVW_TAG_FIRST) {      // a window enters: this lane creates its bulge NEXT step -- its shift pair is on its way meanwhile
            const StreamJob& j = sm.jobs[(tag2 >> AMRVW_TAG_SLOT_SHIFT) & (AMRVW_USTREAM_JOBS - 1)];
            if (lane < j.nb && j.mode == 1) {
                const double2* sp2 = reinterpret_cast<const double2*>(pr.shifts + (long long)j.p * L + lane);
                she1 = __ldcg(sp2); she2 = __ldcg(sp2 + 1);
            }
        }

        const bool cr = (tag1 & AMRVW_TAG_FIRST) != 0u;
        const bool fa = active && (tag2 & AMRVW_TAG_LAST) != 0u;
        const bool fb = pend_b;
        const bool reg = active && !fa;
        if (cr) {              // window [*, start, start+1]: create the bulge one step before its first chase position
            const StreamJob& j = sm.jobs[(tag1 >> AMRVW_TAG_SLOT_SHIFT) & (AMRVW_USTREAM_JOBS - 1)];
            if (lane < j.nb) {
                const Created o = stream_create(q1, q2, b1, b2, c1, c2, j.qm, j.bm, j.cm, she1, she2, j.start, j.stop, j.mode, pr.seed, (uint64_t)j.p, j.ord0 + lane);
                q1 = o.qa; q2 = o.qb; U = o.U; V = o.V; W = o.W;
                active = true;
            }
        }
#ifndef AMRVW_STRICT
        AMRVW_STREAM_COMPUTE(snap_t, NT);       // regular / last-a / last-b lanes in one predicated instruction stream
#else
        if (reg) {
            AMRVW_CHASE_GENERAL();
        } else if (fa) {
            StepState z_ = {q0, q1, q2, b0, b1, b2, c0, c1, c2, U, V, W};
            stream_last_a_general(z_);
            q0 = z_.q0; q1 = z_.q1; q2 = z_.q2; b0 = z_.b0; b1 = z_.b1; b2 = z_.b2; c0 = z_.c0; c1 = z_.c1; c2 = z_.c2; U = z_.U; V = z_.V; W = z_.W;
        } else if (fb) {
            StepState z_ = {q0, q1, q2, b0, b1, b2, c0, c1, c2, U, V, W};
            stream_last_b_general(z_);
            q0 = z_.q0; q1 = z_.q1; q2 = z_.q2; b0 = z_.b0; b1 = z_.b1; b2 = z_.b2; c0 = z_.c0; c1 = z_.c1; c2 = z_.c2; U = z_.U; V = z_.V; W = z_.W;
        }
#endif
        if (fa) { active = false; pend_b = true; }
        else if (fb) pend_b = false;

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
            if (!my_quit) { pn_steps += 1; pn_empty += empty ? 1 : 0; pn_starved += (!empty && f_slot < 0) ? 1 : 0; }
        }
        __syncwarp();          // the retiring lane's writes to the job ring before the next step's reads
        // leave when every pipeline of the warp is through; sleep when none has anything in flight
        const unsigned feeders = __ballot_sync(0xffffffffu, feeds);
        if ((__ballot_sync(0xffffffffu, feeds && my_quit) & feeders) == feeders) break;
        if ((__ballot_sync(0xffffffffu, feeds && empty) & feeders) == feeders) __nanosleep(400);

        // ---------------- every lane of the warp chasing inside its group's current window, and every group's window with
        // triples still to be requested: the lean loop of pipeline.cuh (no tags, no case analysis) until the first group
        // has requested its last triple
        const bool clean = active && (tag0 & tag1 & tag2 & AMRVW_TAG_VALID) != 0u && ((tag0 | tag1 | tag2) & (AMRVW_TAG_FIRST | AMRVW_TAG_LAST)) == 0u;
        int left = 0x7fffffff;
        if (feeds) left = (f_slot >= 0 && f_k - f_start >= 4) ? f_stop1 - f_k + 1 : 0;
        const int fk = __reduce_min_sync(0xffffffffu, left);
        if (fk >= 16 && __all_sync(0xffffffffu, clean)) {
            StreamJob& j = sm.jobs[(tag0 >> AMRVW_TAG_SLOT_SHIFT) & (AMRVW_USTREAM_JOBS - 1)];
            const int idx_next = (int)(tag2 & AMRVW_TAG_INDEX_MASK) + 3 * lane + 1;     // index the group's feeding lane ingests next
            const int t0 = (int)t + 1, t1 = t0 + fk;
            int sp = retires ? j.sp : 0;
            const int role = 1 | (feeds ? 2 : 0) | (retires ? 4 : 0);
            StepState z = {q0, q1, q2, b0, b1, b2, c0, c1, c2, U, V, W};
            sp = lean_chase<L, NT>(z, pr.Q + j.base, pr.B + j.base, pr.Ct + j.base, stage_sa, (unsigned)__cvta_generic_to_shared(snap_t), pr.dstack + j.base, sp,
                                   idx_next - t0, t0, t1, role);
            q0 = z.q0; q1 = z.q1; q2 = z.q2; b0 = z.b0; b1 = z.b1; b2 = z.b2; c0 = z.c0; c1 = z.c1; c2 = z.c2; U = z.U; V = z.V; W = z.W;
            tag0 += (unsigned)fk; tag1 += (unsigned)fk; tag2 += (unsigned)fk;
            if (retires) j.sp = sp;
            if (feeds) {
                f_k += fk;
                sm.feed_left = f_stop1 - f_k + 1;
                const unsigned base_tag = AMRVW_TAG_VALID | ((unsigned)f_slot << AMRVW_TAG_SLOT_SHIFT);
                for (int a = 0; a < AMRVW_RING - 1; ++a) {
                    const int idx = f_k - (AMRVW_RING - 1) + a;
                    sm.ringtag[(unsigned)(t1 + a) & (AMRVW_RING - 1)] = base_tag | (unsigned)idx | (idx == f_stop1 ? AMRVW_TAG_LAST : 0u);
                }
                pn_steps += fk; pn_fast += fk;
            }
            t = (unsigned)(t1 - 1);
            __syncwarp();
        }
    }
    if (feeds) cp_async_wait_all();
    if (pr.prof && feeds) {
        if (lane32 == 0) atomicAdd(pr.prof + 0, (unsigned long long)(clock64() - pc_t0));   // chase loop, per warp
        atomicAdd(pr.prof + 2, (unsigned long long)pn_fast);                                 // lock-steps in the lean loop, per pipeline
        atomicAdd(pr.prof + 4, (unsigned long long)pn_steps);                                // per pipeline
        atomicAdd(pr.prof + 5, (unsigned long long)pn_empty);
        atomicAdd(pr.prof + 6, (unsigned long long)f_started);
        atomicAdd(pr.prof + 7, (unsigned long long)pn_starved);
    }
}

template <int L, int MH, int NW, int NS> constexpr size_t ustream_raw_bytes() { return 12 * 32 * NW * sizeof(double2) + (size_t)NS * (32 / L) * sizeof(ShiftScratch<MH>); }

template <int L, int MH, int NW, int NS>
static cudaError_t launch_ustream(const Problem<double>& pr, const PipeParams& pp, const StreamBufs& sb, int sm_count, cudaStream_t st) {
    StreamQueues sq;
    sq.ready = sb.ready; sq.ready_slots = sb.ready_slots; sq.work = sb.work; sq.work_slots = sb.work_slots; sq.mask = sb.mask; sq.remaining = sb.remaining;
    sq.progress = sb.remaining + 1; sq.abort_flag = pr.counters + CNT_ABORT; sq.t0 = reinterpret_cast<unsigned long long*>(sb.remaining + 2);
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
    if (pp.stream_ctas > 0 && pp.stream_ctas < per_sm) per_sm = pp.stream_ctas;
    constexpr int NP = NW * (32 / L);
    const long long cap = (long long)sm_count * per_sm;
    const long long want = (pr.nbatch + NP - 1) / NP;       // more pipelines than polynomials would only poll
    const int blocks = (int)(want < cap ? want : cap);
    ustream_kernel<L, MH, NW, NS><<<blocks, 32 * (NW + 1 + NS), smem, st>>>(pr, pp, sq);
    return cudaGetLastError();
}

}  // namespace amrvw
