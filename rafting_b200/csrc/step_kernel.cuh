// step_kernel.cuh — the batched "event loop" kernel: one launch drains one batch for every group.
//
// Replaces, for all groups at once, the reference's per-context EventLoop path
// (M/support/EventLoop.java:87-101 + the RaftParticipant handlers and Async callbacks it runs).
// handlers.cuh restates every handler; this file schedules them:
//
//   * one sub-warp of W lanes per group (W = pow2 >= F = R-1); lane f keeps Leadership.State of
//     follower f in registers for the whole batch, every lane keeps a copy of the group scalars;
//   * row r+1's op and lane event are fetched (128-bit loads, coalesced across the warp) before row
//     r is processed, so each thread always has a full row of input in flight;
//   * FAST PATH (leader steady state — configs #2/#4): SUBMIT / keepAlive on a Leader and rows whose
//     lane events are all AE/IS acks run inline out of registers.  Acks of the F lanes are applied
//     concurrently (they touch disjoint Leadership.State objects); the only cross-lane effects —
//     tryCommit after each successful ack — are then evaluated in lane order on the vector
//     (new matchIndex for lanes <= f, old for lanes > f), which is what the serial order would see;
//   * SLOW PATH (everything else: elections, step-downs, inbound requests, flushes, sweeps, term-run
//     pushes, any per-event error) calls the generic handlers through by-value wrappers so the hot
//     state never has its address taken.  Both paths produce identical results by construction of
//     the fast path's entry conditions; tests/test_engine_gpu.py checks them against the oracle.
#pragma once
#include "handlers.cuh"

namespace rafting {

struct KArgs { Tables T; InboxD in; OutboxD out; const CfgD* cfg; };   // block-shared copy of the kernel arguments

struct Lite {                       // immutable per-thread context
    const CfgD* cfg; i64x2* runs; uint32_t gid, F, G; unsigned mask; int lane, sub0; bool lv;
};
__device__ __forceinline__ Ctx make_ctx(const Lite& k, int64_t now, int64_t draw) {
    Ctx c; c.cfg = k.cfg; c.runs = k.runs; c.gid = k.gid; c.F = k.F; c.G = k.G; c.mask = k.mask;
    c.lane = k.lane; c.sub0 = k.sub0; c.lv = k.lv; c.now = now; c.draw = draw;
    return c;
}
template <int W>
__device__ __forceinline__ Lite make_lite(const Tables& T, const CfgD* cfg, uint32_t gid) {
    Lite k;
    k.cfg = cfg; k.gid = gid; k.F = T.F; k.G = T.G; k.runs = T.g_runs + gid;
    k.lane = (int)((blockIdx.x * blockDim.x + threadIdx.x) % W);
    k.sub0 = (int)((threadIdx.x & 31u) & ~(uint32_t)(W - 1));
    k.mask = W == 32 ? 0xffffffffu : (((1u << W) - 1u) << k.sub0);
    k.lv = (uint32_t)k.lane < T.F;
    return k;
}

// ---- state movement between HBM tables and registers (L2-coherent loads: the slow path hands state
//      over through the tables inside one launch) ----
__device__ __forceinline__ void load_hot(const Tables& T, uint32_t gid, GS& g) {
    const uint64_t m = __ldcg(T.g_meta + gid);
    g.word = (uint32_t)m; g.inc = (uint32_t)(m >> 32);
    g.term = __ldcg(T.g_term + gid); g.commit = __ldcg(T.g_commit + gid);
    g.lo = __ldcg(T.g_lo + gid); g.hi = __ldcg(T.g_hi + gid); g.timer = __ldcg(T.g_timer + gid);
    const longlong2 ep = __ldcg((const longlong2*)(T.g_epoch + gid)); g.epochIndex = ep.x; g.epochTerm = ep.y;
    const longlong2 r0 = __ldcg((const longlong2*)(T.g_runs + gid)); g.r0s = r0.x; g.r0t = r0.y;
    g.err = __ldcg(T.g_err + gid);
}
__device__ __forceinline__ void load_cold(const Tables& T, uint32_t gid, GS& g) {
    const longlong2 el = __ldcg((const longlong2*)(T.g_elect + gid));
    g.electTerm = el.x; g.electInc = (uint32_t)(uint64_t)el.y; g.votes = (int32_t)((uint64_t)el.y >> 32);
}
__device__ __forceinline__ void store_hot(const Tables& T, uint32_t gid, const GS& g) {
    T.g_meta[gid] = (uint64_t)g.word | ((uint64_t)g.inc << 32);
    T.g_commit[gid] = g.commit; T.g_hi[gid] = g.hi; T.g_timer[gid] = g.timer; T.g_err[gid] = g.err;
}
__device__ __forceinline__ void store_warm(const Tables& T, uint32_t gid, const GS& g) {
    T.g_term[gid] = g.term; T.g_lo[gid] = g.lo;
    i64x2 v; v.x = g.epochIndex; v.y = g.epochTerm; T.g_epoch[gid] = v;
    v.x = g.r0s; v.y = g.r0t; T.g_runs[gid] = v;
}
__device__ __forceinline__ void store_cold(const Tables& T, uint32_t gid, const GS& g) {
    i64x2 v; v.x = g.electTerm; v.y = (int64_t)((uint64_t)g.electInc | ((uint64_t)(uint32_t)g.votes << 32));
    T.g_elect[gid] = v;
}
__device__ __forceinline__ void load_lane(const Tables& T, size_t li, LS& s) {
    const longlong2 nm = __ldcg((const longlong2*)(T.l_nm + li)), es = __ldcg((const longlong2*)(T.l_es + li)),
                    fr = __ldcg((const longlong2*)(T.l_fr + li));
    const int4 cn = __ldcg(T.l_cnt + li);
    s.next = nm.x; s.match = nm.y; s.lastEpoch = es.x; s.reqSucc = es.y; s.reqFail = fr.x; s.lastReq = fr.y;
    s.inflight = cn.x; s.rej = cn.y; s.fail = cn.z; s.pending = cn.w;
}
__device__ __forceinline__ void store_lane(const Tables& T, size_t li, const LS& s) {
    i64x2 v; int4 cn;
    v.x = s.next; v.y = s.match; T.l_nm[li] = v;
    v.x = s.lastEpoch; v.y = s.reqSucc; T.l_es[li] = v;
    v.x = s.reqFail; v.y = s.lastReq; T.l_fr[li] = v;
    cn.x = s.inflight; cn.y = s.rej; cn.z = s.fail; cn.w = s.pending; T.l_cnt[li] = cn;
}
__device__ __forceinline__ void zero_lane(LS& s) {
    s.next = s.match = s.lastEpoch = s.reqSucc = s.reqFail = s.lastReq = 0; s.inflight = s.rej = s.fail = s.pending = 0;
}

// ---------------------------------------------------------------------------------------------
// SLOW PATH: generic group op of row r (any kind, any role), sequential reference semantics.
// State comes from and goes back to the tables; outputs go straight to the outbox.
// Returns the step's dirty bits.
// ---------------------------------------------------------------------------------------------
template <int W>
__device__ __noinline__ uint32_t slow_op(const KArgs* ka, uint32_t i, uint32_t gid, uint32_t r, uint32_t kind, int64_t sweep, uint32_t dirty) {
    const Tables& T = ka->T; const InboxD& in = ka->in; const OutboxD& out = ka->out;
    const Lite k = make_lite<W>(T, ka->cfg, gid);
    const size_t gi = (size_t)r * in.n + i, li0 = (size_t)gid * T.F + (uint32_t)k.lane;
    GS g; LS s; RowOut ro;
    load_hot(T, gid, g); load_cold(T, gid, g); g.dirty = dirty;
    zero_lane(s); if (k.lv) load_lane(T, li0, s);
    ro.pm = 0; ro.bm = 0; ro.pe = 0; ro.bt = 0; ro.pp.x = ro.pp.y = ro.lc.x = ro.lc.y = ro.bl.x = ro.bl.y = 0;
    const bool alive = (g.word & W_ALIVE) != 0;
    uint32_t meta = 0, entoff = 0; int64_t now = sweep, draw = 0, a = 0, b = 0, cc = 0, d = 0;
    if (sweep == 0) {
        const uint64_t m = in.op_meta[gi]; meta = (uint32_t)m; entoff = (uint32_t)(m >> 32);
        const i64x2 nr = in.op_nr[gi]; now = nr.x; draw = nr.y;
        if (in.op_ab) { const i64x2 v = in.op_ab[gi]; a = v.x; b = v.y; }
        if (in.op_cd) { const i64x2 v = in.op_cd[gi]; cc = v.x; d = v.y; }
    }
    Ctx c = make_ctx(k, now, draw);
    int err = 0; Reply rep = {0, 0, 0};
    const int peer = (int)RAFTING_OP_PEER(meta); const uint32_t count = RAFTING_OP_COUNT(meta);
    if (!alive) err = RAFTING_ERR_CLOSED_GROUP;
    else if (kind == RAFTING_OP_SUBMIT) err = op_submit<W>(g, c, ro, s, count, (uint64_t)a);
    else if (kind == RAFTING_OP_TIMEOUT) err = op_timeout<W>(g, c, ro, s, (uint64_t)a);
    else if (in.flags & RAFTING_INBOX_NO_REQUESTS) err = RAFTING_ERR_BAD_EVENT;   // the caller promised none
    else if (kind == RAFTING_OP_AE_REQUEST) {
        const int64_t first = in.op_e ? in.op_e[gi] : (int64_t)((uint64_t)b + 1u);
        const int64_t* terms = in.ent_terms ? in.ent_terms + entoff : nullptr;
        if (count > 0 && (!terms || (uint64_t)entoff + count > in.ent_count)) err = RAFTING_ERR_BAD_EVENT;
        else err = op_append_entries(g, c, ro, peer, a, b, cc, first, count, terms, d, rep);
    }
    else if (kind == RAFTING_OP_PREVOTE_REQ) err = op_pre_vote(g, c, ro, peer, a, b, cc, rep);
    else if (kind == RAFTING_OP_VOTE_REQ) err = op_request_vote(g, c, ro, peer, a, b, cc, rep);
    else if (kind == RAFTING_OP_IS_REQUEST) err = op_install_snapshot(g, c, ro, a, d != 0, rep);
    else if (kind == RAFTING_OP_FLUSH) err = log_flush(g, c, b, cc);
    else err = RAFTING_ERR_BAD_EVENT;
    if (err) { if (alive) flag_err(g, err); rep.valid = 0; }
    // outputs of this row
    if (k.lv && out.plan_meta) {
        const size_t li = gi * T.F + (uint32_t)k.lane;
        out.plan_meta[li] = ro.pm;
        if (ro.pm != 0) { out.plan_pp[li] = ro.pp; out.plan_lc[li] = ro.lc; out.plan_epoch[li] = ro.pe; }
    }
    if (k.lane == 0) {
        if (out.rep_meta) {
            out.rep_meta[gi] = (uint32_t)(rep.valid ? 1 : 0) | ((uint32_t)(rep.success ? 1 : 0) << 1) | ((uint32_t)err << 8);
            if (rep.valid) out.rep_term[gi] = rep.term;
        }
        if (out.ballot_meta && ro.bm != 0) { out.ballot_meta[gi] = ro.bm; out.ballot_term[gi] = ro.bt; out.ballot_last[gi] = ro.bl; }
        store_hot(T, gid, g); store_warm(T, gid, g); store_cold(T, gid, g);
    }
    if (k.lv) store_lane(T, li0, s);
    return g.dirty;
}

// SLOW PATH: generic lane events of row r, applied strictly in lane order
template <int W>
__device__ __noinline__ uint32_t slow_events(const KArgs* ka, uint32_t i, uint32_t gid, uint32_t r, uint32_t dirty) {
    const Tables& T = ka->T; const InboxD& in = ka->in; const OutboxD& out = ka->out;
    const Lite k = make_lite<W>(T, ka->cfg, gid);
    const size_t gi = (size_t)r * in.n + i, li0 = (size_t)gid * T.F + (uint32_t)k.lane;
    GS g; LS s; RowOut ro;
    load_hot(T, gid, g); load_cold(T, gid, g); g.dirty = dirty;
    zero_lane(s); if (k.lv) load_lane(T, li0, s);
    ro.pm = 0; ro.bm = 0; ro.pe = 0; ro.bt = 0; ro.pp.x = ro.pp.y = ro.lc.x = ro.lc.y = ro.bl.x = ro.bl.y = 0;
    uint64_t em = 0; i64x2 etn = {0, 0}, eel = {0, 0};
    if (k.lv) {
        const size_t li = gi * T.F + (uint32_t)k.lane;
        em = in.ev_meta[li];
        if (RAFTING_EVM_KIND(em) != RAFTING_EV_NONE) { etn = in.ev_tn[li]; if (in.ev_el) eel = in.ev_el[li]; }
    }
    Ctx c = make_ctx(k, 0, 0);
    const int self = (int)k.cfg->local_slot;
    unsigned pending = sub_ballot(c, RAFTING_EVM_KIND(em) != RAFTING_EV_NONE, W);
    while (pending) {
        const int f = __ffs(pending) - 1; pending &= pending - 1;
        const uint64_t m = (uint64_t)shfl64(c.mask, (int64_t)em, f, W);
        const int64_t respTerm = shfl64(c.mask, etn.x, f, W);
        c.now = shfl64(c.mask, etn.y, f, W); c.draw = 0;
        const uint32_t ek = RAFTING_EVM_KIND(m), outcome = RAFTING_EVM_OUTCOME(m), inc = RAFTING_EVM_INC(m);
        const bool success = RAFTING_EVM_SUCCESS(m) != 0;
        int err = 0;
        if (ek == RAFTING_EV_AE_ACK || ek == RAFTING_EV_IS_ACK) {
            // AE-Echo / IS-Echo — Leader.java:174-188,218-237
            if (role_of(g) == RAFTING_ROLE_LEADER && inc == g.inc && (g.word & W_PREPARED)) {
                const bool mine = c.lane == f;
                if (mine) s.inflight--;
                if (outcome == RAFTING_OUT_OK) {
                    if (respTerm > g.term) err = switch_to(g, c, ro, RAFTING_ROLE_FOLLOWER, respTerm, lane_to_slot(c, f));
                    else {
                        const bool snap = ek == RAFTING_EV_IS_ACK;
                        int e = 0;
                        if (mine) {
                            stat_success(s, c.now, !success);
                            e = update_index(s, eel.x, snap ? eel.x : eel.y, success, snap);
                        }
                        e = __shfl_sync(c.mask, e, f, W);
                        if (e) err = e;
                        else if (!snap && success) err = try_commit<W>(g, c, s.match);
                    }
                } else if (mine) stat_failure(s, c.now, outcome == RAFTING_OUT_ERROR, false);
            }
        } else if (ek == RAFTING_EV_PV_REPLY) {
            // PV-Echo — Follower.java:258-270
            if (role_of(g) == RAFTING_ROLE_FOLLOWER && inc == g.inc && (g.word & W_TIMEOUT_DET) && outcome == RAFTING_OUT_OK) {
                const int64_t nextTerm = (int64_t)((uint64_t)g.term + 1u);
                if (respTerm > nextTerm) err = switch_to(g, c, ro, RAFTING_ROLE_FOLLOWER, respTerm, lane_to_slot(c, f));
                else if (success && ++g.votes >= majority(c)) err = switch_to(g, c, ro, RAFTING_ROLE_CANDIDATE, nextTerm, self);
            }
        } else if (ek == RAFTING_EV_RV_REPLY) {
            // RV-Echo — Candidate.java:112-134 (and the elected Candidate's surviving head, :75-80)
            if (role_of(g) == RAFTING_ROLE_CANDIDATE && inc == g.inc) {
                if (outcome == RAFTING_OUT_OK) {
                    if (respTerm > g.term) err = switch_to(g, c, ro, RAFTING_ROLE_FOLLOWER, respTerm, lane_to_slot(c, f));
                    else if (success && ++g.votes >= majority(c)) {
                        g.electInc = g.inc; g.electTerm = g.term; g.word &= ~W_ELECT_ABORT;
                        err = switch_to(g, c, ro, RAFTING_ROLE_LEADER, g.term, self);
                    }
                }
            } else if (g.electInc != 0 && inc == g.electInc && !(g.word & W_ELECT_ABORT) && outcome == RAFTING_OUT_OK) {
                if (respTerm > g.electTerm) {
                    g.word |= W_ELECT_ABORT;
                    err = switch_to(g, c, ro, RAFTING_ROLE_FOLLOWER, respTerm, lane_to_slot(c, f));
                } else if (success) err = switch_to(g, c, ro, RAFTING_ROLE_LEADER, g.electTerm, self);
            }
        } else err = RAFTING_ERR_BAD_EVENT;
        if (err) flag_err(g, err);
    }
    if (k.lane == 0) {
        if (out.ballot_meta && ro.bm != 0) { out.ballot_meta[gi] = ro.bm; out.ballot_term[gi] = ro.bt; out.ballot_last[gi] = ro.bl; }
        store_hot(T, gid, g); store_warm(T, gid, g); store_cold(T, gid, g);
    }
    if (k.lv) store_lane(T, li0, s);
    return g.dirty;
}

// quorum index over (new matchIndex for lanes <= upto, old for lanes > upto): what Leader.tryCommit
// sees right after lane `upto`'s ack in the serial order (Leader.java:247-261, Leadership.java:116-130).
// Executed by the WHOLE warp (full-mask shuffles of width W): every sub-warp gets its own answer.
template <int W>
__device__ __forceinline__ void quorum_after(const Lite& k, int64_t oldM, int64_t newM, int upto, int64_t& full, int64_t& major) {
    const int F = (int)k.F;
    const int64_t mine = k.lane <= upto ? newM : oldM;
    if (W == 1) { full = major = mine; return; }
    if (W == 2) {                                   // R = 3: sorted = [min, max]
        const int64_t other = shfl64(0xffffffffu, mine, k.lane ^ 1, W);
        full = mine < other ? mine : other; major = mine < other ? other : mine;
        return;
    }
    int64_t mn = I64MAX; int rank = 0;
#pragma unroll
    for (int j = 0; j < W; j++) {
        const int64_t mj = shfl64(0xffffffffu, mine, j, W);
        if (j < F) { mn = mj < mn ? mj : mn; rank += (mj < mine) || (mj == mine && j < k.lane); }
    }
    const unsigned b = __ballot_sync(0xffffffffu, k.lv && rank == F / 2);
    const unsigned sel = W == 32 ? b : ((b >> k.sub0) & ((1u << W) - 1u));
    full = mn; major = shfl64(0xffffffffu, mine, __ffs(sel) - 1, W);
}
__device__ __forceinline__ unsigned sub_bits(const Lite& k, unsigned b, int W) {
    return W == 32 ? b : ((b >> k.sub0) & ((1u << W) - 1u));
}

#ifndef RAFTING_MINBLOCKS
#define RAFTING_MINBLOCKS 7
#endif
#ifndef RAFTING_STAGES
#define RAFTING_STAGES 2
#endif
constexpr int TPB = 128;                 // threads per block
constexpr int NST = RAFTING_STAGES;      // input rows staged in shared memory (NST-1 rows in flight per thread)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int W>
struct __align__(16) Stage {             // one staged input row of this block
    i64x2    op_nr[TPB / W];
    i64x2    op_ab[TPB / W];
    i64x2    ev_tn[TPB];
    i64x2    ev_el[TPB];
    uint64_t op_meta[TPB / W];
    uint64_t ev_meta[TPB];
};

// Control flow is WARP-UNIFORM: every collective (ballot / shuffle) is executed by all 32 lanes with
// the full mask, per-group decisions are predicates, and the generic slow path is entered by the whole
// warp when any of its groups needs it (the generic handlers are correct for every group state).
template <int W>
__global__ void __launch_bounds__(TPB, RAFTING_MINBLOCKS)
step_kernel(Tables T, InboxD in, OutboxD out, const CfgD* __restrict__ cfgp, CfgD cfg) {
    __shared__ KArgs ka;
    __shared__ Stage<W> stage[NST];
    if (threadIdx.x == 0) { ka.T = T; ka.in = in; ka.out = out; ka.cfg = cfgp; }
    __syncthreads();
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i = tid / W;
    if ((tid & ~31u) / W >= in.n) return;                    // whole warps beyond the batch leave together
    const uint32_t F = T.F;
    uint32_t gid = 0; bool valid = i < in.n;
    if (valid) { gid = in.gids ? in.gids[i] : i; if (gid >= T.G) { valid = false; gid = 0; } }
    const Lite k = make_lite<W>(T, cfgp, gid);
    const int lane = k.lane; const bool lv = valid && k.lv;
    const uint32_t tl = threadIdx.x, gl = threadIdx.x / W;   // this thread's / this group's slot in a stage
    constexpr unsigned FULL = 0xffffffffu;

    // ---- hot group scalars + this lane's follower slot live in registers for the whole batch ----
    GS g; LS s;
    g.word = 0; g.inc = 0; g.err = 0; g.term = g.commit = g.lo = g.hi = g.timer = g.epochIndex = g.epochTerm = g.r0s = g.r0t = 0;
    if (valid) load_hot(T, gid, g);
    g.dirty = 0; g.electTerm = 0; g.electInc = 0; g.votes = 0;
    const size_t li0 = (size_t)gid * F + (uint32_t)lane;
    zero_lane(s); if (lv) load_lane(T, li0, s);
    const bool hasOps = in.op_meta != nullptr, hasEv = in.ev_meta != nullptr;

    // ---- asynchronous staging of the input rows (cp.async, no registers held): the group op is
    //      fetched by lanes 0/1 of the sub-warp, each lane fetches its own event ----
#define RAFTING_ISSUE(R_)                                                                                  \
    {                                                                                                      \
        const uint32_t r_ = (R_);                                                                          \
        if (r_ < in.rows && valid) {                                                                       \
            Stage<W>& st_ = stage[r_ % NST];                                                               \
            const size_t gi_ = (size_t)r_ * in.n + i;                                                      \
            if (hasOps) {                                                                                  \
                if (lane == 0) { cp_async8(&st_.op_meta[gl], in.op_meta + gi_); cp_async16(&st_.op_nr[gl], in.op_nr + gi_); } \
                if (lane == (W > 1 ? 1 : 0) && in.op_ab) cp_async16(&st_.op_ab[gl], in.op_ab + gi_);       \
            }                                                                                              \
            if (hasEv && lv) {                                                                             \
                const size_t li_ = gi_ * F + (uint32_t)lane;                                               \
                cp_async8(&st_.ev_meta[tl], in.ev_meta + li_);                                             \
                cp_async16(&st_.ev_tn[tl], in.ev_tn + li_);                                                \
                if (in.ev_el) cp_async16(&st_.ev_el[tl], in.ev_el + li_);                                  \
            }                                                                                              \
        }                                                                                                  \
        cp_async_commit();                                                                                 \
    }
#pragma unroll
    for (int p = 0; p < NST - 1; p++) RAFTING_ISSUE((uint32_t)p);

    for (uint32_t r = 0; r < in.rows; r++) {
        const size_t gi = (size_t)r * in.n + i;
        RAFTING_ISSUE(r + NST - 1);
        cp_async_wait<NST - 1>();                            // row r has landed
        __syncwarp();                                        // ... including the parts sibling lanes fetched
        const Stage<W>& st = stage[r % NST];
        const bool alive = valid && (g.word & W_ALIVE) != 0;

        // ================= group op =================
        const int64_t sweep = in.row_now ? in.row_now[r] : 0;             // grid-uniform
        uint32_t meta = 0, kind = RAFTING_OP_NONE; int64_t now = 0; uint64_t unavail = 0;
        if (sweep == 0 && hasOps && valid) {
            meta = (uint32_t)st.op_meta[gl]; kind = RAFTING_OP_KIND(meta); now = st.op_nr[gl].x;
            if (in.op_ab) unavail = (uint64_t)st.op_ab[gl].x;
        }
        bool slowOp;                                          // this group's op needs the generic handler
        if (sweep != 0) {
            const bool due = alive && ((role_of(g) == RAFTING_ROLE_LEADER) ? (g.timer <= sweep)
                                       : (g.timer > 0 && g.timer != I64MAX && g.timer <= sweep));
            kind = due ? (uint32_t)RAFTING_OP_TIMEOUT : (uint32_t)RAFTING_OP_NONE;
            slowOp = due;
        } else {
            const bool hbeat = kind == RAFTING_OP_TIMEOUT;
            const bool fastOk = alive && role_of(g) == RAFTING_ROLE_LEADER && (g.word & W_PREPARED) &&
                                (hbeat || (kind == RAFTING_OP_SUBMIT && nruns_of(g) > 0 && g.r0t == g.term));
            slowOp = kind != RAFTING_OP_NONE && !fastOk;
        }
        const bool opHere = kind != RAFTING_OP_NONE;
        // ---- fast op, computed tentatively (nothing is mutated until the warp agrees to stay fast) ----
        // Leader keepAlive -> replicateLog(true) (RaftRoutine.java:53-62, Leader.java:119-126), or
        // RaftStub.process -> Leader.acceptCommand -> replicateLog(false) (RaftStub.java:79-91, Leader.java:128-140)
        const bool P = opHere && !slowOp;
        const bool hbeat = kind == RAFTING_OP_TIMEOUT;
        const unsigned rb = __ballot_sync(FULL, P && lv && state_ready(s, cfg.avail_critical_point, cfg.recovery_cool_down_ms, now));
        const int rcnt = __popc(sub_bits(k, rb, W));
        const bool ready = rcnt >= 1 && (1 + rcnt > (int)F / 2);             // Leader.isReady, Leader.java:52-64
        const bool go = P && (hbeat || ready);
        uint32_t count = RAFTING_OP_COUNT(meta); if (count == 0) count = 1;
        const int64_t hiN = g.hi + ((go && !hbeat) ? (int64_t)count : 0);   // RocksLog.newEntry x count, same term run
        // Leader.replicateLog for this lane's follower (Leader.java:142-245), pure part
        int e = 0; uint64_t pm = 0; i64x2 pp = {0, 0}, lc = {0, 0}; int dInflight = 0; bool failStat = false;
        if (go && lv) {
            const uint64_t hbBit = hbeat ? (1ull << 4) : 0ull, incBits = (uint64_t)g.inc << 32;
            if ((unavail >> lane) & 1ull) { failStat = true; pm = RAFTING_PLAN_UNAVAILABLE | hbBit | incBits; }
            else if (s.inflight > RAFTING_IN_FLIGHT_LIMIT / (hbeat ? 10 : 1)) pm = RAFTING_PLAN_SKIP_INFLIGHT | hbBit | incBits;
            else if (s.pending) {
                pm = RAFTING_PLAN_IS | hbBit | incBits; pp.x = g.epochIndex; pp.y = g.epochTerm; lc.x = g.epochIndex; lc.y = g.commit;
                dInflight = 1;
            } else {
                int64_t prevTerm = g.epochTerm, prevIndex = g.epochIndex, lastIndex;
                const int64_t nm1 = (int64_t)((uint64_t)s.next - 1u);
                const int64_t nextIndex = nm1 > g.epochIndex ? nm1 : g.epochIndex;
                int64_t idx = nextIndex, len = (RAFTING_REPLICATE_LIMIT >> (hbeat ? 1 : 0)) + 1;
                if (idx == g.epochIndex) { idx++; len--; }                   // RocksLog.batch, RocksLog.java:134-137
                int64_t eFirst = 0, eCount = 0;
                if (len > 0 && nruns_of(g) > 0) {
                    const int64_t hiKey = idx + len - 1;
                    if (idx < g.lo && g.lo <= hiKey) e = RAFTING_ERR_LOG_VACANCY;
                    else {
                        const int64_t a = idx > g.lo ? idx : g.lo, b = hiKey < hiN ? hiKey : hiN;
                        if (a <= b) { eFirst = a; eCount = b - a + 1; }
                    }
                }
                uint32_t cnt = 0;
                if (eCount > 0) {
                    if (eFirst == nextIndex) {                               // Leader.java:198-201
                        int64_t t = g.r0t;
                        if (eFirst < g.r0s) {                                // older term run: walk the table (rare)
                            bool found = false;
#pragma unroll 1
                            for (int q = 1; q < nruns_of(g); q++) {
                                const i64x2 run = k.runs[(size_t)q * k.G];
                                if (!found && eFirst >= run.x) { t = run.y; found = true; }
                            }
                        }
                        prevTerm = t; prevIndex = eFirst; eFirst++; eCount--;
                    } else if (eFirst != g.epochIndex + 1) e = RAFTING_ERR_LOG_START;
                    lastIndex = (eCount == 0) ? prevIndex : eFirst + eCount - 1;
                    cnt = (uint32_t)eCount;
                } else lastIndex = g.epochIndex;
                pm = RAFTING_PLAN_AE | hbBit | ((uint64_t)cnt << 16) | incBits;
                pp.x = prevIndex; pp.y = prevTerm; lc.x = lastIndex; lc.y = g.commit;
                dInflight = 1;
            }
        }
        // the AssertionErrors inside replicateLog abort the follower loop midway: keep those rows serial
        const bool warpSlowOp = __any_sync(FULL, slowOp || e != 0);
        if (!warpSlowOp) {
            // ---- commit the fast op ----
            if (P) {
                if (hbeat) g.timer = (I64MAX - cfg.heartbeat_ms < now) ? I64MAX : now + cfg.heartbeat_ms;   // resetTimer, Leader branch
                else { g.word = ready ? (g.word | W_READY) : (g.word & ~W_READY); if (!ready) flag_err(g, RAFTING_ERR_NOT_READY); }
                g.hi = hiN;
                if (go && lv) {
                    if (now > s.lastReq) s.lastReq = now;                    // Leader.java:158
                    if (failStat) stat_failure(s, now, true, false);         // Leader.java:241-243
                    s.inflight += dInflight;
                }
            }
            if (lv && out.plan_meta) {
                const size_t li = gi * F + (uint32_t)lane;
                out.plan_meta[li] = pm;
                if (pm != 0) { out.plan_pp[li] = pp; out.plan_lc[li] = lc; out.plan_epoch[li] = g.epochIndex; }
            }
            if (valid && lane == 0) {
                if (out.rep_meta) out.rep_meta[gi] = (P && !hbeat && !ready) ? ((uint32_t)RAFTING_ERR_NOT_READY << 8) : 0u;
                if (out.ballot_meta) out.ballot_meta[gi] = 0;
            }
        } else {
            // ---- the whole warp hands its state over through the tables and runs the generic handler ----
            if (valid && lane == 0) { store_hot(T, gid, g); if (out.ballot_meta) out.ballot_meta[gi] = 0; }
            if (lv) store_lane(T, li0, s);
            __syncwarp();
            uint32_t dirty = g.dirty;
            if (valid && opHere) dirty = slow_op<W>(&ka, i, gid, r, kind, sweep, g.dirty);
            else {
                if (lv && out.plan_meta) out.plan_meta[gi * F + (uint32_t)lane] = 0;
                if (valid && lane == 0 && out.rep_meta) out.rep_meta[gi] = 0;
            }
            __syncwarp();
            if (valid) { load_hot(T, gid, g); g.dirty = dirty; }
            if (lv) load_lane(T, li0, s);
        }

        // ================= lane events =================
        if (hasEv) {
            uint64_t em = 0; i64x2 etn = {0, 0}, eel = {0, 0};
            const bool aliveNow = valid && (g.word & W_ALIVE) != 0;
            if (lv && aliveNow) { em = st.ev_meta[tl]; etn = st.ev_tn[tl]; if (in.ev_el) eel = st.ev_el[tl]; }
            const uint32_t ek = RAFTING_EVM_KIND(em);
            const bool present = ek != RAFTING_EV_NONE;
            const bool isAck = ek == RAFTING_EV_AE_ACK || ek == RAFTING_EV_IS_ACK;
            const bool leaderLive = role_of(g) == RAFTING_ROLE_LEADER && (g.word & W_PREPARED);
            // per-lane classification of what the serial order would do with this lane's event
            const bool live = present && isAck && leaderLive && RAFTING_EVM_INC(em) == g.inc;
            const bool okOutcome = RAFTING_EVM_OUTCOME(em) == RAFTING_OUT_OK;
            const bool stepDown = live && okOutcome && etn.x > g.term;      // Leader.java:178-180,224-226
            const bool oddKind = present && !isAck;                         // vote replies / unknown kinds
            // the only per-event error of an ack is "match index should not rollback" (Leadership.java:76-81):
            // known before anything is applied, and rare, so such rows keep the serial bookkeeping
            const bool snap = ek == RAFTING_EV_IS_ACK;
            const bool rollback = live && okOutcome && (snap ? eel.x : eel.y) < s.match;
            const bool warpSlowEv = __any_sync(FULL, oddKind || stepDown || rollback);
            if (!warpSlowEv) {
                // ---- AE-Echo / IS-Echo of all lanes applied concurrently (disjoint Leadership.State objects,
                //      Leader.java:174-188,218-237); acks for a dead Leader object are dropped ----
                const int64_t oldMatch = s.match; bool trig = false;
                if (live) {
                    s.inflight--;
                    if (okOutcome) {
                        const bool success = RAFTING_EVM_SUCCESS(em) != 0;
                        stat_success(s, etn.y, !success);
                        update_index(s, eel.x, snap ? eel.x : eel.y, success, snap);
                        trig = !snap && success;
                    } else stat_failure(s, etn.y, RAFTING_EVM_OUTCOME(em) == RAFTING_OUT_ERROR, false);
                }
                // ---- Leader.tryCommit after each successful AE ack, in lane order (Leader.java:247-280) ----
                const unsigned tb = __ballot_sync(FULL, trig);
                if (tb) {
                    const unsigned myTb = sub_bits(k, tb, W);
#pragma unroll
                    for (int f = 0; f < W; f++) {
                        if (f < (int)F) {
                            int64_t full, major;
                            quorum_after<W>(k, oldMatch, s.match, f, full, major);
                            if ((myTb >> f) & 1u) {
                                int cerr = 0;
                                if (full > major) cerr = RAFTING_ERR_IMPOSSIBLE_REPL;
                                else if (major != 0) {
                                    Ctx c = make_ctx(k, 0, 0);
                                    int64_t t;
                                    if (!term_at(g, c, major, t)) flag_err(g, RAFTING_ERR_TRY_COMMIT_FAILED);
                                    else {
                                        const int64_t ci = (t == g.term) ? major : full;
                                        if (ci != 0 && ci != g.commit) cerr = commit_log(g, ci);
                                    }
                                }
                                if (cerr) flag_err(g, cerr);
                            }
                        }
                    }
                }
            } else {
                if (valid && lane == 0) store_hot(T, gid, g);
                if (lv) store_lane(T, li0, s);
                __syncwarp();
                uint32_t dirty = g.dirty;
                const unsigned pb = sub_bits(k, __ballot_sync(FULL, present), W);
                if (valid && pb != 0) dirty = slow_events<W>(&ka, i, gid, r, g.dirty);
                __syncwarp();
                if (valid) { load_hot(T, gid, g); g.dirty = dirty; }
                if (lv) load_lane(T, li0, s);
            }
        }
        __syncwarp();                                        // siblings are done with this stage before it is refilled
    }
    cp_async_wait<0>();

    // ---- write back: the columns the fast path can change; the slow path stored the rest itself ----
    if (lv) store_lane(T, li0, s);
    if (valid && lane == 0) {
        store_hot(T, gid, g);
        if (out.commit_index) out.commit_index[gid] = g.commit;
        if (out.current_term) out.current_term[gid] = g.term;
        if (out.role_word)
            out.role_word[gid] = (uint32_t)role_of(g) | ((uint32_t)(ballot_of(g) + 1) << 8) | ((uint32_t)(leader_of(g) + 1) << 16) |
                                 ((g.word & W_TIMEOUT_DET) ? (1u << 24) : 0u) | ((g.word & W_READY) ? (1u << 25) : 0u) |
                                 ((g.dirty & 1u) << 30) | (((g.dirty >> 1) & 1u) << 31);
        if (out.incarnation) out.incarnation[gid] = g.inc;
        if (out.err_word) out.err_word[gid] = g.err;
    }
}

}  // namespace rafting
