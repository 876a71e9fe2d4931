// step_kernel.cuh — the batched "event loop" kernel: one launch drains one batch for every group.
//
// Replaces, for all groups at once, the reference's per-context EventLoop path
// (M/support/EventLoop.java:87-101 + the RaftParticipant handlers and Async callbacks it runs).
//
//   K1 ack_quorum_commit : AE-Echo / IS-Echo + Leadership.State.{statSuccess,statFailure,updateIndex,
//                          majorIndices} + Leader.tryCommit + RocksLog.markCommitted
//                          (Leader.java:174-188,218-237,247-280; Leadership.java:53-130; RocksLog.java:100-109)
//   K6 replicate_plan    : Leader.prepareReplication / replicateLog / isReady (Leader.java:30-64,142-245)
//   K2..K5               : handlers.cuh
//
// Mapping (v4, after profiling two sub-warp-per-group versions — profiles/r1a, r1b — which were
// bound by divergence replays, redundant per-lane group work and register spills): ONE THREAD PER
// GROUP.  The thread keeps the group scalars and the Leadership.State of all F followers in registers
// for the whole batch and walks the rows in order, which is exactly the reference's serial order, so
// there is nothing to reconcile across lanes.  Consecutive threads own consecutive groups, so every
// column access is a coalesced 8/16-byte-per-thread load or store.
//   * inputs are staged by cp.async into a per-block shared-memory ring NST rows deep (each thread
//     copies only its own op and its F events, so no barrier is needed, only cp.async.wait_group):
//     NST-1 rows (~120 B per thread at R=3) are always in flight per thread without holding registers;
//   * FAST PATH, inline: SUBMIT / keepAlive on a prepared Leader and rows whose events are acks;
//   * SLOW PATH, out of line: everything else (elections, step-downs, inbound requests, flushes,
//     sweeps, term-run pushes, match-index rollbacks).  State is handed over through the tables so
//     the hot state never has its address taken.
#pragma once
#include "handlers.cuh"

namespace rafting {

struct KArgs { Tables T; InboxD in; OutboxD out; const CfgD* cfg; };   // block-shared copy of the kernel arguments

__device__ __forceinline__ Ctx make_ctx(const Tables& T, const CfgD* cfg, uint32_t gid, int64_t now, int64_t draw) {
    Ctx c; c.cfg = cfg; c.runs = T.g_runs + gid; c.gid = gid; c.F = T.F; c.G = T.G; c.now = now; c.draw = draw;
    return c;
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

// ---------------------------------------------------------------------------------------------
// K6: Leader.isReady / prepareReplication / replicateLog — Leader.java:30-64,142-245
// FT = compile-time bound of the follower loops (arrays stay in registers), F = actual count.
// ---------------------------------------------------------------------------------------------
template <int FT>
__device__ __forceinline__ bool leader_ready(GS& g, const LS (&s)[FT], int F, int32_t crit, int64_t cool, int64_t now) {
    int cnt = 0;
#pragma unroll (FT <= 8 ? FT : 1)
    for (int f = 0; f < FT; f++) if (f < F && state_ready(s[f], crit, cool, now)) cnt++;
    // the Java loop only returns true from inside `isReady(..) && ++ready > half` (Leader.java:55-62)
    const bool ready = (g.word & W_PREPARED) && cnt >= 1 && (1 + cnt > F / 2);
    g.word = ready ? (g.word | W_READY) : (g.word & ~W_READY);
    return ready;
}

struct PlanSink {                   // where replicateLog's per-follower RPC plans of this row go
    uint64_t* pm; i64x2* pp; i64x2* lc; int64_t* pe;      // already offset to (row, group, lane 0); pm may be null
};
__device__ __forceinline__ void put_plan(const PlanSink& o, int f, uint64_t pm, int64_t p0, int64_t p1, int64_t l0, int64_t l1, int64_t pe) {
    if (!o.pm) return;
    o.pm[f] = pm;
    if (pm != 0) { i64x2 v; v.x = p0; v.y = p1; o.pp[f] = v; v.x = l0; v.y = l1; o.lc[f] = v; o.pe[f] = pe; }
}

template <int FT>
__device__ __forceinline__ int replicate_log(GS& g, const Ctx& c, LS (&s)[FT], int F, bool heartbeat, uint64_t unavail, const PlanSink& o) {
    if (!(g.word & W_PREPARED)) {                                            // prepareReplication, Leader.java:30-50
        int64_t li, lt; last_or_epoch(g, li, lt);
#pragma unroll (FT <= 8 ? FT : 1)
        for (int f = 0; f < FT; f++) {
            s[f].next = (int64_t)((uint64_t)li + 1u); s[f].match = 0; s[f].lastEpoch = g.epochIndex;
            s[f].reqSucc = 0; s[f].reqFail = 0; s[f].lastReq = 0; s[f].inflight = 0; s[f].rej = 0; s[f].fail = 0; s[f].pending = 0;
        }
        g.word |= W_PREPARED;
    }
    const int64_t epochIndex = g.epochIndex, epochTerm = g.epochTerm, leaderCommit = g.commit, now = c.now;
    const uint64_t hb = heartbeat ? (1ull << 4) : 0ull, incBits = (uint64_t)g.inc << 32;
    const int limit = RAFTING_IN_FLIGHT_LIMIT / (heartbeat ? 10 : 1);        // :162
    const int64_t fetch = RAFTING_REPLICATE_LIMIT >> (heartbeat ? 1 : 0);    // :194
    const bool nonEmpty = nruns_of(g) > 0;
    int err = 0;
#pragma unroll (FT <= 8 ? FT : 1)
    for (int f = 0; f < FT; f++) {
        if (f >= F) break;
        LS& x = s[f];
        // ---- the AppendEntries plan, computed branch-free (most followers take this path) ----
        // RaftLog.batch(max(nextIndex-1, epoch.index), fetch+1), RocksLog.java:131-166, over the contiguous
        // stored range [lo, hi]; then Leader.java:196-212.  With `has` = batch non-empty, a = first returned
        // index, b = last returned index:  prev = (a, term(a)) if a == nextIndex else epoch;
        // entries = (prev, b];  lastIndex = has ? b : epoch.index.
        const int64_t nm1 = (int64_t)((uint64_t)x.next - 1u);
        const int64_t nextIndex = nm1 > epochIndex ? nm1 : epochIndex;       // :193
        const bool atEpoch = nextIndex == epochIndex;
        const int64_t idx = nextIndex + (atEpoch ? 1 : 0), len = fetch + (atEpoch ? 0 : 1);
        const int64_t hiKey = idx + len - 1;
        const int64_t a = idx > g.lo ? idx : g.lo, b = hiKey < g.hi ? hiKey : g.hi;
        const bool scan = nonEmpty && len > 0;
        const bool vacancy = scan && idx < g.lo && g.lo <= hiKey;            // RocksLog.java:161-163
        const bool has = scan && a <= b;
        const bool isPrev = has && a == nextIndex;                           // :198-201
        const bool badStart = has && !isPrev && a != epochIndex + 1;         // :202-204
        int64_t prevTerm = epochTerm;
        if (isPrev) {
            prevTerm = g.r0t;
            if (a < g.r0s) { int64_t t = 0; term_at(g, c, a, t); prevTerm = t; }   // older term run: walk the table (rare)
        }
        const int64_t prevIndex = isPrev ? a : epochIndex;
        const int64_t lastIndex = has ? b : epochIndex;
        const uint64_t count = has ? (uint64_t)(b - a + (isPrev ? 0 : 1)) : 0ull;
        // ---- which RPC, in the order Leader.replicateLog tests them ----
        const bool unav = ((unavail >> f) & 1ull) != 0;                      // :241-243
        const bool skip = !unav && x.inflight > limit;                       // :163-166
        const bool snap = !unav && !skip && x.pending != 0;                  // :168-190
        const bool ae = !unav && !skip && !snap;
        const int e = ae ? (vacancy ? RAFTING_ERR_LOG_VACANCY : (badStart ? RAFTING_ERR_LOG_START : 0)) : 0;
        if (err) { put_plan(o, f, 0, 0, 0, 0, 0, 0); continue; }             // an Error already aborted the follower loop
        if (now > x.lastReq) x.lastReq = now;                                // :158
        if (e) { err = e; put_plan(o, f, 0, 0, 0, 0, 0, 0); continue; }
        if (unav) stat_failure(x, now, true, false);
        x.inflight += (snap || ae) ? 1 : 0;                                  // :173, :217
        const uint64_t kindBits = unav ? (uint64_t)RAFTING_PLAN_UNAVAILABLE : skip ? (uint64_t)RAFTING_PLAN_SKIP_INFLIGHT
                                  : snap ? (uint64_t)RAFTING_PLAN_IS : (uint64_t)RAFTING_PLAN_AE;
        const uint64_t pm = kindBits | hb | incBits | (ae ? (count << 16) : 0ull);
        const int64_t p0 = ae ? prevIndex : (snap ? epochIndex : 0), p1 = ae ? prevTerm : (snap ? epochTerm : 0);
        const int64_t l0 = ae ? lastIndex : (snap ? epochIndex : 0), l1 = (ae || snap) ? leaderCommit : 0;
        put_plan(o, f, pm, p0, p1, l0, l1, epochIndex);                      // :172, :216
    }
    return err;
}

// K1: Leader.tryCommit + Leadership.State.majorIndices — Leader.java:247-280, Leadership.java:116-130
template <int FT>
__device__ __forceinline__ int try_commit(GS& g, const Ctx& c, const LS (&s)[FT], int F) {
    int64_t full = I64MAX, major = 0;
#pragma unroll (FT <= 8 ? FT : 1)
    for (int a = 0; a < FT; a++) if (a < F) full = s[a].match < full ? s[a].match : full;
    if (FT == 1) major = s[0].match;
    else if (FT == 2) { if (F == 2) major = s[0].match > s[1].match ? s[0].match : s[1].match; else major = s[0].match; }
    else {
        // sorted[F/2] by rank selection (ties broken by position), small F: O(F^2) compares in registers
#pragma unroll (FT <= 8 ? FT : 1)
        for (int a = 0; a < FT; a++) {
            int rank = 0;
#pragma unroll (FT <= 8 ? FT : 1)
            for (int b = 0; b < FT; b++) if (b < F) rank += (s[b].match < s[a].match) || (s[b].match == s[a].match && b < a);
            if (a < F && rank == F / 2) major = s[a].match;
        }
    }
    if (full > major) return RAFTING_ERR_IMPOSSIBLE_REPL;                    // :251-253
    if (major != 0) {
        int64_t t;
        if (!term_at(g, c, major, t)) { flag_err(g, RAFTING_ERR_TRY_COMMIT_FAILED); return 0; }   // NPE, caught + logged :277
        const int64_t ci = (t == g.term) ? major : full;                     // :257-261
        if (ci != 0 && ci != g.commit) return commit_log(g, ci);            // :262-275
    }
    return 0;
}

// RaftStub.process -> Leader.acceptCommand -> RocksLog.newEntry — RaftStub.java:79-91, Leader.java:128-140, RocksLog.java:82-89
template <int FT>
__device__ __forceinline__ int op_submit(GS& g, const Ctx& c, LS (&s)[FT], int F, uint32_t count, uint64_t unavail, const PlanSink& o) {
    if (role_of(g) != RAFTING_ROLE_LEADER) return RAFTING_ERR_NOT_LEADER;
    if (!leader_ready<FT>(g, s, F, c.cfg->avail_critical_point, c.cfg->recovery_cool_down_ms, c.now)) return RAFTING_ERR_NOT_READY;
    if (count == 0) count = 1;
    const bool has = nruns_of(g) > 0;
    if (!has && g.epochIndex != 0) return RAFTING_ERR_LOG_SHAPE;
    if ((!has || g.r0t != g.term) && nruns_of(g) >= KRUNS) return RAFTING_ERR_TERM_RUNS_OVERFLOW;
    const int64_t index = has ? g.hi + 1 : 1;
    if (has && g.r0t == g.term) g.hi = index + count - 1;
    else { push_run(g, c, index, g.term); g.hi = index + count - 1; }
    return replicate_log<FT>(g, c, s, F, false, unavail, o);
}

// RaftRoutine.keepAlive / electionTimeout + onTimeout — RaftRoutine.java:53-77, Leader.java:119-126,
// Follower.java:156-168,223-279, Candidate.java:82-88
template <int FT>
__device__ __forceinline__ int op_timeout(GS& g, const Ctx& c, RowOut& ro, LS (&s)[FT], int F, uint64_t unavail, const PlanSink& o) {
    if (role_of(g) == RAFTING_ROLE_LEADER) {
        reset_timer(g, c, false, false);
        return replicate_log<FT>(g, c, s, F, true, unavail, o);
    }
    if (!(g.timer > 0)) return 0;
    g.timer = RAFTING_TIMER_TIMEOUT;
    if (role_of(g) == RAFTING_ROLE_FOLLOWER && c.cfg->pre_vote) {
        const int64_t t = g.term;
        int err = switch_to(g, c, ro, RAFTING_ROLE_FOLLOWER, g.term, ballot_of(g));
        if (err) return err;
        if (role_of(g) == RAFTING_ROLE_FOLLOWER && g.term == t) {            // prepareElection
            g.word |= W_TIMEOUT_DET;
            g.votes = 1;
            emit_ballot(g, c, ro, RAFTING_BALLOT_PREVOTE, (int64_t)((uint64_t)g.term + 1u));
        }
        return 0;
    }
    return switch_to(g, c, ro, RAFTING_ROLE_CANDIDATE, (int64_t)((uint64_t)g.term + 1u), (int)c.cfg->local_slot);
}

// one lane event, generic (any kind, any role) — the Async callbacks of Leader.java:174-188,218-237,
// Follower.java:258-270, Candidate.java:112-134
template <int FT>
__device__ __forceinline__ int lane_event(GS& g, Ctx& c, RowOut& ro, LS (&s)[FT], int F, int f, uint64_t m, i64x2 etn, i64x2 eel) {
    const uint32_t ek = RAFTING_EVM_KIND(m), outcome = RAFTING_EVM_OUTCOME(m), inc = RAFTING_EVM_INC(m);
    const bool success = RAFTING_EVM_SUCCESS(m) != 0;
    const int64_t respTerm = etn.x;
    const int self = (int)c.cfg->local_slot;
    c.now = etn.y; c.draw = 0;
    if (ek == RAFTING_EV_AE_ACK || ek == RAFTING_EV_IS_ACK) {
        if (!(role_of(g) == RAFTING_ROLE_LEADER && inc == g.inc && (g.word & W_PREPARED))) return 0;   // dead State object
        LS& x = s[f];
        x.inflight--;
        if (outcome == RAFTING_OUT_OK) {
            if (respTerm > g.term) return switch_to(g, c, ro, RAFTING_ROLE_FOLLOWER, respTerm, lane_to_slot(c, f));
            const bool snap = ek == RAFTING_EV_IS_ACK;
            stat_success(x, c.now, !success);
            int e = update_index(x, eel.x, snap ? eel.x : eel.y, success, snap);
            if (e) return e;
            if (!snap && success) return try_commit<FT>(g, c, s, F);
        } else stat_failure(x, c.now, outcome == RAFTING_OUT_ERROR, false);
        return 0;
    }
    if (ek == RAFTING_EV_PV_REPLY) {
        if (role_of(g) == RAFTING_ROLE_FOLLOWER && inc == g.inc && (g.word & W_TIMEOUT_DET) && outcome == RAFTING_OUT_OK) {
            const int64_t nextTerm = (int64_t)((uint64_t)g.term + 1u);
            if (respTerm > nextTerm) return switch_to(g, c, ro, RAFTING_ROLE_FOLLOWER, respTerm, lane_to_slot(c, f));
            if (success && ++g.votes >= majority(c)) return switch_to(g, c, ro, RAFTING_ROLE_CANDIDATE, nextTerm, self);
        }
        return 0;
    }
    if (ek == RAFTING_EV_RV_REPLY) {
        if (role_of(g) == RAFTING_ROLE_CANDIDATE && inc == g.inc) {
            if (outcome != RAFTING_OUT_OK) return 0;
            if (respTerm > g.term) return switch_to(g, c, ro, RAFTING_ROLE_FOLLOWER, respTerm, lane_to_slot(c, f));
            if (success && ++g.votes >= majority(c)) {
                g.electInc = g.inc; g.electTerm = g.term; g.word &= ~W_ELECT_ABORT;
                return switch_to(g, c, ro, RAFTING_ROLE_LEADER, g.term, self);
            }
            return 0;
        }
        // replies to an elected Candidate keep running after it was fenced (Candidate.java:75-80)
        if (g.electInc != 0 && inc == g.electInc && !(g.word & W_ELECT_ABORT) && outcome == RAFTING_OUT_OK) {
            if (respTerm > g.electTerm) {
                g.word |= W_ELECT_ABORT;
                return switch_to(g, c, ro, RAFTING_ROLE_FOLLOWER, respTerm, lane_to_slot(c, f));
            }
            if (success) return switch_to(g, c, ro, RAFTING_ROLE_LEADER, g.electTerm, self);
        }
        return 0;
    }
    return RAFTING_ERR_BAD_EVENT;
}

// ---------------------------------------------------------------------------------------------
// SLOW PATH (out of line): generic handling of row r's group op / lane events for one group.
// State comes from and goes back to the tables; outputs go straight to the outbox.
// `what`: bit0 = run the op, bit1 = run the lane events from lane f0 on.  Returns the step's dirty bits.
// ---------------------------------------------------------------------------------------------
template <int FT>
__device__ __noinline__ uint32_t slow_row(const KArgs* ka, uint32_t i, uint32_t gid, uint32_t r, uint32_t kind, int64_t sweep,
                                          uint32_t what, uint32_t f0, uint32_t dirty) {
    const Tables& T = ka->T; const InboxD& in = ka->in; const OutboxD& out = ka->out;
    const int F = (int)T.F;
    const size_t gi = (size_t)r * in.n + i, li0 = (size_t)gid * T.F;
    GS g; LS s[FT]; RowOut ro;
    load_hot(T, gid, g); load_cold(T, gid, g); g.dirty = dirty;
    for (int f = 0; f < F; f++) load_lane(T, li0 + f, s[f]);
    ro.bm = 0; ro.bt = 0; ro.bl.x = ro.bl.y = 0;
    const bool alive = (g.word & W_ALIVE) != 0;
    Ctx c = make_ctx(T, ka->cfg, gid, 0, 0);
    if (what & 1u) {
        uint32_t meta = 0, entoff = 0; int64_t a = 0, b = 0, cc = 0, d = 0;
        c.now = sweep; c.draw = 0;
        if (sweep == 0) {
            const uint64_t m = in.op_meta[gi]; meta = (uint32_t)m; entoff = (uint32_t)(m >> 32);
            const i64x2 nr = in.op_nr[gi]; c.now = nr.x; c.draw = nr.y;
            if (in.op_ab) { const i64x2 v = in.op_ab[gi]; a = v.x; b = v.y; }
            if (in.op_cd) { const i64x2 v = in.op_cd[gi]; cc = v.x; d = v.y; }
        }
        PlanSink o; o.pm = nullptr; o.pp = nullptr; o.lc = nullptr; o.pe = nullptr;
        if (out.plan_meta) {
            o.pm = out.plan_meta + gi * T.F; o.pp = out.plan_pp + gi * T.F; o.lc = out.plan_lc + gi * T.F; o.pe = out.plan_epoch + gi * T.F;
            for (int f = 0; f < F; f++) o.pm[f] = 0;
        }
        int err = 0; Reply rep = {0, 0, 0};
        const int peer = (int)RAFTING_OP_PEER(meta); const uint32_t count = RAFTING_OP_COUNT(meta);
        if (!alive) err = RAFTING_ERR_CLOSED_GROUP;
        else if (kind == RAFTING_OP_SUBMIT) err = op_submit<FT>(g, c, s, F, count, (uint64_t)a, o);
        else if (kind == RAFTING_OP_TIMEOUT) err = op_timeout<FT>(g, c, ro, s, F, (uint64_t)a, o);
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
        if (out.rep_meta) {
            out.rep_meta[gi] = (uint32_t)(rep.valid ? 1 : 0) | ((uint32_t)(rep.success ? 1 : 0) << 1) | ((uint32_t)err << 8);
            if (rep.valid) out.rep_term[gi] = rep.term;
        }
    }
    if ((what & 2u) && alive) {
        for (int f = (int)f0; f < F; f++) {
            const size_t li = gi * T.F + f;
            const uint64_t m = in.ev_meta[li];
            if (RAFTING_EVM_KIND(m) == RAFTING_EV_NONE) continue;
            i64x2 eel = {0, 0}; const i64x2 etn = in.ev_tn[li]; if (in.ev_el) eel = in.ev_el[li];
            const int err = lane_event<FT>(g, c, ro, s, F, f, m, etn, eel);
            if (err) flag_err(g, err);
        }
    }
    if (out.ballot_meta && ro.bm != 0) { out.ballot_meta[gi] = ro.bm; out.ballot_term[gi] = ro.bt; out.ballot_last[gi] = ro.bl; }
    store_hot(T, gid, g); store_warm(T, gid, g); store_cold(T, gid, g);
    for (int f = 0; f < F; f++) store_lane(T, li0 + f, s[f]);
    return g.dirty;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
#ifndef RAFTING_MINBLOCKS
#define RAFTING_MINBLOCKS 8
#endif
#ifndef RAFTING_TPB
#define RAFTING_TPB 64
#endif
constexpr int TPB = RAFTING_TPB;         // threads (= groups) per block: 64K groups / 64 = 1024 blocks = 6.9 per SM (balanced)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int FT>
struct __align__(16) Stage {             // one staged input row of this block
    i64x2    op_nr[TPB];
    i64x2    op_ab[TPB];
    i64x2    ev_tn[TPB * FT];
    i64x2    ev_el[TPB * FT];
    uint64_t op_meta[TPB];
    uint64_t ev_meta[TPB * FT];
};

// FT follower slots kept in registers (FT >= F).  NST input rows are staged in (dynamic) shared memory,
// NST-1 of them in flight per thread; NST == 0 drops the ring (large FT, where it would not fit) and
// reads the inbox directly.
template <int FT, int NST>
__global__ void __launch_bounds__(TPB, (FT <= 2 ? RAFTING_MINBLOCKS : 1))
step_kernel(Tables T, InboxD in, OutboxD out, const CfgD* __restrict__ cfgp, CfgD cfg) {
    constexpr bool STAGED = NST > 0;
    __shared__ KArgs ka;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Stage<FT>* stage = reinterpret_cast<Stage<FT>*>(smem_raw);
    if (threadIdx.x == 0) { ka.T = T; ka.in = in; ka.out = out; ka.cfg = cfgp; }
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= in.n) return;
    const int F = (int)T.F;
    const uint32_t gid = in.gids ? in.gids[i] : i;
    if (gid >= T.G) return;
    const uint32_t tl = threadIdx.x;

    // ---- the group's hot scalars and every follower slot live in registers for the whole batch ----
    GS g; LS s[FT];
    load_hot(T, gid, g); g.dirty = 0; g.electTerm = 0; g.electInc = 0; g.votes = 0;
    const size_t li0 = (size_t)gid * T.F;
#pragma unroll (FT <= 8 ? FT : 1)
    for (int f = 0; f < FT; f++) {
        s[f].next = s[f].match = s[f].lastEpoch = s[f].reqSucc = s[f].reqFail = s[f].lastReq = 0;
        s[f].inflight = s[f].rej = s[f].fail = s[f].pending = 0;
        if (f < F) load_lane(T, li0 + f, s[f]);
    }
    const bool hasOps = in.op_meta != nullptr, hasEv = in.ev_meta != nullptr;

#define RAFTING_ISSUE(R_)                                                                                  \
    if (STAGED) {                                                                                          \
        const uint32_t r_ = (R_);                                                                          \
        if (r_ < in.rows) {                                                                                \
            auto& st_ = stage[r_ % (NST > 0 ? NST : 1)];                                                                   \
            const uint32_t gi_ = r_ * in.n + i;                                                            \
            if (hasOps) {                                                                                  \
                cp_async8(&st_.op_meta[tl], in.op_meta + gi_); cp_async16(&st_.op_nr[tl], in.op_nr + gi_);  \
                if (in.op_ab) cp_async16(&st_.op_ab[tl], in.op_ab + gi_);                                  \
            }                                                                                              \
            if (hasEv) {                                                                                   \
                _Pragma("unroll 8") for (int f_ = 0; f_ < FT; f_++) if (f_ < F) {                            \
                    const uint32_t li_ = gi_ * T.F + f_;                                                   \
                    cp_async8(&st_.ev_meta[tl * FT + f_], in.ev_meta + li_);                               \
                    cp_async16(&st_.ev_tn[tl * FT + f_], in.ev_tn + li_);                                  \
                    if (in.ev_el) cp_async16(&st_.ev_el[tl * FT + f_], in.ev_el + li_);                    \
                }                                                                                          \
            }                                                                                              \
        }                                                                                                  \
        cp_async_commit();                                                                                 \
    }
#pragma unroll (FT <= 8 ? FT : 1)
    for (int p = 0; p < NST - 1; p++) RAFTING_ISSUE((uint32_t)p);

    for (uint32_t r = 0; r < in.rows; r++) {
        const uint32_t gi = r * in.n + i;                    // rows * n * F < 2^32 is checked by the host
        RAFTING_ISSUE(r + (uint32_t)(NST > 0 ? NST - 1 : 0));
        if (STAGED) cp_async_wait<(NST > 0 ? NST - 1 : 0)>();  // row r has landed (this thread's own copies)
        const auto& st = stage[STAGED ? r % (NST > 0 ? NST : 1) : 0];
        const bool alive = (g.word & W_ALIVE) != 0;

        // ================= group op =================
        const int64_t sweep = in.row_now ? in.row_now[r] : 0;
        uint32_t meta = 0, kind = RAFTING_OP_NONE; int64_t now = 0; uint64_t unavail = 0;
        bool slowOp = false;
        if (sweep != 0) {
            const bool due = alive && ((role_of(g) == RAFTING_ROLE_LEADER) ? (g.timer <= sweep)
                                       : (g.timer > 0 && g.timer != I64MAX && g.timer <= sweep));
            if (due) { kind = RAFTING_OP_TIMEOUT; slowOp = true; }
        } else if (hasOps) {
            if (STAGED) { meta = (uint32_t)st.op_meta[tl]; now = st.op_nr[tl].x; if (in.op_ab) unavail = (uint64_t)st.op_ab[tl].x; }
            else { meta = (uint32_t)in.op_meta[gi]; now = in.op_nr[gi].x; if (in.op_ab) unavail = (uint64_t)in.op_ab[gi].x; }
            kind = RAFTING_OP_KIND(meta);
            const bool fastOk = alive && role_of(g) == RAFTING_ROLE_LEADER && (g.word & W_PREPARED) &&
                                (kind == RAFTING_OP_TIMEOUT || (kind == RAFTING_OP_SUBMIT && nruns_of(g) > 0 && g.r0t == g.term));
            slowOp = kind != RAFTING_OP_NONE && !fastOk;
        }
        PlanSink o; o.pm = nullptr; o.pp = nullptr; o.lc = nullptr; o.pe = nullptr;
        if (out.plan_meta) { const uint32_t pl = gi * T.F; o.pm = out.plan_meta + pl; o.pp = out.plan_pp + pl; o.lc = out.plan_lc + pl; o.pe = out.plan_epoch + pl; }
        uint32_t repMeta = 0;
        if (kind != RAFTING_OP_NONE && !slowOp) {
            // Leader keepAlive -> replicateLog(true) (RaftRoutine.java:53-62, Leader.java:119-126), or
            // RaftStub.process -> Leader.acceptCommand -> replicateLog(false) with no new term run
            // (RaftStub.java:79-91, Leader.java:128-140, RocksLog.java:82-89)
            Ctx c = make_ctx(T, cfgp, gid, now, 0);
            int err = 0; bool go = true;
            if (kind == RAFTING_OP_TIMEOUT) g.timer = (I64MAX - cfg.heartbeat_ms < now) ? I64MAX : now + cfg.heartbeat_ms;   // resetTimer, Leader branch
            else if (!leader_ready<FT>(g, s, F, cfg.avail_critical_point, cfg.recovery_cool_down_ms, now)) { err = RAFTING_ERR_NOT_READY; go = false; }
            else { uint32_t count = RAFTING_OP_COUNT(meta); if (count == 0) count = 1; g.hi += count; }
            if (go) err = replicate_log<FT>(g, c, s, F, kind == RAFTING_OP_TIMEOUT, unavail, o);
            else if (o.pm) {
#pragma unroll (FT <= 8 ? FT : 1)
                for (int f = 0; f < FT; f++) if (f < F) o.pm[f] = 0;
            }
            if (err) flag_err(g, err);
            repMeta = (uint32_t)err << 8;
        } else if (!slowOp && o.pm) {
#pragma unroll (FT <= 8 ? FT : 1)
            for (int f = 0; f < FT; f++) if (f < F) o.pm[f] = 0;
        }

        // ================= lane events: classify =================
        bool anyEv = false, slowEv = false;
        if (hasEv && alive) {
#pragma unroll (FT <= 8 ? FT : 1)
            for (int f = 0; f < FT; f++) {
                if (f >= F) break;
                const uint64_t em = STAGED ? st.ev_meta[tl * FT + f] : in.ev_meta[gi * T.F + f];
                const uint32_t ek = RAFTING_EVM_KIND(em);
                anyEv |= ek != RAFTING_EV_NONE;
                slowEv |= ek > RAFTING_EV_IS_ACK;             // vote replies / unknown kinds
            }
        }

        if (slowOp || slowEv) {
            // ---- hand the state over through the tables and run the generic handlers out of line ----
            if (!slowOp && out.rep_meta) out.rep_meta[gi] = repMeta;      // the op (if any) already ran inline
            if (out.ballot_meta) out.ballot_meta[gi] = 0;
            store_hot(T, gid, g);
#pragma unroll (FT <= 8 ? FT : 1)
            for (int f = 0; f < FT; f++) if (f < F) store_lane(T, li0 + f, s[f]);
            const uint32_t dirty = slow_row<FT>(&ka, i, gid, r, kind, sweep, (slowOp ? 1u : 0u) | (anyEv ? 2u : 0u), 0u, g.dirty);
            load_hot(T, gid, g); g.dirty = dirty;
#pragma unroll (FT <= 8 ? FT : 1)
            for (int f = 0; f < FT; f++) if (f < F) load_lane(T, li0 + f, s[f]);
            continue;
        }
        if (out.rep_meta) out.rep_meta[gi] = repMeta;
        if (out.ballot_meta) out.ballot_meta[gi] = 0;

        // ================= lane events: AE-Echo / IS-Echo in lane order (Leader.java:174-188,218-237) =================
        if (anyEv) {
            const bool leaderLive = role_of(g) == RAFTING_ROLE_LEADER && (g.word & W_PREPARED);
            int bailAt = -1;
#pragma unroll (FT <= 8 ? FT : 1)
            for (int f = 0; f < FT; f++) {
                if (f >= F) break;
                const uint64_t em = STAGED ? st.ev_meta[tl * FT + f] : in.ev_meta[gi * T.F + f];
                if (RAFTING_EVM_KIND(em) == RAFTING_EV_NONE || bailAt >= 0) continue;
                if (!(leaderLive && RAFTING_EVM_INC(em) == g.inc)) continue;      // addressed to a dead Leadership.State: dropped
                const i64x2 etn = STAGED ? st.ev_tn[tl * FT + f] : in.ev_tn[gi * T.F + f];
                i64x2 eel = {0, 0};
                if (in.ev_el) eel = STAGED ? st.ev_el[tl * FT + f] : in.ev_el[gi * T.F + f];
                const bool ok = RAFTING_EVM_OUTCOME(em) == RAFTING_OUT_OK;
                const bool snap = RAFTING_EVM_KIND(em) == RAFTING_EV_IS_ACK;
                if (ok && (etn.x > g.term || (snap ? eel.x : eel.y) < s[f].match)) {
                    // step-down (Leader.java:178-180,224-226) or "match index should not rollback"
                    // (Leadership.java:76-81): the generic handler finishes this row from lane f on
                    bailAt = f;
                    continue;
                }
                LS& x = s[f];
                x.inflight--;
                if (ok) {
                    const bool success = RAFTING_EVM_SUCCESS(em) != 0;
                    stat_success(x, etn.y, !success);
                    update_index(x, eel.x, snap ? eel.x : eel.y, success, snap);
                    if (!snap && success) {
                        Ctx c = make_ctx(T, cfgp, gid, etn.y, 0);
                        const int cerr = try_commit<FT>(g, c, s, F);
                        if (cerr) flag_err(g, cerr);
                    }
                } else stat_failure(x, etn.y, RAFTING_EVM_OUTCOME(em) == RAFTING_OUT_ERROR, false);
            }
            if (bailAt >= 0) {
                store_hot(T, gid, g);
#pragma unroll (FT <= 8 ? FT : 1)
                for (int f = 0; f < FT; f++) if (f < F) store_lane(T, li0 + f, s[f]);
                const uint32_t dirty = slow_row<FT>(&ka, i, gid, r, 0u, 0, 2u, (uint32_t)bailAt, g.dirty);
                load_hot(T, gid, g); g.dirty = dirty;
#pragma unroll (FT <= 8 ? FT : 1)
                for (int f = 0; f < FT; f++) if (f < F) load_lane(T, li0 + f, s[f]);
            }
        }
    }
    if (STAGED) cp_async_wait<0>();

    // ---- write back: the columns the fast path can change; the slow path stored the rest itself ----
#pragma unroll (FT <= 8 ? FT : 1)
    for (int f = 0; f < FT; f++) if (f < F) store_lane(T, li0 + f, s[f]);
    store_hot(T, gid, g);
    if (out.commit_index) out.commit_index[gid] = g.commit;
    if (out.current_term) out.current_term[gid] = g.term;
    if (out.role_word)
        out.role_word[gid] = (uint32_t)role_of(g) | ((uint32_t)(ballot_of(g) + 1) << 8) | ((uint32_t)(leader_of(g) + 1) << 16) |
                             ((g.word & W_TIMEOUT_DET) ? (1u << 24) : 0u) | ((g.word & W_READY) ? (1u << 25) : 0u) |
                             ((g.dirty & 1u) << 30) | (((g.dirty >> 1) & 1u) << 31);
    if (out.incarnation) out.incarnation[gid] = g.inc;
    if (out.err_word) out.err_word[gid] = g.err;
    if (out.last_entry) { i64x2 v; last_or_epoch(g, v.x, v.y); out.last_entry[gid] = v; }
}

}  // namespace rafting
