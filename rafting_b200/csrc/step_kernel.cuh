// step_kernel.cuh — the batched "event loop" kernel: one launch drains one batch for every group.
//
// Replaces, for all groups at once, the reference's per-context EventLoop path
// (M/support/EventLoop.java:87-101 + the RaftParticipant handlers and Async callbacks it runs):
//   K1 ack_quorum_commit : AE-Echo / IS-Echo + Leadership.State.{statSuccess,statFailure,updateIndex,
//                          majorIndices} + Leader.tryCommit + RocksLog.markCommitted
//                          (Leader.java:174-188,218-237,247-280; Leadership.java:53-130; RocksLog.java:100-109)
//   K2 vote_tally        : RV-Echo / PV-Echo + Membership.isBetter (Candidate.java:112-134,
//                          Follower.java:249-270, Membership.java:74-108)
//   K3 ae_request_check  : *.appendEntries, logContains, purgeEntries, RocksLog.conflict/truncate/append
//                          (Follower.java:35-88,177-221; Candidate.java:28-41; Leader.java:66-86; RocksLog.java:169-225)
//   K4 vote_request_check: *.preVote / requestVote / installSnapshot, logUpToDate
//                          (Follower.java:91-153,193-207; Candidate.java:43-72; Leader.java:88-111; RaftMember.java:61-66)
//   K5 timer_sweep       : RaftRoutine.resetTimer / electionTimeout / keepAlive (RaftRoutine.java:53-130)
//   K6 replicate_plan    : Leader.prepareReplication / replicateLog / isReady (Leader.java:30-64,142-245)
// fused into one kernel because they share the same per-group state and the same serial order.
//
// Mapping: one sub-warp of W lanes per group (W = pow2 >= F = R-1).  Lane f owns Leadership.State of
// follower f in registers for the whole batch; every lane carries an identical copy of the group
// scalars.  Lane events are loaded coalesced (one per lane), broadcast inside the sub-warp with
// shuffles and applied in lane order, which is the canonical serial order.  The quorum index is a
// rank-select over the sub-warp's matchIndex registers (shuffles), votes/readiness are ballots.
// Pure integer work: the roofline is HBM bandwidth, no tensor cores.
#pragma once
#include "tables.cuh"

namespace rafting {

constexpr int64_t I64MAX = INT64_MAX;
constexpr int     KRUNS  = RAFTING_TERM_RUNS;

struct GS {                       // group scalars, replicated in every lane of the sub-warp
    uint32_t word, inc, err, dirty;          // dirty: bit0 persist, bit1 commit
    int64_t  term, commit, lo, hi, timer, epochIndex, epochTerm, electTerm;
    uint32_t electInc; int32_t votes;
    int64_t  r0s, r0t;                       // newest term run (start, term)
};
struct LS {                       // Leadership.State of this lane's follower
    int64_t next, match, lastEpoch, reqSucc, reqFail, lastReq;
    int32_t inflight, rej, fail, pending;
};
struct RowOut {                   // outputs of the current row, written once at row end
    uint64_t pm; i64x2 pp, lc; int64_t pe;   // per lane
    uint64_t bm; int64_t bt; i64x2 bl;       // per group
};
struct Ctx {
    const CfgD* cfg;
    i64x2*   runs;                // &g_runs[gid], stride G
    uint32_t gid, F, G;
    unsigned mask;                // member mask of this sub-warp
    int      lane, sub0;          // lane inside the sub-warp, first warp-lane of the sub-warp
    bool     lv;                  // lane < F
    int64_t  now, draw;
    RowOut*  ro;
};
struct Reply { int valid, success; int64_t term; };

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t shfl64(unsigned mask, int64_t v, int src, int W) {
    int lo = __shfl_sync(mask, (int)(uint32_t)(uint64_t)v, src, W);
    int hi = __shfl_sync(mask, (int)(uint32_t)((uint64_t)v >> 32), src, W);
    return (int64_t)(((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo);
}
__device__ __forceinline__ int role_of(const GS& g) { return (int)(g.word & W_ROLE_MASK); }
__device__ __forceinline__ int ballot_of(const GS& g) { return (int)((g.word >> W_BALLOT_SH) & 0xff) - 1; }
__device__ __forceinline__ int leader_of(const GS& g) { return (int)((g.word >> W_LEADER_SH) & 0xff) - 1; }
__device__ __forceinline__ int nruns_of(const GS& g) { return (int)((g.word >> W_NRUNS_SH) & 0xf); }
__device__ __forceinline__ void set_nruns(GS& g, int n) {
    g.word = (g.word & ~(0xfu << W_NRUNS_SH)) | ((uint32_t)n << W_NRUNS_SH);
}
__device__ __forceinline__ void set_leader(GS& g, int slot) {
    g.word = (g.word & ~(0xffu << W_LEADER_SH)) | ((uint32_t)(slot + 1) << W_LEADER_SH);
}
__device__ __forceinline__ int lane_to_slot(const Ctx& c, int f) { return f < (int)c.cfg->local_slot ? f : f + 1; }
__device__ __forceinline__ int majority(const Ctx& c) { return (int)c.cfg->replicas / 2 + 1; }   // RaftContext.java:170
__device__ __forceinline__ void flag_err(GS& g, int code) {
    uint32_t cnt = (g.err >> 16) + 1; if (cnt > 0xffffu) cnt = 0xffffu;
    g.err = (cnt << 16) | (uint32_t)code;
}
__device__ __forceinline__ unsigned sub_ballot(const Ctx& c, bool p, int W) {
    unsigned b = __ballot_sync(c.mask, p);
    return W == 32 ? b : ((b >> c.sub0) & ((1u << W) - 1u));
}

// ---------------------------------------------------------------------------------------------
// run-length term table (replaces RocksLog.get(i).term(), RocksLog.java:122-128)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool term_at(const GS& g, const Ctx& c, int64_t idx, int64_t& t) {
    const int n = nruns_of(g);
    if (n == 0 || idx < g.lo || idx > g.hi) return false;
    if (idx >= g.r0s) { t = g.r0t; return true; }
    bool found = false;
    for (int k = 1; k < n; k++) {                    // n is group-uniform; idx may differ per lane
        i64x2 r = c.runs[(size_t)k * c.G];
        if (!found && idx >= r.x) { t = r.y; found = true; }
    }
    return found;
}
__device__ __noinline__ void push_run(GS& g, const Ctx& c, int64_t start, int64_t term) {
    const int n = nruns_of(g);
    if (c.lane == 0 && n >= 1) {
        for (int k = n - 1; k >= 1; k--) c.runs[(size_t)(k + 1) * c.G] = c.runs[(size_t)k * c.G];
        i64x2 r; r.x = g.r0s; r.y = g.r0t;
        c.runs[(size_t)1 * c.G] = r;
    }
    __syncwarp(c.mask);
    g.r0s = start; g.r0t = term;
    if (n == 0) g.lo = start;
    set_nruns(g, n + 1);
}
__device__ __noinline__ void pop_run(GS& g, const Ctx& c) {
    const int n = nruns_of(g);
    if (n <= 1) { set_nruns(g, 0); return; }
    i64x2 r = c.runs[(size_t)1 * c.G];
    __syncwarp(c.mask);
    if (c.lane == 0)
        for (int k = 1; k <= n - 2; k++) c.runs[(size_t)k * c.G] = c.runs[(size_t)(k + 1) * c.G];
    __syncwarp(c.mask);
    g.r0s = r.x; g.r0t = r.y;
    set_nruns(g, n - 1);
}
// RocksLog.truncate — RocksLog.java:219-225
__device__ __forceinline__ void log_truncate(GS& g, const Ctx& c, int64_t index) {
    if (nruns_of(g) > 0 && g.hi >= index) {
        g.hi = index - 1;
        while (nruns_of(g) > 0 && g.r0s > g.hi) pop_run(g, c);
    }
}
// RocksLog.flush — RocksLog.java:228-242 (deleteRange end-exclusive: the entry at `index` survives)
__device__ __noinline__ int log_flush(GS& g, const Ctx& c, int64_t index, int64_t term) {
    if (index < g.epochIndex) return RAFTING_ERR_FLUSH_RANGE;
    const int n = nruns_of(g);
    if (n > 0) {
        if (index > g.hi) set_nruns(g, 0);
        else if (index > g.lo) {
            if (g.r0s <= index) { g.r0s = index; set_nruns(g, 1); }
            else {
                int keep = n;
                for (int k = 1; k < n; k++) {
                    i64x2 r = c.runs[(size_t)k * c.G];
                    if (keep == n && r.x <= index) {
                        keep = k + 1;
                        if (c.lane == 0) { r.x = index; c.runs[(size_t)k * c.G] = r; }
                    }
                }
                __syncwarp(c.mask);
                set_nruns(g, keep);
            }
            g.lo = index;
        }
    }
    g.epochIndex = index; g.epochTerm = term;
    return 0;
}
__device__ __forceinline__ void log_append_one(GS& g, const Ctx& c, int64_t idx, int64_t term) {
    if (nruns_of(g) > 0 && term == g.r0t) g.hi = idx;
    else { push_run(g, c, idx, term); g.hi = idx; }
}
__device__ __forceinline__ void last_or_epoch(const GS& g, int64_t& idx, int64_t& term) {
    if (nruns_of(g) > 0) { idx = g.hi; term = g.r0t; } else { idx = g.epochIndex; term = g.epochTerm; }
}

// ---------------------------------------------------------------------------------------------
// Leadership.State methods — Leadership.java:40-114
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void stat_success(LS& s, int64_t now, bool reject) {           // :53-63
    if (now > s.reqSucc) s.reqSucc = now;
    s.fail = 0;
    s.rej = reject ? (int32_t)((uint32_t)s.rej + 1u) : 0;
}
__device__ __forceinline__ void stat_failure(LS& s, int64_t now, bool unreachable, bool reject) {   // :65-73
    if (now > s.reqFail) s.reqFail = now;
    if (unreachable) s.fail = (int32_t)((uint32_t)s.fail + 1u);
    if (reject) s.rej = (int32_t)((uint32_t)s.rej + 1u);
}
__device__ __forceinline__ bool state_ready(const LS& s, int32_t crit, int64_t cool, int64_t now) {   // :44-51
    bool unhealthy = (crit > 0 && (uint32_t)s.fail > (uint32_t)crit) ||
                     (cool > 0 && (int64_t)((uint64_t)now - (uint64_t)s.reqFail) < cool);
    return s.reqSucc != 0 && !(s.pending || unhealthy);
}
// round(ln(e + r)) as an integer threshold table (Leadership.java:105); verified against libm by
// tests/test_backoff_table.py for every boundary and every r < 2^16
__device__ __forceinline__ int64_t backoff_step(int32_t r) {
    if (r < 0) return 0;   // unreachable: ln of a negative argument is NaN, Math.round(NaN) == 0
    if (r <= 1) return 1;        if (r <= 9) return 2;         if (r <= 30) return 3;
    if (r <= 87) return 4;       if (r <= 241) return 5;       if (r <= 662) return 6;
    if (r <= 1805) return 7;     if (r <= 4912) return 8;      if (r <= 13357) return 9;
    if (r <= 36312) return 10;   if (r <= 98713) return 11;    if (r <= 268334) return 12;
    if (r <= 729413) return 13;  if (r <= 1982756) return 14;  if (r <= 5389695) return 15;
    if (r <= 14650716) return 16; if (r <= 39824781) return 17; if (r <= 108254985) return 18;
    if (r <= 294267563) return 19; if (r <= 799902174) return 20; if (r <= 2147483647) return 21;
    return 21;
}
__device__ __forceinline__ int update_index(LS& s, int64_t epoch, int64_t index, bool success, bool snapshot) {  // :75-114
    if (index < s.match) return RAFTING_ERR_MATCH_ROLLBACK;
    if (epoch < s.lastEpoch) return 0;
    if (epoch > s.lastEpoch) { s.lastEpoch = epoch; s.next = s.next > epoch ? s.next : epoch; }
    if ((s.pending != 0) != snapshot) return 0;
    const int64_t e1 = (int64_t)((uint64_t)epoch + 1u);
    if (s.pending) {
        if (success) { s.next = s.next > e1 ? s.next : e1; s.pending = 0; }
    } else {
        if (success) {
            if (index > s.match) { s.next = (int64_t)((uint64_t)index + 1u); s.match = index; }
        } else if (s.match == 0) {
            int64_t step = backoff_step(s.rej);
            int64_t a = (int64_t)((uint64_t)s.next - (uint64_t)step);
            int64_t nx = a > e1 ? a : e1;
            int64_t b = (int64_t)((uint64_t)s.next - 1u);
            s.next = b < nx ? b : nx;
        }
    }
    if (s.next <= epoch && !s.pending) s.pending = 1;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Membership.isBetter — Membership.java:74-108 (the filter is never null once a group is open)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int is_better(int nr, int64_t nt, int nb, int cr, int64_t ct, int cb) {
    if (nt != ct) return nt > ct;
    if (nr != cr) {
        if (nr == RAFTING_ROLE_LEADER) return cr == RAFTING_ROLE_CANDIDATE ? 1 : -RAFTING_ERR_LEADER_UNCHANGED;
        return nr == RAFTING_ROLE_FOLLOWER;
    }
    if (nr == RAFTING_ROLE_LEADER) return 0;
    if (nr == RAFTING_ROLE_FOLLOWER) return 1;
    if (nb != cb) return -RAFTING_ERR_BALLOT_MISMATCH;
    return 0;
}

// RaftRoutine.resetTimer — RaftRoutine.java:86-130.  g.timer: deadline (non-leader) / next keepAlive (leader)
__device__ __forceinline__ bool reset_timer(GS& g, const Ctx& c, bool muted, bool ticketNull) {
    const bool leader = role_of(g) == RAFTING_ROLE_LEADER;
    const int64_t moment = ticketNull ? 0 : (leader ? I64MAX : g.timer);
    if (!ticketNull && moment < 0) return false;
    int64_t draw = c.draw != 0 ? c.draw : rafting_draw(c.cfg->timer_seed, c.gid, g.inc, c.cfg->election_ms);
    const int64_t timeout = leader ? c.cfg->heartbeat_ms : (muted ? I64MAX : draw);
    const int64_t a = (moment == I64MAX) ? 0 : ((moment < I64MAX - 1 ? moment : I64MAX - 1) + 1);
    const int64_t b = (I64MAX - timeout < c.now) ? I64MAX : c.now + timeout;
    if (leader) {
        const int64_t delay = ticketNull ? 0 : timeout;
        g.timer = (I64MAX - delay < c.now) ? I64MAX : c.now + delay;
    } else {
        g.timer = a > b ? a : b;
    }
    return true;
}

__device__ __forceinline__ void emit_ballot(const GS& g, const Ctx& c, int kind, int64_t term) {
    int64_t li, lt; last_or_epoch(g, li, lt);
    c.ro->bm = (uint64_t)kind | ((uint64_t)g.inc << 32);
    c.ro->bt = term; c.ro->bl.x = li; c.ro->bl.y = lt;
}

// RaftRoutine.trySwitch + switchTo + convertTo — RaftRoutine.java:140-216
__device__ __forceinline__ int switch_to(GS& g, const Ctx& c, int role, int64_t term, int ballot) {
    int better = is_better(role, term, ballot, role_of(g), g.term, ballot_of(g));
    if (better < 0) return -better;
    if (!better) return 0;
    // convertTo: the old ticket is fenced and dropped; the old role's AsyncHead dies with its incarnation
    g.word = (g.word & ~(W_ROLE_MASK | W_TIMEOUT_DET | W_PREPARED | (0xffu << W_BALLOT_SH) | (0xffu << W_LEADER_SH)))
             | (uint32_t)role | ((uint32_t)(ballot + 1) << W_BALLOT_SH);
    g.term = term;
    g.inc++;
    g.dirty |= 1u;                                           // RaftMember ctor persists (term, lastCandidate)
    g.votes = 0;
    if (role == RAFTING_ROLE_CANDIDATE) {                    // Candidate ctor -> startElection (Candidate.java:22-25,90-143)
        g.votes = 1;
        emit_ballot(g, c, RAFTING_BALLOT_VOTE, term);
    }
    reset_timer(g, c, false, true);
    return 0;
}

// RaftContext.commitLog + RocksLog.markCommitted — RaftContext.java:244-255, RocksLog.java:100-109
__device__ __forceinline__ int commit_log(GS& g, int64_t ci) {
    if (ci < g.commit) return RAFTING_ERR_COMMIT_ROLLBACK;
    if (ci > g.commit) { g.commit = ci; g.dirty |= 2u; }
    return 0;
}

// Leader.tryCommit + Leadership.State.majorIndices — Leader.java:247-280, Leadership.java:116-130.
// Rank-select over the sub-warp's matchIndex registers: the element of rank F/2 (ties broken by
// lane) is sorted[F/2]; the minimum is sorted[0].
template <int W>
__device__ __forceinline__ int try_commit(GS& g, const Ctx& c, int64_t myMatch) {
    const int F = (int)c.F;
    int64_t full = I64MAX; int rank = 0;
#pragma unroll
    for (int j = 0; j < W; j++) {
        int64_t mj = shfl64(c.mask, myMatch, j, W);
        if (j < F) {
            full = mj < full ? mj : full;
            rank += (mj < myMatch) || (mj == myMatch && j < c.lane);
        }
    }
    unsigned sel = sub_ballot(c, c.lv && rank == F / 2, W);
    int64_t major = shfl64(c.mask, myMatch, __ffs(sel) - 1, W);
    if (full > major) return RAFTING_ERR_IMPOSSIBLE_REPL;
    if (major != 0) {
        int64_t t;
        if (!term_at(g, c, major, t)) { flag_err(g, RAFTING_ERR_TRY_COMMIT_FAILED); return 0; }
        int64_t ci = (t == g.term) ? major : full;
        if (ci != 0 && ci != g.commit) return commit_log(g, ci);
    }
    return 0;
}

// Leader.prepareReplication + replicateLog — Leader.java:30-50,142-245.  Lane-parallel: each lane
// plans its own follower; the (unreachable under the store invariants) AssertionErrors of
// RocksLog.batch abort the followers after the failing one, as the sequential loop would.
template <int W>
__device__ __forceinline__ int replicate_log(GS& g, const Ctx& c, LS& s, bool heartbeat, uint64_t unavail) {
    if (!(g.word & W_PREPARED)) {
        int64_t li, lt; last_or_epoch(g, li, lt);
        s.next = (int64_t)((uint64_t)li + 1u); s.match = 0; s.lastEpoch = g.epochIndex;
        s.reqSucc = 0; s.reqFail = 0; s.lastReq = 0; s.inflight = 0; s.rej = 0; s.fail = 0; s.pending = 0;
        g.word |= W_PREPARED;
    }
    const int64_t epochIndex = g.epochIndex, epochTerm = g.epochTerm, leaderCommit = g.commit, now = c.now;
    const uint64_t hb = heartbeat ? (1ull << 4) : 0ull;
    const uint64_t incBits = (uint64_t)g.inc << 32;
    // tentative per-lane result
    int e = 0; uint64_t pm = 0; i64x2 pp = {0, 0}, lc = {0, 0};
    int dInflight = 0; bool fail = false;
    if (c.lv) {
        if ((unavail >> c.lane) & 1ull) { fail = true; pm = RAFTING_PLAN_UNAVAILABLE | hb | incBits; }
        else if (s.inflight > RAFTING_IN_FLIGHT_LIMIT / (heartbeat ? 10 : 1)) pm = RAFTING_PLAN_SKIP_INFLIGHT | hb | incBits;
        else if (s.pending) {
            pm = RAFTING_PLAN_IS | hb | incBits; pp.x = epochIndex; pp.y = epochTerm; lc.x = epochIndex; lc.y = leaderCommit;
            dInflight = 1;
        } else {
            int64_t prevTerm = epochTerm, prevIndex = epochIndex, lastIndex;
            const int64_t nm1 = (int64_t)((uint64_t)s.next - 1u);
            const int64_t nextIndex = nm1 > epochIndex ? nm1 : epochIndex;
            const int fetch = RAFTING_REPLICATE_LIMIT >> (heartbeat ? 1 : 0);
            int64_t idx = nextIndex, len = fetch + 1;
            if (idx == epochIndex) { idx++; len--; }                         // RocksLog.java:134-137
            int64_t eFirst = 0, eCount = 0;
            if (len > 0 && nruns_of(g) > 0) {
                const int64_t hiKey = idx + len - 1;
                if (idx < g.lo && g.lo <= hiKey) e = RAFTING_ERR_LOG_VACANCY; // RocksLog.java:161-163
                else {
                    const int64_t a = idx > g.lo ? idx : g.lo, b = hiKey < g.hi ? hiKey : g.hi;
                    if (a <= b) { eFirst = a; eCount = b - a + 1; }
                }
            }
            if (!e) {
                uint32_t count = 0;
                if (eCount > 0) {
                    if (eFirst == nextIndex) {                               // Leader.java:198-201
                        int64_t t = 0; term_at(g, c, eFirst, t);
                        prevTerm = t; prevIndex = eFirst; eFirst++; eCount--;
                    } else if (eFirst != epochIndex + 1) e = RAFTING_ERR_LOG_START;   // :202-204
                    lastIndex = (eCount == 0) ? prevIndex : eFirst + eCount - 1;
                    count = (uint32_t)eCount;
                } else lastIndex = epochIndex;                               // :210-212
                if (!e) {
                    pm = RAFTING_PLAN_AE | hb | ((uint64_t)count << 16) | incBits;
                    pp.x = prevIndex; pp.y = prevTerm; lc.x = lastIndex; lc.y = leaderCommit;
                    dInflight = 1;
                }
            }
        }
    }
    const unsigned errs = sub_ballot(c, c.lv && e != 0, W);
    const int fe = errs ? __ffs(errs) - 1 : W;                               // first failing follower
    if (c.lv && c.lane <= fe) {
        if (now > s.lastReq) s.lastReq = now;                                // Leader.java:158
        if (c.lane < fe) {
            if (fail) stat_failure(s, now, true, false);                     // :241-243
            s.inflight += dInflight;
            c.ro->pm = pm; c.ro->pp = pp; c.ro->lc = lc; c.ro->pe = epochIndex;
        }
    }
    if (errs) return __shfl_sync(c.mask, e, fe, W);
    return 0;
}

template <int W>
__device__ __forceinline__ bool leader_ready(GS& g, const Ctx& c, const LS& s) {          // Leader.java:52-64
    bool r = c.lv && state_ready(s, c.cfg->avail_critical_point, c.cfg->recovery_cool_down_ms, c.now);
    int cnt = __popc(sub_ballot(c, r, W));
    // the Java loop only returns true from inside `isReady(..) && ++ready > half`: at least one follower must be ready
    bool ready = (g.word & W_PREPARED) && cnt >= 1 && (1 + cnt > (int)c.F / 2);
    g.word = ready ? (g.word | W_READY) : (g.word & ~W_READY);
    return ready;
}

// RaftStub.process -> Leader.acceptCommand -> RocksLog.newEntry — RaftStub.java:79-91, Leader.java:128-140, RocksLog.java:82-89
template <int W>
__device__ __forceinline__ int op_submit(GS& g, const Ctx& c, LS& s, uint32_t count, uint64_t unavail) {
    if (role_of(g) != RAFTING_ROLE_LEADER) return RAFTING_ERR_NOT_LEADER;
    if (!leader_ready<W>(g, c, s)) return RAFTING_ERR_NOT_READY;
    if (count == 0) count = 1;
    const bool has = nruns_of(g) > 0;
    if (!has && g.epochIndex != 0) return RAFTING_ERR_LOG_SHAPE;
    if ((!has || g.r0t != g.term) && nruns_of(g) >= KRUNS) return RAFTING_ERR_TERM_RUNS_OVERFLOW;
    const int64_t index = has ? g.hi + 1 : 1;
    if (has && g.r0t == g.term) g.hi = index + count - 1;
    else { push_run(g, c, index, g.term); g.hi = index + count - 1; }
    return replicate_log<W>(g, c, s, false, unavail);
}

// RaftRoutine.keepAlive / electionTimeout + onTimeout — RaftRoutine.java:53-77, Leader.java:119-126,
// Follower.java:156-168,223-279, Candidate.java:82-88
template <int W>
__device__ __forceinline__ int op_timeout(GS& g, const Ctx& c, LS& s, uint64_t unavail) {
    if (role_of(g) == RAFTING_ROLE_LEADER) {
        reset_timer(g, c, false, false);
        return replicate_log<W>(g, c, s, true, unavail);
    }
    if (!(g.timer > 0)) return 0;
    g.timer = RAFTING_TIMER_TIMEOUT;
    if (role_of(g) == RAFTING_ROLE_FOLLOWER && c.cfg->pre_vote) {
        const int64_t t = g.term;
        int err = switch_to(g, c, RAFTING_ROLE_FOLLOWER, g.term, ballot_of(g));
        if (err) return err;
        if (role_of(g) == RAFTING_ROLE_FOLLOWER && g.term == t) {            // prepareElection
            g.word |= W_TIMEOUT_DET;
            g.votes = 1;
            emit_ballot(g, c, RAFTING_BALLOT_PREVOTE, (int64_t)((uint64_t)g.term + 1u));
        }
        return 0;
    }
    return switch_to(g, c, RAFTING_ROLE_CANDIDATE, (int64_t)((uint64_t)g.term + 1u), (int)c.cfg->local_slot);
}

// ---------------------------------------------------------------------------------------------
// inbound requests (compiled only into the REQ variant of the kernel)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int log_contains(const GS& g, const Ctx& c, int64_t index, int64_t term) {   // Follower.java:177-191
    if (index == 0 && term == 0) return 1;
    if (index == 0 || term == 0) return -RAFTING_ERR_INDEX_TERM_ZERO;
    if (index <= g.epochIndex) {
        if (index == g.epochIndex && term != g.epochTerm) return -RAFTING_ERR_EPOCH_TERM_MISMATCH;
        return 1;
    }
    int64_t t;
    return term_at(g, c, index, t) && t == term;
}
__device__ __forceinline__ int log_up_to_date(const GS& g, int64_t index, int64_t term) {              // Follower.java:193-207
    if (nruns_of(g) > 0) return term > g.r0t || (term == g.r0t && index >= g.hi);
    if ((index > g.epochIndex && term < g.epochTerm) || (index == g.epochIndex && term != g.epochTerm))
        return -RAFTING_ERR_IMPOSSIBLE_LOG;
    return index >= g.epochIndex;
}

// Follower.appendEntries body — Follower.java:52-87 + RocksLog.conflict/truncate/append (RocksLog.java:169-225)
__device__ __noinline__ int follower_append(GS& g, const Ctx& c, int peer, int64_t term, int64_t prevIndex,
                                            int64_t prevTerm, int64_t first, uint32_t n, const int64_t* terms,
                                            int64_t leaderCommit, Reply& rep) {
    set_leader(g, peer);
    int err = 0;
    int lc = log_contains(g, c, prevIndex, prevTerm);
    if (lc < 0) err = -lc;
    else if (!lc) { rep.valid = 1; rep.success = 0; rep.term = g.term; }
    else {
        if (n > 0 && first <= g.epochIndex) {                                // purgeEntries :209-221
            int64_t skip = g.epochIndex - first + 1;
            if ((uint64_t)skip >= n) n = 0; else { first += skip; terms += skip; n -= (uint32_t)skip; }
        }
        if (n > 0) {
            int64_t conflictIndex = 0;                                       // RocksLog.conflict :199-216
            for (uint32_t i = 0; i < n; i++) {
                int64_t t;
                if (!term_at(g, c, first + i, t)) break;
                if (t != terms[i]) { conflictIndex = first + i; break; }
            }
            const bool nonEmpty = nruns_of(g) > 0;
            const int64_t hiAfter = (conflictIndex != 0 && nonEmpty && g.hi >= conflictIndex) ? conflictIndex - 1 : g.hi;
            const bool emptyAfter = !nonEmpty || hiAfter < g.lo;
            int64_t prevLogIndex = g.epochIndex;
            if (!emptyAfter && g.lo <= first) prevLogIndex = first < hiAfter ? first : hiAfter;   // seekForPrev
            else if (!emptyAfter) err = RAFTING_ERR_LOG_SHAPE;
            if (!err) {
                // capacity pre-check of the run table: reject before mutating anything
                const bool firstPutOk = emptyAfter ? (first == g.epochIndex + 1) : true;
                const bool contOk = !(first > prevLogIndex + 1);
                if (firstPutOk && contOk) {
                    uint32_t runs = 0; int64_t lastT = 0; bool have = false;
                    if (!emptyAfter) {
                        const int nr = nruns_of(g);
                        if (g.r0s <= hiAfter) { runs = (uint32_t)nr; lastT = g.r0t; have = true; }
                        else {
                            for (int k = 1; k < nr; k++) {
                                i64x2 r = c.runs[(size_t)k * c.G];
                                if (!have && r.x <= hiAfter) { runs = (uint32_t)(nr - k); lastT = r.y; have = true; }
                            }
                        }
                    }
                    for (uint32_t i = 0; i < n; i++)
                        if (first + i > prevLogIndex && (!have || terms[i] != lastT)) { runs++; lastT = terms[i]; have = true; }
                    if (runs > (uint32_t)KRUNS) err = RAFTING_ERR_TERM_RUNS_OVERFLOW;
                }
            }
            if (!err) {
                if (conflictIndex != 0) log_truncate(g, c, conflictIndex);   // Follower.java:70-72
                if (emptyAfter && first != g.epochIndex + 1) err = RAFTING_ERR_LOG_NOT_FOLLOW_EPOCH;
                else if (first > prevLogIndex + 1) err = RAFTING_ERR_LOG_NOT_CONTINUOUS;
                else
                    for (uint32_t i = 0; i < n; i++) {
                        // RocksLog.java:183-191: put everything above prevLogIndex; keys that already
                        // exist there hold the same term (no conflict was found), so only the tail grows
                        const int64_t idx = first + i;
                        if (idx > prevLogIndex && (nruns_of(g) == 0 || idx > g.hi)) log_append_one(g, c, idx, terms[i]);
                    }
            }
        }
        if (!err && leaderCommit > g.epochIndex && nruns_of(g) > 0)         // Follower.java:76-82
            err = commit_log(g, leaderCommit < g.hi ? leaderCommit : g.hi);
    }
    reset_timer(g, c, false, false);                                         // finally :83-85
    if (!err && !rep.valid) { rep.valid = 1; rep.success = 1; rep.term = term; }
    return err;
}
// Follower.appendEntries — Follower.java:35-88
__device__ __forceinline__ int follower_append_entries(GS& g, const Ctx& c, int peer, int64_t term, int64_t prevIndex,
                                                       int64_t prevTerm, int64_t first, uint32_t n, const int64_t* terms,
                                                       int64_t leaderCommit, Reply& rep) {
    if (term < g.term) { rep.valid = 1; rep.success = 0; rep.term = g.term; return 0; }
    reset_timer(g, c, true, false);
    if (term > g.term || (g.word & W_TIMEOUT_DET)) {
        int err = switch_to(g, c, RAFTING_ROLE_FOLLOWER, term, ballot_of(g));
        if (err) return err;
        reset_timer(g, c, true, false);
    } else if (leader_of(g) != -1 && peer != leader_of(g)) {
        return RAFTING_ERR_FOLLOWER_TWO_LEADERS;
    }
    return follower_append(g, c, peer, term, prevIndex, prevTerm, first, n, terms, leaderCommit, rep);
}
__device__ __forceinline__ int op_append_entries(GS& g, const Ctx& c, int peer, int64_t term, int64_t prevIndex,
                                                 int64_t prevTerm, int64_t first, uint32_t n, const int64_t* terms,
                                                 int64_t leaderCommit, Reply& rep) {
    const int role = role_of(g);
    if (role == RAFTING_ROLE_LEADER) {                                       // Leader.java:66-86
        if (peer == (int)c.cfg->local_slot) return RAFTING_ERR_LEADER_SELF_AE;
        if (term < g.term) { rep.valid = 1; rep.success = 0; rep.term = g.term; return 0; }
        if (term == g.term) return RAFTING_ERR_TWO_LEADERS;
        int err = switch_to(g, c, RAFTING_ROLE_FOLLOWER, g.term, ballot_of(g));
        if (err) return err;
    } else if (role == RAFTING_ROLE_CANDIDATE) {                             // Candidate.java:28-41
        if (term < g.term) { rep.valid = 1; rep.success = 0; rep.term = g.term; return 0; }
        int err = switch_to(g, c, RAFTING_ROLE_FOLLOWER, term, ballot_of(g));
        if (err) return err;
    }
    return follower_append_entries(g, c, peer, term, prevIndex, prevTerm, first, n, terms, leaderCommit, rep);
}
// Follower.requestVote — Follower.java:108-127
__device__ __forceinline__ int follower_request_vote(GS& g, const Ctx& c, int peer, int64_t term, int64_t lastIndex,
                                                     int64_t lastTerm, Reply& rep) {
    if (term < g.term) { rep.valid = 1; rep.success = 0; rep.term = g.term; return 0; }
    if (term == g.term) { rep.valid = 1; rep.success = peer == ballot_of(g); rep.term = g.term; return 0; }
    reset_timer(g, c, true, false);
    int up = log_up_to_date(g, lastIndex, lastTerm);
    if (up < 0) return -up;
    int err = switch_to(g, c, RAFTING_ROLE_FOLLOWER, term, up ? peer : -1);
    if (err) return err;
    rep.valid = 1; rep.success = peer == ballot_of(g); rep.term = g.term;
    return 0;
}
__device__ __forceinline__ int op_request_vote(GS& g, const Ctx& c, int peer, int64_t term, int64_t lastIndex,
                                               int64_t lastTerm, Reply& rep) {
    const int role = role_of(g), self = (int)c.cfg->local_slot;
    if (role == RAFTING_ROLE_LEADER) {                                       // Leader.java:93-111
        if (term < g.term) { rep.valid = 1; rep.success = 0; rep.term = g.term; return 0; }
        if (term == g.term) {
            if (ballot_of(g) == self) { rep.valid = 1; rep.success = 0; rep.term = g.term; return 0; }
            return RAFTING_ERR_LEADER_VOTE_SELF;
        }
        int err = switch_to(g, c, RAFTING_ROLE_FOLLOWER, g.term, peer);
        if (err) return err;
    } else if (role == RAFTING_ROLE_CANDIDATE) {                             // Candidate.java:49-72
        if (peer == self) return RAFTING_ERR_CANDIDATE_SELF_RV;
        if (term < g.term) { rep.valid = 1; rep.success = 0; rep.term = g.term; return 0; }
        if (term == g.term) {
            if (peer != ballot_of(g)) { rep.valid = 1; rep.success = 0; rep.term = g.term; return 0; }
            else if (ballot_of(g) != self) return RAFTING_ERR_CANDIDATE_VOTE_SELF;
        }
        int err = switch_to(g, c, RAFTING_ROLE_FOLLOWER, term, peer);
        if (err) return err;
    }
    return follower_request_vote(g, c, peer, term, lastIndex, lastTerm, rep);
}
__device__ __forceinline__ int op_pre_vote(GS& g, const Ctx& c, int peer, int64_t term, int64_t lastIndex,
                                           int64_t lastTerm, Reply& rep) {
    const int role = role_of(g);
    if (role == RAFTING_ROLE_LEADER) { rep.valid = 1; rep.success = 0; rep.term = g.term; return 0; }   // Leader.java:88-91
    if (role == RAFTING_ROLE_CANDIDATE) return op_request_vote(g, c, peer, term, lastIndex, lastTerm, rep);  // Candidate.java:43-46
    if (term <= g.term || !(g.word & W_TIMEOUT_DET)) { rep.valid = 1; rep.success = 0; rep.term = g.term; return 0; }
    reset_timer(g, c, true, false);
    int up = log_up_to_date(g, lastIndex, lastTerm);
    reset_timer(g, c, false, false);
    if (up < 0) return -up;
    rep.valid = 1; rep.success = up; rep.term = g.term;
    return 0;
}
// installSnapshot — RaftMember.java:61-66, Follower.java:130-153
__device__ __forceinline__ int op_install_snapshot(GS& g, const Ctx& c, int64_t term, int hostResult, Reply& rep) {
    if (role_of(g) != RAFTING_ROLE_FOLLOWER) {
        if (term >= g.term) return RAFTING_ERR_IS_BEFORE_AE;
        rep.valid = 1; rep.success = 0; rep.term = g.term; return 0;
    }
    reset_timer(g, c, true, false);
    if (term < g.term) { rep.valid = 1; rep.success = 0; rep.term = g.term; return 0; }
    if (term > g.term) return RAFTING_ERR_IS_BEFORE_AE;
    if (g.word & W_TIMEOUT_DET) {
        int err = switch_to(g, c, RAFTING_ROLE_FOLLOWER, g.term, ballot_of(g));
        if (err) return err;
        reset_timer(g, c, true, false);
    }
    rep.valid = 1; rep.success = hostResult; rep.term = g.term;
    reset_timer(g, c, false, false);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int W, bool REQ>
__global__ void __launch_bounds__(256)
step_kernel(Tables T, InboxD in, OutboxD out, CfgD cfg) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i = tid / W;
    if (i >= in.n) return;                                   // whole sub-warps leave together
    const uint32_t F = T.F, G = T.G;
    const uint32_t gid = in.gids ? in.gids[i] : i;
    if (gid >= G) return;
    RowOut ro;
    Ctx c;
    c.cfg = &cfg; c.gid = gid; c.F = F; c.G = G; c.runs = T.g_runs + gid;
    c.lane = (int)(tid % W);
    c.sub0 = (int)((threadIdx.x & 31u) & ~(uint32_t)(W - 1));
    c.mask = W == 32 ? 0xffffffffu : (((1u << W) - 1u) << c.sub0);
    c.lv = (uint32_t)c.lane < F;
    c.now = 0; c.draw = 0; c.ro = &ro;

    // ---- load the group's scalars (every lane) and this lane's follower slot ----
    GS g;
    {
        const uint64_t m = T.g_meta[gid];
        g.word = (uint32_t)m; g.inc = (uint32_t)(m >> 32);
        g.term = T.g_term[gid]; g.commit = T.g_commit[gid];
        g.lo = T.g_lo[gid]; g.hi = T.g_hi[gid]; g.timer = T.g_timer[gid];
        const i64x2 ep = T.g_epoch[gid]; g.epochIndex = ep.x; g.epochTerm = ep.y;
        const i64x2 el = T.g_elect[gid]; g.electTerm = el.x; g.electInc = (uint32_t)(uint64_t)el.y; g.votes = (int32_t)((uint64_t)el.y >> 32);
        const i64x2 r0 = T.g_runs[gid]; g.r0s = r0.x; g.r0t = r0.y;
        g.err = T.g_err[gid]; g.dirty = 0;
    }
    const GS g0 = g;
    LS s; s.next = s.match = s.lastEpoch = s.reqSucc = s.reqFail = s.lastReq = 0; s.inflight = s.rej = s.fail = s.pending = 0;
    const size_t li0 = (size_t)gid * F + (uint32_t)c.lane;
    if (c.lv) {
        const i64x2 nm = T.l_nm[li0], es = T.l_es[li0], fr = T.l_fr[li0]; const int4 cn = T.l_cnt[li0];
        s.next = nm.x; s.match = nm.y; s.lastEpoch = es.x; s.reqSucc = es.y; s.reqFail = fr.x; s.lastReq = fr.y;
        s.inflight = cn.x; s.rej = cn.y; s.fail = cn.z; s.pending = cn.w;
    }
    const bool alive = (g.word & W_ALIVE) != 0;
    const int self = (int)cfg.local_slot;

    for (uint32_t r = 0; r < in.rows; r++) {
        const size_t gi = (size_t)r * in.n + i;
        ro.pm = 0; ro.bm = 0; ro.pe = 0; ro.bt = 0;
        uint32_t repMeta = 0; int64_t repTerm = 0;

        // ---- group op, or the implied TIMEOUT of a sweep row ----
        const int64_t sweep = in.row_now ? in.row_now[r] : 0;
        uint32_t kind = RAFTING_OP_NONE, meta = 0, entoff = 0;
        c.draw = 0;
        if (sweep != 0) {
            if (alive) {
                const bool due = (role_of(g) == RAFTING_ROLE_LEADER) ? (g.timer <= sweep)
                                 : (g.timer > 0 && g.timer != I64MAX && g.timer <= sweep);
                if (due) { kind = RAFTING_OP_TIMEOUT; c.now = sweep; }
            }
        } else if (in.op_meta) {
            const uint64_t m = in.op_meta[gi];
            meta = (uint32_t)m; entoff = (uint32_t)(m >> 32);
            kind = RAFTING_OP_KIND(meta);
            if (kind != RAFTING_OP_NONE) { const i64x2 nr = in.op_nr[gi]; c.now = nr.x; c.draw = nr.y; }
        }
        if (kind != RAFTING_OP_NONE) {
            int err = 0; Reply rep = {0, 0, 0};
            int64_t a = 0, b = 0, cc = 0, d = 0;
            if (sweep == 0) {
                if (in.op_ab) { const i64x2 v = in.op_ab[gi]; a = v.x; b = v.y; }
                if (REQ && in.op_cd) { const i64x2 v = in.op_cd[gi]; cc = v.x; d = v.y; }
            }
            const int peer = (int)RAFTING_OP_PEER(meta); const uint32_t count = RAFTING_OP_COUNT(meta);
            if (!alive) err = RAFTING_ERR_CLOSED_GROUP;
            else if (kind == RAFTING_OP_SUBMIT) err = op_submit<W>(g, c, s, count, (uint64_t)a);
            else if (kind == RAFTING_OP_TIMEOUT) err = op_timeout<W>(g, c, s, (uint64_t)a);
            else if (REQ && kind == RAFTING_OP_AE_REQUEST) {
                const int64_t first = in.op_e ? in.op_e[gi] : (int64_t)((uint64_t)b + 1u);
                const int64_t* terms = in.ent_terms ? in.ent_terms + entoff : nullptr;
                if (count > 0 && (!terms || (uint64_t)entoff + count > in.ent_count)) err = RAFTING_ERR_BAD_EVENT;
                else err = op_append_entries(g, c, peer, a, b, cc, first, count, terms, d, rep);
            }
            else if (REQ && kind == RAFTING_OP_PREVOTE_REQ) err = op_pre_vote(g, c, peer, a, b, cc, rep);
            else if (REQ && kind == RAFTING_OP_VOTE_REQ) err = op_request_vote(g, c, peer, a, b, cc, rep);
            else if (REQ && kind == RAFTING_OP_IS_REQUEST) err = op_install_snapshot(g, c, a, d != 0, rep);
            else if (REQ && kind == RAFTING_OP_FLUSH) err = log_flush(g, c, b, cc);
            else err = RAFTING_ERR_BAD_EVENT;
            if (err) { if (alive) flag_err(g, err); rep.valid = 0; }
            repMeta = (uint32_t)(rep.valid ? 1 : 0) | ((uint32_t)(rep.success ? 1 : 0) << 1) | ((uint32_t)err << 8);
            repTerm = rep.valid ? rep.term : 0;
        }

        // ---- lane events: loaded one per lane (coalesced), applied in lane order ----
        if (in.ev_meta) {
            uint64_t em = 0; i64x2 etn = {0, 0}, eel = {0, 0};
            if (c.lv) {
                const size_t li = gi * F + (uint32_t)c.lane;
                em = in.ev_meta[li];
                if (RAFTING_EVM_KIND(em) != RAFTING_EV_NONE) { etn = in.ev_tn[li]; if (in.ev_el) eel = in.ev_el[li]; }
            }
            unsigned pending = sub_ballot(c, RAFTING_EVM_KIND(em) != RAFTING_EV_NONE, W);
            if (!alive) pending = 0;
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
                            if (respTerm > g.term) err = switch_to(g, c, RAFTING_ROLE_FOLLOWER, respTerm, lane_to_slot(c, f));
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
                        if (respTerm > nextTerm) err = switch_to(g, c, RAFTING_ROLE_FOLLOWER, respTerm, lane_to_slot(c, f));
                        else if (success && ++g.votes >= majority(c)) err = switch_to(g, c, RAFTING_ROLE_CANDIDATE, nextTerm, self);
                    }
                } else if (ek == RAFTING_EV_RV_REPLY) {
                    // RV-Echo — Candidate.java:112-134 (and the elected Candidate's surviving head, :75-80)
                    if (role_of(g) == RAFTING_ROLE_CANDIDATE && inc == g.inc) {
                        if (outcome == RAFTING_OUT_OK) {
                            if (respTerm > g.term) err = switch_to(g, c, RAFTING_ROLE_FOLLOWER, respTerm, lane_to_slot(c, f));
                            else if (success && ++g.votes >= majority(c)) {
                                g.electInc = g.inc; g.electTerm = g.term; g.word &= ~W_ELECT_ABORT;
                                err = switch_to(g, c, RAFTING_ROLE_LEADER, g.term, self);
                            }
                        }
                    } else if (g.electInc != 0 && inc == g.electInc && !(g.word & W_ELECT_ABORT) && outcome == RAFTING_OUT_OK) {
                        if (respTerm > g.electTerm) {
                            g.word |= W_ELECT_ABORT;
                            err = switch_to(g, c, RAFTING_ROLE_FOLLOWER, respTerm, lane_to_slot(c, f));
                        } else if (success) err = switch_to(g, c, RAFTING_ROLE_LEADER, g.electTerm, self);
                    }
                } else err = RAFTING_ERR_BAD_EVENT;
                if (err) flag_err(g, err);
            }
        }

        // ---- row outputs ----
        if (c.lv && out.plan_meta) {
            const size_t li = gi * F + (uint32_t)c.lane;
            out.plan_meta[li] = ro.pm;
            if (ro.pm != 0) { out.plan_pp[li] = ro.pp; out.plan_lc[li] = ro.lc; out.plan_epoch[li] = ro.pe; }
        }
        if (c.lane == 0) {
            if (out.rep_meta) { out.rep_meta[gi] = repMeta; if (repMeta & 1u) out.rep_term[gi] = repTerm; }
            if (out.ballot_meta) {
                out.ballot_meta[gi] = ro.bm;
                if (ro.bm != 0) { out.ballot_term[gi] = ro.bt; out.ballot_last[gi] = ro.bl; }
            }
        }
    }

    // ---- write back what changed ----
    if (c.lv) {
        i64x2 v; int4 cn;
        v.x = s.next; v.y = s.match; T.l_nm[li0] = v;
        v.x = s.lastEpoch; v.y = s.reqSucc; T.l_es[li0] = v;
        v.x = s.reqFail; v.y = s.lastReq; T.l_fr[li0] = v;
        cn.x = s.inflight; cn.y = s.rej; cn.z = s.fail; cn.w = s.pending; T.l_cnt[li0] = cn;
    }
    if (c.lane == 0) {
        if (g.word != g0.word || g.inc != g0.inc) T.g_meta[gid] = (uint64_t)g.word | ((uint64_t)g.inc << 32);
        if (g.term != g0.term) T.g_term[gid] = g.term;
        if (g.commit != g0.commit) T.g_commit[gid] = g.commit;
        if (g.lo != g0.lo) T.g_lo[gid] = g.lo;
        if (g.hi != g0.hi) T.g_hi[gid] = g.hi;
        if (g.timer != g0.timer) T.g_timer[gid] = g.timer;
        if (g.epochIndex != g0.epochIndex || g.epochTerm != g0.epochTerm) { i64x2 v; v.x = g.epochIndex; v.y = g.epochTerm; T.g_epoch[gid] = v; }
        if (g.electTerm != g0.electTerm || g.electInc != g0.electInc || g.votes != g0.votes) {
            i64x2 v; v.x = g.electTerm; v.y = (int64_t)((uint64_t)g.electInc | ((uint64_t)(uint32_t)g.votes << 32)); T.g_elect[gid] = v;
        }
        if (g.r0s != g0.r0s || g.r0t != g0.r0t) { i64x2 v; v.x = g.r0s; v.y = g.r0t; T.g_runs[gid] = v; }
        if (g.err != g0.err) T.g_err[gid] = g.err;
        // end-of-step snapshot columns (RaftLog.lastCommitted / RaftParticipant.currentTerm / role)
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
