// handlers.cuh — device-side restatement of the reference's per-event handlers (used by step_kernel.cuh).
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
#pragma unroll 1
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

__device__ __forceinline__ void emit_ballot(const GS& g, const Ctx& c, RowOut& ro, int kind, int64_t term) {
    int64_t li, lt; last_or_epoch(g, li, lt);
    ro.bm = (uint64_t)kind | ((uint64_t)g.inc << 32);
    ro.bt = term; ro.bl.x = li; ro.bl.y = lt;
}

// RaftRoutine.trySwitch + switchTo + convertTo — RaftRoutine.java:140-216
__device__ __forceinline__ int switch_to(GS& g, const Ctx& c, RowOut& ro, int role, int64_t term, int ballot) {
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
        emit_ballot(g, c, ro, RAFTING_BALLOT_VOTE, term);
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
__device__ __forceinline__ int replicate_log(GS& g, const Ctx& c, RowOut& ro, LS& s, bool heartbeat, uint64_t unavail) {
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
            ro.pm = pm; ro.pp = pp; ro.lc = lc; ro.pe = epochIndex;
        }
    }
    if (errs) return __shfl_sync(c.mask, e, fe, W);
    return 0;
}

template <int W>
__device__ __forceinline__ bool leader_ready(GS& g, const Ctx& c, const LS& s, int32_t crit, int64_t cool) {   // Leader.java:52-64
    bool r = c.lv && state_ready(s, crit, cool, c.now);
    int cnt = __popc(sub_ballot(c, r, W));
    // the Java loop only returns true from inside `isReady(..) && ++ready > half`: at least one follower must be ready
    bool ready = (g.word & W_PREPARED) && cnt >= 1 && (1 + cnt > (int)c.F / 2);
    g.word = ready ? (g.word | W_READY) : (g.word & ~W_READY);
    return ready;
}

// RaftStub.process -> Leader.acceptCommand -> RocksLog.newEntry — RaftStub.java:79-91, Leader.java:128-140, RocksLog.java:82-89
template <int W>
__device__ __forceinline__ int op_submit(GS& g, const Ctx& c, RowOut& ro, LS& s, uint32_t count, uint64_t unavail) {
    if (role_of(g) != RAFTING_ROLE_LEADER) return RAFTING_ERR_NOT_LEADER;
    if (!leader_ready<W>(g, c, s, c.cfg->avail_critical_point, c.cfg->recovery_cool_down_ms)) return RAFTING_ERR_NOT_READY;
    if (count == 0) count = 1;
    const bool has = nruns_of(g) > 0;
    if (!has && g.epochIndex != 0) return RAFTING_ERR_LOG_SHAPE;
    if ((!has || g.r0t != g.term) && nruns_of(g) >= KRUNS) return RAFTING_ERR_TERM_RUNS_OVERFLOW;
    const int64_t index = has ? g.hi + 1 : 1;
    if (has && g.r0t == g.term) g.hi = index + count - 1;
    else { push_run(g, c, index, g.term); g.hi = index + count - 1; }
    return replicate_log<W>(g, c, ro, s, false, unavail);
}

// RaftRoutine.keepAlive / electionTimeout + onTimeout — RaftRoutine.java:53-77, Leader.java:119-126,
// Follower.java:156-168,223-279, Candidate.java:82-88
template <int W>
__device__ __forceinline__ int op_timeout(GS& g, const Ctx& c, RowOut& ro, LS& s, uint64_t unavail) {
    if (role_of(g) == RAFTING_ROLE_LEADER) {
        reset_timer(g, c, false, false);
        return replicate_log<W>(g, c, ro, s, true, unavail);
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
__device__ __noinline__ int follower_append(GS& g, const Ctx& c, RowOut& ro, int peer, int64_t term, int64_t prevIndex,
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
__device__ __forceinline__ int follower_append_entries(GS& g, const Ctx& c, RowOut& ro, int peer, int64_t term, int64_t prevIndex,
                                                       int64_t prevTerm, int64_t first, uint32_t n, const int64_t* terms,
                                                       int64_t leaderCommit, Reply& rep) {
    if (term < g.term) { rep.valid = 1; rep.success = 0; rep.term = g.term; return 0; }
    reset_timer(g, c, true, false);
    if (term > g.term || (g.word & W_TIMEOUT_DET)) {
        int err = switch_to(g, c, ro, RAFTING_ROLE_FOLLOWER, term, ballot_of(g));
        if (err) return err;
        reset_timer(g, c, true, false);
    } else if (leader_of(g) != -1 && peer != leader_of(g)) {
        return RAFTING_ERR_FOLLOWER_TWO_LEADERS;
    }
    return follower_append(g, c, ro, peer, term, prevIndex, prevTerm, first, n, terms, leaderCommit, rep);
}
__device__ __forceinline__ int op_append_entries(GS& g, const Ctx& c, RowOut& ro, int peer, int64_t term, int64_t prevIndex,
                                                 int64_t prevTerm, int64_t first, uint32_t n, const int64_t* terms,
                                                 int64_t leaderCommit, Reply& rep) {
    const int role = role_of(g);
    if (role == RAFTING_ROLE_LEADER) {                                       // Leader.java:66-86
        if (peer == (int)c.cfg->local_slot) return RAFTING_ERR_LEADER_SELF_AE;
        if (term < g.term) { rep.valid = 1; rep.success = 0; rep.term = g.term; return 0; }
        if (term == g.term) return RAFTING_ERR_TWO_LEADERS;
        int err = switch_to(g, c, ro, RAFTING_ROLE_FOLLOWER, g.term, ballot_of(g));
        if (err) return err;
    } else if (role == RAFTING_ROLE_CANDIDATE) {                             // Candidate.java:28-41
        if (term < g.term) { rep.valid = 1; rep.success = 0; rep.term = g.term; return 0; }
        int err = switch_to(g, c, ro, RAFTING_ROLE_FOLLOWER, term, ballot_of(g));
        if (err) return err;
    }
    return follower_append_entries(g, c, ro, peer, term, prevIndex, prevTerm, first, n, terms, leaderCommit, rep);
}
// Follower.requestVote — Follower.java:108-127
__device__ __forceinline__ int follower_request_vote(GS& g, const Ctx& c, RowOut& ro, int peer, int64_t term, int64_t lastIndex,
                                                     int64_t lastTerm, Reply& rep) {
    if (term < g.term) { rep.valid = 1; rep.success = 0; rep.term = g.term; return 0; }
    if (term == g.term) { rep.valid = 1; rep.success = peer == ballot_of(g); rep.term = g.term; return 0; }
    reset_timer(g, c, true, false);
    int up = log_up_to_date(g, lastIndex, lastTerm);
    if (up < 0) return -up;
    int err = switch_to(g, c, ro, RAFTING_ROLE_FOLLOWER, term, up ? peer : -1);
    if (err) return err;
    rep.valid = 1; rep.success = peer == ballot_of(g); rep.term = g.term;
    return 0;
}
__device__ __forceinline__ int op_request_vote(GS& g, const Ctx& c, RowOut& ro, int peer, int64_t term, int64_t lastIndex,
                                               int64_t lastTerm, Reply& rep) {
    const int role = role_of(g), self = (int)c.cfg->local_slot;
    if (role == RAFTING_ROLE_LEADER) {                                       // Leader.java:93-111
        if (term < g.term) { rep.valid = 1; rep.success = 0; rep.term = g.term; return 0; }
        if (term == g.term) {
            if (ballot_of(g) == self) { rep.valid = 1; rep.success = 0; rep.term = g.term; return 0; }
            return RAFTING_ERR_LEADER_VOTE_SELF;
        }
        int err = switch_to(g, c, ro, RAFTING_ROLE_FOLLOWER, g.term, peer);
        if (err) return err;
    } else if (role == RAFTING_ROLE_CANDIDATE) {                             // Candidate.java:49-72
        if (peer == self) return RAFTING_ERR_CANDIDATE_SELF_RV;
        if (term < g.term) { rep.valid = 1; rep.success = 0; rep.term = g.term; return 0; }
        if (term == g.term) {
            if (peer != ballot_of(g)) { rep.valid = 1; rep.success = 0; rep.term = g.term; return 0; }
            else if (ballot_of(g) != self) return RAFTING_ERR_CANDIDATE_VOTE_SELF;
        }
        int err = switch_to(g, c, ro, RAFTING_ROLE_FOLLOWER, term, peer);
        if (err) return err;
    }
    return follower_request_vote(g, c, ro, peer, term, lastIndex, lastTerm, rep);
}
__device__ __forceinline__ int op_pre_vote(GS& g, const Ctx& c, RowOut& ro, int peer, int64_t term, int64_t lastIndex,
                                           int64_t lastTerm, Reply& rep) {
    const int role = role_of(g);
    if (role == RAFTING_ROLE_LEADER) { rep.valid = 1; rep.success = 0; rep.term = g.term; return 0; }   // Leader.java:88-91
    if (role == RAFTING_ROLE_CANDIDATE) return op_request_vote(g, c, ro, peer, term, lastIndex, lastTerm, rep);  // Candidate.java:43-46
    if (term <= g.term || !(g.word & W_TIMEOUT_DET)) { rep.valid = 1; rep.success = 0; rep.term = g.term; return 0; }
    reset_timer(g, c, true, false);
    int up = log_up_to_date(g, lastIndex, lastTerm);
    reset_timer(g, c, false, false);
    if (up < 0) return -up;
    rep.valid = 1; rep.success = up; rep.term = g.term;
    return 0;
}
// installSnapshot — RaftMember.java:61-66, Follower.java:130-153
__device__ __forceinline__ int op_install_snapshot(GS& g, const Ctx& c, RowOut& ro, int64_t term, int hostResult, Reply& rep) {
    if (role_of(g) != RAFTING_ROLE_FOLLOWER) {
        if (term >= g.term) return RAFTING_ERR_IS_BEFORE_AE;
        rep.valid = 1; rep.success = 0; rep.term = g.term; return 0;
    }
    reset_timer(g, c, true, false);
    if (term < g.term) { rep.valid = 1; rep.success = 0; rep.term = g.term; return 0; }
    if (term > g.term) return RAFTING_ERR_IS_BEFORE_AE;
    if (g.word & W_TIMEOUT_DET) {
        int err = switch_to(g, c, ro, RAFTING_ROLE_FOLLOWER, g.term, ballot_of(g));
        if (err) return err;
        reset_timer(g, c, true, false);
    }
    rep.valid = 1; rep.success = hostResult; rep.term = g.term;
    reset_timer(g, c, false, false);
    return 0;
}


}  // namespace rafting
