// workload.cu — synthetic replication streams (the "peers" of every Raft group) for tests and bench.
//
// The reference has no simulator (three JVMs over loopback TCP, asserts nothing);
// BASELINE.json's configs are one-line descriptions.  This file is their concretisation: a
// counter-based generator (SplitMix64 keyed by (seed, gid, tick, lane, salt)) that plays the remote
// followers/voters of each group.  It is a pure function of the previous step's outbox, so the very
// same stream can be produced on the host (CPU tests, CPU baseline) and on the device (bench at full
// size, where the stream must already be resident in HBM) and any shard of it can be regenerated
// independently.  It is workload infrastructure, not part of the engine and not part of the oracle.
//
// Leader steady state (configs #2 / #4): per tick (= one inbox row) and group
//   op    : SUBMIT(count = a) with a in U{0..max_submit}; a == 0 -> TIMEOUT (the Leader's keepAlive)
//   lanes : the reply to the plan this lane emitted one STEP earlier in the same row:
//           AE plan -> AE ack, IS plan -> IS ack, nothing for SKIP / UNAVAILABLE / NONE;
//           outcome ok+success / ok+reject / error (RPC timeout) / canceled with the configured ppm.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rafting_b200.h"
#include "../../include/rafting_workload.h"

namespace {

struct Cols {
    // inbox (to fill)
    uint64_t* op_meta; rafting_i64x2_t* op_nr; rafting_i64x2_t* op_ab; rafting_i64x2_t* op_cd; int64_t* op_e;
    uint64_t* ev_meta; rafting_i64x2_t* ev_tn; rafting_i64x2_t* ev_el;
    // previous outbox (read)
    const uint64_t* plan_meta; const rafting_i64x2_t* plan_lc; const int64_t* plan_epoch;
    const uint64_t* ballot_meta; const int64_t* ballot_term;
    const int64_t* current_term; const int64_t* commit_index; const uint32_t* incarnation; const uint32_t* role_word;
    const rafting_i64x2_t* last_entry;
    const uint32_t* gids;
    bool compact;           // RAFTING_INBOX_COMPACT_GROUPS: the previous outbox's group columns are indexed by position
};

RAFTING_HD inline uint64_t key(uint64_t seed, uint32_t gid, uint64_t tick, uint32_t lane, uint32_t salt) {
    uint64_t h = rafting_splitmix64(seed ^ ((uint64_t)gid * 0x9E3779B97F4A7C15ull));
    h = rafting_splitmix64(h ^ (tick * 0xC2B2AE3D27D4EB4Full));
    return rafting_splitmix64(h ^ ((uint64_t)lane << 32) ^ salt);
}

RAFTING_HD inline void leader_cell(const rafting_wl_cfg_t& w, const Cols& c, uint64_t step, uint32_t r, uint32_t i) {
    const uint32_t gid = c.gids ? c.gids[i] : w.gid_base + i;    // global group id keys the RNG
    const uint32_t lgid = (c.gids && !c.compact) ? c.gids[i] : i;                // local id indexes the snapshot columns
    const uint64_t tick = step * w.rows + r;
    const size_t gi = (size_t)r * w.n + i;
    const int64_t now = w.t0 + 10 * (int64_t)tick + (int64_t)(gid % 10u);
    if (c.op_meta) {
        const uint32_t a = (uint32_t)(key(w.seed, gid, tick, 0, 1) % (uint64_t)(w.max_submit + 1));
        c.op_meta[gi] = a ? RAFTING_OP_MAKE(RAFTING_OP_SUBMIT, 0, a) : RAFTING_OP_MAKE(RAFTING_OP_TIMEOUT, 0, 0);
        c.op_nr[gi].x = now; c.op_nr[gi].y = 0;
        if (c.op_ab) { c.op_ab[gi].x = 0; c.op_ab[gi].y = 0; }     // no unavailable follower in this stream: the column may be omitted
    }
    if (c.ev_meta) {
        for (uint32_t f = 0; f < w.F; f++) {
            const size_t li = gi * w.F + f;
            uint64_t em = 0; rafting_i64x2_t tn = {0, 0}, el = {0, 0};
            if (c.plan_meta) {
                const uint64_t pm = c.plan_meta[li];
                const uint32_t pk = RAFTING_PLM_KIND(pm);
                if (pk == RAFTING_PLAN_AE || pk == RAFTING_PLAN_IS) {
                    const uint32_t u = (uint32_t)(key(w.seed, gid, tick, f, 2) % 1000000ull);
                    uint32_t outcome = RAFTING_OUT_OK; int success = 1;
                    if (u < w.p_error_ppm) outcome = RAFTING_OUT_ERROR;
                    else if (u < w.p_error_ppm + w.p_cancel_ppm) outcome = RAFTING_OUT_CANCELED;
                    else if (u < w.p_error_ppm + w.p_cancel_ppm + w.p_reject_ppm) success = 0;
                    if (pk == RAFTING_PLAN_IS && outcome == RAFTING_OUT_OK) success = (key(w.seed, gid, tick, f, 3) & 1ull) != 0;
                    if (outcome != RAFTING_OUT_OK) success = 0;
                    em = RAFTING_EVM_MAKE(pk == RAFTING_PLAN_IS ? RAFTING_EV_IS_ACK : RAFTING_EV_AE_ACK, outcome, success, RAFTING_PLM_INC(pm));
                    tn.x = outcome == RAFTING_OUT_OK ? c.current_term[lgid] : 0;     // the follower answers with the leader's term
                    tn.y = now + 5;
                    el.x = c.plan_epoch[li]; el.y = c.plan_lc[li].x;
                }
            }
            c.ev_meta[li] = em;
            if (em) { c.ev_tn[li] = tn; c.ev_el[li] = el; }
        }
    }
}

// election warm-up: phase 0 = TIMEOUT for every group; phase 1/2 = every lane grants the
// PreVote / RequestVote of the role object that asked (incarnation from the previous outbox)
RAFTING_HD inline void election_cell(const rafting_wl_cfg_t& w, const Cols& c, uint32_t phase, uint32_t i) {
    const uint32_t gid = c.gids ? c.gids[i] : w.gid_base + i;
    const uint32_t lgid = (c.gids && !c.compact) ? c.gids[i] : i;
    const size_t gi = i;                                           // single row
    const int64_t now = w.t0 - 1000 + 10 * (int64_t)phase + (int64_t)(gid % 10u);
    if (c.op_meta) {
        c.op_meta[gi] = phase == 0 ? RAFTING_OP_MAKE(RAFTING_OP_TIMEOUT, 0, 0) : 0;
        c.op_nr[gi].x = now; c.op_nr[gi].y = 0; if (c.op_ab) { c.op_ab[gi].x = 0; c.op_ab[gi].y = 0; }
    }
    if (c.ev_meta) {
        for (uint32_t f = 0; f < w.F; f++) {
            const size_t li = gi * w.F + f;
            uint64_t em = 0;
            if (phase == 1 || phase == 2) {
                em = RAFTING_EVM_MAKE(phase == 1 ? RAFTING_EV_PV_REPLY : RAFTING_EV_RV_REPLY, RAFTING_OUT_OK, 1, c.incarnation[lgid]);
                c.ev_tn[li].x = c.current_term[lgid]; c.ev_tn[li].y = now;
                c.ev_el[li].x = 0; c.ev_el[li].y = 0;
            }
            c.ev_meta[li] = em;
        }
    }
}

// ---- shared pieces ---------------------------------------------------------------------------------
RAFTING_HD inline void put_op(const Cols& c, size_t gi, uint32_t kind, uint32_t peer, uint32_t count, uint32_t ent,
                              int64_t now, int64_t a, int64_t b, int64_t cc, int64_t d, int64_t e) {
    c.op_meta[gi] = (uint64_t)RAFTING_OP_MAKE(kind, peer, count) | ((uint64_t)ent << 32);
    c.op_nr[gi].x = now; c.op_nr[gi].y = 0;
    c.op_ab[gi].x = a; c.op_ab[gi].y = b;
    if (c.op_cd) { c.op_cd[gi].x = cc; c.op_cd[gi].y = d; }
    if (c.op_e) c.op_e[gi] = e;
}
// replies of the F peers to the ballot this group emitted one step earlier in the same row
RAFTING_HD inline bool ballot_replies(const rafting_wl_cfg_t& w, const Cols& c, uint32_t gid, uint64_t tick, size_t gi, int64_t now,
                                      uint32_t p_grant_pv_ppm, uint32_t p_grant_rv_ppm, uint32_t p_higher_ppm, uint32_t p_err_ppm) {
    if (!c.ballot_meta) return false;
    const uint64_t bm = c.ballot_meta[gi];
    const uint32_t bk = (uint32_t)(bm & 0xf);
    if (bk == RAFTING_BALLOT_NONE) return false;
    const bool pre = bk == RAFTING_BALLOT_PREVOTE;
    const int64_t bt = c.ballot_term[gi];
    for (uint32_t f = 0; f < w.F; f++) {
        const size_t li = gi * w.F + f;
        const uint32_t u = (uint32_t)(key(w.seed, gid, tick, f, 7) % 1000000ull);
        uint32_t outcome = RAFTING_OUT_OK; int granted = 0; int64_t rt = pre ? bt - 1 : bt;
        if (u < p_err_ppm) outcome = RAFTING_OUT_ERROR;
        else if (u < p_err_ppm + p_higher_ppm) rt = bt + 2;
        else granted = (key(w.seed, gid, tick, f, 8) % 1000000ull) < (pre ? p_grant_pv_ppm : p_grant_rv_ppm);
        c.ev_meta[li] = RAFTING_EVM_MAKE(pre ? RAFTING_EV_PV_REPLY : RAFTING_EV_RV_REPLY, outcome, granted, (uint32_t)(bm >> 32));
        c.ev_tn[li].x = outcome == RAFTING_OUT_OK ? rt : 0; c.ev_tn[li].y = now + 5;
        c.ev_el[li].x = 0; c.ev_el[li].y = 0;
    }
    return true;
}
RAFTING_HD inline void no_events(const rafting_wl_cfg_t& w, const Cols& c, size_t gi) {
    for (uint32_t f = 0; f < w.F; f++) c.ev_meta[gi * w.F + f] = 0;
}
RAFTING_HD inline uint32_t lane_slot(uint32_t f, uint32_t local) { return f < local ? f : f + 1; }

// ---- config #3: RequestVote storm with PreVote ---------------------------------------------------------
// Every group that has no ballot in flight times out with probability p (all of them at tick 0);
// peers grant PreVote w.p. 0.7 / RequestVote w.p. 0.6, answer with a higher term w.p. 0.02 and time out
// w.p. 0.05; 30 % of the groups also receive an inbound PreVote / RequestVote from a peer candidate.
RAFTING_HD inline void vote_cell(const rafting_wl_cfg_t& w, const Cols& c, uint64_t step, uint32_t r, uint32_t i) {
    const uint32_t gid = c.gids ? c.gids[i] : w.gid_base + i;
    const uint32_t lgid = (c.gids && !c.compact) ? c.gids[i] : i;
    const uint64_t tick = step * w.rows + r;
    const size_t gi = (size_t)r * w.n + i;
    const int64_t now = w.t0 + 10 * (int64_t)tick + (int64_t)(gid % 10u);
    const bool answered = ballot_replies(w, c, gid, tick, gi, now, 700000, 600000, 20000, 50000);
    if (!answered) no_events(w, c, gi);
    const uint32_t role = c.role_word ? (c.role_word[lgid] & 3u) : RAFTING_ROLE_FOLLOWER;
    const int64_t term = c.current_term ? c.current_term[lgid] : 0;
    const int64_t li = c.last_entry ? c.last_entry[lgid].x : 0, lt = c.last_entry ? c.last_entry[lgid].y : 0;
    const uint32_t u = (uint32_t)(key(w.seed, gid, tick, 0, 9) % 1000u);
    if (u < 300u && c.current_term) {                       // inbound vote request from a peer candidate
        const uint32_t v = (uint32_t)(key(w.seed, gid, tick, 0, 10) % 100u);
        const uint32_t peer = lane_slot((uint32_t)(key(w.seed, gid, tick, 0, 11) % w.F), w.local_slot);
        const int64_t t = term + (v < 20 ? 0 : (v < 80 ? 1 : 2));
        const int64_t qi = li + (int64_t)(key(w.seed, gid, tick, 0, 12) % 3u) - 1;
        const int64_t qt = lt + ((key(w.seed, gid, tick, 0, 13) % 10u) == 0 ? 1 : 0);
        put_op(c, gi, (v & 1u) ? RAFTING_OP_PREVOTE_REQ : RAFTING_OP_VOTE_REQ, peer, 0, 0, now, t, qi < 0 ? 0 : qi, qt, 0, 0);
    } else if (!answered && (tick == 0 || (role == RAFTING_ROLE_LEADER ? u < 650u : u < 800u))) {
        put_op(c, gi, RAFTING_OP_TIMEOUT, 0, 0, 0, now, 0, 0, 0, 0, 0);    // election timeout / Leader keepAlive
    } else {
        put_op(c, gi, RAFTING_OP_NONE, 0, 0, 0, now, 0, 0, 0, 0, 0);
    }
}

// ---- config #5: mixed leader churn + InstallSnapshot catch-up -------------------------------------------
// class = hash(gid) % 10: 0..7 steady leaders (as config #2); 8 churn; 9 catch-up.
//   churn    (64-tick cycle): a Leader at phase 0 receives an AppendEntries from a higher-term leader and
//            steps down; while Follower, phases 1..32 bring AppendEntries requests (0..50 entries, 1 %
//            mismatching prev, 1 % conflicting suffix); afterwards the election timer fires and every
//            peer grants, so it is re-elected and resumes leader work.
//   catch-up (128-tick cycle): lane 1 stops answering (RPC errors) during phases 28..127, the log is
//            flushed to commitIndex at phase 0, so the straggler falls behind the epoch, gets
//            InstallSnapshot plans (IS-Echo false/true at random) and then catches up.
// Entry terms of an inbound request all equal its term; the pool holds 50 copies of every term < 256.
RAFTING_HD inline void mixed_cell(const rafting_wl_cfg_t& w, const Cols& c, uint64_t step, uint32_t r, uint32_t i) {
    const uint32_t gid = c.gids ? c.gids[i] : w.gid_base + i;
    const uint32_t lgid = (c.gids && !c.compact) ? c.gids[i] : i;
    const uint64_t tick = step * w.rows + r;
    const size_t gi = (size_t)r * w.n + i;
    const int64_t now = w.t0 + 10 * (int64_t)tick + (int64_t)(gid % 10u);
    const uint32_t cls = (uint32_t)(rafting_splitmix64(w.seed ^ ((uint64_t)gid * 0x2545F4914F6CDD1Dull)) % 10u);
    const uint32_t role = c.role_word ? (c.role_word[lgid] & 3u) : RAFTING_ROLE_FOLLOWER;
    const int64_t term = c.current_term ? c.current_term[lgid] : 0;
    const int64_t li = c.last_entry ? c.last_entry[lgid].x : 0, lt = c.last_entry ? c.last_entry[lgid].y : 0;
    if (ballot_replies(w, c, gid, tick, gi, now, 1000000, 1000000, 0, 0)) {     // mid-election: everybody grants
        put_op(c, gi, RAFTING_OP_NONE, 0, 0, 0, now, 0, 0, 0, 0, 0);
        return;
    }
    if (cls == 8u) {
        const uint32_t ph = (uint32_t)((tick + gid) % 64u);
        const uint32_t leader = lane_slot(gid % w.F, w.local_slot);
        if (role == RAFTING_ROLE_LEADER && ph == 0 && c.current_term) {
            no_events(w, c, gi);
            put_op(c, gi, RAFTING_OP_AE_REQUEST, leader, 0, 0, now, term + 1, li, lt, 0, li + 1);
            return;
        }
        if (role != RAFTING_ROLE_LEADER) {
            no_events(w, c, gi);
            if (ph >= 1 && ph <= 32 && c.current_term) {
                const uint32_t v = (uint32_t)(key(w.seed, gid, tick, 0, 20) % 100u);
                uint32_t n = (uint32_t)(key(w.seed, gid, tick, 0, 21) % 51u);
                int64_t pi = li, pt = lt;
                if (v == 0) { pi = li + 3; pt = term; }                      // prev the follower does not hold
                else if (v == 1 && li >= 3) { pi = li - 2; pt = lt; if (n == 0) n = 2; }   // rewrite the tail
                const uint32_t tt = (uint32_t)(term < 255 ? term : 255);
                put_op(c, gi, RAFTING_OP_AE_REQUEST, leader, n, tt * 50u, now, term, pi, pt, li, pi + 1);
            } else if (ph > 32) {
                put_op(c, gi, RAFTING_OP_TIMEOUT, 0, 0, 0, now, 0, 0, 0, 0, 0);
            } else put_op(c, gi, RAFTING_OP_NONE, 0, 0, 0, now, 0, 0, 0, 0, 0);
            return;
        }
    }
    // leader work (classes 0..7, re-elected churn groups, catch-up groups)
    leader_cell(w, c, step, r, i);
    if (c.op_cd) { c.op_cd[gi].x = 0; c.op_cd[gi].y = 0; }
    if (c.op_e) c.op_e[gi] = 0;
    if (cls == 9u && role == RAFTING_ROLE_LEADER) {
        const uint32_t ph = (uint32_t)((tick + gid) % 128u);
        if (ph == 0 && c.commit_index && c.commit_index[lgid] > 0)
            put_op(c, gi, RAFTING_OP_FLUSH, 0, 0, 0, now, 0, c.commit_index[lgid], lt, 0, 0);
        if (ph >= 28 && w.F > 1) {                                           // lane 1 is unreachable: its RPCs time out
            const size_t l1 = gi * w.F + 1;
            if (c.ev_meta[l1] != 0) {
                c.ev_meta[l1] = (c.ev_meta[l1] & ~((uint64_t)0x70)) | ((uint64_t)RAFTING_OUT_ERROR << 4);
                c.ev_tn[l1].x = 0;
            }
        }
    }
}

__global__ void vote_kernel(rafting_wl_cfg_t w, Cols c, uint64_t step) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)w.rows * w.n) return;
    vote_cell(w, c, step, (uint32_t)(t / w.n), (uint32_t)(t % w.n));
}
__global__ void mixed_kernel(rafting_wl_cfg_t w, Cols c, uint64_t step) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)w.rows * w.n) return;
    mixed_cell(w, c, step, (uint32_t)(t / w.n), (uint32_t)(t % w.n));
}
__global__ void pool_kernel(int64_t* pool) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 256u * 50u) pool[t] = (int64_t)(t / 50u);
}

__global__ void leader_kernel(rafting_wl_cfg_t w, Cols c, uint64_t step) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t total = (uint64_t)w.rows * w.n;
    if (t >= total) return;
    leader_cell(w, c, step, (uint32_t)(t / w.n), (uint32_t)(t % w.n));
}
__global__ void election_kernel(rafting_wl_cfg_t w, Cols c, uint32_t phase) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.n) return;
    election_cell(w, c, phase, i);
}

Cols make_cols(const rafting_inbox_t* in, const rafting_outbox_t* prev) {
    Cols c;
    c.op_meta = (uint64_t*)in->op_meta; c.op_nr = (rafting_i64x2_t*)in->op_nr; c.op_ab = (rafting_i64x2_t*)in->op_ab;
    c.op_cd = (rafting_i64x2_t*)in->op_cd; c.op_e = (int64_t*)in->op_e;
    c.ballot_meta = prev ? prev->ballot_meta : nullptr; c.ballot_term = prev ? prev->ballot_term : nullptr;
    c.commit_index = prev ? prev->commit_index : nullptr; c.last_entry = prev ? prev->last_entry : nullptr;
    c.ev_meta = (uint64_t*)in->ev_meta; c.ev_tn = (rafting_i64x2_t*)in->ev_tn; c.ev_el = (rafting_i64x2_t*)in->ev_el;
    c.plan_meta = prev ? prev->plan_meta : nullptr; c.plan_lc = prev ? prev->plan_lc : nullptr;
    c.plan_epoch = prev ? prev->plan_epoch : nullptr; c.current_term = prev ? prev->current_term : nullptr;
    c.incarnation = prev ? prev->incarnation : nullptr; c.role_word = prev ? prev->role_word : nullptr;
    c.gids = in->gids; c.compact = in->gids && (in->flags & RAFTING_INBOX_COMPACT_GROUPS);
    return c;
}

}  // namespace

extern "C" int rafting_wl_leader_step(const rafting_wl_cfg_t* w, uint64_t step, const rafting_outbox_t* prev_out,
                                      const rafting_inbox_t* in, int on_device, void* stream) {
    if (!w || !in || w->n == 0 || w->rows == 0) return RAFTING_E_INVAL;
    Cols c = make_cols(in, prev_out);
    if (on_device) {
        const uint64_t total = (uint64_t)w->rows * w->n;
        leader_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*w, c, step);
        return cudaGetLastError() == cudaSuccess ? RAFTING_OK : RAFTING_E_CUDA;
    }
    for (uint32_t r = 0; r < w->rows; r++)
        for (uint32_t i = 0; i < w->n; i++) leader_cell(*w, c, step, r, i);
    return RAFTING_OK;
}

extern "C" int rafting_wl_election_step(const rafting_wl_cfg_t* w, uint32_t phase, const rafting_outbox_t* prev_out,
                                        const rafting_inbox_t* in, int on_device, void* stream) {
    if (!w || !in || w->n == 0 || phase > 2) return RAFTING_E_INVAL;
    if (phase > 0 && (!prev_out || !prev_out->incarnation || !prev_out->current_term)) return RAFTING_E_INVAL;
    Cols c = make_cols(in, prev_out);
    if (on_device) {
        election_kernel<<<(w->n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*w, c, phase);
        return cudaGetLastError() == cudaSuccess ? RAFTING_OK : RAFTING_E_CUDA;
    }
    for (uint32_t i = 0; i < w->n; i++) election_cell(*w, c, phase, i);
    return RAFTING_OK;
}

typedef void (*cell_fn)(const rafting_wl_cfg_t&, const Cols&, uint64_t, uint32_t, uint32_t);
static int run_cells(const rafting_wl_cfg_t* w, uint64_t step, const rafting_outbox_t* prev_out, const rafting_inbox_t* in,
                     int on_device, void* stream, int which) {
    if (!w || !in || w->n == 0 || w->rows == 0 || !in->op_meta || !in->op_nr || !in->op_ab || !in->ev_meta || !in->ev_tn || !in->ev_el)
        return RAFTING_E_INVAL;
    Cols c = make_cols(in, prev_out);
    if (on_device) {
        const uint64_t total = (uint64_t)w->rows * w->n;
        if (which == 0) vote_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*w, c, step);
        else mixed_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*w, c, step);
        return cudaGetLastError() == cudaSuccess ? RAFTING_OK : RAFTING_E_CUDA;
    }
    for (uint32_t r = 0; r < w->rows; r++)
        for (uint32_t i = 0; i < w->n; i++) { if (which == 0) vote_cell(*w, c, step, r, i); else mixed_cell(*w, c, step, r, i); }
    return RAFTING_OK;
}
extern "C" int rafting_wl_vote_step(const rafting_wl_cfg_t* w, uint64_t step, const rafting_outbox_t* prev_out,
                                    const rafting_inbox_t* in, int on_device, void* stream) {
    return run_cells(w, step, prev_out, in, on_device, stream, 0);
}
extern "C" int rafting_wl_mixed_step(const rafting_wl_cfg_t* w, uint64_t step, const rafting_outbox_t* prev_out,
                                     const rafting_inbox_t* in, int on_device, void* stream) {
    return run_cells(w, step, prev_out, in, on_device, stream, 1);
}
extern "C" int rafting_wl_fill_term_pool(int64_t* pool, uint32_t capacity, int on_device, void* stream) {
    if (!pool || capacity < RAFTING_WL_POOL_TERMS) return RAFTING_E_INVAL;
    if (on_device) { pool_kernel<<<(RAFTING_WL_POOL_TERMS + 255) / 256, 256, 0, (cudaStream_t)stream>>>(pool); return cudaGetLastError() == cudaSuccess ? RAFTING_OK : RAFTING_E_CUDA; }
    for (uint32_t t = 0; t < RAFTING_WL_POOL_TERMS; t++) pool[t] = (int64_t)(t / 50u);
    return RAFTING_OK;
}
