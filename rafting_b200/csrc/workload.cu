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
    uint64_t* op_meta; rafting_i64x2_t* op_nr; rafting_i64x2_t* op_ab;
    uint64_t* ev_meta; rafting_i64x2_t* ev_tn; rafting_i64x2_t* ev_el;
    // previous outbox (read)
    const uint64_t* plan_meta; const rafting_i64x2_t* plan_lc; const int64_t* plan_epoch;
    const int64_t* current_term; const uint32_t* incarnation; const uint32_t* role_word;
    const uint32_t* gids;
};

RAFTING_HD inline uint64_t key(uint64_t seed, uint32_t gid, uint64_t tick, uint32_t lane, uint32_t salt) {
    uint64_t h = rafting_splitmix64(seed ^ ((uint64_t)gid * 0x9E3779B97F4A7C15ull));
    h = rafting_splitmix64(h ^ (tick * 0xC2B2AE3D27D4EB4Full));
    return rafting_splitmix64(h ^ ((uint64_t)lane << 32) ^ salt);
}

RAFTING_HD inline void leader_cell(const rafting_wl_cfg_t& w, const Cols& c, uint64_t step, uint32_t r, uint32_t i) {
    const uint32_t gid = c.gids ? c.gids[i] : w.gid_base + i;    // global group id keys the RNG
    const uint32_t lgid = c.gids ? c.gids[i] : i;                // local id indexes the snapshot columns
    const uint64_t tick = step * w.rows + r;
    const size_t gi = (size_t)r * w.n + i;
    const int64_t now = w.t0 + 10 * (int64_t)tick + (int64_t)(gid % 10u);
    if (c.op_meta) {
        const uint32_t a = (uint32_t)(key(w.seed, gid, tick, 0, 1) % (uint64_t)(w.max_submit + 1));
        c.op_meta[gi] = a ? RAFTING_OP_MAKE(RAFTING_OP_SUBMIT, 0, a) : RAFTING_OP_MAKE(RAFTING_OP_TIMEOUT, 0, 0);
        c.op_nr[gi].x = now; c.op_nr[gi].y = 0;
        c.op_ab[gi].x = 0; c.op_ab[gi].y = 0;
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
    const uint32_t lgid = c.gids ? c.gids[i] : i;
    const size_t gi = i;                                           // single row
    const int64_t now = w.t0 - 1000 + 10 * (int64_t)phase + (int64_t)(gid % 10u);
    if (c.op_meta) {
        c.op_meta[gi] = phase == 0 ? RAFTING_OP_MAKE(RAFTING_OP_TIMEOUT, 0, 0) : 0;
        c.op_nr[gi].x = now; c.op_nr[gi].y = 0; c.op_ab[gi].x = 0; c.op_ab[gi].y = 0;
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
    c.ev_meta = (uint64_t*)in->ev_meta; c.ev_tn = (rafting_i64x2_t*)in->ev_tn; c.ev_el = (rafting_i64x2_t*)in->ev_el;
    c.plan_meta = prev ? prev->plan_meta : nullptr; c.plan_lc = prev ? prev->plan_lc : nullptr;
    c.plan_epoch = prev ? prev->plan_epoch : nullptr; c.current_term = prev ? prev->current_term : nullptr;
    c.incarnation = prev ? prev->incarnation : nullptr; c.role_word = prev ? prev->role_word : nullptr;
    c.gids = in->gids;
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
