/*
 * rafting_workload.h — synthetic replication streams for tests and bench (rafting_b200/csrc/workload.cu).
 * Not part of the drop-in boundary: the reference has no counterpart; these entry
 * points play the remote peers of every group so the BASELINE.json configs can be driven at full size.
 */
#ifndef RAFTING_WORKLOAD_H
#define RAFTING_WORKLOAD_H
#include "rafting_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct rafting_wl_cfg {
    uint64_t seed;
    uint32_t rows;          /* inbox rows per step (ticks per step)                       */
    uint32_t n;             /* groups in the step (dense: the shard's group count)        */
    uint32_t F;             /* follower lanes                                             */
    uint32_t gid_base;      /* global id of local group 0 (shard offset) — keys the RNG   */
    uint32_t max_submit;    /* a in U{0..max_submit} entries per tick                     */
    uint32_t p_reject_ppm;  /* ok + success=false                                         */
    uint32_t p_error_ppm;   /* RPC error / timeout                                        */
    uint32_t p_cancel_ppm;  /* canceled                                                   */
    int64_t  t0;            /* wall clock of tick 0, ms                                   */
    uint32_t local_slot;    /* this node's slot (peer ids of inbound requests avoid it)   */
    uint32_t _pad;
} rafting_wl_cfg_t;

#define RAFTING_WL_POOL_TERMS (256u * 50u)   /* entry-term pool of the mixed stream: 50 copies of every term < 256 */

/* fills in->{op_meta,op_nr,op_ab,ev_meta,ev_tn,ev_el} (those that are non-NULL) for step `step`
   from the previous step's outbox (NULL: no acks).  on_device != 0: all pointers are device
   pointers and the generator runs as a kernel on `stream`.  With in->gids the outbox's per-group columns
   are read by gid, or by position when in->flags has RAFTING_INBOX_COMPACT_GROUPS (prev_out must come
   from a step with the same flag). */
int rafting_wl_leader_step(const rafting_wl_cfg_t* w, uint64_t step, const rafting_outbox_t* prev_out,
                           const rafting_inbox_t* in, int on_device, void* stream);

/* single-row election warm-up: phase 0 TIMEOUT everywhere, 1 grant every PreVote, 2 grant every RequestVote */
int rafting_wl_election_step(const rafting_wl_cfg_t* w, uint32_t phase, const rafting_outbox_t* prev_out,
                             const rafting_inbox_t* in, int on_device, void* stream);

/* config #3 — RequestVote storm with PreVote; config #5 — mixed leader churn + InstallSnapshot catch-up.
   Both need in->op_cd / in->op_e as well; the mixed stream also needs in->ent_terms filled by
   rafting_wl_fill_term_pool and in->ent_count = RAFTING_WL_POOL_TERMS. */
int rafting_wl_vote_step (const rafting_wl_cfg_t* w, uint64_t step, const rafting_outbox_t* prev_out,
                          const rafting_inbox_t* in, int on_device, void* stream);
int rafting_wl_mixed_step(const rafting_wl_cfg_t* w, uint64_t step, const rafting_outbox_t* prev_out,
                          const rafting_inbox_t* in, int on_device, void* stream);
int rafting_wl_fill_term_pool(int64_t* pool, uint32_t capacity, int on_device, void* stream);

#ifdef __cplusplus
}
#endif
#endif
