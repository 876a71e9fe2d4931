/*
 * raft_oracle.h — CPU restatement of curioloop/rafting's per-context EventLoop path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in rafting_b200/ (the product) may include, link or call
 * this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, as the checker / the timed CPU baseline.
 *
 * PARITY UNPINNED BY UPSTREAM: the reference is 100 % Java (no JDK in this image, so it cannot be
 * run here) and its own tests hold no golden vector / known-answer test for this path
 * (SURVEY.md §4, §8c).  The only upstream known-answer material is the comment table at
 * M/context/member/Leadership.java:120-126, which tests/test_oracle_kat.py checks.  Everything
 * else is pinned by line-by-line transliteration (each function cites the lines it follows) and
 * hand-derived KATs per branch.
 *
 * It speaks the same batch format as the engine (include/rafting_b200.h) so a test can hand the
 * very same inbox to both and compare outboxes and exported state byte for byte.
 */
#ifndef RAFT_ORACLE_H
#define RAFT_ORACLE_H

#include "../include/rafting_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_engine orc_engine_t;

orc_engine_t* orc_create(const rafting_cfg_t* cfg);
void          orc_destroy(orc_engine_t* e);
int           orc_group_open(orc_engine_t* e, uint32_t gid, const rafting_group_init_t* init);
int           orc_group_open_bulk(orc_engine_t* e, uint32_t first, uint32_t count,
                                  const rafting_group_init_t* inits);
int           orc_group_load_runs(orc_engine_t* e, uint32_t gid, const rafting_i64x2_t* runs, uint32_t n);
int           orc_group_close(orc_engine_t* e, uint32_t gid);

/* one batch, canonical serial order, all pointers are host pointers.
   threads <= 1: single thread.  threads = T: groups round-robined over T loop threads exactly
   like M/support/EventLoopGroup.java:77-80 binds contexts to loops (ContextManager.java:46 uses 3). */
int           orc_step(orc_engine_t* e, const rafting_inbox_t* in, const rafting_outbox_t* out,
                       int threads);

int           orc_state_export(orc_engine_t* e, uint32_t gid, rafting_group_state_t* out);
int           orc_log_term(orc_engine_t* e, uint32_t gid, int64_t index, int64_t* term);
uint64_t      orc_events_processed(orc_engine_t* e);

/* stand-alone pieces exposed for known-answer tests */
int           orc_state_apply(int64_t st[10], int op, const int64_t* args, int64_t ret[2]);   /* Leadership.State methods, test hook */
void          orc_major_indices(const int64_t* match, int n, int64_t out[2]); /* Leadership.java:116-130 */
int64_t       orc_backoff_step(int32_t recent_rejection);  /* round(ln(e + r)), Leadership.java:105 (libm) */
int           orc_is_better(int new_role, int64_t new_term, int new_ballot,
                            int cur_role, int64_t cur_term, int cur_ballot, int cur_null); /* Membership.java:74-108 */
int64_t       orc_draw(uint64_t seed, uint32_t gid, uint32_t incarnation, int64_t election_ms);

#ifdef __cplusplus
}
#endif
#endif
