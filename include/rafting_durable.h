/*
 * rafting_durable.h — batched durability records for the engine's outbox.  HOST code, no CUDA.
 *
 * What it replaces in the reference: every role object persists (currentTerm, votedFor) with its own fsync
 * (RaftMember.java:20-26 -> StableLock.persist(term, candidate), support/StableLock.java:69-79), and every snapshot
 * milestone with another (StableLock.persist(Snapshot), :81-90).  With one engine step covering thousands of groups the
 * pump thread needs ONE durable barrier per step instead: every group whose outbox role_word carries the persist-dirty
 * bit (bit 30) contributes a record to a write-ahead journal, the journal is fdatasync'ed once, and only then are the
 * step's replies released — the reference's persist-before-reply order, batched.
 *
 * Files in the journal directory:
 *   stable.tbl   fixed table, 32 bytes per group: milestone.index, milestone.term, term (int64 LE), ballot (int32 LE), flags
 *   stable.wal   append-only batches: { magic, n, seq, crc32c(records) } + n records { gid, kind, ballot, a, b }
 * Recovery = table + every complete batch of the journal in order; a torn tail (bad magic / crc / short read) ends it.
 * rafting_journal_checkpoint folds the journal into the table (write, fsync, truncate the journal).
 *
 * rafting_stable_image writes one group's state in the REFERENCE's StableLock file layout (big-endian, StableLock.java:
 * 52-67): [0,8) milestone.index  [8,16) milestone.term  [16,24) term  [24,28) id length  [28..) serialised candidate ID
 * (the Kryo bytes of the ID are opaque here: the caller passes the bytes of the node the ballot slot stands for).
 *
 * All functions return 0 or a negative rafting_status_t value (same numbering as rafting_b200.h).
 */
#ifndef RAFTING_DURABLE_H
#define RAFTING_DURABLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rafting_journal rafting_journal_t;

typedef struct rafting_stable {
    int64_t term;             /* Persistence.term                                  */
    int32_t ballot;           /* Persistence.ballot as node slot, -1 = null        */
    int32_t _pad;
    int64_t milestone_index;  /* Persistence.milestone (snapshot lastIncludedIndex) */
    int64_t milestone_term;
} rafting_stable_t;

int rafting_journal_open(const char* dir, uint32_t max_groups, rafting_journal_t** out);
int rafting_journal_close(rafting_journal_t* j);

/* One step's persist records + ONE fdatasync.  gids == NULL: position i is group i (dense step); with an active list,
 * role_word / current_term are indexed by gid unless `compact` != 0 (RAFTING_INBOX_COMPACT_GROUPS), then by position.
 * votedFor is taken from role_word bits 8..15 (slot + 1).  *n_records (optional) = records written (0 = no sync issued).
 * Failure contract: a batch is durable as a whole or absent.  If the write or the fdatasync fails (ENOSPC, EIO, ...) the file
 * is cut back to the pre-batch offset, the batch sequence number is not consumed and RAFTING_E_IO (-3) is returned: the step's
 * replies must NOT be released, the same step may be committed again.  If the cut-back fails too the journal is sticky-failed:
 * every later commit / milestone / checkpoint returns -3 until the journal is reopened (recovery then cuts the torn tail). */
int rafting_journal_commit_step(rafting_journal_t* j, const uint32_t* gids, uint32_t n, int compact,
                                const uint32_t* role_word, const int64_t* current_term, uint64_t* n_records);
/* StableLock.persist(Snapshot): durable before it returns */
int rafting_journal_milestone(rafting_journal_t* j, uint32_t gid, int64_t index, int64_t term);
/* StableLock.restore() */
int rafting_journal_restore(rafting_journal_t* j, uint32_t gid, rafting_stable_t* out);
/* fold the journal into the table, truncate it */
int rafting_journal_checkpoint(rafting_journal_t* j);
/* counters: batches committed, records written, fdatasync calls, current journal bytes */
int rafting_journal_stats(rafting_journal_t* j, uint64_t out[4]);
/* reference-format image of one group's stable state (see above); *len = bytes written */
int rafting_stable_image(rafting_journal_t* j, uint32_t gid, const void* id_bytes, uint32_t id_len,
                         void* out, uint32_t cap, uint32_t* len);
const char* rafting_durable_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
