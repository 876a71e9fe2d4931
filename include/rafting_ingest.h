/*
 * rafting_ingest.h — transport-side framing for the engine's pump thread.  HOST code, no CUDA.
 *
 * What it replaces in the reference: EventCodec.FrameDecoder / FrameEncoder
 * (src/main/java/io/lubricant/consensus/raft/transport/EventCodec.java:169-201,222-334) — the byte framing every Raft
 * RPC travels in — and the scope parsing of NettyNode.parseContextId (transport/NettyNode.java:92-107), so that a receive
 * buffer filled by the transport can be cut into frames and routed to group ids without creating one Java object per frame.
 *
 * Wire format, exactly as FrameEncoder writes it (all integers big-endian, Netty ByteBuf.writeInt):
 *     SOH(0x01) TYPE [SEQUENCE i32]  STX(0x02) HEAD_LEN i32  HEAD utf-8  BODY_LEN i32  BODY  ETX(0x03) [EOT(0x04)]
 *   TYPE      ENQ 0x05 PingEvent (RPC request) | ACK 0x06 PongEvent (RPC reply) | SYN 0x16 ShakeHandEvent |
 *             MW 0x95 WaitSnapEvent | PM 0x9E TransSnapEvent          (EventCodec.java:32-41)
 *   SEQUENCE  present for ENQ / ACK only (EventCodec.java:239-244)
 *   HEAD      ENQ / ACK: the scope "<RaftService method name>:<contextId>" (NettyNode.java:55-75); others: the message
 *   BODY      Kryo bytes of the RPC arguments / the RaftResponse — OPAQUE here (kryo 4.0.2 is not in the reference tree);
 *             the pump hands AppendEntries bodies to rafting_log_append and reply bodies to the Java-side decoder
 *   limits    HEAD_LEN <= 128, BODY_LEN <= 64 MiB (EventCodec.java:25-26); a violation, or a byte other than
 *             SOH/EOT at a frame start, STX after the type, ETX after the body, is the decoder's DecoderException:
 *             RAFTING_E_INVAL, the channel is to be closed (EventCodec.java:326-330)
 *   EOT       at a frame start switches the channel to transparent mode (snapshot stream): scanning stops there
 *
 * One extension, used only between two engines: TYPE SUB 0x1A carries a BATCH — one frame per peer and step instead of
 * one frame per group (Leader.java:216 sends one RPC per follower and group).  Its BODY is an array of fixed 40-byte
 * little-endian records (rafting_batch_rec_t); rafting_batch_to_inbox writes reply records straight into the SoA inbox
 * columns of a leased step.  A peer that does not know the type rejects the frame like any unknown type.
 */
#ifndef RAFTING_INGEST_H
#define RAFTING_INGEST_H

#include <stddef.h>
#include <stdint.h>

#include "rafting_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { RAFTING_FRAME_ENQ = 0x05, RAFTING_FRAME_ACK = 0x06, RAFTING_FRAME_SYN = 0x16, RAFTING_FRAME_MW = 0x95,
       RAFTING_FRAME_PM = 0x9E, RAFTING_FRAME_BATCH = 0x1A };
#define RAFTING_FRAME_MAX_HEAD 128u           /* EventCodec.java:25 */
#define RAFTING_FRAME_MAX_BODY (1u << 26)     /* EventCodec.java:26 */

typedef struct rafting_frame {
    uint8_t  type;
    uint8_t  has_sequence;
    uint8_t  ending;          /* an EOT follows the frame's ETX (Event.Ending) — consumed with the frame               */
    uint8_t  _pad;
    int32_t  sequence;
    uint32_t head_off, head_len;   /* offsets into the scanned buffer                                                  */
    uint32_t body_off, body_len;
} rafting_frame_t;

/* Cuts buf[0..len) into complete frames (at most cap).  *consumed = bytes of complete frames (the caller keeps the rest
 * for the next call, exactly like ByteToMessageDecoder's cumulation); *transparent = 1 when an EOT was met at a frame
 * start (the bytes after it belong to the snapshot stream).  Returns 0, or RAFTING_E_INVAL on malformed input with the
 * frames before the error reported. */
int rafting_frame_scan(const uint8_t* buf, size_t len, rafting_frame_t* out, uint32_t cap, uint32_t* n,
                       size_t* consumed, int* transparent);

/* Byte-for-byte what FrameEncoder writes.  Returns the frame's size, or 0 when cap is too small / the limits are exceeded. */
size_t rafting_frame_encode(uint8_t* dst, size_t cap, uint8_t type, int has_sequence, int32_t sequence,
                            const char* head, uint32_t head_len, const void* body, uint32_t body_len, int ending);

/* scope "<method>:<contextId>" -> RAFTING_OP_AE_REQUEST / _PREVOTE_REQ / _VOTE_REQ / _IS_REQUEST and the offset of the
 * context id inside head; RAFTING_E_INVAL for an unknown method ("unknown context", NettyNode.java:106) */
int rafting_scope_parse(const char* head, uint32_t head_len, uint32_t* op_kind, uint32_t* ctx_off);

/* context id -> gid registry (ContextManager keeps the same map by name, ContextManager.java:57-106) */
typedef struct rafting_ctxmap rafting_ctxmap_t;
int rafting_ctxmap_create (rafting_ctxmap_t** out);
int rafting_ctxmap_destroy(rafting_ctxmap_t* m);
int rafting_ctxmap_put    (rafting_ctxmap_t* m, const char* ctx, uint32_t len, uint32_t gid);
int rafting_ctxmap_get    (const rafting_ctxmap_t* m, const char* ctx, uint32_t len, uint32_t* gid);   /* -1 if unknown */

/* ---- engine-to-engine batch records (TYPE 0x1A) ---- */
typedef struct rafting_batch_rec {
    uint32_t gid;             /* both ends open a context under the same gid (rafting_group_open takes the gid)         */
    uint8_t  kind;            /* RAFTING_EV_* : reply to an RPC of that kind                                            */
    uint8_t  lane;            /* follower lane of the RECEIVER the reply belongs to                                     */
    uint8_t  flags;           /* bits 0..1 outcome (RAFTING_OUT_*), bit 2 success                                       */
    uint8_t  row;             /* row of the receiver's step the event goes to (assigned by the receiver's pump)         */
    uint32_t incarnation;     /* echo of plan_meta / ballot_meta bits 32..63                                            */
    uint32_t _pad;
    int64_t  term;            /* RaftResponse.term()                                                                    */
    int64_t  epoch_at_send;   /* echo of plan_epoch                                                                     */
    int64_t  last_at_send;    /* echo of plan_lc.x                                                                      */
} rafting_batch_rec_t;      /* 40 bytes */

/* Writes n reply records into the lane-event columns of a (leased, host) inbox of `rows` x `n_groups` x F: ev_meta, ev_tn =
 * (term, now_ms), ev_el = (epoch, last).  Dense steps only (position == gid).  A slot that is already occupied, a gid /
 * lane / row out of range -> RAFTING_E_INVAL with *n_done records written (the caller defers the rest to the next step:
 * one event per (row, group, lane)). */
int rafting_batch_to_inbox(const rafting_batch_rec_t* recs, uint32_t n, int64_t now_ms, const rafting_inbox_t* in,
                           uint32_t n_groups, uint32_t F, uint32_t* n_done);

#ifdef __cplusplus
}
#endif
#endif
