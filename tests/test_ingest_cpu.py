"""Transport framing (include/rafting_ingest.h): the reference's EventCodec frame layout
(transport/EventCodec.java:169-201 encoder, :222-334 decoder) restated in C.  The expected bytes below are written out by
hand from the layout constants of EventCodec.java:25-41, not produced by the code under test."""
import os
import struct

import numpy as np
import pytest

from rafting_b200 import abi, ingest

SOH, STX, ETX, EOT = b"\x01", b"\x02", b"\x03", b"\x04"


def hand_frame(ftype, head, body, seq=None, ending=False):
    """|SOH|TYPE|{SEQUENCE}|STX|HEAD_LEN|HEAD|BODY_LEN|{BODY}|ETX[|EOT] — EventCodec.java:297, big-endian ints (ByteBuf.writeInt)"""
    out = SOH + bytes([ftype])
    if seq is not None:
        out += struct.pack(">i", seq)
    out += STX + struct.pack(">i", len(head)) + head + struct.pack(">i", len(body)) + body + ETX
    return out + (EOT if ending else b"")


def test_encoder_is_byte_exact_against_the_hand_written_layout():
    cases = [(ingest.ENQ, b"appendEntries:ctx-7", b"\x01\x00kryo-bytes", 42, False),
             (ingest.ACK, b"requestVote:group/with:colon", b"", -5, False),          # sequence is a signed int, body may be empty
             (ingest.SYN, b"node-1:8080", b"", None, False),
             (ingest.PM, b"snap@ctx", b"", None, True)]                              # Event.Ending -> trailing EOT
    for ftype, head, body, seq, ending in cases:
        assert ingest.encode(ftype, head, body, seq, ending) == hand_frame(ftype, head, body, seq, ending)
    # the single-frame overhead FrameEncoder.allocateBuffer budgets for (12 + head + body, :206) holds for StrEvents
    assert len(ingest.encode(ingest.SYN, b"abc")) == 12 + 3


def test_scan_round_trip_and_partial_input():
    frames = [hand_frame(ingest.ENQ, b"appendEntries:g%d" % k, bytes([k]) * (k * 3), seq=1000 + k) for k in range(5)]
    frames.append(hand_frame(ingest.SYN, b"hello", b""))
    stream = b"".join(frames)
    rc, fr, used, tr = ingest.scan(stream)
    assert rc == 0 and used == len(stream) and not tr and len(fr) == 6
    for k in range(5):
        f = fr[k]
        assert f["type"] == ingest.ENQ and f["has_sequence"] == 1 and f["sequence"] == 1000 + k
        assert stream[f["head_off"]:f["head_off"] + f["head_len"]] == b"appendEntries:g%d" % k
        assert stream[f["body_off"]:f["body_off"] + f["body_len"]] == bytes([k]) * (k * 3)
    assert fr[5]["type"] == ingest.SYN and fr[5]["has_sequence"] == 0
    # every prefix: only complete frames are consumed, the rest stays for the next read (ByteToMessageDecoder cumulation)
    bounds = np.cumsum([len(f) for f in frames])
    for cut in range(len(stream) + 1):
        rc, fr, used, tr = ingest.scan(stream[:cut])
        assert rc == 0 and not tr
        assert used == ([0] + list(bounds[bounds <= cut]))[-1] and len(fr) == int((bounds <= cut).sum())
    # re-encoding what was scanned gives the same bytes back
    rc, fr, used, _ = ingest.scan(stream)
    again = b"".join(ingest.encode(int(f["type"]), stream[f["head_off"]:f["head_off"] + f["head_len"]],
                                   stream[f["body_off"]:f["body_off"] + f["body_len"]],
                                   int(f["sequence"]) if f["has_sequence"] else None) for f in fr)
    assert again == stream


def test_eot_switches_to_transparent_and_malformed_input_is_refused():
    a = hand_frame(ingest.PM, b"snap", b"", ending=True)
    rc, fr, used, tr = ingest.scan(a + b"raw snapshot bytes")
    assert rc == 0 and len(fr) == 1 and fr[0]["ending"] == 1 and tr and used == len(a)
    rc, fr, used, tr = ingest.scan(EOT + b"xyz")                          # EOT at a frame start (verify(buf, SOH, EOT), :299)
    assert rc == 0 and len(fr) == 0 and tr and used == 1
    good = hand_frame(ingest.ACK, b"preVote:c", b"\x07", seq=3)
    for bad in (b"\x09" + good,                                                        # neither SOH nor EOT
                good[:6] + b"\x09" + good[7:],                                         # STX missing after the sequence
                good[:-1] + b"\x09",                                                   # ETX missing
                SOH + bytes([ingest.SYN]) + STX + struct.pack(">i", 129) + b"x" * 140,  # head longer than MAX_HEAD_SIZE = 128
                SOH + bytes([ingest.SYN]) + STX + struct.pack(">i", -1) + b"x" * 8,
                SOH + bytes([ingest.SYN]) + STX + struct.pack(">i", 1) + b"h" + struct.pack(">i", (1 << 26) + 1) + b"y" * 8):
        rc, fr, used, tr = ingest.scan(good + bad)
        assert rc == -1 and len(fr) == 1 and used == len(good)             # the frame before the error is still delivered
    with pytest.raises(ValueError):
        ingest.encode(ingest.SYN, b"h" * 129)


def test_scope_and_context_registry():
    assert ingest.scope_parse(b"appendEntries:ctx-1") == (abi.OP_AE_REQUEST, b"ctx-1")
    assert ingest.scope_parse(b"preVote:a:b") == (abi.OP_PREVOTE_REQ, b"a:b")
    assert ingest.scope_parse(b"requestVote:") == (abi.OP_VOTE_REQ, b"")
    assert ingest.scope_parse(b"installSnapshot:@raft") == (abi.OP_IS_REQUEST, b"@raft")
    for bad in (b"obtainSnapshot:x", b"appendEntries", b""):
        with pytest.raises(ValueError):
            ingest.scope_parse(bad)
    m = ingest.CtxMap()
    for k in range(1000):
        m.put(b"ctx-%d" % k, 5000 + k)
    assert m.get(b"ctx-0") == 5000 and m.get(b"ctx-999") == 5999 and m.get(b"ctx-1000") is None


def test_batch_records_land_in_the_inbox_columns():
    rows, n, F = 3, 16, 2
    ib = abi.Inbox(rows, n, F)
    recs = np.zeros(5, dtype=ingest.BATCH_REC)
    recs[0] = (3, abi.EV_AE_ACK, 1, abi.OUT_OK | 4, 0, 77, 0, 9, 100, 140)
    recs[1] = (3, abi.EV_AE_ACK, 0, abi.OUT_OK, 0, 77, 0, 9, 100, 141)               # success = false
    recs[2] = (15, abi.EV_RV_REPLY, 1, abi.OUT_ERROR, 2, 5, 0, 0, 0, 0)
    recs[3] = (3, abi.EV_AE_ACK, 1, abi.OUT_OK | 4, 0, 77, 0, 9, 100, 150)           # slot (0, 3, 1) is taken: stops here
    recs[4] = (4, abi.EV_IS_ACK, 0, abi.OUT_OK | 4, 1, 1, 0, 2, 30, 30)
    rc, done = ingest.batch_to_inbox(recs, 1234, ib)
    assert rc == -1 and done == 3
    assert ib.ev_meta[0, 3, 1] == abi.evm_make(abi.EV_AE_ACK, abi.OUT_OK, True, 77)
    assert ib.ev_meta[0, 3, 0] == abi.evm_make(abi.EV_AE_ACK, abi.OUT_OK, False, 77)
    assert tuple(ib.ev_tn[0, 3, 1]) == (9, 1234) and tuple(ib.ev_el[0, 3, 1]) == (100, 140)
    assert ib.ev_meta[2, 15, 1] == abi.evm_make(abi.EV_RV_REPLY, abi.OUT_ERROR, False, 5)
    rc, done = ingest.batch_to_inbox(recs[4:], 1300, ib)
    assert rc == 0 and done == 1 and tuple(ib.ev_el[1, 4, 0]) == (30, 30)
    bad = recs[:1].copy(); bad["gid"] = 16
    assert ingest.batch_to_inbox(bad, 0, ib)[0] == -1


# ---- reply bodies: kryo.writeClassAndObject(RaftResponse) restated (parity unpinned: no JVM here; oracle/java/KryoGen.java) ----
def _kryo_response(term: int, success: bool) -> bytes:
    """An independent restatement, straight from the format notes in include/rafting_ingest.h."""
    name = b"io.lubricant.consensus.raft.RaftResponse"
    out = bytearray([0x01, 0x00]) + name[:-1] + bytes([name[-1] | 0x80]) + bytes([0x01, 1 if success else 0])
    z = ((term << 1) ^ (term >> 63)) & 0xFFFFFFFFFFFFFFFF
    for k in range(9):
        if k == 8:
            out.append(z & 0xFF)
            break
        if z >> 7 == 0:
            out.append(z)
            break
        out.append((z & 0x7F) | 0x80)
        z >>= 7
    return bytes(out)


def test_reply_body_is_the_kryo_layout_of_a_raft_response():
    assert ingest.reply_body_encode(5, True) == \
        b"\x01\x00io.lubricant.consensus.raft.RaftRespons" + bytes([ord("e") | 0x80, 0x01, 0x01, 0x0A])
    assert ingest.reply_body_encode(0, False)[-3:] == b"\x01\x00\x00" and len(ingest.reply_body_encode(0, False)) == 45
    assert ingest.reply_body_encode(-1, True)[-1] == 0x01                            # zig-zag: -1 -> 1
    assert ingest.reply_body_encode(64, True)[-2:] == b"\x80\x01"                     # 128 after zig-zag: two bytes
    rng = np.random.default_rng(7)
    terms = [0, 1, -1, 63, 64, -64, -65, 2**31, 2**55, 2**56 - 1, 2**62, 2**63 - 1, -2**63] + \
        [int(x) for x in rng.integers(-2**63, 2**63 - 1, 200, dtype=np.int64)] + [int(x) for x in rng.integers(0, 1 << 20, 200)]
    for t in terms:
        for ok in (False, True):
            b = ingest.reply_body_encode(t, ok)
            assert b == _kryo_response(t, ok), (t, ok)
            assert ingest.reply_body_decode(b) == (t, ok)
    assert len(ingest.reply_body_encode(-2**63, True)) == 53 == max(len(ingest.reply_body_encode(t, True)) for t in terms)


def test_reply_body_decoder_accepts_nothing_but_that_shape():
    good = ingest.reply_body_encode(1234567, True)
    assert ingest.reply_body_decode(good) == (1234567, True)
    for cut in range(len(good)):
        assert ingest.reply_body_decode(good[:cut]) is None, cut                        # every truncation
    assert ingest.reply_body_decode(good + b"\x00") is None                             # trailing byte
    for pos, val in ((0, 0x02), (1, 0x01), (10, ord("X")), (41, ord("e")), (42, 0x02), (42, 0x00), (43, 0x02)):
        bad = bytearray(good); bad[pos] = val
        assert ingest.reply_body_decode(bytes(bad)) is None, pos                        # registered id, nameId, class, flag, ref, bool
    ten = good[:44] + b"\x80" * 9 + b"\x01"
    assert ingest.reply_body_decode(ten) is None                                        # a varlong never has a tenth byte
    nine = good[:44] + b"\xff" * 9                                                      # nine bytes: the last carries 8 bits
    assert ingest.reply_body_decode(nine) == (-2**63, True) and nine == ingest.reply_body_encode(-2**63, True)


def test_ack_frames_become_lane_event_fields():
    cm = ingest.CtxMap()
    cm.put(b"ctx-7", 7), cm.put(b"ctx-9", 9)
    kinds = {b"appendEntries": abi.EV_AE_ACK, b"installSnapshot": abi.EV_IS_ACK, b"preVote": abi.EV_PV_REPLY, b"requestVote": abi.EV_RV_REPLY}
    stream, want = b"", []
    for i, (m, k) in enumerate(kinds.items()):
        gid = 7 if i % 2 == 0 else 9
        stream += ingest.encode(ingest.ACK, m + b":ctx-%d" % gid, ingest.reply_body_encode(100 + i, i % 2 == 1), sequence=1000 + i)
        want.append((gid, k, 1000 + i, 100 + i, i % 2 == 1))
    stream += ingest.encode(ingest.ACK, b"appendEntries:nobody", ingest.reply_body_encode(1, True), sequence=1)     # unknown context
    stream += ingest.encode(ingest.ENQ, b"appendEntries:ctx-7", b"opaque", sequence=2)                                # a request
    stream += ingest.encode(ingest.ACK, b"appendEntries:ctx-7", b"\x01\x00not-a-response", sequence=3)                # another body
    rc, frames, used, _ = ingest.scan(stream)
    assert rc == 0 and used == len(stream) and len(frames) == 7
    got = [ingest.ack_frame_decode(stream, f, cm) for f in frames]
    assert got[:4] == want and got[4:] == [None, None, None]
    recs = ingest.ack_frames_decode(stream, frames, cm)                                 # the whole buffer in one call
    assert [(int(r["gid"]), int(r["kind"]), int(r["sequence"]), int(r["term"]), bool(r["success"])) for r in recs] == want
    assert recs["frame"].tolist() == [0, 1, 2, 3]


UPSTREAM_KRYO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "upstream_kryo.json")


@pytest.mark.skipif(not os.path.exists(UPSTREAM_KRYO), reason="tests/golden/upstream_kryo.json absent: no JDK / kryo jar here (make -C oracle/java kryo)")
def test_reply_body_matches_the_bytes_the_reference_emits():
    import json
    for v in json.load(open(UPSTREAM_KRYO))["vectors"]:
        want = bytes.fromhex(v["hex"])
        assert ingest.reply_body_encode(v["term"], v["success"]) == want, v
        assert ingest.reply_body_decode(want) == (v["term"], v["success"]), v


def test_mutated_streams_never_crash_and_never_report_bytes_outside_the_buffer():
    """Fuzz: valid streams with random byte flips / truncations.  Whatever the scanner reports must lie inside the buffer
    and inside `consumed`; the ACK decoder must take or refuse each frame without touching anything else (run under
    ASan / UBSan for profiles/r2_host_sanitizers.txt)."""
    rng = np.random.default_rng(20260923)
    cm = ingest.CtxMap()
    for g in range(32):
        cm.put(b"c%d" % g, g)
    methods = [b"appendEntries", b"preVote", b"requestVote", b"installSnapshot", b"bogus"]
    for trial in range(300):
        parts = []
        for i in range(int(rng.integers(1, 12))):
            t = [ingest.ACK, ingest.ENQ, ingest.SYN, ingest.BATCH][int(rng.integers(0, 4))]
            head = methods[int(rng.integers(0, 5))] + b":c%d" % int(rng.integers(0, 40))
            body = ingest.reply_body_encode(int(rng.integers(-2**40, 2**40)), bool(rng.integers(0, 2))) if rng.random() < 0.7 \
                else bytes(rng.integers(0, 256, int(rng.integers(0, 60)), dtype=np.uint8))
            parts.append(ingest.encode(t, head, body, sequence=i if t in (ingest.ACK, ingest.ENQ) else None, ending=rng.random() < 0.05))
        data = bytearray(b"".join(parts))
        for _ in range(int(rng.integers(0, 4))):
            data[int(rng.integers(0, len(data)))] = int(rng.integers(0, 256))
        data = bytes(data[:int(rng.integers(1, len(data) + 1))])
        rc, frames, used, tr = ingest.scan(data, cap=64)
        assert rc in (0, -1) and used <= len(data)                                      # RAFTING_OK / RAFTING_E_INVAL
        for f in frames:
            assert f["head_off"] + f["head_len"] <= used and f["body_off"] + f["body_len"] <= used
            assert f["head_len"] <= ingest.MAX_HEAD and f["body_len"] <= ingest.MAX_BODY
        recs = ingest.ack_frames_decode(data, frames, cm)
        for r in recs:
            f = frames[int(r["frame"])]
            assert f["type"] == ingest.ACK and r["gid"] < 32 and r["kind"] in (abi.EV_AE_ACK, abi.EV_IS_ACK, abi.EV_PV_REPLY, abi.EV_RV_REPLY)
            assert ingest.reply_body_decode(data[f["body_off"]:f["body_off"] + f["body_len"]]) == (int(r["term"]), bool(r["success"]))


def test_replies_become_the_compact_words_the_python_encoder_writes():
    """ACK records + the pending-invocation table -> ev_c words / escape records of a compact inbox, against
    rafting_b200.compact.encode_inbox (the encoder the GPU parity tests of the compact path use) on the same events."""
    from rafting_b200 import compact
    rng = np.random.default_rng(99)
    rows, G, F, NPEER = 3, 64, 2, 2
    sent_term = rng.integers(5, 9, G).astype(np.int64)
    sent_inc = rng.integers(1, 4, G).astype(np.uint32)
    base = 1_000_000
    ib = abi.Inbox(rows, G, F)
    tags = np.full((rows, G, F), abi.CTAG_NONE, dtype=np.uint8)
    pend = ingest.Pending()
    per_call = {}                                                   # (row, peer, now) -> ack records, in arrival order
    seq = 0
    for r in range(rows):
        for g in range(G):
            for f in range(F):
                if rng.random() < 0.25:
                    continue
                seq += 1
                kind = [abi.EV_AE_ACK, abi.EV_AE_ACK, abi.EV_IS_ACK, abi.EV_RV_REPLY][int(rng.integers(0, 4))]
                tag = int(rng.integers(0, 32)) if rng.random() < 0.9 else abi.CTAG_NONE
                term = int(sent_term[g]) if rng.random() < 0.9 else int(sent_term[g]) + 1          # a newer term: travels in full
                ok = bool(rng.integers(0, 2))
                epoch, last = int(rng.integers(0, 50)), int(rng.integers(50, 90))
                peer = f                                             # lane f <-> one connection
                now = base + 1000 * r + (0 if peer == 0 else 7)      # the two peers' buffers are drained 7 ms apart
                # an escaped plan has no tag and may come from another role object: its reply carries that incarnation in full
                inc = int(sent_inc[g]) if tag != abi.CTAG_NONE else int(sent_inc[g]) + int(rng.integers(0, 2))
                pend.put(peer, seq, kind, g, f, tag, inc, int(sent_term[g]), epoch, last)
                rec = np.zeros(1, dtype=ingest.ACK_REC)[0]
                rec["gid"], rec["kind"], rec["success"], rec["sequence"], rec["term"] = g, kind, ok, seq, term
                per_call.setdefault((r, peer, now), []).append(rec)
                tags[r, g, f] = tag
                if kind in (abi.EV_AE_ACK, abi.EV_IS_ACK):
                    ib.ack(r, g, f, now, inc, term, ok, epoch, last, snapshot=kind == abi.EV_IS_ACK)
                else:
                    ib.vote_reply(r, g, f, now, inc, term, ok)
                    ib.ev_el[r, g, f] = (epoch, last)
    want = compact.encode_inbox(ib, tags, sent_term, sent_inc)
    got = compact.CompactInbox(rows, G, F)
    esc = np.zeros(rows * G * F, dtype=abi.CESC_IN)
    n_esc = 0
    assert len(pend) == seq
    for (r, peer, now), recs in sorted(per_call.items(), key=lambda kv: kv[0]):
        got.row_base[r] = base + 1000 * r                            # the row's base: the earliest drain of that row
        rc, n_esc, deferred, unknown = pend.acks_to_cinbox(peer, np.array(recs, dtype=ingest.ACK_REC), now, r, got, esc, n_esc)
        assert rc == 0 and len(deferred) == 0 and unknown == 0
    assert len(pend) == 0                                            # every invocation was completed and removed
    assert np.array_equal(got.row_base, want.row_base) and np.array_equal(got.ev_c, want.ev_c)
    a, b = np.sort(esc[:n_esc], order="slot"), np.sort(want.esc, order="slot")
    assert len(a) == len(b) > 0 and a.tobytes() == b.tobytes()
    assert ((got.ev_c & 0xF) == abi.CEV_ESCAPED).sum() == n_esc and (got.ev_c != 0).sum() == seq
    # a second reply for a lane slot that is taken is deferred (its invocation stays pending); an unknown sequence is dropped
    pend.put(0, 9001, abi.EV_AE_ACK, 3, 0, 1, int(sent_inc[3]), int(sent_term[3]), 1, 2)
    pend.put(0, 9002, abi.EV_AE_ACK, 3, 0, 2, int(sent_inc[3]), int(sent_term[3]), 1, 2)
    two = np.zeros(3, dtype=ingest.ACK_REC)
    two["gid"], two["kind"], two["success"], two["term"] = 3, abi.EV_AE_ACK, 1, int(sent_term[3])
    two["sequence"] = [9001, 9002, 7777]
    fresh = compact.CompactInbox(1, G, F)
    fresh.row_base[0] = base
    rc, n2, deferred, unknown = pend.acks_to_cinbox(0, two, base + 5, 0, fresh, esc, 0)
    assert rc == 0 and n2 == 0 and deferred.tolist() == [1] and unknown == 1 and len(pend) == 1
    assert fresh.ev_c[0, 3, 0] == (abi.EV_AE_ACK | (1 << 6) | (1 << 7) | (1 << 8) | (5 << 16))
    assert pend.remove(0, 9002) and not pend.remove(0, 9002) and len(pend) == 0


def test_pending_table_behaves_like_a_dict_under_churn():
    """put / complete / time-out in random order, growth from a tiny table: the open-addressing table (backward-shift deletion)
    must agree with a dict on every lookup — checked through acks_to_cinbox's "unknown sequence" count."""
    rng = np.random.default_rng(5)
    pend, model = ingest.Pending(4), {}
    G, F = 32, 2
    from rafting_b200 import compact
    esc = np.zeros(8, dtype=abi.CESC_IN)
    for step in range(4000):
        op = rng.random()
        peer, seq = int(rng.integers(0, 3)), int(rng.integers(-50, 400))
        if op < 0.5:
            g, f = int(rng.integers(0, G)), int(rng.integers(0, F))
            pend.put(peer, seq, abi.EV_AE_ACK, g, f, int(rng.integers(0, 32)), 1, 7, 0, 0)
            model[(peer, seq)] = (g, f)
        elif op < 0.75:
            assert pend.remove(peer, seq) == ((peer, seq) in model)
            model.pop((peer, seq), None)
        else:
            rec = np.zeros(1, dtype=ingest.ACK_REC)
            known = (peer, seq) in model
            rec["gid"], rec["kind"], rec["sequence"], rec["term"] = (model[(peer, seq)][0] if known else 0), abi.EV_AE_ACK, seq, 7
            cin = compact.CompactInbox(1, G, F)
            rc, _, deferred, unknown = pend.acks_to_cinbox(peer, rec, 10, 0, cin, esc, 0)
            assert rc == 0 and unknown == (0 if known else 1) and len(deferred) == 0
            if known:
                g, f = model.pop((peer, seq))
                assert cin.ev_c[0, g, f] != 0 and (cin.ev_c != 0).sum() == 1
        assert len(pend) == len(model)
