"""The compact wire format is lossless — checked on CPU with a numpy restatement of the engine's unpack / pack kernels
(tests/compact_model.py) around the oracle: every dense outbox must survive pack -> decode, every dense inbox encode -> unpack,
through the settling phase (thousands of escapes), steady state (none), unavailable followers, a late reply and a tag shortage."""
import numpy as np

from oracle import binding
from rafting_b200 import abi, compact, workload
from tests import compact_model as model
from tests import harness


def _fields_equal(a: abi.Inbox, b: abi.Inbox):
    ek = (a.ev_meta & np.uint64(0xF)) != 0
    assert np.array_equal(a.ev_meta, b.ev_meta)
    for name in ("ev_tn", "ev_el"):
        assert np.array_equal(getattr(a, name)[ek], getattr(b, name)[ek]), name
    ok = (a.op_meta & np.uint64(0xFF)) != 0
    assert np.array_equal(a.op_meta, b.op_meta) and np.array_equal(a.op_nr["x"][ok], b.op_nr["x"][ok])
    ua = a.op_ab["x"] if a.op_ab is not None else 0
    ub = b.op_ab["x"] if b.op_ab is not None else 0
    assert np.array_equal(np.broadcast_to(ua, a.op_meta.shape), np.broadcast_to(ub, a.op_meta.shape))


def test_codec_and_kernel_model_are_lossless_around_the_oracle():
    G, R, rows = 96, 3, 8
    F = R - 1
    cfg = abi.make_cfg(replicas=R, max_groups=G, max_rows=rows)
    o = binding.Oracle(cfg)
    o.open_bulk(0, harness.init_array(G, terms=np.arange(G) % 7))
    w1 = workload.make_wl(17, 1, G, F)
    w = workload.make_wl(17, rows, G, F, p_reject_ppm=60_000, p_error_ppm=30_000, p_cancel_ppm=20_000)
    harness.elect_all(o, w1)
    st = model.InFlight(G, F)
    g_term = np.array([o.export(g).current_term for g in range(G)], dtype=np.int64)
    prev, tags, sent_term, sent_inc = None, None, None, None
    escapes_late = 0
    for k in range(14):
        ib = workload.leader_inbox_host(w, k, prev)
        if k % 3 == 2:
            ib.op_ab["x"][:, ::7] = 2                                  # follower lane 1 unavailable for every seventh group
        if k == 9:
            ev = np.argwhere((ib.ev_meta & np.uint64(0xF)) != 0)[0]
            ib.ev_tn["y"][tuple(ev)] += 80_000                         # a reply 80 s late: escape record
        ci = compact.encode_inbox(ib, tags, sent_term, sent_inc)
        back = model.unpack(ci, g_term, st)
        _fields_equal(ib, back)
        if k >= 10:
            escapes_late += len(ci.esc)
        dense = o.step(ib)
        epoch = np.zeros(G, dtype=abi.I64X2)
        for g in range(G):
            s = o.export(g); epoch[g] = (s.epoch_index, s.epoch_term)
        co = model.pack(dense, epoch, st, esc_cap=rows * G * F + 2 * rows * G)
        harness.assert_outbox_equal(dense, compact.decode_outbox(co), where=f"pack -> decode, step {k}")
        if k >= 10:
            escapes_late += int(co.counts[0])
        prev, tags, sent_term, sent_inc = dense, co.tags(), co.current_term.copy(), co.incarnation.copy()
        g_term = dense.current_term.copy()
    assert (st.bits != 0).any() and int(np.bitwise_count(st.bits).max()) <= 21          # IN_FLIGHT_LIMIT + 1 RPCs outstanding per lane
    assert escapes_late < 40                                                      # steady state: only the irregular records


def test_a_lane_without_a_free_tag_falls_back_to_escape_records():
    G, F, rows = 2, 2, 1
    st = model.InFlight(G, F)
    st.bits[:] = 0xFFFFFFFF                                                      # every tag of every lane taken
    dense = abi.Outbox(rows, G, F, G)
    dense.current_term[:] = 3; dense.incarnation[:] = 5; dense.commit_index[:] = 40; dense.last_entry["x"] = 50; dense.last_entry["y"] = 3
    dense.plan_meta[0, 0, 0] = abi.PLAN_AE | (2 << 16) | (5 << 32)
    dense.plan_pp[0, 0, 0] = (45, 3); dense.plan_lc[0, 0, 0] = (47, 38); dense.plan_epoch[0, 0, 0] = 0
    co = model.pack(dense, np.zeros(G, dtype=abi.I64X2), st, esc_cap=8)
    assert int(co.tags()[0, 0, 0]) == abi.CTAG_NONE and int(co.counts[0]) == 0  # still a compact plan, just untagged
    harness.assert_outbox_equal(dense, compact.decode_outbox(co))
    ib = abi.Inbox(rows, G, F)
    ib.ack(0, 0, 0, 1000, 5, 3, True, 0, 47)
    ci = compact.encode_inbox(ib, co.tags(), co.current_term, co.incarnation)
    assert len(ci.esc) == 1 and int(ci.ev_c[0, 0, 0]) == abi.CEV_ESCAPED          # no tag: the reply carries its echo pair itself
    back = model.unpack(ci, co.current_term, st)
    assert tuple(back.ev_el[0, 0, 0]) == (0, 47) and int(back.ev_meta[0, 0, 0]) == int(ib.ev_meta[0, 0, 0])


def test_reference_format_reply_frames_reach_the_compact_inbox_through_the_native_chain():
    """SURVEY §8(f)-2 end to end on the host, on the real leader stream: every reply of the stream travels as ONE frame of the
    reference's wire layout (EventCodec framing, scope head, Kryo RaftResponse body, sequence), and the native chain —
    rafting_frame_scan -> rafting_ack_frames_decode -> rafting_acks_to_cinbox with the pending table filled when the plans left
    — must build the compact inbox compact.encode_inbox builds from the same dense events, word for word and escape record
    for escape record (the encoder the GPU parity tests of the compact path feed the engine with)."""
    from rafting_b200 import ingest
    G, R, rows = 64, 3, 6
    F = R - 1
    cfg = abi.make_cfg(replicas=R, max_groups=G, max_rows=rows)
    o = binding.Oracle(cfg)
    o.open_bulk(0, harness.init_array(G, terms=np.arange(G) % 7))
    w1 = workload.make_wl(23, 1, G, F)
    w = workload.make_wl(23, rows, G, F, p_reject_ppm=80_000, p_error_ppm=30_000, p_cancel_ppm=20_000)   # rejections, time-outs, cancellations
    harness.elect_all(o, w1)
    st = model.InFlight(G, F)
    cm = ingest.CtxMap()
    for g in range(G):
        cm.put(b"ctx-%d" % g, g)
    pend = ingest.Pending()
    seq_of = {}                                                                      # (row, g, f) of the planning step -> sequence
    next_seq = [1] * F
    prev, tags, sent_term, sent_inc = None, None, None, None
    frames_total = words = escapes = failures = 0
    for k in range(10):
        ib = workload.leader_inbox_host(w, k, prev)
        ek = (ib.ev_meta & np.uint64(0xF)).astype(np.int64)
        if k in (4, 7):
            for ev in np.argwhere(ek != 0)[:3]:
                ib.ev_tn["y"][tuple(ev)] += 80_000                                   # replies 80 s late: they travel as escape records
        if prev is not None:
            want = compact.encode_inbox(ib, tags, sent_term, sent_inc)
            got = compact.CompactInbox(rows, G, F)
            got.row_base[:] = want.row_base                                         # the pump's clock: the first drain of each row
            esc = np.zeros(rows * G * F, dtype=abi.CESC_IN)
            n_esc = 0
            for r in range(rows):
                for f in range(F):                                                   # one connection per follower lane
                    at = [(g, int(ib.ev_tn["y"][r, g, f])) for g in range(G) if ek[r, g, f] != 0]
                    for now in sorted(set(t for _, t in at)):                        # one receive buffer per (connection, drain time)
                        buf = b""
                        failed = {abi.OUT_ERROR: [], abi.OUT_CANCELED: []}
                        for g, t in at:
                            if t != now:
                                continue
                            outcome = (int(ib.ev_meta[r, g, f]) >> 4) & 3
                            if outcome != abi.OUT_OK:                                # no frame: the invocation's Async failed
                                failed[outcome].append(seq_of[(r, g, f)])
                                continue
                            method = b"installSnapshot" if ek[r, g, f] == abi.EV_IS_ACK else b"appendEntries"
                            body = ingest.reply_body_encode(int(ib.ev_tn["x"][r, g, f]), bool((int(ib.ev_meta[r, g, f]) >> 6) & 1))
                            buf += ingest.encode(ingest.ACK, method + b":ctx-%d" % g, body, sequence=seq_of[(r, g, f)])
                        rc, frames, used, _ = ingest.scan(buf, cap=G + 1)
                        assert rc == 0 and used == len(buf)
                        acks = ingest.ack_frames_decode(buf, frames, cm)
                        assert len(acks) == len(frames)
                        frames_total += len(frames)
                        rc, n_esc, deferred, unknown = pend.acks_to_cinbox(f, acks, now, r, got, esc, n_esc)
                        assert rc == 0 and len(deferred) == 0 and unknown == 0
                        for outcome, seqs in failed.items():
                            rc, n_esc, deferred, unknown = pend.failures_to_cinbox(f, seqs, outcome, now, r, got, esc, n_esc)
                            assert rc == 0 and len(deferred) == 0 and unknown == 0
                            failures += len(seqs)
            assert np.array_equal(got.ev_c, want.ev_c), f"step {k}: ev_c words differ"
            a, b = np.sort(esc[:n_esc], order="slot"), np.sort(want.esc, order="slot")
            assert a.tobytes() == b.tobytes(), f"step {k}: escape records differ"
            words += int(((got.ev_c & 0xF) != 0).sum()) - n_esc
            escapes += n_esc
            assert len(pend) == 0                                                    # every plan of the previous step was answered
        dense = o.step(ib)
        epoch = np.zeros(G, dtype=abi.I64X2)
        for g in range(G):
            s = o.export(g); epoch[g] = (s.epoch_index, s.epoch_term)
        co = model.pack(dense, epoch, st, esc_cap=rows * G * F + 2 * rows * G)
        model.unpack(compact.encode_inbox(ib, tags, sent_term, sent_inc), dense.current_term, st) if prev is not None else None   # frees the tags
        prev, tags, sent_term, sent_inc = dense, co.tags(), co.current_term.copy(), co.incarnation.copy()
        # the plans leave: remember what each was sent with, under the sequence of its connection
        seq_of = {}
        pk = (dense.plan_meta & np.uint64(0xF)).astype(np.int64)
        for r, g, f in np.argwhere((pk == abi.PLAN_AE) | (pk == abi.PLAN_IS)):
            r, g, f = int(r), int(g), int(f)
            seq_of[(r, g, f)] = next_seq[f]
            pend.put(f, next_seq[f], abi.EV_IS_ACK if pk[r, g, f] == abi.PLAN_IS else abi.EV_AE_ACK, g, f, int(tags[r, g, f]),
                     int(dense.plan_meta[r, g, f]) >> 32, int(sent_term[g]),
                     int(dense.plan_epoch[r, g, f]), int(dense.plan_lc[r, g, f]["x"]))
            next_seq[f] += 1
    assert frames_total > 3000 and words > 2500 and escapes >= 6 and failures > 100
