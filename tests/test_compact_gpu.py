"""The compact host path (`-m gpu`): a step fed through rafting_step_begin_compact — narrow wire columns, the
(epochAtSend, lastIndexAtSend) echo kept in the HBM in-flight table under a tag — must be, once decoded, bit for bit the
dense step of the oracle on the same logical stream.  Election phases run through the dense path (vote replies), then
the leader stream through the compact one, including rejected / failed / cancelled acks, unavailable followers, the
matchIndex-rollback error of out-of-order heartbeat acks, and forced escapes."""
import numpy as np
import pytest

from oracle import binding
from rafting_b200 import abi, compact, workload
from tests import harness

pytestmark = pytest.mark.gpu


def _run(G, R, rows, steps, seed, esc_cap=1 << 16, mangle=None):
    from rafting_b200 import engine
    cfg = abi.make_cfg(replicas=R, max_groups=G, max_rows=rows)
    e, o = engine.Engine(cfg), binding.Oracle(cfg)
    init = harness.init_array(G, terms=np.arange(G) % 7)
    e.open_bulk(0, init), o.open_bulk(0, init)
    w1 = workload.make_wl(seed, 1, G, R - 1)
    w = workload.make_wl(seed, rows, G, R - 1, p_reject_ppm=60_000, p_error_ppm=20_000, p_cancel_ppm=20_000)
    harness.assert_outbox_equal(harness.elect_all(o, w1), harness.elect_all(e, w1), where="election (dense path)")
    prev, tags, sent_term, sent_inc = None, None, None, None
    up = down = escapes = 0
    overflows = []
    for k in range(steps):
        ib = workload.leader_inbox_host(w, k, prev)
        if mangle:
            mangle(k, ib)
        want = o.step(ib, threads=8)
        cin = compact.encode_inbox(ib, tags, sent_term, sent_inc)
        cout = compact.CompactOutbox(rows, G, R - 1, esc_cap=esc_cap)
        e.step_compact(cin, cout)
        try:
            got = compact.decode_outbox(cout)
        except OverflowError:                                     # more escapes than esc_cap: the lossless fallback
            got = abi.Outbox(rows, G, R - 1, G)
            e.step_fetch_dense(0, got.as_c())
            overflows.append(k)
        harness.assert_outbox_equal(want, got, where=f"compact step {k}")
        assert np.array_equal(cout.epoch["x"], np.array([0] * G))         # nothing was compacted in this stream
        prev, tags, sent_term, sent_inc = want, cout.tags(), cout.current_term.copy(), cout.incarnation.copy()
        up += cin.nbytes(); down += cout.nbytes(); escapes += len(cin.esc) + int(cout.counts[0])
    harness.assert_states_equal(o, e, list(range(0, G, max(1, G // 64))) + [G - 1], R - 1, where="compact end")
    return up, down, escapes, ib, cout, overflows


def test_compact_leader_stream_is_the_dense_step_bit_for_bit():
    # a fresh leader's first replicateLog sends prevLogIndex 0 / prevLogTerm 0 for every follower: none of those plans is
    # compact, so the first steps need an escape list as large as the plan matrix
    G, R, rows = 8192, 3, 16
    up, down, escapes, ib, cout, overflows = _run(G=G, R=R, rows=rows, steps=8, seed=0x5EED0002, esc_cap=rows * G * R)
    assert not overflows
    # the byte cut that motivates the format (dense = 65 B up + 69 B down per ack), on the last, steady-state step
    acks = int(((ib.ev_meta & np.uint64(0xF)) != 0).sum())
    cin_bytes = rows * 8 + rows * G * 4 + rows * G * (R - 1) * 4
    assert cin_bytes / acks < 10 and cout.nbytes() / acks < 10, (cin_bytes / acks, cout.nbytes() / acks)
    assert int(cout.counts[0]) < rows * G // 20                      # steady state: a few per cent of escapes at most


def test_compact_with_unavailable_followers_three_lanes_and_escapes():
    def mangle(k, ib):
        if k % 2 == 1:
            ib.op_ab["x"][:, ::5] = 1                                   # follower lane 0 unavailable for every fifth group
        if k == 3:
            ib.ev_tn["y"][0, 7, 0] += 70_000                            # a reply 70 s late: beyond the 16-bit offset -> escape record
            ib.ev_tn["x"][1, 9, 1] += 0                                 # (unchanged term: stays compact)
    up, down, escapes, ib, cout, overflows = _run(G=2048, R=4, rows=8, steps=7, seed=0x5EED0004, mangle=mangle, esc_cap=4096)
    assert escapes > 0 and overflows and overflows[0] == 0            # the first steps overflow 4096 records: dense fallback


def test_compact_escape_overflow_falls_back_to_the_dense_outbox():
    from rafting_b200 import engine
    G, R, rows = 1024, 3, 4
    cfg = abi.make_cfg(replicas=R, max_groups=G, max_rows=rows)
    e, o = engine.Engine(cfg), binding.Oracle(cfg)
    init = harness.init_array(G, terms=np.arange(G) % 7)
    e.open_bulk(0, init), o.open_bulk(0, init)
    w1 = workload.make_wl(7, 1, G, R - 1); w = workload.make_wl(7, rows, G, R - 1)
    harness.elect_all(o, w1), harness.elect_all(e, w1)
    ib = workload.leader_inbox_host(w, 0, None)
    want = o.step(ib)
    # the first replicateLog of a fresh leader sends prevLogIndex 0 / prevLogTerm 0: not compact -> thousands of escapes
    cout = compact.CompactOutbox(rows, G, R - 1, esc_cap=8)
    e.step_compact(compact.encode_inbox(ib, None, None, None), cout)
    assert int(cout.counts[0]) > 8
    with pytest.raises(OverflowError):
        compact.decode_outbox(cout)
    dense = abi.Outbox(rows, G, R - 1, G)
    e.step_fetch_dense(0, dense.as_c())
    harness.assert_outbox_equal(want, dense, where="dense fallback")
