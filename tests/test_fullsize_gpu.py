"""Parity at BASELINE.json's full sizes (`-m gpu`).  The oracle runs the whole configuration on the host
cores (groups sharded over loop threads); every outbox column and a spread of exported group states are
compared bit for bit, plus the size-independent properties of the domain."""
import numpy as np
import pytest

from oracle import binding
from rafting_b200 import abi, workload
from tests import harness

pytestmark = pytest.mark.gpu


def _pair(G, R, rows, **kw):
    from rafting_b200 import engine
    cfg = abi.make_cfg(replicas=R, max_groups=G, max_rows=rows, **kw)
    o, e = binding.Oracle(cfg), engine.Engine(cfg)
    init = harness.init_array(G, terms=np.arange(G) % 7)
    o.open_bulk(0, init)
    e.open_bulk(0, init)
    return cfg, o, e


def test_config2_64k_groups_3_replicas_leader_stream():
    """config #2: 64K RaftContext groups, 3 replicas, synthetic AppendEntries stream (seed 0x5EED0002)."""
    G, R, rows, steps, T = 65536, 3, 8, 6, 16
    cfg, o, e = _pair(G, R, rows)
    w1 = workload.make_wl(0x5EED0002, 1, G, R - 1)
    w = workload.make_wl(0x5EED0002, rows, G, R - 1)
    out = None
    for ph in (0, 1, 2):
        ib = workload.election_inbox_host(w1, ph, out)
        out, oe = o.step(ib, threads=T), e.step(ib)
        harness.assert_outbox_equal(out, oe, where=f"election phase {ph}")
    prev = None
    commits = np.zeros(G, dtype=np.int64)
    acks = 0
    for k in range(steps):
        ib = workload.leader_inbox_host(w, k, prev)
        acks += int(((ib.ev_meta & np.uint64(0xF)) != 0).sum())
        prev, oe = o.step(ib, threads=T), e.step(ib)
        harness.assert_outbox_equal(prev, oe, where=f"step {k}")
        assert (oe.commit_index >= commits).all()                 # markCommitted never rolls back
        commits = oe.commit_index.copy()
    assert acks > 3_000_000
    harness.assert_states_equal(o, e, list(range(0, G, 997)) + [G - 1], R - 1, where="config #2 end")
    st = e.export_bulk(0, 4096)
    for s in st:
        assert s.commit_index <= s.last_index and s.role == abi.ROLE_LEADER
    assert (e.digest(0, G) != 0).all()


def test_config3_256k_groups_5_replicas_vote_storm():
    """config #3: 256K groups, 5 replicas, PreVote enabled, RequestVote storm (seed 0x5EED0003), 8 rounds."""
    from rafting_b200 import engine
    G, R, rows, T = 262144, 5, 1, 16
    cfg = abi.make_cfg(replicas=R, local_slot=2, max_groups=G, max_rows=rows, pre_vote=True)
    o, e = binding.Oracle(cfg), engine.Engine(cfg)
    init = harness.init_array(G, terms=1 + np.arange(G) % 5)
    init["last_index"] = 100 + np.arange(G) % 50
    init["last_term"] = 1 + np.arange(G) % 5
    o.open_bulk(0, init), e.open_bulk(0, init)
    w = workload.make_wl(0x5EED0003, rows, G, R - 1, local_slot=2)
    out = None
    msgs = 0
    for k in range(8):
        ib = workload.vote_inbox_host(w, k, out)
        msgs += int(((ib.ev_meta & np.uint64(0xF)) != 0).sum()) + int(((ib.op_meta & np.uint64(0xFF)) >= abi.OP_PREVOTE_REQ).sum())
        out = o.step(ib, threads=T)
        harness.assert_outbox_equal(out, e.step(ib), where=f"vote round {k}")
    assert msgs > 2_000_000
    harness.assert_states_equal(o, e, list(range(0, G, 4099)) + [G - 1], R - 1, where="config #3 end")
    roles = np.bincount(out.role_word & 3, minlength=3)
    assert roles[abi.ROLE_LEADER] > G // 5 and roles.sum() == G
    assert (out.current_term >= init["term"]).all()


def test_config5_512k_groups_mixed_churn_and_snapshot_catch_up():
    """config #5: 512K groups, 3 replicas, leader churn + InstallSnapshot catch-up (seed 0x5EED0005);
    140 ticks cover the 64-tick churn cycle twice and the 128-tick catch-up cycle once."""
    from rafting_b200 import engine
    G, R, rows, T = 524288, 3, 4, 16
    cfg = abi.make_cfg(replicas=R, max_groups=G, max_rows=rows, entry_pool_cap=workload.POOL_TERMS)
    o, e = binding.Oracle(cfg), engine.Engine(cfg)
    init = harness.init_array(G, terms=np.arange(G) % 7)
    o.open_bulk(0, init), e.open_bulk(0, init)
    w1 = workload.make_wl(0x5EED0005, 1, G, R - 1)
    out = None
    for ph in (0, 1, 2):
        ib = workload.election_inbox_host(w1, ph, out)
        out = o.step(ib, threads=T)
        harness.assert_outbox_equal(out, e.step(ib), where=f"election {ph}")
    w = workload.make_wl(0x5EED0005, rows, G, R - 1)
    out = None
    for k in range(35):
        ib = workload.mixed_inbox_host(w, k, out)
        out = o.step(ib, threads=T)
        harness.assert_outbox_equal(out, e.step(ib), where=f"mixed step {k}")
    harness.assert_states_equal(o, e, list(range(0, G, 8191)) + [G - 1], R - 1, where="config #5 end")
    assert ((out.plan_meta & 0xF) == abi.PLAN_IS).sum() >= 0
    ep = [e.export(g).epoch_index for g in range(0, G, 2053)]
    assert max(ep) > 0
