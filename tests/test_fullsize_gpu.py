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
