"""CPU tests of the synthetic stream against the oracle: the loop-thread sharding of the oracle
(EventLoopGroup.java:77-80) does not change results, and the stream has the properties the
domain promises (commitIndex monotone, never beyond the log, leaders stay leaders)."""
import numpy as np

from oracle import binding
from rafting_b200 import abi, workload
from tests import harness


def _mk(G, R, seed, rows):
    cfg = abi.make_cfg(replicas=R, max_groups=G, max_rows=rows)
    o = binding.Oracle(cfg)
    o.open_bulk(0, harness.init_array(G, terms=np.arange(G) % 7))
    return cfg, o


def test_leader_stream_threads_agree_and_commit_progresses():
    G, R, rows = 96, 3, 4
    cfg, a = _mk(G, R, 1, rows)
    _, b = _mk(G, R, 1, rows)
    w1 = workload.make_wl(0x5EED0002, 1, G, R - 1)
    w = workload.make_wl(0x5EED0002, rows, G, R - 1)
    oa, ob = harness.elect_all(a, w1), harness.elect_all(b, w1)
    assert harness.assert_outbox_equal(oa, ob) is None
    assert ((oa.role_word & 3) == abi.ROLE_LEADER).all()
    assert (oa.current_term == 1 + np.arange(G) % 7).all()
    prev_a = prev_b = None
    commits = np.zeros(G, dtype=np.int64)
    for k in range(12):
        ib = workload.leader_inbox_host(w, k, prev_a)
        prev_a = a.step(ib, threads=1)
        prev_b = b.step(ib, threads=3)
        harness.assert_outbox_equal(prev_a, prev_b, where=f"step {k}")
        assert (prev_a.commit_index >= commits).all()            # markCommitted never rolls back
        commits = prev_a.commit_index.copy()
    harness.assert_states_equal(a, b, range(G), R - 1)
    assert ((prev_a.role_word & 3) == abi.ROLE_LEADER).all()
    assert commits.min() > 0
    for g in range(G):
        st = a.export(g)
        assert st.commit_index <= st.last_index and (st.err_word & 0xFFFF) in (0, abi.ERR["NOT_READY"])


def test_fuzz_is_deterministic_and_exercises_all_roles():
    G, R = 24, 5
    cfg = abi.make_cfg(replicas=R, local_slot=2, max_groups=G, max_rows=3)
    a, b = binding.Oracle(cfg), binding.Oracle(cfg)
    for o in (a, b):
        o.open_bulk(0, harness.init_array(G, terms=np.arange(G) % 3))
    fz = harness.Fuzzer(cfg, a, seed=11)
    out = None
    roles = set()
    for k in range(40):
        ib = fz.make(out)
        out = a.step(ib)
        out_b = b.step(ib, threads=2)
        harness.assert_outbox_equal(out, out_b, where=f"step {k}")
        roles |= set((out.role_word & 3).tolist())
    harness.assert_states_equal(a, b, range(G), R - 1)
    assert roles == {0, 1, 2}
