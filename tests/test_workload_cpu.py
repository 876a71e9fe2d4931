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


def _vote_setup(G, R, local_slot=2):
    cfg = abi.make_cfg(replicas=R, local_slot=local_slot, max_groups=G, max_rows=2, entry_pool_cap=workload.POOL_TERMS)
    init = harness.init_array(G, terms=1 + np.arange(G) % 5)
    init["last_index"] = 100 + np.arange(G) % 50
    init["last_term"] = 1 + np.arange(G) % 5
    return cfg, init


def test_vote_storm_stream_threads_agree_and_elects_leaders():
    """config #3 shape at reduced size: PreVote on, 5 replicas, grant/higher-term/timeout mix."""
    G, R = 600, 5
    cfg, init = _vote_setup(G, R)
    a, b = binding.Oracle(cfg), binding.Oracle(cfg)
    a.open_bulk(0, init), b.open_bulk(0, init)
    w = workload.make_wl(0x5EED0003, 2, G, R - 1, local_slot=2)
    out = None
    seen_roles = set()
    for k in range(8):
        ib = workload.vote_inbox_host(w, k, out)
        out = a.step(ib, threads=1)
        harness.assert_outbox_equal(out, b.step(ib, threads=3), where=f"vote round {k}")
        seen_roles |= set((out.role_word & 3).tolist())
    harness.assert_states_equal(a, b, range(G), R - 1)
    assert seen_roles == {0, 1, 2}
    leaders = ((out.role_word & 3) == abi.ROLE_LEADER).sum()
    assert G * 0.2 < leaders < G
    assert (out.current_term >= init["term"]).all()              # terms never go back


def test_mixed_churn_stream_threads_agree_and_covers_paths():
    """config #5 shape at reduced size: steady leaders + churn (step-down, follower AE, re-election) +
    catch-up (flush, InstallSnapshot)."""
    G, R, rows = 700, 3, 4
    cfg = abi.make_cfg(replicas=R, max_groups=G, max_rows=rows, entry_pool_cap=workload.POOL_TERMS)
    a, b = binding.Oracle(cfg), binding.Oracle(cfg)
    init = harness.init_array(G, terms=np.arange(G) % 7)
    a.open_bulk(0, init), b.open_bulk(0, init)
    w1 = workload.make_wl(0x5EED0005, 1, G, R - 1)
    harness.elect_all(a, w1), harness.elect_all(b, w1)
    w = workload.make_wl(0x5EED0005, rows, G, R - 1)
    out = None
    op_kinds, plan_kinds = set(), set()
    for k in range(40):
        ib = workload.mixed_inbox_host(w, k, out)
        out = a.step(ib, threads=1)
        harness.assert_outbox_equal(out, b.step(ib, threads=2), where=f"mixed step {k}")
        op_kinds |= set((ib.op_meta & 0xFF).ravel().tolist())
        plan_kinds |= set((out.plan_meta & 0xF).ravel().tolist())
    harness.assert_states_equal(a, b, range(G), R - 1)
    assert {abi.OP_SUBMIT, abi.OP_TIMEOUT, abi.OP_AE_REQUEST, abi.OP_FLUSH} <= op_kinds
    assert {abi.PLAN_AE, abi.PLAN_IS} <= plan_kinds
    assert any(a.export(g).epoch_index > 0 for g in range(G))


def test_oracle_compact_group_columns():
    """RAFTING_INBOX_COMPACT_GROUPS in the oracle: same step, group columns indexed by position in gids[]."""
    G, R, rows = 96, 3, 2
    cfg = abi.make_cfg(replicas=R, max_groups=G, max_rows=rows)
    a, b = binding.Oracle(cfg), binding.Oracle(cfg)
    init = harness.init_array(G, terms=np.arange(G) % 5)
    a.open_bulk(0, init), b.open_bulk(0, init)
    w1 = workload.make_wl(5, 1, G, R - 1)
    harness.elect_all(a, w1), harness.elect_all(b, w1)
    gids = np.arange(1, G, 5, dtype=np.uint32)
    ib = abi.Inbox(rows, len(gids), R - 1, gids=gids)
    for r in range(rows):
        for i in range(len(gids)):
            ib.submit(r, i, harness.T0 + r, 2)
    wide = a.step(ib)
    ib.flags |= abi.INBOX_COMPACT_GROUPS
    tight = b.step(ib)
    assert tight.commit_index.shape == (len(gids),)
    assert not wide.equal(tight, gids=slice(0, 0))          # row columns identical
    for name, _ in abi.Outbox.GROUP_COLS:
        assert np.array_equal(getattr(tight, name), getattr(wide, name)[gids]), name
    assert tight.current_term.min() >= 1 and tight.incarnation.min() >= 1
