"""BASELINE config #1 on the CPU oracle: three nodes (three oracle instances with local_slot 0/1/2) exchange real
AppendEntries / PreVote / RequestVote traffic through tests/cluster_sim.py.  The reference's own check for this
configuration is "the three files are identical" (README.md:28-33)."""
import pytest

from oracle import binding
from tests.cluster_sim import Cluster


def _oracle(cfg):
    return binding.Oracle(cfg)


def test_three_node_cluster_elects_replicates_and_files_match():
    c = Cluster(_oracle, G=6, seed=11)
    c.run(120)
    assert all(c.leader_of(g) is not None for g in range(c.G))
    c.run(150)
    c.run(60, submit=False)                       # drain
    c.check(converged=True)
    assert min(len(nd.file[g]) for nd in c.nodes for g in range(c.G)) > 20
    assert c.counts["ae_ok"] > 500 and c.counts["vote"] >= c.G


@pytest.mark.parametrize("seed", [3, 4])
def test_lossy_links_and_leader_isolation(seed):
    """2 % message loss, then the leader of every group is cut off: a new leader is elected, the old one keeps
    accepting commands it can never commit; after healing its conflicting suffix is truncated and the files agree."""
    c = Cluster(_oracle, G=5, seed=seed, drop_ppm=20_000)
    c.run(150)
    c.check(converged=False)
    victim = c.leader_of(0)
    assert victim is not None
    c.cut = {victim}
    c.run(150)
    assert any(c.leader_of(g) not in (None, victim) for g in range(c.G))
    c.cut = set()
    c.run(200)
    c.drop_ppm = 0
    c.run(100, submit=False)
    c.check(converged=True)
    assert min(len(nd.file[0]) for nd in c.nodes) > 10


def test_five_node_cluster_two_nodes_down():
    """R = 5: a majority of three keeps committing while two nodes are cut off, then everyone catches up."""
    c = Cluster(_oracle, G=4, R=5, seed=9, drop_ppm=5_000)
    c.run(150)
    lead = c.leader_of(0)
    c.cut = {(lead + 1) % 5, (lead + 2) % 5}
    before = len(c.nodes[lead].file[0])
    c.run(150)
    assert len(c.nodes[c.leader_of(0)].file[0]) > before + 10
    c.cut = set()
    c.run(150)
    c.drop_ppm = 0
    c.run(100, submit=False)
    c.check(converged=True)


def test_compaction_and_snapshot_catch_up():
    """Every node compacts its log every 25 applied commands (RaftRoutine.compactLog -> RaftLog.flush).  A follower
    that was cut off long enough falls behind the leader's epoch and is caught up through InstallSnapshot
    (pendingInstallation -> IS RPC answered false until the download has finished -> flush to the milestone ->
    AppendEntries from there)."""
    c = Cluster(_oracle, G=4, seed=21, compact_every=25)
    c.run(120)
    lead = c.leader_of(0)
    lagger = (lead + 1) % 3
    c.cut = {lagger}
    c.run(220)
    c.cut = set()
    c.run(200)
    c.run(80, submit=False)
    c.check(converged=True)
    assert c.counts["compactions"] > 10 and c.counts["is_sent"] > 0 and c.counts["snapshots_installed"] > 0
    assert all(nd.sut.export(g).epoch_index > 0 for nd in c.nodes for g in range(c.G))


@pytest.mark.parametrize("R,pre_vote,seed", [(3, True, 101), (3, False, 102), (5, True, 103), (5, False, 104), (2, True, 105),
                                             (3, False, 106), (5, False, 107), (3, True, 108)])
def test_random_partitions_keep_raft_safe(R, pre_vote, seed):
    """Jepsen-style: every 40 ticks a random minority (possibly the leader) is cut off, 3 % of the messages are lost,
    logs are compacted every 30 commands.  Election safety is asserted on every step inside the simulator, prefix
    agreement of the applied files after every phase, identical files after healing.  The simulated network drops the
    vote requests the reference's Candidate would wrongly grant (cluster_sim._vote_request_is_unsafe)."""
    c = _jepsen(R, pre_vote, seed, guard=True)
    assert c.counts["votes_granted_to_a_stale_log"] == 0
    assert min(len(nd.file[g]) for nd in c.nodes for g in range(c.G)) > (5 if R == 2 else 30)


def _jepsen(R, pre_vote, seed, guard, keep=None, flags=0, shadow_native=False):
    import numpy as np
    rng = np.random.default_rng(seed)
    c = Cluster(_oracle, G=4, R=R, seed=seed, drop_ppm=30_000, compact_every=30, pre_vote=pre_vote, guard_candidate_votes=guard,
                cfg_flags=flags, shadow_native=shadow_native)
    if keep is not None:
        keep.append(c)
    c.run(80)
    for phase in range(10):
        k = int(rng.integers(0, (R - 1) // 2 + 1))
        c.cut = set(int(x) for x in rng.choice(R, size=k, replace=False))
        c.run(40)
        c.check(converged=False)
    c.cut = set()
    c.run(250)
    c.drop_ppm = 0
    c.run(120, submit=False)
    c.check(converged=True)
    return c


def test_upstream_candidate_votes_without_log_check():
    """The flaw itself, first as a known answer: a Candidate of term 3 whose log ends at (124, 1) receives
    RequestVote(term 4) from a peer whose log ends at (84, 1) and GRANTS it (Candidate.java:68-71 — switchTo(Follower,
    term, candidateId) with no logUpToDate), where a Follower in the same position refuses (Follower.java:118-122).
    Then end to end: the unguarded run of seed 102 elects that peer and the applied files diverge."""
    from rafting_b200 import abi
    from tests import harness
    for role_first, want in (("candidate", 1), ("follower", 0)):
        cfg = abi.make_cfg(replicas=3, local_slot=2, max_groups=1, max_rows=1, pre_vote=False, election_ms=900)
        o = binding.Oracle(cfg)
        o.open_group(0, term=2, first_index=1, last_index=124, last_term=1, now_ms=harness.T0, rand_ms=1000)
        if role_first == "candidate":
            ib = abi.Inbox(1, 1, 2); ib.timeout(0, 0, harness.T0 + 1000, rand=1000); o.step(ib)      # election timeout -> Candidate, term 3
            assert o.export(0).role == abi.ROLE_CANDIDATE and o.export(0).current_term == 3
        ib = abi.Inbox(1, 1, 2); ib.vote_request(0, 0, harness.T0 + 1100, 1, 4, 84, 1)
        out = o.step(ib)
        m = int(out.rep_meta[0, 0])
        assert (m & 1, (m >> 1) & 1) == (1, want), role_first
    keep = []
    with pytest.raises(AssertionError):
        _jepsen(3, False, 102, guard=False, keep=keep)
    assert keep[0].counts["votes_granted_to_a_stale_log"] > 0


def test_upstream_commit_rollback_can_stall_a_group():
    """Second upstream finding, a LIVENESS one, mirrored faithfully: commitIndex is neither persisted nor carried by votes,
    so a freshly elected leader may know a lower commitIndex than a majority of its followers.  Every AppendEntries it
    sends then makes those followers throw "rollback is not allowed" (RocksLog.java:100-103 via Follower.java:76-84) —
    AFTER resetting their election timers in the `finally` — so no ack ever reaches the leader, its commitIndex never
    moves, and nobody times out: the group is stuck for good.  Seed 1030 of the Jepsen run ends in exactly that state."""
    import numpy as np
    R, seed = 5, 1030
    rng = np.random.default_rng(seed)
    c = Cluster(_oracle, G=4, R=R, seed=seed, drop_ppm=30_000, compact_every=30, pre_vote=True, guard_candidate_votes=True)
    c.run(80)
    for phase in range(10):
        k = int(rng.integers(0, (R - 1) // 2 + 1))
        c.cut = set(int(x) for x in rng.choice(R, size=k, replace=False))
        c.run(40)
        c.check(converged=False)                      # safety holds throughout
    c.cut = set()
    c.run(250)
    c.drop_ppm = 0
    c.run(200, submit=False)
    before = c.counts["commit_rollback"]
    c.run(200, submit=False)                          # a perfect network, and still ...
    g = 1
    lead = c.leader_of(g)
    commits = [nd.sut.export(g).commit_index for nd in c.nodes]
    assert lead is not None and sum(x > commits[lead] for x in commits) >= R // 2 + 1 - 1   # a blocking set is ahead of the leader
    assert c.counts["commit_rollback"] - before > 50                                          # ... every heartbeat is thrown away
    assert len({len(nd.file[g]) for nd in c.nodes}) > 1                                       # ... and the group never converges
    c.check(converged=False)                                                                  # (what was applied still agrees)


def test_gpu_scenarios_converge_on_the_oracle(tmp_path):
    """tests/test_cluster_gpu.py runs these scenarios with engine and oracle in lock-step; the trajectory is the oracle's,
    so their convergence can (and must) be checked here without a GPU."""
    from tests import cluster_scenarios as scenarios
    for R, G, seed in scenarios.GPU_ISOLATION:
        scenarios.isolation(_oracle, R, G, seed)
    scenarios.compaction(_oracle)
    scenarios.restart(_oracle, tmp_path)
    for R, pre_vote, seed in scenarios.GPU_JEPSEN:
        scenarios.jepsen(_oracle, R, pre_vote, seed)


@pytest.mark.parametrize("R,pre_vote,seed", [(3, False, 102), (5, True, 1030), (3, False, 1007), (5, False, 1032), (5, False, 1031)])
def test_opt_in_fixes_make_the_unguarded_runs_safe_and_live(R, pre_vote, seed):
    """RAFTING_CFG_STRICT_CANDIDATE_VOTE | RAFTING_CFG_LENIENT_FOLLOWER_COMMIT (oracle only in this version; the engine
    rejects a non-zero cfg.flags): the very runs that diverge (102, 1031), stall (1030) or crawl (1007, 1032) with the
    reference's behaviour converge with no guard on the network.  Offline sweep: 0 failures in 160 runs."""
    from rafting_b200 import abi
    c = _jepsen(R, pre_vote, seed, guard=False, flags=abi.CFG_STRICT_CANDIDATE_VOTE | abi.CFG_LENIENT_FOLLOWER_COMMIT)
    assert c.counts["votes_granted_to_a_stale_log"] == 0 and c.counts["commit_rollback"] == 0


@pytest.mark.parametrize("R,pre_vote,seed", [(3, True, 301), (5, False, 302)])
def test_native_pump_dispatch_matches_the_python_pump_under_partitions(R, pre_vote, seed):
    """SURVEY §8(f)-2, outbound half: the C dispatch loop (rafting_outbox_to_requests: plans and vote broadcasts -> request
    records, one batch per peer and step instead of one frame per group) and the C placement (rafting_request_to_inbox:
    record -> op slot + entry terms) run beside the simulator's Python pump on EVERY step of a Jepsen-style run — elections,
    heartbeats, appends, compaction, InstallSnapshot, partitions, restarts of role objects — and must emit exactly the same
    requests in the same order and build exactly the same op columns."""
    c = _jepsen(R, pre_vote, seed, guard=True, shadow_native=True)
    assert c.counts["native_requests_checked"] > 2500 and c.counts["native_placements_checked"] > 3000
    assert c.counts["native_replies_checked"] > 2500 and c.counts["is_sent"] > 0 and c.counts["vote"] > 0
    assert c.counts["native_apply_ranges_checked"] > 500          # commit-dirty groups -> (gid, first, last), same as the Python loop


@pytest.mark.parametrize("R,pre_vote,seed", [(3, True, 401), (5, False, 402), (3, False, 403)])
def test_native_inbox_builder_matches_the_python_queues(R, pre_vote, seed):
    """The third phase of the pump in C — rafting_builder_*: per-group FIFOs of requests, replies and submits -> the rows of a
    step — receives every item the simulator queues (lossy links, partitions, time-outs, leader changes) and must build the
    identical inbox, column for column, and keep the identical leftovers, on every step.  (Runs without compaction and without
    the vote guard: both are host-side decisions taken while placing, outside the builder; safety is not asserted here.)"""
    import numpy as np
    rng = np.random.default_rng(seed)
    c = Cluster(_oracle, G=5, R=R, seed=seed, drop_ppm=30_000, pre_vote=pre_vote, shadow_native=True)
    assert c.shadow_builder
    c.run(60)
    for phase in range(6):
        k = int(rng.integers(0, (R - 1) // 2 + 1))
        c.cut = set(int(x) for x in rng.choice(R, size=k, replace=False))
        c.run(40)
    c.cut = set()
    c.run(100)
    assert c.counts["native_builds_checked"] == R * c.tick and c.counts["native_build_items"] > 2500
    assert c.counts["native_requests_checked"] > 800 and c.counts["native_replies_checked"] > 600
