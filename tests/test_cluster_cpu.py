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
