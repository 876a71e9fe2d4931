"""BASELINE config #1 on the GPU: the cluster of tests/cluster_sim.py with every node running BOTH the CUDA engine
and the CPU oracle in lock-step.  Each step's outbox must be identical (so the two clusters exchange the same
messages for the whole run), the exported state must match at the end, and the engine cluster must pass the same
"the files are identical" check."""
import numpy as np
import pytest

from oracle import binding
from tests import harness
from tests.cluster_sim import Cluster

pytestmark = pytest.mark.gpu


class Pair:
    """Steps the engine and the oracle with the same inbox and insists on identical answers."""

    def __init__(self, engine_mod, cfg):
        self.e, self.o = engine_mod.Engine(cfg), binding.Oracle(cfg)
        self.F = cfg.replicas - 1
        self.steps = 0

    def open_bulk(self, first, init):
        self.e.open_bulk(first, init), self.o.open_bulk(first, init)

    def step(self, ib):
        oe, oo = self.e.step(ib), self.o.step(ib)
        harness.assert_outbox_equal(oo, oe, where=f"cluster step {self.steps}")
        self.steps += 1
        return oe

    def export(self, g):
        se, so = self.e.export(g), self.o.export(g)
        assert harness.state_bytes(se, self.F) == harness.state_bytes(so, self.F)
        return se

    def load_runs(self, g, runs):
        self.e.load_runs(g, runs), self.o.load_runs(g, runs)

    def log_term(self, g, i):
        t = self.e.log_term(g, i)
        assert t == self.o.log_term(g, i)
        return t


@pytest.fixture(scope="module")
def engine_mod():
    from rafting_b200 import engine
    return engine


@pytest.mark.parametrize("R,G,seed", [(3, 24, 5), (5, 8, 6)])
def test_cluster_engine_matches_oracle_and_files_agree(engine_mod, R, G, seed):
    c = Cluster(lambda cfg: Pair(engine_mod, cfg), G=G, R=R, seed=seed, drop_ppm=15_000)
    c.run(140)
    victim = c.leader_of(0)
    assert victim is not None
    c.cut = {victim}
    c.run(120)
    c.cut = set()
    c.run(150)
    c.drop_ppm = 0
    c.run(80, submit=False)
    c.check(converged=True)
    assert min(len(nd.file[g]) for nd in c.nodes for g in range(G)) > 10
    for nd in c.nodes:
        harness.assert_states_equal(nd.sut.o, nd.sut.e, range(G), R - 1, where=f"node {nd.slot}")


def test_cluster_with_compaction_and_snapshot_install(engine_mod):
    c = Cluster(lambda cfg: Pair(engine_mod, cfg), G=12, R=3, seed=21, compact_every=25, drop_ppm=5_000)
    c.run(120)
    c.cut = {(c.leader_of(0) + 1) % 3}
    c.run(220)
    c.cut = set()
    c.run(200)
    c.drop_ppm = 0
    c.run(80, submit=False)
    c.check(converged=True)
    assert c.counts["snapshots_installed"] > 0 and c.counts["is_sent"] > 0
    for nd in c.nodes:
        harness.assert_states_equal(nd.sut.o, nd.sut.e, range(c.G), 2, where=f"node {nd.slot}")


def test_node_crash_and_restart_from_the_journal(engine_mod, tmp_path):
    """Durability journal + rafting_group_open + rafting_group_load_runs: a killed node comes back from what the
    reference keeps on disk; engine and oracle stay identical through the restart and the files converge."""
    from rafting_b200 import durable
    G = 8
    journals = [durable.Journal(str(tmp_path / f"n{k}"), G) for k in range(3)]
    c = Cluster(lambda cfg: Pair(engine_mod, cfg), G=G, seed=31, drop_ppm=10_000)
    c.on_outbox = lambda nd, ob: journals[nd.slot].commit_step(ob.role_word, ob.current_term)
    c.run(150)
    victim = c.leader_of(0)
    c.cut = {victim}
    c.run(2)
    journals[victim].close()
    journals[victim] = durable.Journal(str(tmp_path / f"n{victim}"), G)
    c.restart(victim, lambda cfg: Pair(engine_mod, cfg), lambda g: journals[victim].restore(g))
    c.run(60)
    c.cut = set()
    c.run(250)
    c.drop_ppm = 0
    c.run(80, submit=False)
    c.check(converged=True)
    for nd in c.nodes:
        harness.assert_states_equal(nd.sut.o, nd.sut.e, range(G), 2, where=f"node {nd.slot}")
    for j in journals:
        j.close()


@pytest.mark.parametrize("R,pre_vote,seed", [(3, False, 106), (5, True, 103)])
def test_random_partitions_engine_in_lock_step(engine_mod, R, pre_vote, seed):
    """The Jepsen-style run of tests/test_cluster_cpu.py with every node stepping engine and oracle together."""
    rng = np.random.default_rng(seed)
    c = Cluster(lambda cfg: Pair(engine_mod, cfg), G=4, R=R, seed=seed, drop_ppm=30_000, compact_every=30, pre_vote=pre_vote,
                guard_candidate_votes=True)
    c.run(80)
    for phase in range(10):
        k = int(rng.integers(0, (R - 1) // 2 + 1))
        c.cut = set(int(x) for x in rng.choice(R, size=k, replace=False))
        c.run(40)
    c.cut = set()
    c.run(250)
    c.drop_ppm = 0
    c.run(120, submit=False)
    c.check(converged=True)
    for nd in c.nodes:
        harness.assert_states_equal(nd.sut.o, nd.sut.e, range(c.G), R - 1, where=f"node {nd.slot}")
