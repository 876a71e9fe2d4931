"""BASELINE config #1 on the GPU: the cluster of tests/cluster_sim.py with every node running BOTH the CUDA engine
and the CPU oracle in lock-step.  Each step's outbox must be identical (so the two clusters exchange the same
messages for the whole run), the exported state must match at the end, and the engine cluster must pass the same
"the files are identical" check."""
import numpy as np
import pytest

from oracle import binding
from tests import harness
from tests import cluster_scenarios as scenarios

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


@pytest.mark.parametrize("R,G,seed", scenarios.GPU_ISOLATION)
def test_cluster_engine_matches_oracle_and_files_agree(engine_mod, R, G, seed):
    c = scenarios.isolation(lambda cfg: Pair(engine_mod, cfg), R, G, seed)
    _states_equal(c, R)


def test_cluster_with_compaction_and_snapshot_install(engine_mod):
    _states_equal(scenarios.compaction(lambda cfg: Pair(engine_mod, cfg)), 3)


def test_node_crash_and_restart_from_the_journal(engine_mod, tmp_path):
    """Durability journal + rafting_group_open + rafting_group_load_runs: a killed node comes back from what the
    reference keeps on disk; engine and oracle stay identical through the restart and the files converge."""
    _states_equal(scenarios.restart(lambda cfg: Pair(engine_mod, cfg), tmp_path), 3)


@pytest.mark.parametrize("R,pre_vote,seed", scenarios.GPU_JEPSEN)
def test_random_partitions_engine_in_lock_step(engine_mod, R, pre_vote, seed):
    """The Jepsen-style run of tests/test_cluster_cpu.py with every node stepping engine and oracle together."""
    _states_equal(scenarios.jepsen(lambda cfg: Pair(engine_mod, cfg), R, pre_vote, seed), R)


def _states_equal(c, R):
    for nd in c.nodes:
        harness.assert_states_equal(nd.sut.o, nd.sut.e, range(c.G), R - 1, where=f"node {nd.slot}")


def test_opt_in_protocol_fixes_on_the_device_in_lock_step():
    """The device branches of RAFTING_CFG_STRICT_CANDIDATE_VOTE / _LENIENT_FOLLOWER_COMMIT live in a library of their own
    (-DRAFTING_ENABLE_CFG_FLAGS; the default build rejects cfg.flags != 0 and keeps its SASS): the unguarded runs that
    diverge (seed 102) or stall (1030) with the reference's behaviour run on it in lock-step with the oracle and converge."""
    import os
    import subprocess
    import sys
    from rafting_b200 import _build
    lib = _build.build_flags()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RAFTING_B200_LIB=lib)
    res = subprocess.run([sys.executable, os.path.join(root, "tests", "run_flagged_cluster.py")], capture_output=True, text=True,
                         timeout=900, env=env, cwd=root)
    assert res.returncode == 0 and "FLAGGED-OK" in res.stdout, (res.stdout[-1500:], res.stderr[-3000:])
