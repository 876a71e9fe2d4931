"""The cluster scenarios of tests/test_cluster_gpu.py as functions of the SUT factory, so that the CPU suite can replay
them on the oracle alone (tests/test_cluster_cpu.py::test_gpu_scenarios_converge_on_the_oracle): a change of the simulator
that breaks their convergence shows up without a GPU."""
import os

import numpy as np

from tests.cluster_sim import Cluster


def isolation(make, R, G, seed):
    c = Cluster(make, G=G, R=R, seed=seed, drop_ppm=15_000)
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
    return c


def compaction(make):
    c = Cluster(make, G=12, R=3, seed=21, compact_every=25, drop_ppm=5_000)
    c.run(120)
    c.cut = {(c.leader_of(0) + 1) % 3}
    c.run(220)
    c.cut = set()
    c.run(200)
    c.drop_ppm = 0
    c.run(80, submit=False)
    c.check(converged=True)
    assert c.counts["snapshots_installed"] > 0 and c.counts["is_sent"] > 0
    return c


def restart(make, tmp_path):
    from rafting_b200 import durable
    G = 8
    journals = [durable.Journal(os.path.join(str(tmp_path), f"n{k}"), G) for k in range(3)]
    c = Cluster(make, G=G, seed=31, drop_ppm=10_000)
    c.on_outbox = lambda nd, ob: journals[nd.slot].commit_step(ob.role_word, ob.current_term)
    c.run(150)
    victim = c.leader_of(0)
    c.cut = {victim}
    c.run(2)
    journals[victim].close()
    journals[victim] = durable.Journal(os.path.join(str(tmp_path), f"n{victim}"), G)
    c.restart(victim, make, lambda g: journals[victim].restore(g))
    c.run(60)
    c.cut = set()
    c.run(250)
    c.drop_ppm = 0
    c.run(80, submit=False)
    c.check(converged=True)
    for j in journals:
        j.close()
    return c


def jepsen(make, R, pre_vote, seed):
    rng = np.random.default_rng(seed)
    c = Cluster(make, G=4, R=R, seed=seed, drop_ppm=30_000, compact_every=30, pre_vote=pre_vote, guard_candidate_votes=True)
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
    return c


GPU_ISOLATION = [(3, 24, 5), (5, 8, 6)]
GPU_JEPSEN = [(3, False, 106), (5, True, 103)]
