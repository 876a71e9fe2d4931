"""Runs in a subprocess with RAFTING_B200_LIB pointing at the -DRAFTING_ENABLE_CFG_FLAGS build (tests/test_cluster_gpu.py):
the UNGUARDED Jepsen runs that diverge / stall with the reference's behaviour, with both opt-in fixes switched on, every node
stepping the CUDA engine and the CPU oracle in lock-step (identical outboxes for the whole run, identical exported state)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import binding  # noqa: E402
from rafting_b200 import abi, engine  # noqa: E402
from tests import harness  # noqa: E402
from tests.cluster_sim import Cluster  # noqa: E402
from tests.test_cluster_gpu import Pair  # noqa: E402

FLAGS = abi.CFG_STRICT_CANDIDATE_VOTE | abi.CFG_LENIENT_FOLLOWER_COMMIT


def main():
    assert "flags" in os.path.basename(engine.lib_path()), engine.lib_path()
    for R, pre_vote, seed in ((3, False, 102), (5, True, 1030)):
        rng = np.random.default_rng(seed)
        c = Cluster(lambda cfg: Pair(engine, cfg), G=4, R=R, seed=seed, drop_ppm=30_000, compact_every=30, pre_vote=pre_vote,
                    guard_candidate_votes=False, cfg_flags=FLAGS)
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
        assert c.counts["votes_granted_to_a_stale_log"] == 0 and c.counts["commit_rollback"] == 0
        for nd in c.nodes:
            harness.assert_states_equal(nd.sut.o, nd.sut.e, range(c.G), R - 1, where=f"node {nd.slot}")
        print(f"flagged build: R={R} pre_vote={pre_vote} seed={seed}: {sum(nd.sut.steps for nd in c.nodes)} lock-step steps, converged")
    print("FLAGGED-OK")


if __name__ == "__main__":
    main()
