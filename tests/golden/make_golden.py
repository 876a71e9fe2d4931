#!/usr/bin/env python
"""Writes the golden fixtures of this directory by running the CPU oracle on the seeded streams.

Upstream has no golden vector for this path, so these pin OUR restatement: any later change
to the oracle, the streams or the batch format that alters results shows up as a diff of these files.
Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding  # noqa: E402
from rafting_b200 import abi, workload  # noqa: E402
from tests import harness  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def state_hash(sut, G, F):
    h = hashlib.sha256()
    for g in range(G):
        h.update(harness.state_bytes(sut.export(g), F))
    return h.hexdigest()


def run_cluster(sut_factory, snapshots=False):
    """BASELINE config #1 (three real nodes, tests/cluster_sim.py): loss, leader isolation, healing; with
    `snapshots` also log compaction and a follower that has to be caught up through InstallSnapshot."""
    from tests.cluster_sim import Cluster
    c = Cluster(sut_factory, G=6, R=3, seed=0x5EED0001, drop_ppm=10_000, compact_every=25 if snapshots else 0)
    c.run(120)
    lead = c.leader_of(0)
    c.cut = {(lead + 1) % 3} if snapshots else {lead}
    c.run(200 if snapshots else 100)
    c.cut = set()
    c.run(160 if snapshots else 120)
    c.drop_ppm = 0
    c.run(60, submit=False)
    c.check(converged=True)
    h = hashlib.sha256()
    for g in range(c.G):
        for term, payload in c.nodes[0].file[g]:
            h.update(f"{g}|{term}|{payload}\n".encode())
    sh = hashlib.sha256()
    for nd in c.nodes:
        for g in range(c.G):
            sh.update(harness.state_bytes(nd.sut.export(g), 2))
    return {
        "case": "cluster_snap_r3" if snapshots else "cluster_r3", "groups": c.G, "replicas": 3, "ticks": c.tick, "seed": 0x5EED0001,
        "file_lengths": [len(c.nodes[0].file[g]) for g in range(c.G)],
        "terms": [int(c.nodes[0].sut.export(g).current_term) for g in range(c.G)],
        "counts": dict(sorted(c.counts.items())),
        "files_sha256": h.hexdigest(), "state_sha256": sh.hexdigest(),
    }


def run_case(name, sut_factory=binding.Oracle):
    """Returns the record of one golden case; shared by the generator and by tests/test_golden.py."""
    if name == "cluster_r3":
        return run_cluster(sut_factory)
    if name == "cluster_snap_r3":
        return run_cluster(sut_factory, snapshots=True)
    if name == "leader_r3":
        G, R, rows, steps, seed = 512, 3, 4, 12, 0x5EED0002
        cfg = abi.make_cfg(replicas=R, max_groups=G, max_rows=rows)
        sut = sut_factory(cfg)
        sut.open_bulk(0, harness.init_array(G, terms=np.arange(G) % 7))
        harness.elect_all(sut, workload.make_wl(seed, 1, G, R - 1))
        w = workload.make_wl(seed, rows, G, R - 1)
        out = None
        for k in range(steps):
            out = sut.step(workload.leader_inbox_host(w, k, out))
    elif name == "votes_r5":
        G, R, rows, steps, seed = 384, 5, 2, 8, 0x5EED0003
        cfg = abi.make_cfg(replicas=R, local_slot=2, max_groups=G, max_rows=rows)
        sut = sut_factory(cfg)
        init = harness.init_array(G, terms=1 + np.arange(G) % 5)
        init["last_index"] = 100 + np.arange(G) % 50
        init["last_term"] = 1 + np.arange(G) % 5
        sut.open_bulk(0, init)
        w = workload.make_wl(seed, rows, G, R - 1, local_slot=2)
        out = None
        for k in range(steps):
            out = sut.step(workload.vote_inbox_host(w, k, out))
    elif name == "mixed_r3":
        G, R, rows, steps, seed = 640, 3, 4, 40, 0x5EED0005
        cfg = abi.make_cfg(replicas=R, max_groups=G, max_rows=rows, entry_pool_cap=workload.POOL_TERMS)
        sut = sut_factory(cfg)
        sut.open_bulk(0, harness.init_array(G, terms=np.arange(G) % 7))
        harness.elect_all(sut, workload.make_wl(seed, 1, G, R - 1))
        w = workload.make_wl(seed, rows, G, R - 1)
        out = None
        for k in range(steps):
            out = sut.step(workload.mixed_inbox_host(w, k, out))
    else:
        raise KeyError(name)
    return {
        "case": name, "groups": G, "replicas": R, "rows": rows, "steps": steps, "seed": seed,
        "commit_index_sum": int(out.commit_index.sum()), "commit_index_max": int(out.commit_index.max()),
        "term_sum": int(out.current_term.sum()),
        "roles": np.bincount(out.role_word & 3, minlength=3).tolist(),
        "errors": int((out.err_word & 0xFFFF != 0).sum()),
        "commit_index_head": out.commit_index[:16].tolist(),
        "state_sha256": state_hash(sut, G, R - 1),
    }


CASES = ("leader_r3", "votes_r5", "mixed_r3", "cluster_r3", "cluster_snap_r3")

if __name__ == "__main__":
    recs = {c: run_case(c) for c in CASES}
    with open(os.path.join(HERE, "streams.json"), "w") as f:
        json.dump(recs, f, indent=1)
    print(json.dumps(recs, indent=1))
