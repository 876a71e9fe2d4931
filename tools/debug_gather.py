"""debug: table commit column vs outbox commit column vs gathered vector (world = 1)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import binding
from rafting_b200 import abi, engine, workload
from tests import harness

G = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
rows, R, T = 2, 3, 32
cfg = abi.make_cfg(replicas=R, max_groups=G, max_rows=rows)
e, o = engine.Engine(cfg), binding.Oracle(cfg)
init = harness.init_array(G, terms=np.arange(G) % 7)
e.open_bulk(0, init); o.open_bulk(0, init)
engine.Engine.comm_init_all([e])
w1 = workload.make_wl(0x5EED0004, 1, G, R - 1); w = workload.make_wl(0x5EED0004, rows, G, R - 1)
out = None
for ph in (0, 1, 2):
    ib = workload.election_inbox_host(w1, ph, out)
    out = o.step(ib, threads=T); eo = e.step(ib)
prev = None
for k in range(3):
    ib = workload.leader_inbox_host(w, k, prev)
    prev = o.step(ib, threads=T); eo = e.step(ib)
    bad = prev.equal(eo)
    got = engine.Engine.allgather_commit_all([e])[0]
    got2 = e.allgather_commit(to_host=True)
    tab = np.array([s.commit_index for s in e.export_bulk(0, min(G, 65536))], dtype=np.int64)
    d1 = np.flatnonzero(got != prev.commit_index); d2 = np.flatnonzero(got2 != prev.commit_index)
    d3 = np.flatnonzero(tab != prev.commit_index[:len(tab)])
    print(f"step {k}: outbox diff cols {bad}; gather_all mismatches {len(d1)} first {d1[:8].tolist()}; gather mismatches {len(d2)}; table(export) mismatches {len(d3)} first {d3[:8].tolist()}")
    if len(d1):
        j = d1[0]; print("   at", j, "got", got[j], "oracle", prev.commit_index[j], "engine outbox", eo.commit_index[j])
