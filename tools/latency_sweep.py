#!/usr/bin/env python
"""p50 / p99 commitIndex latency as a function of the batch size ("batch interval is a swept
parameter: 1 k ... 1 M events/step").  One synchronous host-path step per sample: acks are written into pinned
buffers, rafting_step_begin_host + rafting_step_wait_slot run, and the clock stops when the commit column is
readable on the host.  The stream is the closed-loop leader stream restricted to an active list of groups.
Prints one JSON line per (groups in the step, rows)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rafting_b200 import abi, engine, workload  # noqa: E402


def caller_owned(e, w, rows, n, F, G, gids, lat):
    import torch
    # pinned host batch buffers owned by the caller
    def pinned(shape, dt):
        t = torch.zeros(int(np.prod(shape)) * np.dtype(dt).itemsize, dtype=torch.uint8).pin_memory()
        return t, np.frombuffer(t.numpy(), dtype=dt).reshape(shape)
    ib = abi.Inbox(rows, n, F, gids=gids); ib.flags = abi.INBOX_NO_REQUESTS
    ob = abi.Outbox(rows, n, F, G)
    keep = []
    for obj, names in ((ib, ("op_meta", "op_nr", "op_ab", "ev_meta", "ev_tn", "ev_el")),
                       (ob, [c[0] for c in abi.Outbox.ROW_COLS] + [c[0] for c in abi.Outbox.GROUP_COLS])):
        for name in names:
            a = getattr(obj, name)
            t, v = pinned(a.shape, a.dtype)
            keep.append(t); setattr(obj, name, v)
    ib.op_cd = None; ib.op_e = None
    prev_out = None
    for k in range(60):
        ic = ib.as_c(); ic.op_cd = None; ic.op_e = None
        workload.leader_step(w, k, None if prev_out is None else prev_out.as_c(), ic)      # the peers (not timed)
        oc = ob.as_c()
        t0 = time.perf_counter()
        e.step_begin_host(0, ic, oc)
        e.step_wait_slot(0)
        _ = int(ob.commit_index[0])                                                      # commit column readable
        dt = time.perf_counter() - t0
        if k >= 10:
            lat.append(dt * 1e3)
        prev_out = ob
    return int(((ib.ev_meta & np.uint64(0xF)) != 0).sum())


LEASE = "--lease" in sys.argv


def main():
    G, R = 65536, 3
    F = R - 1
    out = []
    for n_act, rows in ((512, 1), (4096, 1), (65536, 1), (4096, 16), (65536, 4), (65536, 16)):
        cfg = abi.make_cfg(replicas=R, max_groups=G, max_rows=rows)
        e = engine.Engine(cfg)
        init = np.zeros(G, dtype=abi.GROUP_INIT_DTYPE)
        init["ballot"] = -1; init["first_index"] = 1; init["now_ms"] = workload.T0_MS - 2000
        init["term"] = np.arange(G) % 7
        e.open_bulk(0, init)
        gids = None if n_act == G else np.arange(0, G, G // n_act, dtype=np.uint32)[:n_act]
        n = G if gids is None else len(gids)
        w1 = workload.make_wl(0x5EED0002, 1, n, F)
        w = workload.make_wl(0x5EED0002, rows, n, F)
        prev = None
        for ph in (0, 1, 2):                                   # election through the lease path
            ib = abi.Inbox(1, n, F, gids=gids); ib.flags = abi.INBOX_NO_REQUESTS
            ic = ib.as_c(); ic.op_cd = None; ic.op_e = None
            workload.election_step(w1, ph, None if prev is None else prev.as_c(), ic)
            ib.op_cd = None; ib.op_e = None
            prev = e.step(ib)
        lat = []
        if LEASE:
            # engine-owned pinned columns: one contiguous block per direction, so the step is 1 (+1 for an active
            # list) copy up and 1 copy down; with an active list the group columns are compact (n entries, not G)
            prev_c = None
            ev_meta = None
            for k in range(60):
                fl = abi.INBOX_NO_REQUESTS | (0 if gids is None else abi.INBOX_COMPACT_GROUPS)
                L = e.lease(rows, 0 if gids is None else n, 0, fl)
                if gids is not None:
                    L.gids[:] = gids
                L.use(ops=True, events=True, flags=fl)
                ic = L.c.inbox; ic.op_cd = None; ic.op_e = None
                workload.leader_step(w, k, prev_c, ic)                                        # the peers (not timed)
                t0 = time.perf_counter()
                L.begin(); L.wait()
                _ = int(L.out.commit_index[0])
                dt = time.perf_counter() - t0
                if k >= 10:
                    lat.append(dt * 1e3)
                prev_keep = L.outbox_copy(); prev_c = prev_keep.as_c()
                ev_meta = L.ev_meta
            acks = int(((ev_meta & np.uint64(0xF)) != 0).sum())
        else:
            acks = caller_owned(e, w, rows, n, F, G, gids, lat)
        rec = {"buffers": "lease" if LEASE else "caller", "groups_in_step": n, "rows": rows, "acks_per_step": acks, "p50_ms": float(np.percentile(lat, 50)),
               "p99_ms": float(np.percentile(lat, 99)), "acks_per_s_at_p50": acks / (np.percentile(lat, 50) * 1e-3)}
        print(json.dumps(rec), flush=True)
        out.append(rec)
        e.close()


if __name__ == "__main__":
    main()
