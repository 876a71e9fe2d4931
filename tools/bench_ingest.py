#!/usr/bin/env python
"""Host-side rates of the transport half (include/rafting_ingest.h), one core, no GPU: cutting a receive buffer of ACK frames
in the reference's wire layout (EventCodec) and turning them into (gid, kind, sequence, term, success) records
(rafting_ack_frames_decode: scope -> context registry -> Kryo RaftResponse body).  Prints one JSON line."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rafting_b200 import ingest  # noqa: E402


def main(groups=65536, frames_n=400000, reps=7):
    cm = ingest.CtxMap()
    for g in range(groups):
        cm.put(b"ctx-%d" % g, g)
    rng = np.random.default_rng(1)
    gids = rng.integers(0, groups, frames_n)
    stream = b"".join(ingest.encode(ingest.ACK, b"appendEntries:ctx-%d" % g, ingest.reply_body_encode(1000 + i % 7, True), sequence=i)
                      for i, g in enumerate(gids))
    buf = np.frombuffer(stream, dtype=np.uint8)
    frames = np.zeros(frames_n, dtype=ingest.FRAME)
    recs = np.zeros(frames_n, dtype=ingest.ACK_REC)
    n, used, tr, m = C.c_uint32(), C.c_size_t(), C.c_int(), C.c_uint32()
    L = ingest.lib()
    best = [1e9, 1e9]
    for _ in range(reps):
        t0 = time.perf_counter()
        L.rafting_frame_scan(buf.ctypes.data, len(stream), frames.ctypes.data, frames_n, C.byref(n), C.byref(used), C.byref(tr))
        t1 = time.perf_counter()
        L.rafting_ack_frames_decode(buf.ctypes.data, frames.ctypes.data, n.value, cm._h, recs.ctypes.data, C.byref(m))
        t2 = time.perf_counter()
        best = [min(best[0], t1 - t0), min(best[1], t2 - t1)]
    assert n.value == frames_n == m.value and (recs["gid"] == gids).all()
    # replies -> compact words: one reply per lane slot of a row (every follower of every group answers once), in random order;
    # each sequence is pending under a tag
    from rafting_b200 import compact, abi as _abi
    Fl = 2
    n_slots = groups * Fl
    order = np.random.default_rng(2).permutation(n_slots)
    acks = np.zeros(n_slots, dtype=ingest.ACK_REC)
    acks["gid"], acks["kind"], acks["success"], acks["sequence"], acks["term"] = order // Fl, _abi.EV_AE_ACK, 1, np.arange(n_slots), 7
    esc = np.zeros(1024, dtype=_abi.CESC_IN)
    deferred = np.zeros(n_slots, dtype=np.uint32)
    t_c = 1e9
    for _ in range(3):
        pend = ingest.Pending(n_slots)
        hp = pend._h
        for sq, slot in enumerate(order):
            L.rafting_pending_put(hp, 0, sq, _abi.EV_AE_ACK, int(slot) // Fl, int(slot) % Fl, sq % 32, 1, 7, 0, 0)
        cin = compact.CompactInbox(1, groups, Fl)
        cin.row_base[0] = 5
        cc = cin.as_c()
        ne, nd, nu = C.c_uint32(), C.c_uint32(), C.c_uint32()
        t0 = time.perf_counter()
        rc = L.rafting_acks_to_cinbox(hp, 0, acks.ctypes.data, n_slots, 9, 0, C.byref(cc), groups, Fl, esc.ctypes.data, len(esc),
                                      C.byref(ne), deferred.ctypes.data, C.byref(nd), C.byref(nu))
        t_c = min(t_c, time.perf_counter() - t0)
        assert rc == 0 and nu.value == 0 and nd.value == 0 and ne.value == 0 and len(pend) == 0 and int((cin.ev_c != 0).sum()) == n_slots
    # the dispatch loop over a config-#2-sized dense outbox (64 K groups x 2 followers x 16 rows, 80 % of the lane slots planned)
    from rafting_b200 import abi
    G, F, rows = 65536, 2, 16
    ob = abi.Outbox(rows, G, F, G)
    ob.incarnation[:] = 1
    ob.current_term[:] = 7
    planned = np.random.default_rng(3).random((rows, G, F)) < 0.8
    ob.plan_meta[:] = np.where(planned, np.uint64(abi.PLAN_AE | (1 << 16) | (1 << 32)), np.uint64(0))
    disp = ingest.Dispatch(G, F, 0)
    cap = int(planned.sum()) + 16
    reqs = np.zeros(cap, dtype=ingest.REQ_REC)
    nr, unk, oc = C.c_uint32(), C.c_uint32(), ob.as_c()
    t_disp = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        rc = L.rafting_outbox_to_requests(disp._h, C.byref(oc), rows, reqs.ctypes.data, cap, C.byref(nr), C.byref(unk))
        t_disp = min(t_disp, time.perf_counter() - t0)
    assert rc == 0 and nr.value == int(planned.sum())
    print(json.dumps({"what": "ACK frames in the reference's wire layout, one core", "contexts": groups, "frames": frames_n,
                      "bytes_per_frame": len(stream) / frames_n,
                      "frame_scan": {"frames_per_s": frames_n / best[0], "GB_per_s": len(stream) / best[0] / 1e9},
                      "ack_frames_decode": {"acks_per_s": frames_n / best[1]},
                      "both": {"acks_per_s": frames_n / (best[0] + best[1])},
                      "acks_to_cinbox": {"acks_per_s": n_slots / t_c, "acks": n_slots, "note": "one reply per lane slot of a 64 K x 2 row, random order, every sequence pending under a tag"},
                      "outbox_to_requests": {"plans": nr.value, "ms": t_disp * 1e3, "records_per_s": nr.value / t_disp,
                                             "outbox": "64 K groups x 2 x 16 rows, dense"}}))


if __name__ == "__main__":
    main()
