#!/usr/bin/env python
"""Measurement of the HBM segmented entry buffer: append throughput from host payloads and
the gather kernels' achieved bandwidth against the HBM roofline (bytes read from the arena + bytes written to the
send buffer, over the CUDA-event time of the two gather kernels).  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rafting_b200 import abi, engine  # noqa: E402


def main():
    G, per_group, payload = 65536, 32, 256
    cfg = abi.make_cfg(replicas=3, max_groups=G, max_rows=1)
    e = engine.Engine(cfg)
    init = np.zeros(G, dtype=abi.GROUP_INIT_DTYPE)
    init["ballot"] = -1; init["first_index"] = 1; init["last_index"] = per_group; init["last_term"] = 1; init["term"] = 1
    e.open_bulk(0, init)
    e.log_config(segment_bytes=1 << 22, hbm_segments=256, ring_slots=64)      # 1 GiB arena
    refs = np.zeros(G * per_group, dtype=engine.Engine.ENTRY_REF)
    refs["gid"] = np.repeat(np.arange(G, dtype=np.uint32), per_group)
    refs["index"] = np.tile(np.arange(1, per_group + 1, dtype=np.int64), G)
    refs["term"] = 1
    refs["len"] = payload
    refs["blob_off"] = np.arange(G * per_group, dtype=np.uint64) * payload
    blob = np.random.default_rng(1).integers(0, 256, size=G * per_group * payload, dtype=np.uint8)
    L = engine.lib()
    t0 = time.perf_counter()
    rc = L.rafting_log_append(e._h, refs.ctypes.data, len(refs), blob.ctypes.data, blob.nbytes)
    assert rc == 0
    import ctypes as C
    # a read forces completion of the enqueued copies
    e.log_read(0, 1, 1)
    t_append = time.perf_counter() - t0
    # gather: every group's 16 newest entries (what a step's AE plans would ask for)
    n_ranges = G
    g = np.arange(G, dtype=np.uint32); f = np.full(G, per_group - 15, dtype=np.int64); c = np.full(G, 16, dtype=np.uint32)
    cap = G * 16 + 1
    out_refs = np.zeros(cap, dtype=engine.Engine.ENTRY_REF)
    out_blob = np.zeros(G * 16 * (payload + 16) + 4096, dtype=np.uint8)
    n, nb = C.c_uint32(), C.c_size_t()
    best = None
    for _ in range(5):
        rc = L.rafting_log_gather(e._h, n_ranges, g.ctypes.data, f.ctypes.data, c.ctypes.data, out_refs.ctypes.data, cap,
                                  out_blob.ctypes.data, out_blob.nbytes, C.byref(n), C.byref(nb))
        assert rc == 0
        st = e.log_stats()
        ms = st["gather_kernel_ns"] / 1e6
        best = ms if best is None else min(best, ms)
    # verify a sample
    k = 12345
    r = out_refs[k]
    src = blob[(int(r["gid"]) * per_group + int(r["index"]) - 1) * payload:][:payload]
    assert np.array_equal(out_blob[int(r["blob_off"]):int(r["blob_off"]) + payload], src)
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    moved = 2 * nb.value + n.value * (32 + 8 + 24 + 4)   # payload read + written; header, ring slot, request, length per entry
    print(json.dumps({
        "what": "HBM segmented entry buffer", "groups": G, "entries": int(len(refs)), "payload_bytes": payload,
        "append": {"seconds": t_append, "entries_per_s": len(refs) / t_append, "GBps_host_to_hbm": blob.nbytes / t_append / 1e9,
                   "note": "layout arithmetic + host index on the host, blob H2D + scatter kernel on the device, wall clock"},
        "gather": {"entries": int(n.value), "payload_bytes": int(nb.value), "kernel_ms": best,
                   "achieved_GBps": moved / (best * 1e-3) / 1e9, "peak_GBps": peak, "frac": moved / (best * 1e-3) / 1e9 / peak,
                   "bound": "hbm", "note": "seglog_gather_kernel, CUDA events on the engine stream, best of 5"},
        "stats": st,
    }))


if __name__ == "__main__":
    main()
