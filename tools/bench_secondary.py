#!/usr/bin/env python
"""Secondary rates of SURVEY.md §8(d), device-resident, at BASELINE.json's full sizes on one GPU:

  config #3  256 K groups x 5 replicas, PreVote on, RequestVote storm  -> vote replies + vote requests per second
  config #5  512 K groups x 3 replicas, leader churn + InstallSnapshot catch-up
                                                                        -> acks, follower-side AE requests, entries per second

The peers are the closed-loop generators of workload.cu running on the device (the same streams the full-size
parity tests replay against the oracle); every step is generator kernel -> step kernel on the engine's stream, and
only the step kernel is inside the CUDA-event pairs (second pass over the same stream after a rollback of the tables).  Algorithmic bytes per unit are SURVEY §8(d)'s:
B_vote = 64, B_ack(R) = 184 + 8(R-2), B_req(n) = 160 + 32 n.  Prints one JSON line per configuration."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rafting_b200 import abi, devbatch, engine, workload  # noqa: E402
from rafting_b200.workload import _bind  # noqa: E402
import ctypes as C  # noqa: E402


def peak_gbs():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 7700.0 * 0.92, "fallback (B200_PROFILING.md: 92 % of 7.7 TB/s)"


def init_array(G, **cols):
    a = np.zeros(G, dtype=abi.GROUP_INIT_DTYPE)
    a["ballot"] = -1; a["first_index"] = 1; a["now_ms"] = workload.T0_MS - 2000
    for k, v in cols.items():
        a[k] = v
    return a


def run(name, G, R, rows, steps, seed, gen, local_slot=0, elect=False, pool=False, quiet=False):
    import torch
    F = R - 1
    dev = torch.device("cuda:0")
    cfg = abi.make_cfg(replicas=R, local_slot=local_slot, max_groups=G, max_rows=rows, pre_vote=True,
                       entry_pool_cap=workload.POOL_TERMS if pool else 0)
    e = engine.Engine(cfg)
    if name == "config3":
        e.open_bulk(0, init_array(G, term=1 + np.arange(G) % 5, last_index=100 + np.arange(G) % 50,
                                  last_term=1 + np.arange(G) % 5))
    else:
        e.open_bulk(0, init_array(G, term=np.arange(G) % 7))
    st = torch.cuda.ExternalStream(e.stream(), device=dev)
    ob = [devbatch.DevOutbox(rows, G, F, G, dev) for _ in range(2)]
    ib = devbatch.DevInbox(rows, G, F, dev, requests=True)
    pool_t = None
    if pool:
        pool_t = torch.zeros(workload.POOL_TERMS, dtype=torch.int64, device=dev)
        _bind().rafting_wl_fill_term_pool(C.c_void_p(pool_t.data_ptr()), workload.POOL_TERMS, 1, C.c_void_p(e.stream()))
    prev = None
    if elect:
        w1 = workload.make_wl(seed, 1, G, F, local_slot=local_slot)
        ib1 = devbatch.DevInbox(1, G, F, dev, requests=False)
        ob1 = [devbatch.DevOutbox(1, G, F, G, dev) for _ in range(2)]
        for ph in (0, 1, 2):
            workload.election_step(w1, ph, prev, ib1.as_c(), on_device=True, stream=e.stream())
            oc = ob1[ph & 1].as_c()
            e.step_device(ib1.as_c(), oc, e.stream())
            prev = oc
        prev = None
    w = workload.make_wl(seed, rows, G, F, local_slot=local_slot)
    # pass 0 warms the kernels of this configuration up (module load, first-launch costs); the tables are then rolled back
    # (rafting_checkpoint / rafting_restore) and pass 1 — the identical stream, the generators are pure functions of the
    # previous outbox — is the timed one
    e.checkpoint()
    for timed in (False, True):
        if timed:
            e.restore()
        prev = None
        tot = dict(acks=0, votes=0, vote_requests=0, ae_requests=0, is_requests=0, entries=0, ops_other=0)
        ms = []
        with torch.cuda.stream(st):
            for k in range(steps):
                ic = ib.as_c()
                if pool:
                    ic.ent_terms = pool_t.data_ptr(); ic.ent_count = workload.POOL_TERMS
                rc = gen(C.byref(w), k, None if prev is None else C.byref(prev), C.byref(ic), 1, C.c_void_p(e.stream()))
                assert rc == 0, rc
                oc = ob[k & 1].as_c()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(st)
                e.step_device(ic, oc, e.stream())
                b.record(st)
                b.synchronize()
                ms.append(a.elapsed_time(b))
                prev = oc
                evk = ib.t["ev_meta"].view(torch.int64) & 0xF
                opm = ib.t["op_meta"].view(torch.int64)
                opk = opm & 0xFF
                tot["acks"] += int(((evk == abi.EV_AE_ACK) | (evk == abi.EV_IS_ACK)).sum())
                tot["votes"] += int(((evk == abi.EV_PV_REPLY) | (evk == abi.EV_RV_REPLY)).sum())
                tot["vote_requests"] += int(((opk == abi.OP_PREVOTE_REQ) | (opk == abi.OP_VOTE_REQ)).sum())
                ae = opk == abi.OP_AE_REQUEST
                tot["ae_requests"] += int(ae.sum())
                tot["entries"] += int((((opm >> 16) & 0xFFFF) * ae).sum())
                tot["is_requests"] += int((opk == abi.OP_IS_REQUEST).sum())
                tot["ops_other"] += int(((opk == abi.OP_SUBMIT) | (opk == abi.OP_TIMEOUT) | (opk == abi.OP_FLUSH)).sum())
    role = ob[(steps - 1) & 1].t["role_word"].view(torch.int32) & 3
    err = (ob[(steps - 1) & 1].t["err_word"].view(torch.int32) & 0xFFFF) != 0
    secs = sum(ms) * 1e-3
    peak, src = peak_gbs()
    b_ack = 184 + 8 * (R - 2)
    alg = tot["acks"] * b_ack + (tot["votes"] + tot["vote_requests"]) * 64 + \
        (tot["ae_requests"] + tot["is_requests"]) * 160 + tot["entries"] * 32
    rec = {
        "config": name, "groups": G, "replicas": R, "rows_per_step": rows, "steps": steps, "seed": hex(seed),
        "kernel_ms_total": sum(ms), "kernel_ms_per_step_median": float(np.median(ms)),
        "units": tot,
        "rates_per_s": {k: v / secs for k, v in tot.items() if v},
        "roofline": {"bound": "hbm", "algorithmic_bytes": alg, "achieved": alg / secs / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg / secs / 1e9 / peak, "peak_source": src,
                     "bytes_per_unit": {"ack": b_ack, "vote": 64, "request": "160 + 32 n"}},
        "end_state": {"roles": torch.bincount(role.long(), minlength=3).tolist(), "groups_with_error": int(err.sum())},
    }
    if not quiet:
        print(json.dumps(rec), flush=True)
    e.close()
    return rec


def run_all(quiet=False, steps5=64):
    L = _bind()
    return [run("config3", 262144, 5, 1, 8, 0x5EED0003, L.rafting_wl_vote_step, local_slot=2, quiet=quiet),
            run("config5", 524288, 3, 4, steps5, 0x5EED0005, L.rafting_wl_mixed_step, elect=True, pool=True, quiet=quiet)]


def main():
    run_all()


if __name__ == "__main__":
    main()
