#!/usr/bin/env python
"""bench.py — AppendEntries/sec across Raft groups on B200 (BASELINE.json metric).

A "step" is one rafting_step over one batch: ROWS ticks x G groups, each tick carrying one group op
(SUBMIT / heartbeat) and one AppendEntries ack per follower lane.  Unit of work = one AE ack
consumed by the leader path (ack -> Leadership.State update -> quorum index -> commitIndex).

  value     device-resident: inboxes already in HBM (a different, freshly generated >L2 batch each
            step), timed with CUDA events on the engine's stream, max over ranks.
  e2e       the same stream through the C-ABI host path (rafting_lease + rafting_step): pinned host
            inbox -> H2D -> kernel -> D2H of the outbox, every step inside the timed region.
  roofline  algorithmic bytes per ack (192 B at R=3) x acks per launch / mean kernel
            time, against MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline / --impl reference
            the CPU restatement of the reference's EventLoop path (oracle/, "port": the reference is
            Java and there is no JDK in this image) on the host cores, on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED = 0x5EED0002
WORKLOAD = "64K RaftContext groups, 3 replicas, synthetic AppendEntries stream on 1xB200"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--groups", type=int, default=65536, help="groups per GPU")
    p.add_argument("--replicas", type=int, default=3)
    p.add_argument("--rows", type=int, default=16, help="ticks per step")
    p.add_argument("--cpu-groups", type=int, default=65536, help="groups in the CPU baseline sample")
    p.add_argument("--cpu-seconds", type=float, default=12.0)
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--no-secondary", action="store_true", help="skip the vote / follower-request rates (configs #3, #5)")
    return p.parse_args()


def b_ack(R):                      # B_ack(R) = 184 + 8 (R - 2)
    return 184 + 8 * (R - 2)


def measured_traffic(G, R, rows):
    """DRAM bytes per launch from the committed ncu --set full capture (only valid for the profiled shape)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        if (G, R, rows) == (65536, 3, 16):
            return float(t["traffic_bytes_per_launch"]), t["source"]
    except Exception:
        pass
    return None, None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of the reference's EventLoop path) on a bounded sample
# ------------------------------------------------------------------------------------------------
def run_cpu_sample(args, seconds, threads, steps=None, warmup=1):
    from oracle import binding
    from rafting_b200 import abi, workload
    G, R, rows = args.cpu_groups, args.replicas, args.rows
    cfg = abi.make_cfg(replicas=R, max_groups=G, max_rows=rows)
    o = binding.Oracle(cfg)
    init = np.zeros(G, dtype=abi.GROUP_INIT_DTYPE)
    init["ballot"] = -1; init["first_index"] = 1; init["now_ms"] = workload.T0_MS - 2000
    init["term"] = np.arange(G) % 7
    o.open_bulk(0, init)
    w1 = workload.make_wl(SEED, 1, G, R - 1)
    w = workload.make_wl(SEED, rows, G, R - 1)
    out = None
    for ph in (0, 1, 2):
        out = o.step(workload.election_inbox_host(w1, ph, out), threads=threads)
    prev, acks, spent, k, times = None, 0, 0.0, 0, []
    while True:
        ib = workload.leader_inbox_host(w, k, prev)
        n_acks = int(((ib.ev_meta & np.uint64(0xF)) != 0).sum())
        t0 = time.perf_counter()
        prev = o.step(ib, threads=threads)
        dt = time.perf_counter() - t0
        if k >= warmup:
            acks += n_acks; spent += dt; times.append(dt)
        k += 1
        if steps is not None:
            if k >= warmup + steps:
                break
        elif spent >= seconds:
            break
    return {"value": acks / spent if spent > 0 else 0.0, "acks": acks, "seconds": spent, "steps": len(times),
            "ms_per_step": 1e3 * spent / max(1, len(times)),
            "sample": f"{G} groups x {rows} ticks/step x {len(times)} steps of the same keyed stream (first {G} group ids), "
                      f"in-memory log, {threads} loop threads (groups round-robined like EventLoopGroup)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # bound the per-step sample so K steps finish in a few minutes
    res = run_cpu_sample(args, seconds=0, threads=cores, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "AppendEntries/sec across Raft groups", "value": res["value"], "unit": "acks/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "replicas": args.replicas, "rows_per_step": args.rows},
        "cpu_baseline": {"value": res["value"], "unit": "acks/s", "cores": cores, "kind": "port", "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": "acks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is Java; no JDK in this image -> oracle/ (C port of the reference's EventLoop path) is timed",
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_engine(args):
    import torch
    import torch.distributed as dist
    from rafting_b200 import abi, devbatch, engine, workload

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    G, R, rows = args.groups, args.replicas, args.rows
    F = R - 1
    K, W = args.steps, max(args.warmup, 3)
    cfg = abi.make_cfg(replicas=R, max_groups=G, max_rows=rows, device=local)
    e = engine.Engine(cfg)
    gid_base = rank * G
    init = np.zeros(G, dtype=abi.GROUP_INIT_DTYPE)
    init["ballot"] = -1; init["first_index"] = 1; init["now_ms"] = workload.T0_MS - 2000
    init["term"] = (gid_base + np.arange(G)) % 7
    e.open_bulk(0, init)
    if world > 1:
        box = [engine.Engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        e.comm_init(rank, world, box[0])
    else:
        e.comm_init(0, 1, None)
    stream_ptr = e.stream()
    ext = torch.cuda.ExternalStream(stream_ptr, device=dev)

    # ---- election warm-up + settle, all on the device ------------------------------------------
    w1 = workload.make_wl(SEED, 1, G, F, gid_base=gid_base)
    w = workload.make_wl(SEED, rows, G, F, gid_base=gid_base)
    in1, out1 = devbatch.DevInbox(1, G, F, dev), devbatch.DevOutbox(1, G, F, G, dev)
    prev_c = None
    for ph in (0, 1, 2):
        ic = in1.as_c()
        workload.election_step(w1, ph, prev_c, ic, on_device=True, stream=stream_ptr)
        oc = out1.as_c()
        e.step_device(ic, oc, stream_ptr)
        prev_c = oc
    outs = [devbatch.DevOutbox(rows, G, F, G, dev) for _ in range(2)]
    n_rec = W + K
    # the leader stream never marks a follower unavailable: the op_ab column (its only field used by SUBMIT / TIMEOUT)
    # is omitted from the batch, as a shim would do
    inboxes = [devbatch.DevInbox(rows, G, F, dev, unavail=False) for _ in range(n_rec)]
    settle = devbatch.DevInbox(rows, G, F, dev, unavail=False)
    SETTLE = 3
    prev_out = None
    for k in range(SETTLE):
        ic = settle.as_c()
        workload.leader_step(w, k, None if prev_out is None else prev_out.as_c(), ic, on_device=True, stream=stream_ptr)
        prev_out = outs[k % 2]
        e.step_device(ic, prev_out.as_c(), stream_ptr)
    torch.cuda.synchronize()
    # keep the outbox the recorded stream starts from, then checkpoint the tables
    start_out = devbatch.DevOutbox(rows, G, F, G, dev)
    for name in start_out.t:
        start_out.t[name].copy_(prev_out.t[name])
    torch.cuda.synchronize()
    e.checkpoint()

    # ---- phase A: generate + record the stream closed-loop (untimed) ------------------------------
    prev_out = start_out
    acks_per_step = []
    for k in range(n_rec):
        ic = inboxes[k].as_c()
        workload.leader_step(w, SETTLE + k, prev_out.as_c(), ic, on_device=True, stream=stream_ptr)
        prev_out = outs[k % 2]
        e.step_device(ic, prev_out.as_c(), stream_ptr)
    torch.cuda.synchronize()
    for k in range(n_rec):
        m = inboxes[k].t["ev_meta"].view(torch.int64)
        acks_per_step.append(int(((m & 0xF) != 0).sum().item()))
    digest_a = e.digest(0, G)

    # ---- phase B: timed replay, inputs resident in HBM ---------------------------------------------
    e.restore()
    sampler = ClockSampler(local); sampler.start()
    out_b = outs[0]
    ics = [ib.as_c() for ib in inboxes]
    ocs = [outs[k % 2].as_c() for k in range(n_rec)]
    def warm():
        for k in range(W):
            e.step_device(ics[k], ocs[k], stream_ptr)
            if world > 1:
                e.allgather_commit(to_host=False)
        if world > 1:
            e.allgather_join()
        torch.cuda.synchronize(); barrier()

    # pass 1 — the metric: K steps back to back, two events around the whole region
    warm()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(ext)
    for j in range(K):
        e.step_device(ics[W + j], ocs[W + j], stream_ptr)
        if world > 1:
            e.allgather_commit(to_host=False)
    if world > 1:
        e.allgather_join()                      # the timed region ends when the last summary has been gathered
    ev1.record(ext)
    torch.cuda.synchronize(); barrier()
    total_ms = ev0.elapsed_time(ev1)
    digest_b = e.digest(0, G)
    # pass 2 — the same K steps again with an event pair around every kernel (roofline of the dominant kernel)
    e.restore()
    warm()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(2 * K)]
    for j in range(K):
        evs[2 * j].record(ext)
        e.step_device(ics[W + j], ocs[W + j], stream_ptr)
        evs[2 * j + 1].record(ext)
    torch.cuda.synchronize(); barrier()
    kern_ms = [evs[2 * j].elapsed_time(evs[2 * j + 1]) for j in range(K)]

    replay_ok = bool((digest_a == digest_b).all())
    acks_timed = sum(acks_per_step[W:])
    launches0, _ = e.counters()

    # ---- e2e: the same stream through the C-ABI host path, HOST buffers in, HOST buffers out ---------
    e2e = None
    lat_ms = []
    if not args.no_e2e:
        # the transport's pinned receive buffers: one pinned inbox per timed step (filled before the clock
        # starts, as Netty would have decoded them), two pinned outboxes (one per slot)
        K2 = min(K, 24)                                   # pinned host memory is bounded: at most 24 timed e2e steps
        host_in = []
        for k in range(W, W + K2):
            cols = {name: t.cpu().pin_memory() for name, t in inboxes[k].t.items()}
            ic = abi.InboxC()
            ic.rows, ic.n_active, ic.flags = rows, 0, abi.INBOX_NO_REQUESTS
            for name, t in cols.items():
                setattr(ic, name, t.data_ptr())
            host_in.append((cols, ic))
        NSL = 3                                           # steps in flight on the host path
        host_out = []
        for sl in range(NSL):
            cols = {name: torch.zeros(t.numel(), dtype=torch.uint8).pin_memory() for name, t in outs[0].t.items()}
            oc = abi.OutboxC()
            for name, t in cols.items():
                setattr(oc, name, t.data_ptr())
            host_out.append((cols, oc))
        h2d = sum(t.numel() for t in host_in[0][0].values())
        sparse = ("rep_term", "ballot_term", "ballot_last")        # copied down only when a step produced replies / ballots
        d2h = sum(t.numel() for name, t in host_out[0][0].items() if name not in sparse) + 16

        def rewind():
            e.restore()
            for k in range(W):
                e.step_device(ics[k], ocs[k], stream_ptr)
            torch.cuda.synchronize()

        # untimed: touch both slots once so their device staging exists before the clock starts
        rewind()
        for sl in range(NSL):
            e.step_begin_host(sl, host_in[sl % len(host_in)][1], host_out[sl][1])
        for sl in range(NSL):
            e.step_wait_slot(sl)
        # (1) throughput: two slots in flight — H2D of step j+1, kernel of step j and D2H of step j-1 overlap
        rewind(); barrier()
        t0 = time.perf_counter()
        for j in range(K2):
            sl = j % NSL
            if j >= NSL:
                e.step_wait_slot(sl)                      # outbox of step j-NSL is readable on the host
            e.step_begin_host(sl, host_in[j][1], host_out[sl][1])
        for sl in range(NSL):
            e.step_wait_slot(sl)
        spent = time.perf_counter() - t0
        acks_e2e = sum(acks_per_step[W:W + K2])
        e2e_ok = None                                     # verified only when the e2e pass replays all K steps
        if K2 == K:
            digest_c = e.digest(0, G)
            e2e_ok = bool((digest_a == digest_c).all())
        # (2) latency: one step at a time, host ack in -> commit record readable out
        rewind()
        for j in range(min(K2, 12)):
            t1 = time.perf_counter()
            e.step_begin_host(0, host_in[j][1], host_out[0][1])
            e.step_wait_slot(0)
            lat_ms.append((time.perf_counter() - t1) * 1e3)
        t = torch.tensor([spent], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ae = torch.tensor([acks_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ae, op=dist.ReduceOp.SUM)
        e2e = {"spent": float(t.item()), "h2d": int(h2d), "d2h": int(d2h), "ok": e2e_ok, "acks": float(ae.item()), "steps": K2}
    sampler.stop_flag = True

    # ---- reduce over ranks ------------------------------------------------------------------------
    tt = torch.tensor([total_ms, float(np.mean(kern_ms))], dtype=torch.float64, device=dev)
    aa = torch.tensor([acks_timed], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(aa, op=dist.ReduceOp.SUM)
    total_ms_max, kern_ms_max = float(tt[0].item()), float(tt[1].item())
    acks_all = float(aa.item())

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        cpu = run_cpu_sample(args, seconds=args.cpu_seconds, threads=cores)
        cpu3 = run_cpu_sample(args, seconds=min(4.0, args.cpu_seconds), threads=3)
        cpu["cores"] = cores; cpu["t3"] = cpu3["value"]

    # vote replies (config #3) and follower-side AppendEntries requests (config #5) are separate rates,
    # device-resident at full size; reported next to the headline, not part of `value`
    secondary = None
    if rank == 0 and world == 1 and not args.no_secondary and not args.no_e2e:
        import importlib.util
        spec = importlib.util.spec_from_file_location("bench_secondary", os.path.join(ROOT, "tools", "bench_secondary.py"))
        mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
        secondary = []
        for r in mod.run_all(quiet=True, steps5=64):
            secondary.append({"config": r["config"], "groups": r["groups"], "replicas": r["replicas"], "rows_per_step": r["rows_per_step"],
                              "steps": r["steps"], "kernel_ms_per_step": r["kernel_ms_per_step_median"],
                              "rates_per_s": r["rates_per_s"], "roofline_frac": r["roofline"]["frac"]})

    if rank == 0:
        peak, peak_src = measured_peak()
        acks_per_launch = acks_timed / K
        # dominant kernel's launch duration: at N=1 the timed region holds nothing but the K step kernels, so the
        # region's own CUDA-event time / K is the unperturbed figure; with N>1 the region also holds the gathers, so
        # the per-kernel event pairs of pass 2 are used (they include ~2 us of event overhead per launch)
        k_ms = total_ms / K if world == 1 else float(np.mean(kern_ms))
        achieved = acks_per_launch * b_ack(R) / (k_ms * 1e-3) / 1e9
        traffic, traffic_src = measured_traffic(G, R, rows)
        value = acks_all / (total_ms_max * 1e-3)
        line = {
            "metric": "AppendEntries/sec across Raft groups", "value": value, "unit": "acks/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": total_ms_max / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": WORKLOAD if world == 1 else
                       f"{world * G // 1024}K RaftContext groups, 3 replicas, synthetic AppendEntries stream sharded across {world}xB200 "
                       f"with NCCL commitIndex all-gather (64K groups per GPU, weak scaling of the 1xB200 configuration)",
                       "groups_per_gpu": G, "replicas": R, "rows_per_step": rows,
                       "acks_per_step_per_gpu": acks_per_launch,
                       "inputs": f"every step reads a distinct pre-generated inbox resident in HBM "
                                 f"({inboxes[0].nbytes() / 1e6:.0f} MB inbox + {outs[0].nbytes() / 1e6:.0f} MB outbox per step, > L2), no L2 flush needed",
                       "collective": "ncclAllGather of commitIndex[G] after every step, on its own stream behind the producing kernel" if world > 1 else "none (1 GPU)",
                       "bit_exact_replay": replay_ok},
            "gpu_launches": K,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "bytes_per_ack": b_ack(R), "algorithmic_bytes_per_launch": acks_per_launch * b_ack(R),
                         "kernel_ms": k_ms, "kernel_ms_event_pair_per_launch": float(np.mean(kern_ms)),
                         "kernel": "rafting::unrolled::step_kernel<FT=R-1,NST=3>"},
            "clocks": sampler.summary(),
        }
        if e2e:
            ev = e2e["acks"] / e2e["spent"]
            line["e2e"] = {"value": ev, "unit": "acks/s", "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                           "bit_exact_replay": e2e["ok"], "steps": e2e["steps"],
                           "note": "wall clock around K x rafting_step_begin_host/rafting_step_wait_slot with caller-owned pinned "
                                   "buffers, three slots in flight (H2D / kernel / D2H of successive steps overlap); every step's inbox "
                                   "crosses PCIe up and its outbox crosses PCIe down inside the timed region (the payload columns of "
                                   "replies / ballots only when the step produced any)"}
            line["commit_latency_ms"] = {"p50": float(np.percentile(lat_ms, 50)), "p99": float(np.percentile(lat_ms, 99)),
                                         "what": "one synchronous step: host ack in pinned inbox -> commit record readable in pinned outbox"}
        if cpu:
            line["cpu_baseline"] = {"value": cpu["value"], "unit": "acks/s", "cores": cpu["cores"], "kind": "port",
                                    "sample": cpu["sample"], "t3_loop_threads_value": cpu["t3"]}
        if secondary:
            line["secondary_rates"] = secondary
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The one JSON line, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # stdout carries exactly one JSON line: anything libraries print on fd 1 while the bench runs (NCCL prints its
    # version banner there) is diverted to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
