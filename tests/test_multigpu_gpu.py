"""Cross-shard commitIndex summary on real GPUs (`-m gpu`; SURVEY.md §8(e), BASELINE config #4).

Groups shard by contiguous gid blocks; the only exchange is one all-gather of commitIndex[G_local] per step.  The pass
criterion is SURVEY §8(d) #4's: the gathered [world * G_local] vector on EVERY rank equals the oracle's commitIndex
over all groups.  Three shapes:
  * config #4 at its stated size (1 M groups, 3 replicas) over every visible GPU (1, 2, 4 or 8 shards) in ONE process
    (rafting_comm_init_all — the reference's host is one JVM);
  * two shards, one process, gathering after every step from the step's outbox commit column (device path, overlapped);
  * two shards, one PROCESS per GPU (the torchrun shape bench.py uses, rafting_comm_init + NCCL unique id).
The two-shard tests skip on a box with fewer than 2 GPUs.
"""
import os
import socket

import numpy as np
import pytest

from oracle import binding
from rafting_b200 import abi, workload
from tests import harness

pytestmark = pytest.mark.gpu

SEED = 0x5EED0004
R = 3


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _open_shard(engine_mod, rank, g_local, rows, device):
    cfg = abi.make_cfg(replicas=R, max_groups=g_local, max_rows=rows, device=device)
    e = engine_mod.Engine(cfg)
    o = binding.Oracle(cfg)
    init = harness.init_array(g_local, terms=(rank * g_local + np.arange(g_local)) % 7)
    e.open_bulk(0, init), o.open_bulk(0, init)
    return cfg, e, o


def _elect(e, o, w1, threads):
    out = None
    for ph in (0, 1, 2):
        ib = workload.election_inbox_host(w1, ph, out)
        out = o.step(ib, threads=threads)
        harness.assert_outbox_equal(out, e.step(ib), where=f"election {ph}")


def test_config4_1m_groups_sharded_over_every_visible_gpu_one_process():
    from rafting_b200 import engine
    G_TOTAL, rows, steps, T = 1 << 20, 2, 3, 32
    world = max(w for w in (1, 2, 4, 8) if w <= max(1, _ngpu()))
    g_local = G_TOTAL // world
    shards = [_open_shard(engine, r, g_local, rows, r) for r in range(world)]
    engines = [s[1] for s in shards]
    engine.Engine.comm_init_all(engines)
    wl1 = [workload.make_wl(SEED, 1, g_local, R - 1, gid_base=r * g_local) for r in range(world)]
    wl = [workload.make_wl(SEED, rows, g_local, R - 1, gid_base=r * g_local) for r in range(world)]
    for r, (_, e, o) in enumerate(shards):
        _elect(e, o, wl1[r], T)
    prev = [None] * world
    for k in range(steps):
        want = []
        for r, (_, e, o) in enumerate(shards):
            ib = workload.leader_inbox_host(wl[r], k, prev[r])
            prev[r] = o.step(ib, threads=T)
            harness.assert_outbox_equal(prev[r], e.step(ib), where=f"shard {r} step {k}")
            want.append(prev[r].commit_index)
        want = np.concatenate(want)
        got = engine.Engine.allgather_commit_all(engines)                    # source: the live table column
        for r in range(world):
            assert np.array_equal(got[r], want), f"step {k}: rank {r}'s gathered commitIndex[{G_TOTAL}] differs from the oracle"
    assert want.shape == (G_TOTAL,) and want.max() > 0 and (want > 0).mean() > 0.05      # (six ticks in: the first commits)
    for _, e, _ in shards:
        e.close()


@pytest.mark.skipif("_ngpu() < 2")
def test_two_shards_one_process_gather_from_the_outbox_column_device_path():
    """Device path: every step is followed by a gather of ITS outbox commit column while the next step is already
    enqueued (two outboxes rotating); each gathered vector must be the state after exactly that step."""
    import torch
    from rafting_b200 import devbatch, engine
    g_local, rows, steps, world, T = 8192, 4, 6, 2, 8
    shards = [_open_shard(engine, r, g_local, rows, r) for r in range(world)]
    engines = [s[1] for s in shards]
    engine.Engine.comm_init_all(engines)
    wl1 = [workload.make_wl(SEED, 1, g_local, R - 1, gid_base=r * g_local) for r in range(world)]
    wl = [workload.make_wl(SEED, rows, g_local, R - 1, gid_base=r * g_local) for r in range(world)]
    for r, (_, e, o) in enumerate(shards):
        _elect(e, o, wl1[r], T)
    # oracle stream first (host), recorded per shard and step
    inboxes, wants = [[] for _ in range(world)], []
    prev = [None] * world
    for k in range(steps):
        for r, (_, _, o) in enumerate(shards):
            ib = workload.leader_inbox_host(wl[r], k, prev[r])
            prev[r] = o.step(ib, threads=T)
            inboxes[r].append(ib)
        wants.append(np.concatenate([prev[r].commit_index for r in range(world)]))
    # engines: device-resident inboxes, two rotating outboxes, gather k overlaps step k+1
    dev_in, dev_out = [], []
    for r in range(world):
        dev = torch.device("cuda", r)
        dev_in.append([devbatch.DevInbox.from_host(inboxes[r][k], dev) for k in range(steps)])
        dev_out.append([devbatch.DevOutbox(rows, g_local, R - 1, g_local, dev) for _ in range(2)])
    gathered = []
    for k in range(steps):
        srcs = []
        for r, e in enumerate(engines):
            oc = dev_out[r][k % 2].as_c()
            e.step_device(dev_in[r][k].as_c(), oc, 0)
            srcs.append(oc.commit_index)
        gathered.append(engine.Engine.allgather_commit_all(engines, srcs))
    for k in range(steps):
        for r in range(world):
            assert np.array_equal(gathered[k][r], wants[k]), f"gather after step {k} on rank {r} is not the state after step {k}"
    for e in engines:
        e.close()


def _proc_per_gpu(rank, world, port, g_local, rows, steps, out_dir):
    import torch
    import torch.distributed as dist
    from rafting_b200 import engine
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(rank)
    _, e, o = _open_shard(engine, rank, g_local, rows, rank)
    box = [engine.Engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    e.comm_init(rank, world, box[0])
    w1 = workload.make_wl(SEED, 1, g_local, R - 1, gid_base=rank * g_local)
    w = workload.make_wl(SEED, rows, g_local, R - 1, gid_base=rank * g_local)
    _elect(e, o, w1, 4)
    prev = None
    for k in range(steps):
        ib = workload.leader_inbox_host(w, k, prev)
        prev = o.step(ib, threads=4)
        harness.assert_outbox_equal(prev, e.step(ib), where=f"rank {rank} step {k}")
        got = e.allgather_commit(to_host=True)
        mine = torch.from_numpy(prev.commit_index.copy())
        parts = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)                                          # the ORACLE's shards, over gloo
        assert np.array_equal(got, torch.cat(parts).numpy()), f"rank {rank} step {k}: NCCL-gathered vector != oracle shards"
    np.save(os.path.join(out_dir, f"ok_{rank}.npy"), got)
    e.close()
    dist.destroy_process_group()


@pytest.mark.skipif("_ngpu() < 2")
def test_two_shards_one_process_per_gpu_nccl_unique_id(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_proc_per_gpu, args=(world, port, 4096, 4, 5, str(tmp_path)), nprocs=world, join=True)
    a, b = np.load(tmp_path / "ok_0.npy"), np.load(tmp_path / "ok_1.npy")
    assert np.array_equal(a, b) and a.shape == (2 * 4096,) and a.max() > 0
