"""world_size-2 test of the N>1 host logic on CPU (gloo): groups shard by contiguous gid blocks, each
rank drives its shard with the stream keyed by GLOBAL group id, and the all-gathered commitIndex
vector equals what one process computes over all groups (SURVEY.md §8e, config #4 shape)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import binding
from rafting_b200 import abi, workload
from tests import harness

G_LOCAL, R, ROWS, STEPS, SEED = 96, 3, 3, 6, 0x5EED0004


def _run_shard(gid_base, n):
    cfg = abi.make_cfg(replicas=R, max_groups=n, max_rows=ROWS)
    o = binding.Oracle(cfg)
    o.open_bulk(0, harness.init_array(n, terms=(gid_base + np.arange(n)) % 7))
    w1 = workload.make_wl(SEED, 1, n, R - 1, gid_base=gid_base)
    w = workload.make_wl(SEED, ROWS, n, R - 1, gid_base=gid_base)
    harness.elect_all(o, w1)
    last = harness.run_leader_workload([o], w, steps=STEPS, compare=False)
    return last.commit_index.copy()


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = torch.from_numpy(_run_shard(rank * G_LOCAL, G_LOCAL))
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    np.save(os.path.join(out_dir, f"gather_{rank}.npy"), torch.cat(gathered).numpy())
    dist.destroy_process_group()


def test_two_rank_shards_match_single_process(tmp_path):
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    whole = _run_shard(0, world * G_LOCAL)
    for rank in range(world):
        got = np.load(tmp_path / f"gather_{rank}.npy")
        assert np.array_equal(got, whole), f"rank {rank}: gathered commitIndex differs from the single-process run"
    assert whole.max() > 0
