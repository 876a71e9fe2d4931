"""HBM segmented entry buffer (payload side of RaftLog) against the dict restatement of RocksLog
(oracle/payload_model.py): batched appends with overwrites, reads bounded by the group's stored key range,
truncation through a conflicting AppendEntries, and enough data to wrap the HBM segment ring so that old
entries come back from the pinned-host cold tier — byte for byte."""
import numpy as np
import pytest

from oracle.payload_model import PayloadLog
from rafting_b200 import abi
from tests import harness

pytestmark = pytest.mark.gpu


def _payload(rng, gid, index, term):
    n = int(rng.integers(0, 200))
    return bytes(rng.integers(0, 256, size=n, dtype=np.uint8)) + f"|{gid}:{index}:{term}".encode()


def test_append_read_gather_truncate_and_spill():
    from rafting_b200 import engine
    G, R = 64, 3
    cfg = abi.make_cfg(replicas=R, max_groups=G, max_rows=1, entry_pool_cap=64)
    e = engine.Engine(cfg)
    init = harness.init_array(G, terms=1)
    init["last_index"] = 300                     # stored keys 1..300, one term run
    init["last_term"] = 1
    e.open_bulk(0, init)
    e.log_config(segment_bytes=8192, hbm_segments=4, ring_slots=16)     # tiny arena: 32 KiB, forces spills
    rng = np.random.default_rng(5)
    models = [PayloadLog() for _ in range(G)]
    # batched appends, several rounds, with overwrites of the tail (a re-sent AppendEntries suffix)
    for rnd in range(6):
        batch = []
        for gid in range(G):
            lo = rnd * 40 + 1
            for index in range(max(1, lo - 5), lo + 40):
                p = _payload(rng, gid, index, 1)
                batch.append((gid, index, 1, p))
                models[gid].put(index, 1, p)
        rng.shuffle(batch)                       # order inside a batch is irrelevant except for duplicates
        seen = {}
        for k, (gid, index, _, p) in enumerate(batch):
            seen[(gid, index)] = p
        for (gid, index), p in seen.items():
            models[gid].put(index, 1, p)
        e.log_append(batch)
    st = e.log_stats()
    assert st["spilled_bytes"] > 0 and st["appended"] == G * (40 + 5 * 45)
    # point/batch reads, everywhere in the log (old indexes live in the cold tier by now)
    for gid in (0, 7, 63):
        for first, n in ((1, 10), (37, 50), (230, 20), (236, 80)):
            assert e.log_read(gid, first, n) == models[gid].batch(first, n), (gid, first, n)
    assert e.log_stats()["cold_hits"] > 0
    # gather = the AE plans of a step, many groups at once; beyond the stored keys -> not stored
    ranges = [(gid, int(rng.integers(1, 230)), int(rng.integers(1, 12))) for gid in range(G)] + [(3, 238, 10)]
    got = e.log_gather(ranges)
    want = []
    for gid, first, cnt in ranges:
        for i in range(first, first + cnt):
            want.append((gid, i, models[gid].kv[i][0], models[gid].kv[i][1]) if i in models[gid].kv else (gid, i, 0, None))
    assert got == want
    assert e.log_stats()["hbm_hits"] > 0
    # a conflicting AppendEntries truncates the follower's log: the payload of the dropped suffix disappears
    ib = abi.Inbox(1, G, R - 1, ent_cap=8)
    ib.ae_request(0, 5, harness.T0, 1, 2, 100, 1, [2, 2], leader_commit=0)      # entries 101,102 at term 2: conflict at 101
    e.step(ib)
    s = e.export(5)
    assert (s.last_index, s.last_term) == (102, 2)
    models[5].truncate(101)
    p101, p102 = b"new-101", b"new-102"
    e.log_append([(5, 101, 2, p101), (5, 102, 2, p102)])
    models[5].put(101, 2, p101), models[5].put(102, 2, p102)
    assert e.log_read(5, 95, 20) == models[5].batch(95, 20)
    assert e.log_read(5, 103, 5) == []


def test_trim_after_compaction_frees_index_and_cold_tier():
    """RaftLog.flush moves the epoch (RocksLog.java:228-242); rafting_log_trim then drops the index entries below every
    group's lowest stored key and frees cold segments without a live record; later segments that hold only dead
    records are never spilled at all.  Reads of the surviving range stay byte-exact."""
    from rafting_b200 import engine
    G, R = 16, 3
    cfg = abi.make_cfg(replicas=R, max_groups=G, max_rows=1, entry_pool_cap=8)
    e = engine.Engine(cfg)
    init = harness.init_array(G, terms=1)
    init["last_index"] = 400
    init["last_term"] = 1
    e.open_bulk(0, init)
    e.log_config(segment_bytes=8192, hbm_segments=4, ring_slots=16)
    rng = np.random.default_rng(9)
    models = [PayloadLog() for _ in range(G)]
    for lo in range(1, 201, 40):
        batch = []
        for gid in range(G):
            for index in range(lo, lo + 40):
                p = _payload(rng, gid, index, 1)
                batch.append((gid, index, 1, p)); models[gid].put(index, 1, p)
        e.log_append(batch)
    before = e.log_stats()
    assert before["spilled_bytes"] > 0 and before["indexed"] == G * 200
    # compaction: every group flushes to index 150 (the entry AT the index survives: deleteRange is end-exclusive)
    ib = abi.Inbox(1, G, R - 1)
    for gid in range(G):
        ib.flush(0, gid, harness.T0, 150, 1)
    e.step(ib)
    for m in models:
        m.flush(0, 150)
    assert e.export(3).epoch_index == 150
    dropped, freed = e.log_trim()
    assert dropped == G * 149 and freed > 0
    st = e.log_stats()
    assert st["indexed"] == G * 51 and st["trimmed"] == dropped and st["cold_freed_bytes"] == freed
    for gid in (0, 9, 15):
        assert e.log_read(gid, 150, 60) == models[gid].batch(150, 60)
        assert e.log_read(gid, 149, 5) == []
    got = e.log_gather([(2, 150, 51), (7, 190, 11)])
    assert [(g, i, t, p) for g, i, t, p in got] == [(2, i, 1, models[2].kv[i][1]) for i in range(150, 201)] + \
        [(7, i, 1, models[7].kv[i][1]) for i in range(190, 201)]
    # keep appending in rounds smaller than the arena, compacting after each: by the time a segment is recycled every
    # record in it has been trimmed -> it is not spilled at all
    for lo in range(201, 401, 10):
        batch = []
        for gid in range(G):
            for index in range(lo, lo + 10):
                p = _payload(rng, gid, index, 1)
                batch.append((gid, index, 1, p)); models[gid].put(index, 1, p)
        e.log_append(batch)
        ib = abi.Inbox(1, G, R - 1)
        for gid in range(G):
            ib.flush(0, gid, harness.T0, lo + 9, 1)
        e.step(ib)
        for m in models:
            m.flush(0, lo + 9)
        e.log_trim()
    st2 = e.log_stats()
    assert st2["spills_skipped"] > 0
    for gid in (1, 14):
        assert e.log_read(gid, 400, 20) == models[gid].batch(400, 20) and len(models[gid].batch(400, 20)) == 1
    assert e.log_read(1, 300, 5) == [] and e.log_stats()["indexed"] == G


def test_entry_file_survives_a_kill_and_bounds_the_pinned_pool(tmp_path):
    """the crash-durable tier.  Appends are framed into the entry file before they reach HBM, one
    rafting_log_sync per step is the durability barrier, truncations are logged as range marks.  The process is "killed"
    (the engine is dropped without any shutdown call, the file loses a torn tail), a new engine replays the file, the
    groups are re-opened from what rafting_log_recovered reports, and every payload is back byte for byte — served from
    the file tier.  The pinned cold pool stays within its bound while the file holds the evicted segments."""
    import struct
    from rafting_b200 import engine
    G, R = 8, 3
    path = str(tmp_path / "entries.wal")
    cfg = abi.make_cfg(replicas=R, max_groups=G, max_rows=1, entry_pool_cap=8)
    rng = np.random.default_rng(11)
    models = [PayloadLog() for _ in range(G)]

    def boot():
        e = engine.Engine(cfg)
        e.log_config(segment_bytes=8192, hbm_segments=4, ring_slots=16)
        return e

    e = boot()
    assert e.log_store_open(path, cold_max_segments=3) == 0
    init = harness.init_array(G, terms=1)
    init["last_index"] = 120; init["last_term"] = 2
    e.open_bulk(0, init)
    for lo in range(1, 121, 30):                                             # 4 steps of appends, one barrier each
        batch = []
        for gid in range(G):
            for index in range(lo, lo + 30):
                term = 1 if index <= 70 else 2                               # two term runs
                p = _payload(rng, gid, index, term)
                batch.append((gid, index, term, p)); models[gid].put(index, term, p)
        e.log_append(batch)
        e.log_sync()
    st = e.log_store_stats()
    assert st["syncs"] == 4 and st["synced_bytes"] == st["file_bytes"] > 0
    assert st["cold_resident"] <= 3 and st["cold_evicted"] > 0               # bounded pool: the rest lives in the file only
    for gid in (0, 5):
        assert e.log_read(gid, 1, 120) == models[gid].batch(1, 120)          # HBM + pinned + FILE tiers together
    assert e.log_store_stats()["file_hits"] > 0
    # group 3 truncates its suffix from 101 (a conflicting AppendEntries): logged as a range mark, then durable
    e.log_mark(3, 1, 100, 0, 0); models[3].truncate(101)
    e.log_sync()
    # an append that never got its barrier, then the kill: the tail of the file is torn
    e.log_append([(6, 121, 2, b"never acknowledged")])
    e.log_sync()
    size = __import__("os").path.getsize(path)
    key, val = e.log_export_kv(2, 71)                                         # the reference's RocksDB layout
    assert key == struct.pack(">q", 71) and val == struct.pack(">q", 2) + models[2].kv[71][1]
    del e                                                                     # no shutdown path: like kill -9
    with open(path, "r+b") as f:
        f.truncate(size - 7)
    # ---- restart ----
    e = boot()
    n = e.log_store_open(path, cold_max_segments=3)
    assert n == G * 120 + 2                                                   # every complete frame; the torn PUT is gone
    for gid in range(G):
        gi, runs = e.log_recovered(gid)
        last = 100 if gid == 3 else 120
        assert (gi.first_index, gi.last_index, gi.last_term) == (1, last, 2) and runs == [(1, 1), (71, 2)]
        e.open_group(gid, term=2, first_index=gi.first_index, last_index=gi.last_index, last_term=gi.last_term,
                     epoch_index=gi.epoch_index, epoch_term=gi.epoch_term, now_ms=harness.T0)
        e.load_runs(gid, runs)
    for gid in range(G):
        assert e.log_read(gid, 1, 130) == models[gid].batch(1, 130), gid
    assert e.log_term(3, 100) == 2 and e.log_term(3, 101) == -1 and e.log_term(0, 70) == 1
    # life goes on: new appends land in HBM and in the file, reads mix all tiers
    e.log_append([(0, 121, 2, b"after restart")]); models[0].put(121, 2, b"after restart")
    ib = abi.Inbox(1, G, R - 1, ent_cap=4)
    ib.ae_request(0, 0, harness.T0 + 5, 1, 2, 120, 2, [2], leader_commit=0)
    e.step(ib)
    e.log_sync()
    assert e.log_read(0, 119, 5) == models[0].batch(119, 5) and len(models[0].batch(119, 5)) == 3
