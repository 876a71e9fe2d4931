"""Durability journal (include/rafting_durable.h): one fdatasync per step for the (term, votedFor)
records of the groups whose outbox carries the persist-dirty bit; crash recovery; the reference's StableLock layout."""
import os
import struct

import numpy as np
import pytest

from oracle import binding
from rafting_b200 import abi, durable
from tests.cluster_sim import Cluster, T0


def _role_word(role, ballot, dirty):
    return np.uint32(role | ((ballot + 1) << 8) | ((1 << 30) if dirty else 0))


def test_commit_restore_and_reference_image(tmp_path):
    j = durable.Journal(str(tmp_path / "j"), 8)
    rw = np.array([_role_word(0, -1, False), _role_word(0, 2, True), _role_word(2, 0, True), _role_word(1, 1, False)] + [0] * 4, dtype=np.uint32)
    term = np.array([9, 5, 7, 3, 0, 0, 0, 0], dtype=np.int64)
    assert j.commit_step(rw, term) == 2
    assert j.commit_step(np.zeros(8, np.uint32), term) == 0                    # nothing dirty: no record, no sync
    j.milestone(1, 40, 4)
    s = j.stats()
    assert s["batches"] == 2 and s["records"] == 3 and s["syncs"] == 2
    st = j.restore(1)
    assert (st.term, st.ballot, st.milestone_index, st.milestone_term) == (5, 2, 40, 4)
    assert j.restore(0).term == 0 and j.restore(0).ballot == -1               # not dirty: never written
    # the reference's file layout (StableLock.java:52-67,69-90): big-endian epoch, term, id length, id bytes
    img = j.image(1, b"node-2")
    assert img == struct.pack(">qqqi", 40, 4, 5, 6) + b"node-2"
    assert j.image(0, b"whatever") == struct.pack(">qqqi", 0, 0, 0, 0)         # null ballot: length 0, no bytes
    j.close()
    j = durable.Journal(str(tmp_path / "j"), 8)                                # clean reopen: journal replay
    st = j.restore(2)
    assert (st.term, st.ballot) == (7, 0) and j.restore(1).milestone_index == 40
    j.close()


def test_torn_tail_is_cut_and_checkpoint_folds(tmp_path):
    d = str(tmp_path / "j")
    j = durable.Journal(d, 4)
    term = np.array([1, 1, 1, 1], dtype=np.int64)
    for k in range(1, 6):
        rw = np.array([_role_word(0, k % 3, True), 0, 0, _role_word(0, -1, k == 5)], dtype=np.uint32)
        assert j.commit_step(rw, term * k) >= 1
    size = os.path.getsize(os.path.join(d, "stable.wal"))
    j.close()
    # crash in the middle of the last batch: everything before it survives, the torn batch does not
    with open(os.path.join(d, "stable.wal"), "r+b") as f:
        f.truncate(size - 10)
    j = durable.Journal(d, 4)
    assert j.restore(0).term == 4 and j.restore(0).ballot == 1 and j.restore(3).term == 0
    assert os.path.getsize(os.path.join(d, "stable.wal")) < size - 10          # the torn tail was cut off
    # a corrupted byte in the middle: replay stops at that batch
    assert j.commit_step(np.array([_role_word(0, 0, True), 0, 0, 0], dtype=np.uint32), term * 6) == 1
    j.checkpoint()
    assert os.path.getsize(os.path.join(d, "stable.wal")) == 0
    assert j.commit_step(np.array([0, _role_word(1, 1, True), 0, 0], dtype=np.uint32), term * 8) == 1
    j.close()
    with open(os.path.join(d, "stable.wal"), "r+b") as f:
        f.seek(30); b = f.read(1); f.seek(30); f.write(bytes([b[0] ^ 0xFF]))
    j = durable.Journal(d, 4)
    assert j.restore(0).term == 6 and j.restore(1).term == 0                   # table survived, the bad batch did not
    j.close()


def test_active_list_indexing(tmp_path):
    j = durable.Journal(str(tmp_path / "j"), 16)
    gids = np.array([3, 9, 12], dtype=np.uint32)
    wide_rw = np.zeros(16, np.uint32); wide_t = np.zeros(16, np.int64)
    wide_rw[9] = _role_word(0, 1, True); wide_t[9] = 77
    assert j.commit_step(wide_rw, wide_t, gids=gids) == 1 and j.restore(9).term == 77
    tight_rw = np.array([0, 0, _role_word(0, 0, True)], np.uint32); tight_t = np.array([0, 0, 88], np.int64)
    assert j.commit_step(tight_rw, tight_t, gids=gids, compact=True) == 1 and j.restore(12).term == 88
    with pytest.raises(RuntimeError):
        j.commit_step(tight_rw, tight_t, gids=np.array([3, 9, 99], dtype=np.uint32), compact=True)
    j.close()


def test_node_crash_and_restart_in_the_cluster(tmp_path):
    """The whole loop: every node journals its persist-dirty groups before its replies leave (Cluster.on_outbox); one
    node is then killed, loses its engine, and comes back from journal + payload store with rafting_group_open
    (StableLock.restore + RaftLog state, ContextManager.java:57-106).  Safety must survive: no second leader in a term
    the node already voted in, and the files converge."""
    G = 5
    journals = [durable.Journal(str(tmp_path / f"n{k}"), G) for k in range(3)]
    c = Cluster(lambda cfg: binding.Oracle(cfg), G=G, seed=31, drop_ppm=10_000)
    c.on_outbox = lambda nd, ob: journals[nd.slot].commit_step(ob.role_word, ob.current_term)
    c.run(150)
    victim = c.leader_of(0)
    before = [(int(c.nodes[victim].snap.current_term[g]), (int(c.nodes[victim].snap.role_word[g]) >> 8 & 0xFF) - 1) for g in range(G)]
    # crash: the engine state is gone, messages in flight to and from the node are lost
    c.cut = {victim}
    c.run(2)
    journals[victim].close()
    journals[victim] = durable.Journal(str(tmp_path / f"n{victim}"), G)
    for g in range(G):
        st = journals[victim].restore(g)
        assert (st.term, st.ballot) == before[g], "what was acknowledged before the crash is what comes back"
    c.restart(victim, lambda cfg: binding.Oracle(cfg), lambda g: journals[victim].restore(g))
    c.run(60)                                    # still cut off: it times out, asks for votes nobody hears
    c.cut = set()
    c.run(250)
    c.drop_ppm = 0
    c.run(80, submit=False)
    c.check(converged=True)
    assert sum(j.stats()["records"] for j in journals) > 3 * G
    assert all(j.stats()["syncs"] <= j.stats()["batches"] for j in journals)   # one barrier per step, never per group
    for j in journals:
        j.close()


def _fsize_limited_commit(d, limit):
    """Child process: commits batches against an RLIMIT_FSIZE so that a real write(2) fails in the middle of a batch."""
    import resource
    import signal
    signal.signal(signal.SIGXFSZ, signal.SIG_IGN)
    j = durable.Journal(d, 64)
    term = np.arange(64, dtype=np.int64) + 100
    rw = np.array([_role_word(0, 1, True)] * 64, dtype=np.uint32)
    assert j.commit_step(rw[:2], term[:2]) == 2                                # batch 1: 24 + 48 bytes
    resource.setrlimit(resource.RLIMIT_FSIZE, (limit, resource.getrlimit(resource.RLIMIT_FSIZE)[1]))
    failed = False
    try:
        j.commit_step(rw, term * 2)                                            # batch 2 would need 24 + 64*24 bytes: torn by EFBIG
    except RuntimeError as ex:
        failed = "rolled back" in str(ex)
    assert failed
    size_after_failure = os.path.getsize(os.path.join(d, "stable.wal"))
    assert size_after_failure == 72, size_after_failure                        # cut back to the pre-batch offset
    assert j.restore(5).term == 0                                              # nothing of the failed batch was applied
    # a smaller batch fits under the limit: it is appended right behind batch 1 with the NEXT sequence number
    assert j.commit_step(rw[:3], term[:3] * 3) == 3
    j.close()


def test_failed_batch_is_rolled_back_and_later_batches_survive_recovery(tmp_path):
    """a partial write used to leave torn bytes + a consumed sequence number, so batches committed after the
    failure were dropped by recovery (acknowledged votes lost).  Now the batch is cut back and the journal stays consistent."""
    import multiprocessing as mp
    d = str(tmp_path / "j")
    ctx = mp.get_context("fork")
    p = ctx.Process(target=_fsize_limited_commit, args=(d, 400))
    p.start(); p.join()
    assert p.exitcode == 0
    j = durable.Journal(d, 64)                                                 # recovery sees batch 1 and batch 3, both complete
    assert [j.restore(g).term for g in (0, 1, 2, 5)] == [300, 303, 306, 0]
    assert j.stats()["journal_bytes"] == 72 + 24 + 72
    j.close()
