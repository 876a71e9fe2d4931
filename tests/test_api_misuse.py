"""Error behaviour of the C ABI that needs no GPU: every entry point rejects a null engine with
RAFTING_E_INVAL instead of crashing, and bad configs are refused before any CUDA call."""
import ctypes as C

import pytest

from rafting_b200 import abi, engine


def test_null_engine_is_rejected_everywhere():
    L = engine.lib()
    lease, inbox, outbox, st = abi.LeaseC(), abi.InboxC(), abi.OutboxC(), abi.GroupState()
    gi = abi.GroupInit()
    calls = [
        lambda: L.rafting_group_open(None, 0, C.byref(gi)),
        lambda: L.rafting_group_open_bulk(None, 0, 1, None),
        lambda: L.rafting_group_close(None, 0),
        lambda: L.rafting_lease(None, 1, 0, 0, C.byref(lease)),
        lambda: L.rafting_step(None, C.byref(lease)),
        lambda: L.rafting_step_begin(None, C.byref(lease)),
        lambda: L.rafting_step_wait(None, C.byref(lease)),
        lambda: L.rafting_step_device(None, C.byref(inbox), C.byref(outbox), None),
        lambda: L.rafting_step_begin_host(None, 0, C.byref(inbox), C.byref(outbox)),
        lambda: L.rafting_step_wait_slot(None, 0),
        lambda: L.rafting_state_export(None, 0, C.byref(st)),
        lambda: L.rafting_checkpoint(None),
        lambda: L.rafting_restore(None),
        lambda: L.rafting_allgather_commit(None, None, None),
        lambda: L.rafting_allgather_join(None),
        lambda: L.rafting_log_config(None, 1 << 16, 4, 16),
        lambda: L.rafting_log_append(None, None, 1, None, 0),
        lambda: L.rafting_lease_release(None, C.byref(lease)),
        lambda: L.rafting_step_begin_compact(None, 0, C.byref(abi.CInboxC()), C.byref(abi.COutboxC())),
        lambda: L.rafting_step_wait_compact(None, 0),
        lambda: L.rafting_step_fetch_dense(None, 0, C.byref(outbox)),
        lambda: L.rafting_step_device_seq(None, None, None, 1, 0, None),
        lambda: L.rafting_allgather_commit_from(None, None, None, None),
        lambda: L.rafting_allgather_commit_all(None, 1, None, None, None),
        lambda: L.rafting_comm_init_all(None, 1),
        lambda: L.rafting_allgather_last(None, None),
        lambda: L.rafting_restore_async(None),
        lambda: L.rafting_state_save(None, b"/tmp/x"),
        lambda: L.rafting_state_load(None, b"/tmp/x"),
        lambda: L.rafting_log_store_open(None, b"/tmp/x", 0, None),
        lambda: L.rafting_log_sync(None),
        lambda: L.rafting_log_mark(None, 0, 1, 0, 0, 0),
        lambda: L.rafting_compact_layout(1, 1, 1, 0, 0, None, None),
    ]
    for k, call in enumerate(calls):
        assert call() == -1, f"call #{k} did not return RAFTING_E_INVAL"
    assert L.rafting_engine_destroy(None) == 0
    assert b"" != L.rafting_last_error()


@pytest.mark.parametrize("kw", [dict(replicas=1), dict(replicas=34), dict(replicas=3, local_slot=3), dict(max_groups=0)])
def test_bad_config_is_refused_before_cuda(kw):
    cfg = abi.make_cfg(**kw)
    with pytest.raises(engine.RaftingError) as ei:
        engine.Engine(cfg)
    assert ei.value.rc == -1

def test_struct_size_guard():
    cfg = abi.make_cfg()
    cfg.struct_size = 12
    with pytest.raises(engine.RaftingError) as ei:
        engine.Engine(cfg)
    assert ei.value.rc == -1
