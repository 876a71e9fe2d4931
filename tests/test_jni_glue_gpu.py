"""The device natives of the JNI glue, executed on a GPU (`-m gpu`) against the stand-in JNIEnv of tests/jni_stub/: what
GpuContextManager's pump does through NativeEngine — create, groupOpenBulk, stepBeginHost /
stepWaitSlot with caller-owned buffers, stateExport, checkpoint / restore, destroy — must give the oracle's outboxes and
states bit for bit, and a misuse must come back as the promised exception, not as a crash."""
import ctypes as C

import numpy as np
import pytest

from oracle import binding
from rafting_b200 import abi, workload
from tests import harness, jni_exec
from tests.jni_exec import Buf, buf_of_array, buf_of_struct, fn

pytestmark = pytest.mark.gpu


class GlueEngine:
    """abi.Inbox in, abi.Outbox out — every call goes through a Java_..._NativeEngine_* native."""

    def __init__(self, L, cfg, G, F):
        self.L, self.cfg, self.G, self.F = L, cfg, G, F
        L.fake_reset()
        self.h = fn(L, "create", C.c_int64, C.POINTER(Buf))(C.byref(buf_of_struct(cfg)))
        assert self.h != 0 and L.fake_throws() == 0, L.fake_thrown_message()
        self._begin = fn(L, "stepBeginHost", None, C.c_int64, C.c_int32, C.POINTER(Buf), C.POINTER(Buf))
        self._wait = fn(L, "stepWaitSlot", None, C.c_int64, C.c_int32)

    def open_bulk(self, first, init):
        fn(self.L, "groupOpenBulk", None, C.c_int64, C.c_int32, C.c_int32, C.POINTER(Buf))(self.h, first, len(init), C.byref(buf_of_array(init)))
        assert self.L.fake_throws() == 0, self.L.fake_thrown_message()

    def step(self, ib, slot=0):
        out = abi.Outbox(ib.rows, self.G, self.F, self.G)
        ic, oc = ib.as_c(), out.as_c()
        self._begin(self.h, slot, C.byref(buf_of_struct(ic)), C.byref(buf_of_struct(oc)))
        self._wait(self.h, slot)
        assert self.L.fake_throws() == 0, self.L.fake_thrown_message()
        return out

    def export(self, gid):
        st = abi.GroupState()
        fn(self.L, "stateExport", None, C.c_int64, C.c_int32, C.POINTER(Buf))(self.h, gid, C.byref(buf_of_struct(st)))
        assert self.L.fake_throws() == 0, self.L.fake_thrown_message()
        return st

    def call(self, name):
        fn(self.L, name, None, C.c_int64)(self.h)


def test_pump_cycle_through_the_natives_matches_the_oracle(tmp_path):
    L = jni_exec.build(str(tmp_path))
    if L is None:
        pytest.skip("no gcc")
    G, R, rows, steps = 1024, 3, 4, 6
    cfg = abi.make_cfg(replicas=R, max_groups=G, max_rows=rows)
    o, e = binding.Oracle(cfg), GlueEngine(L, cfg, G, R - 1)
    init = harness.init_array(G, terms=np.arange(G) % 7)
    o.open_bulk(0, init), e.open_bulk(0, init)
    w1, w = workload.make_wl(11, 1, G, R - 1), workload.make_wl(11, rows, G, R - 1)
    harness.elect_all(o, w1), harness.elect_all(e, w1)
    prev = None
    for k in range(steps):
        if k == 3:
            e.call("checkpoint")
            mark, mark_prev = k, prev
        ib = workload.leader_inbox_host(w, k, prev)
        prev = o.step(ib)
        harness.assert_outbox_equal(prev, e.step(ib), where=f"step {k} through the natives")
    harness.assert_states_equal(o, e, range(0, G, 37), R - 1, where="after the stream")
    # roll the tables back to the checkpoint and replay its successor: the same outbox again
    e.call("restore")
    o2 = binding.Oracle(cfg)
    o2.open_bulk(0, init)
    harness.elect_all(o2, w1)
    p2 = None
    for k in range(mark + 1):
        ib = workload.leader_inbox_host(w, k, p2)
        p2 = o2.step(ib)
    harness.assert_outbox_equal(p2, e.step(workload.leader_inbox_host(w, mark, mark_prev)), where="replay after restore")
    # misuse: a slot that does not exist -> IllegalArgumentException, the engine stays usable
    L.fake_reset()
    e._wait(e.h, 99)
    assert L.fake_throws() == 1 and L.fake_thrown_class() == b"java/lang/IllegalArgumentException"
    L.fake_reset()
    assert e.export(5).current_term == o2.export(5).current_term
    fn(L, "destroy", None, C.c_int64)(e.h)
