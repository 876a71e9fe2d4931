"""The JNI glue (rafting_b200/csrc/jni/rafting_jni.c) cannot be loaded into a JVM here (no JDK), but it must not rot: it is
type-checked with gcc against the real C-ABI headers and a stand-in jni.h, every native method INTEGRATION.md's NativeEngine
declares must have its Java_..._NativeEngine_<name> definition in the glue, and the natives are EXECUTED against a stand-in
JNIEnv (tests/jni_stub/fake_env.c): frame scan and the journal for real, the device natives down their error paths."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GLUE = os.path.join(ROOT, "rafting_b200", "csrc", "jni", "rafting_jni.c")


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no gcc")
def test_glue_type_checks_against_the_c_abi():
    res = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "jni_stub"),
                          "-I", os.path.join(ROOT, "include"), GLUE], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]


def test_every_native_method_of_the_integration_guide_is_defined():
    guide = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    declared = set(re.findall(r"static native [\w\[\]]+ (\w+)\(", guide))
    defined = set(re.findall(r"FN\((\w+)\)", open(GLUE).read()))
    assert len(declared) >= 30 and declared <= defined, sorted(declared - defined)


# ---- the natives EXECUTED against a stand-in JNIEnv (tests/jni_stub/fake_env.c) -------------------------------------------
# The glue is compiled with the stand-in jni.h and linked with the in-tree libraries; a fake function table plays the JVM
# (direct buffers = {addr, cap}, strings = char*, long[] = {n, p}, exceptions recorded).  Host-only natives run for real;
# the device natives are driven as far as their error path (no GPU on this box -> the status must surface as the exception
# class INTEGRATION.md promises).
import ctypes as C  # noqa: E402

import numpy as np  # noqa: E402

from tests import jni_exec  # noqa: E402
from tests.jni_exec import Buf as _Buf, Longs as _Longs, buf_of_array as _buf, fn as _fn, PFX  # noqa: E402


@pytest.fixture(scope="module")
def glue(tmp_path_factory):
    L = jni_exec.build(str(tmp_path_factory.mktemp("jni")))
    if L is None:
        pytest.skip("no gcc")
    return L


def test_frame_scan_native_cuts_a_receive_buffer_like_the_c_entry_point(glue):
    from rafting_b200 import ingest
    stream = ingest.encode(ingest.ENQ, b"appendEntries:ctx-7", b"\x01\x02\x03", sequence=41) + \
        ingest.encode(ingest.ACK, b"appendEntries:ctx-7", b"", sequence=41) + \
        ingest.encode(ingest.SYN, b"node-1")[:9]                                     # a frame cut off mid-way stays unconsumed
    data = np.frombuffer(stream, dtype=np.uint8).copy()
    frames = np.zeros(8, dtype=ingest.FRAME)
    meta = np.zeros(3, dtype=np.int64)
    glue.fake_reset()
    n = _fn(glue, "frameScan", C.c_int32, C.POINTER(_Buf), C.c_int64, C.POINTER(_Buf), C.c_int32, C.POINTER(_Buf))(
        C.byref(_buf(data)), len(data), C.byref(_buf(frames)), 8, C.byref(_buf(meta)))
    rc, want, used, tr = ingest.scan(stream, cap=8)
    assert n == 2 == len(want) and meta.tolist() == [used, int(tr), rc] and used == len(stream) - 9
    assert frames[:2].tobytes() == want.tobytes() and frames["sequence"][0] == 41 and glue.fake_throws() == 0


def test_ack_decode_natives_turn_reply_frames_into_records(glue):
    from rafting_b200 import abi, ingest
    m = _fn(glue, "ctxmapCreate", C.c_int64)()
    put = _fn(glue, "ctxmapPut", None, C.c_int64, C.c_char_p, C.c_int32)
    glue.fake_reset()
    put(m, b"ctx-3", 3), put(m, b"ctx-4", 4)
    stream = ingest.encode(ingest.ACK, b"appendEntries:ctx-3", ingest.reply_body_encode(17, True), sequence=5) + \
        ingest.encode(ingest.ENQ, b"appendEntries:ctx-3", b"opaque", sequence=6) + \
        ingest.encode(ingest.ACK, b"requestVote:ctx-4", ingest.reply_body_encode(18, False), sequence=7)
    data = np.frombuffer(stream, dtype=np.uint8).copy()
    rc, frames, used, _ = ingest.scan(stream)
    frames = frames.copy()
    recs = np.zeros(len(frames), dtype=ingest.ACK_REC)
    n = _fn(glue, "acksDecode", C.c_int32, C.POINTER(_Buf), C.POINTER(_Buf), C.c_int32, C.c_int64, C.POINTER(_Buf))(
        C.byref(_buf(data)), C.byref(_buf(frames)), len(frames), m, C.byref(_buf(recs)))
    assert n == 2 and glue.fake_throws() == 0 and glue.fake_balance() == 0
    assert [(int(r["gid"]), int(r["kind"]), int(r["sequence"]), int(r["term"]), int(r["success"]), int(r["frame"])) for r in recs[:2]] == \
        [(3, abi.EV_AE_ACK, 5, 17, 1, 0), (4, abi.EV_RV_REPLY, 7, 18, 0, 2)]
    _fn(glue, "ctxmapDestroy", None, C.c_int64)(m)


def test_journal_natives_persist_and_restore_through_the_glue(glue, tmp_path):
    from rafting_b200 import durable
    jopen = _fn(glue, "journalOpen", C.c_int64, C.c_char_p, C.c_int32)
    commit = _fn(glue, "journalCommitStep", C.c_int64, C.c_int64, C.POINTER(_Buf), C.c_int32, C.c_uint8, C.POINTER(_Buf), C.POINTER(_Buf))
    glue.fake_reset()
    j = jopen(str(tmp_path / "wal").encode(), 8)
    assert j != 0 and glue.fake_throws() == 0 and glue.fake_balance() == 0          # GetStringUTFChars was released
    rw = np.zeros(8, dtype=np.uint32)
    rw[1] = 0 | ((2 + 1) << 8) | (1 << 30)                                            # follower, votedFor slot 2, persist-dirty
    rw[5] = 2 | ((0 + 1) << 8) | (1 << 30)
    term = np.arange(8, dtype=np.int64) * 10
    assert commit(j, None, 8, 0, C.byref(_buf(rw)), C.byref(_buf(term))) == 2
    _fn(glue, "journalMilestone", None, C.c_int64, C.c_int32, C.c_int64, C.c_int64)(j, 5, 1234, 50)
    _fn(glue, "journalCheckpoint", None, C.c_int64)(j)
    st = np.zeros(1, dtype=np.dtype([("term", "<i8"), ("ballot", "<i4"), ("_pad", "<i4"), ("mi", "<i8"), ("mt", "<i8")]))
    restore = _fn(glue, "journalRestore", None, C.c_int64, C.c_int32, C.POINTER(_Buf))
    restore(j, 5, C.byref(_buf(st)))
    assert (st["term"][0], st["ballot"][0], st["mi"][0], st["mt"][0]) == (50, 0, 1234, 50)
    _fn(glue, "journalClose", None, C.c_int64)(j)
    assert glue.fake_throws() == 0
    # a second process (the plain binding) finds what the natives made durable
    j2 = durable.Journal(str(tmp_path / "wal"), 8)
    assert (j2.restore(1).term, j2.restore(1).ballot) == (10, 2) and j2.restore(5).milestone_index == 1234
    j2.close()
    # failure path: the journal directory cannot be created -> java/io/IOException with the library's message
    glue.fake_reset()
    blocker = tmp_path / "file"
    blocker.write_text("x")
    assert jopen(str(blocker / "sub").encode(), 8) == 0
    assert glue.fake_throws() == 1 and glue.fake_thrown_class() == b"java/io/IOException" and glue.fake_thrown_message()
    assert glue.fake_balance() == 0


def test_status_codes_surface_as_the_promised_exception_classes(glue):
    import torch
    from rafting_b200 import abi
    glue.fake_reset()
    # wrap(): a DirectByteBuffer over native memory
    raw = np.zeros(64, np.uint8)
    glue_wrap = getattr(glue, PFX + "wrap")
    glue_wrap.restype = C.POINTER(_Buf)
    glue_wrap.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
    b = glue_wrap(glue.env, None, raw.ctypes.data, 64)
    assert b.contents.addr == raw.ctypes.data and b.contents.cap == 64
    glue.fake_free_buf(b)
    # commInitAll over more shards than one process can hold: RAFTING_E_CAPACITY -> RaftException; the long[] is released
    hs = (C.c_int64 * 65)()
    longs = _Longs(65, C.cast(hs, C.POINTER(C.c_int64)))
    _fn(glue, "commInitAll", None, C.POINTER(_Longs))(C.byref(longs))
    assert glue.fake_throws() == 1 and glue.fake_thrown_class() == b"io/lubricant/consensus/raft/support/RaftException"
    assert glue.fake_balance() == 0
    # an invalid configuration -> IllegalArgumentException (checked before any device is touched)
    glue.fake_reset()
    cfg = abi.make_cfg(replicas=3, max_groups=16, max_rows=2)
    cfg.replicas = 99
    cbuf = _Buf(C.addressof(cfg), C.sizeof(cfg))
    h = _fn(glue, "create", C.c_int64, C.POINTER(_Buf))(C.byref(cbuf))
    assert h == 0 and glue.fake_throws() == 1 and glue.fake_thrown_class() == b"java/lang/IllegalArgumentException"
    if not torch.cuda.is_available():
        # a valid configuration on a box without a GPU: the CUDA status comes back as RaftException, never as a crash
        glue.fake_reset()
        cfg = abi.make_cfg(replicas=3, max_groups=16, max_rows=2)
        cbuf = _Buf(C.addressof(cfg), C.sizeof(cfg))
        h = _fn(glue, "create", C.c_int64, C.POINTER(_Buf))(C.byref(cbuf))
        assert h == 0 and glue.fake_throws() == 1
        assert glue.fake_thrown_class() == b"io/lubricant/consensus/raft/support/RaftException" and glue.fake_thrown_message()


def test_dispatch_natives_turn_an_outbox_into_request_and_reply_records(glue):
    """outboxToRequests / requestToInbox / outboxToReplies through the glue, against the plain C-ABI binding."""
    from rafting_b200 import abi, ingest
    G, F, rows = 4, 2, 2
    ob = abi.Outbox(rows, G, F, G)
    ob.incarnation[:] = [5, 6, 7, 8]
    ob.current_term[:] = [50, 60, 70, 80]
    ob.plan_meta[1, 2, 0] = abi.PLAN_AE | (3 << 16) | (7 << 32)
    ob.plan_pp[1, 2, 0] = (100, 69); ob.plan_lc[1, 2, 0] = (103, 99); ob.plan_epoch[1, 2, 0] = 40
    ob.ballot_meta[0, 1] = abi.BALLOT_VOTE | (6 << 32); ob.ballot_term[0, 1] = 61; ob.ballot_last[0, 1] = (12, 59)
    glue.fake_reset()
    d = _fn(glue, "dispatchCreate", C.c_int64, C.c_int32, C.c_int32, C.c_int32)(G, F, 1)
    recs = np.zeros(16, dtype=ingest.REQ_REC)
    oc = ob.as_c()
    n = _fn(glue, "outboxToRequests", C.c_int32, C.c_int64, C.POINTER(_Buf), C.c_int32, C.POINTER(_Buf), C.c_int32)(
        d, C.byref(jni_exec.buf_of_struct(oc)), rows, C.byref(_buf(recs)), 16)
    want, unknown = ingest.Dispatch(G, F, 1).requests(ob, rows)
    assert n == 3 == len(want) and unknown == 0 and recs[:3].tobytes() == want.tobytes() and glue.fake_throws() == 0
    assert [(int(q["kind"]), int(q["dst_slot"]), int(q["term"])) for q in recs[:3]] == \
        [(abi.OP_VOTE_REQ, 0, 61), (abi.OP_VOTE_REQ, 2, 61), (abi.OP_AE_REQUEST, 0, 70)]
    # the AE record lands in a receiver's inbox (node slot 0), its three entry terms appended to the pool
    ib = abi.Inbox(rows, G, F, ent_cap=64)
    ic = ib.as_c()
    terms = np.array([69, 70, 70], dtype=np.int64)
    cnt = _fn(glue, "requestToInbox", C.c_int32, C.POINTER(_Buf), C.POINTER(_Buf), C.c_int32, C.c_int64, C.c_uint8, C.POINTER(_Buf),
              C.c_int32, C.c_int32, C.c_int32)(C.byref(_buf(recs[2:3])), C.byref(_buf(terms)), 1, 123456, 0,
                                               C.byref(jni_exec.buf_of_struct(ic)), G, 64, 0)
    ref = abi.Inbox(rows, G, F, ent_cap=64)
    ref.ae_request(1, 2, 123456, 1, 70, 100, 69, [69, 70, 70], 99)
    assert cnt == 3 and glue.fake_throws() == 0
    for col in ("op_meta", "op_nr", "op_ab", "op_cd", "op_e"):
        assert np.array_equal(getattr(ib, col), getattr(ref, col)), col
    assert np.array_equal(ib.ent_terms[:3], ref.ent_terms[:3])
    # ... the receiver's step answered it: one reply record for the sender's lane
    rob = abi.Outbox(rows, G, F, G)
    rob.rep_meta[1, 2] = 1 | 2; rob.rep_term[1, 2] = 70
    reps = np.zeros(4, dtype=ingest.BATCH_REC)
    roc = rob.as_c()
    prow = np.array([1], dtype=np.uint8)
    m = _fn(glue, "outboxToReplies", C.c_int32, C.POINTER(_Buf), C.c_int32, C.c_int32, C.POINTER(_Buf), C.POINTER(_Buf), C.c_int32,
            C.POINTER(_Buf))(C.byref(jni_exec.buf_of_struct(roc)), G, 0, C.byref(_buf(recs[2:3])), C.byref(_buf(prow)), 1, C.byref(_buf(reps)))
    assert m == 1 and glue.fake_throws() == 0
    r = reps[0]
    assert (int(r["gid"]), int(r["kind"]), int(r["lane"]), int(r["flags"]), int(r["incarnation"]), int(r["term"]), int(r["epoch_at_send"]),
            int(r["last_at_send"])) == (2, abi.EV_AE_ACK, 0, abi.OUT_OK | 4, 7, 70, 40, 103)
    # commit records -> apply ranges
    cob = abi.Outbox(rows, G, F, G)
    cob.commit_index[:] = [10, 20, 30, 40]
    cob.role_word[:] = [1 << 31, 0, 1 << 31, 1 << 31]
    applied = np.array([4, 0, 30, 35], dtype=np.int64)
    ranges = np.zeros(4, dtype=ingest.APPLY_REC)
    coc = cob.as_c()
    k = _fn(glue, "applyRanges", C.c_int32, C.POINTER(_Buf), C.POINTER(_Buf), C.c_int32, C.POINTER(_Buf), C.c_int32, C.POINTER(_Buf), C.c_int32)(
        C.byref(jni_exec.buf_of_struct(coc)), None, G, C.byref(_buf(applied)), G, C.byref(_buf(ranges)), 4)
    assert k == 2 and [(int(a["gid"]), int(a["first"]), int(a["last"])) for a in ranges[:2]] == [(0, 5, 10), (3, 36, 40)]
    assert applied.tolist() == [10, 0, 30, 40] and glue.fake_throws() == 0
    # capacity error surfaces as RaftException
    glue.fake_reset()
    _fn(glue, "outboxToRequests", C.c_int32, C.c_int64, C.POINTER(_Buf), C.c_int32, C.POINTER(_Buf), C.c_int32)(
        d, C.byref(jni_exec.buf_of_struct(oc)), rows, C.byref(_buf(recs)), 1)
    assert glue.fake_throws() == 1 and glue.fake_thrown_class() == b"io/lubricant/consensus/raft/support/RaftException"
    _fn(glue, "dispatchDestroy", None, C.c_int64)(d)


# ---- the Java declarations against the glue's C signatures, native by native ---------------------------------------------------
JAVA = os.path.join(ROOT, "rafting_b200", "java", "io", "lubricant", "consensus", "raft", "gpu", "NativeEngine.java")
_JTYPE = {"long": "jlong", "int": "jint", "boolean": "jboolean", "void": "void", "ByteBuffer": "jobject", "String": "jstring",
          "long[]": "jlongArray"}


def _java_natives(text):
    out = {}
    for ret, name, params in re.findall(r"static native ([\w\[\]]+)\s+(\w+)\(([^)]*)\)", text, re.S):
        types = [" ".join(p.split()[:-1]) for p in (q.strip() for q in params.split(",")) if p]
        out[name] = (_JTYPE[ret], [_JTYPE[t] for t in types])
    return out


def _glue_natives(text):
    out = {}
    for ret, name, params in re.findall(r"JNIEXPORT (\w+) JNICALL FN\((\w+)\)\(([^)]*)\)", text, re.S):
        ps = [" ".join(p.split()[:-1]) for p in (q.strip() for q in params.split(","))]
        assert ps[:2] == ["JNIEnv*", "jclass"], (name, ps[:2])
        out[name] = (ret, ps[2:])
    return out


def test_java_declarations_match_the_glue_signatures():
    """A JNI mismatch (a missing argument, an int where the glue takes a jlong) is undefined behaviour that no compiler sees:
    NativeEngine.java and rafting_jni.c are compared native by native — same set, same return type, same argument types in
    the same order — and the listing in INTEGRATION.md must be the file's."""
    java = _java_natives(open(JAVA).read())
    glue = _glue_natives(open(GLUE).read())
    assert set(java) == set(glue), (sorted(set(java) - set(glue)), sorted(set(glue) - set(java)))
    for name in sorted(java):
        assert java[name] == glue[name], f"{name}: Java {java[name]} vs glue {glue[name]}"
    assert len(java) >= 50
    guide = _java_natives(open(os.path.join(ROOT, "INTEGRATION.md")).read())
    assert guide == java
    assert "package io.lubricant.consensus.raft.gpu;" in open(JAVA).read()


def test_builder_natives_fill_an_inbox_like_the_c_entry_points(glue):
    from rafting_b200 import abi, ingest
    G, F, rows = 4, 2, 4
    glue.fake_reset()
    b = _fn(glue, "builderCreate", C.c_int64, C.c_int32, C.c_int32)(G, F)
    ref = ingest.Builder(G, F)
    req = np.zeros(1, dtype=ingest.REQ_REC)
    req["gid"], req["kind"], req["src_slot"], req["term"], req["a"], req["b"], req["commit"], req["count"] = 2, abi.OP_AE_REQUEST, 1, 9, 40, 8, 39, 2
    terms = np.array([9, 9], dtype=np.int64)
    rep = np.zeros(1, dtype=ingest.BATCH_REC)
    rep["gid"], rep["kind"], rep["lane"], rep["flags"], rep["incarnation"], rep["term"], rep["epoch_at_send"], rep["last_at_send"] = 2, abi.EV_AE_ACK, 1, 4, 3, 9, 0, 41
    _fn(glue, "builderPushRequest", None, C.c_int64, C.POINTER(_Buf), C.POINTER(_Buf))(b, C.byref(_buf(req)), C.byref(_buf(terms)))
    _fn(glue, "builderPushReply", None, C.c_int64, C.POINTER(_Buf))(b, C.byref(_buf(rep)))
    _fn(glue, "builderPushSubmit", None, C.c_int64, C.c_int32, C.c_int32, C.c_int32)(b, 2, 3, 0)
    ref.push_request(req[0], terms), ref.push_reply(rep[0]), ref.push_submit(2, 3)
    got, want = (abi.Inbox(rows, G, F, ent_cap=16, sweep=True) for _ in range(2))
    gc = got.as_c()
    placed, prow, counters = np.zeros(8, dtype=ingest.REQ_REC), np.zeros(8, dtype=np.uint8), np.zeros(2, dtype=np.int64)
    n = _fn(glue, "builderBuild", C.c_int32, C.c_int64, C.c_int64, C.POINTER(_Buf), C.c_int32, C.POINTER(_Buf), C.POINTER(_Buf), C.c_int32,
            C.POINTER(_Buf))(b, 5000, C.byref(jni_exec.buf_of_struct(gc)), 16, C.byref(_buf(placed)), C.byref(_buf(prow)), 8, C.byref(_buf(counters)))
    wp, wr = ref.build(5000, want)
    assert n == 2 == len(wp) and counters.tolist() == [2, 0] and glue.fake_throws() == 0
    assert placed[:2].tobytes() == wp.tobytes() and prow[:2].tolist() == wr.tolist() == [1, 2]     # the submit leads, then the request
    for col in ("row_now", "op_meta", "op_nr", "op_ab", "op_cd", "op_e", "ev_meta", "ev_tn", "ev_el"):
        assert np.array_equal(getattr(got, col), getattr(want, col)), col
    assert got.ev_meta[2, 2, 1] != 0 and np.array_equal(got.ent_terms[:2], [9, 9])                  # the reply follows the request's row
    _fn(glue, "builderDestroy", None, C.c_int64)(b)
