"""ctypes binding of the transport framing library (include/rafting_ingest.h, rafting_b200/csrc/ingest.cpp): the reference's
EventCodec frame layout restated in C so that the pump thread can cut a receive buffer into frames and route them to group
ids.  Host only — no CUDA, no torch."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _build, abi

_LIB = None

ENQ, ACK, SYN, MW, PM, BATCH = 0x05, 0x06, 0x16, 0x95, 0x9E, 0x1A
MAX_HEAD, MAX_BODY = 128, 1 << 26

FRAME = np.dtype([("type", "u1"), ("has_sequence", "u1"), ("ending", "u1"), ("_pad", "u1"), ("sequence", "<i4"),
                  ("head_off", "<u4"), ("head_len", "<u4"), ("body_off", "<u4"), ("body_len", "<u4")])
BATCH_REC = np.dtype([("gid", "<u4"), ("kind", "u1"), ("lane", "u1"), ("flags", "u1"), ("row", "u1"), ("incarnation", "<u4"),
                      ("_pad", "<u4"), ("term", "<i8"), ("epoch_at_send", "<i8"), ("last_at_send", "<i8")])
assert FRAME.itemsize == 24 and BATCH_REC.itemsize == 40


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(_build.build_ingest())
        L.rafting_frame_scan.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32),
                                         C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
        L.rafting_frame_encode.restype = C.c_size_t
        L.rafting_frame_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_uint8, C.c_int, C.c_int32, C.c_char_p, C.c_uint32,
                                           C.c_char_p, C.c_uint32, C.c_int]
        L.rafting_scope_parse.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.rafting_ctxmap_create.argtypes = [C.POINTER(C.c_void_p)]
        L.rafting_ctxmap_destroy.argtypes = [C.c_void_p]
        L.rafting_ctxmap_put.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint32]
        L.rafting_ctxmap_get.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.rafting_batch_to_inbox.argtypes = [C.c_void_p, C.c_uint32, C.c_int64, C.POINTER(abi.InboxC), C.c_uint32, C.c_uint32,
                                             C.POINTER(C.c_uint32)]
        _LIB = L
    return _LIB


def encode(ftype: int, head: bytes, body: bytes = b"", sequence: int | None = None, ending: bool = False) -> bytes:
    buf = C.create_string_buffer(16 + len(head) + len(body))
    n = lib().rafting_frame_encode(buf, len(buf), ftype, 0 if sequence is None else 1, sequence or 0, head, len(head), body, len(body),
                                   1 if ending else 0)
    if n == 0:
        raise ValueError("frame does not fit / exceeds the codec limits")
    return buf.raw[:n]


def scan(data: bytes, cap: int = 1024):
    """-> (rc, frames (structured array), consumed bytes, transparent)"""
    out = np.zeros(cap, dtype=FRAME)
    n, used, tr = C.c_uint32(), C.c_size_t(), C.c_int()
    buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
    rc = lib().rafting_frame_scan(buf.ctypes.data, len(data), out.ctypes.data, cap, C.byref(n), C.byref(used), C.byref(tr))
    return rc, out[:n.value], used.value, bool(tr.value)


def scope_parse(head: bytes):
    k, off = C.c_uint32(), C.c_uint32()
    rc = lib().rafting_scope_parse(head, len(head), C.byref(k), C.byref(off))
    if rc:
        raise ValueError("unknown context " + head.decode(errors="replace"))
    return k.value, head[off.value:]


class CtxMap:
    def __init__(self):
        self._h = C.c_void_p()
        lib().rafting_ctxmap_create(C.byref(self._h))

    def put(self, ctx: bytes, gid: int):
        lib().rafting_ctxmap_put(self._h, ctx, len(ctx), gid)

    def get(self, ctx: bytes):
        g = C.c_uint32()
        return g.value if lib().rafting_ctxmap_get(self._h, ctx, len(ctx), C.byref(g)) == 0 else None

    def __del__(self):
        try:
            lib().rafting_ctxmap_destroy(self._h)
        except Exception:
            pass


def batch_to_inbox(recs: np.ndarray, now_ms: int, inbox: abi.Inbox) -> tuple[int, int]:
    recs = np.ascontiguousarray(recs, dtype=BATCH_REC)
    done = C.c_uint32()
    ic = inbox.as_c()
    rc = lib().rafting_batch_to_inbox(recs.ctypes.data, len(recs), now_ms, C.byref(ic), inbox.n, inbox.F, C.byref(done))
    return rc, done.value
