"""ctypes binding of the durability journal (include/rafting_durable.h): one fdatasync per engine step for the
(term, votedFor) records of every group whose outbox role_word carries the persist-dirty bit."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _build

_LIB = None


class Stable(C.Structure):
    _fields_ = [("term", C.c_int64), ("ballot", C.c_int32), ("_pad", C.c_int32),
                ("milestone_index", C.c_int64), ("milestone_term", C.c_int64)]


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(_build.build_durable())
        L.rafting_journal_open.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_void_p)]
        L.rafting_journal_close.argtypes = [C.c_void_p]
        L.rafting_journal_commit_step.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p,
                                                  C.POINTER(C.c_uint64)]
        L.rafting_journal_milestone.argtypes = [C.c_void_p, C.c_uint32, C.c_int64, C.c_int64]
        L.rafting_journal_restore.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(Stable)]
        L.rafting_journal_checkpoint.argtypes = [C.c_void_p]
        L.rafting_journal_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64 * 4)]
        L.rafting_stable_image.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32,
                                           C.POINTER(C.c_uint32)]
        L.rafting_durable_last_error.restype = C.c_char_p
        _LIB = L
    return _LIB


def _check(rc, what):
    if rc:
        raise RuntimeError(f"{what}: rc={rc} {lib().rafting_durable_last_error().decode()}")


class Journal:
    def __init__(self, directory: str, max_groups: int):
        self._h = C.c_void_p()
        _check(lib().rafting_journal_open(directory.encode(), max_groups, C.byref(self._h)), "rafting_journal_open")

    def close(self):
        if self._h:
            lib().rafting_journal_close(self._h)
            self._h = C.c_void_p()

    def commit_step(self, role_word: np.ndarray, current_term: np.ndarray, gids: np.ndarray | None = None,
                    compact: bool = False) -> int:
        """Persist-before-reply barrier of one step; returns the number of records made durable."""
        role_word = np.ascontiguousarray(role_word, dtype=np.uint32)
        current_term = np.ascontiguousarray(current_term, dtype=np.int64)
        n = len(role_word) if gids is None else len(gids)
        g = None if gids is None else np.ascontiguousarray(gids, dtype=np.uint32)
        cnt = C.c_uint64()
        _check(lib().rafting_journal_commit_step(self._h, None if g is None else g.ctypes.data, n, 1 if compact else 0,
                                                 role_word.ctypes.data, current_term.ctypes.data, C.byref(cnt)),
               "rafting_journal_commit_step")
        return cnt.value

    def milestone(self, gid: int, index: int, term: int):
        _check(lib().rafting_journal_milestone(self._h, gid, index, term), "rafting_journal_milestone")

    def restore(self, gid: int) -> Stable:
        st = Stable()
        _check(lib().rafting_journal_restore(self._h, gid, C.byref(st)), "rafting_journal_restore")
        return st

    def checkpoint(self):
        _check(lib().rafting_journal_checkpoint(self._h), "rafting_journal_checkpoint")

    def stats(self) -> dict:
        a = (C.c_uint64 * 4)()
        _check(lib().rafting_journal_stats(self._h, C.byref(a)), "rafting_journal_stats")
        return dict(batches=a[0], records=a[1], syncs=a[2], journal_bytes=a[3])

    def image(self, gid: int, id_bytes: bytes = b"") -> bytes:
        buf = C.create_string_buffer(28 + len(id_bytes))
        n = C.c_uint32()
        _check(lib().rafting_stable_image(self._h, gid, id_bytes, len(id_bytes), buf, len(buf), C.byref(n)), "rafting_stable_image")
        return buf.raw[:n.value]
