"""HBM-resident batches for the device path (rafting_step_device): torch is used ONLY to own device
memory and streams; the columns are raw bytes whose addresses go into the C structs."""
from __future__ import annotations

import torch

from . import abi

_IN_COLS = (("op_meta", 8, False), ("op_nr", 16, False), ("op_ab", 16, False), ("op_cd", 16, False), ("op_e", 8, False),
            ("ev_meta", 8, True), ("ev_tn", 16, True), ("ev_el", 16, True))
_OUT_ROW = (("rep_meta", 4, False), ("rep_term", 8, False), ("plan_meta", 8, True), ("plan_pp", 16, True),
            ("plan_lc", 16, True), ("plan_epoch", 8, True), ("ballot_meta", 8, False), ("ballot_term", 8, False),
            ("ballot_last", 16, False))
_OUT_GRP = (("commit_index", 8), ("current_term", 8), ("role_word", 4), ("incarnation", 4), ("err_word", 4), ("last_entry", 16))


class DevInbox:
    def __init__(self, rows, n, F, device, requests=False, unavail=True):
        """requests=False omits op_cd/op_e (no inbound-request ops); unavail=False also omits op_ab (its only use in
        such a step is the unavailable-follower mask of SUBMIT / TIMEOUT: absent == nobody unavailable)."""
        self.rows, self.n, self.F = rows, n, F
        self.t = {}
        for name, sz, lane in _IN_COLS:
            if not requests and name in ("op_cd", "op_e"):
                continue
            if not unavail and name == "op_ab":
                continue
            cnt = rows * n * (F if lane else 1)
            self.t[name] = torch.zeros(cnt * sz, dtype=torch.uint8, device=device)
        self.flags = 0 if requests else abi.INBOX_NO_REQUESTS

    @classmethod
    def from_host(cls, ib: abi.Inbox, device) -> "DevInbox":
        """Device-resident copy of a dense host inbox (columns absent on the host stay absent)."""
        import numpy as np
        self = cls.__new__(cls)
        self.rows, self.n, self.F = ib.rows, ib.n, ib.F
        self.t = {}
        for name, _, _ in _IN_COLS:
            col = getattr(ib, name)
            if col is not None:
                self.t[name] = torch.from_numpy(np.ascontiguousarray(col).view(np.uint8).reshape(-1)).to(device)
        self.flags = ib.flags
        return self

    def nbytes(self):
        return sum(t.numel() for t in self.t.values())

    def as_c(self) -> abi.InboxC:
        c = abi.InboxC()
        c.rows, c.n_active = self.rows, 0
        for name, _, _ in _IN_COLS:
            setattr(c, name, self.t[name].data_ptr() if name in self.t else None)
        c.flags = self.flags
        return c


class DevOutbox:
    def __init__(self, rows, n, F, G, device):
        self.rows, self.n, self.F, self.G = rows, n, F, G
        self.t = {}
        for name, sz, lane in _OUT_ROW:
            self.t[name] = torch.zeros(rows * n * (F if lane else 1) * sz, dtype=torch.uint8, device=device)
        for name, sz in _OUT_GRP:
            self.t[name] = torch.zeros(G * sz, dtype=torch.uint8, device=device)

    def nbytes(self):
        return sum(t.numel() for t in self.t.values())

    def as_c(self) -> abi.OutboxC:
        c = abi.OutboxC()
        for name in self.t:
            setattr(c, name, self.t[name].data_ptr())
        return c
