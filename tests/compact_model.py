"""numpy restatement of the engine's unpack / pack kernels (rafting_b200/csrc/compact.cuh) — TEST INFRASTRUCTURE: the CPU
suite uses it to check that the compact wire format is lossless (codec of rafting_b200/compact.py + the format rules of
include/rafting_b200.h) without a GPU; under -m gpu the real kernels are checked against the oracle (tests/test_compact_gpu.py)."""
from __future__ import annotations

import numpy as np

from rafting_b200 import abi, compact

CTAGS = 32
U64 = np.uint64


class InFlight:
    """The HBM in-flight table: per (group, follower) 32 tags -> (epochAtSend, lastIndexAtSend, incarnation) + a tag bitmap."""

    def __init__(self, G, F):
        self.el = np.zeros((CTAGS, G, F), dtype=abi.I64X2)
        self.inc = np.zeros((CTAGS, G, F), dtype=np.uint32)
        self.bits = np.zeros((G, F), dtype=np.uint32)


def pack(dense: abi.Outbox, epoch: np.ndarray, st: InFlight, esc_cap: int) -> compact.CompactOutbox:
    """pack_kernel: dense outbox -> compact outbox; tags handed out per lane in row order, lowest free first."""
    rows, n, F = dense.rows, dense.n, dense.F
    co = compact.CompactOutbox(rows, n, F, esc_cap=esc_cap)
    for name in ("commit_index", "current_term", "role_word", "incarnation", "err_word", "last_entry"):
        getattr(co, name)[:] = getattr(dense, name)[:n]
    co.epoch[:] = epoch
    esc = []
    for g in range(n):
        commit_end, term_end, inc_end = int(dense.commit_index[g]), int(dense.current_term[g]), int(dense.incarnation[g])
        last_end, ep = int(dense.last_entry[g]["x"]), (int(epoch[g]["x"]), int(epoch[g]["y"]))
        for f in range(F):
            b = int(st.bits[g, f])
            for r in range(rows):
                pm = int(dense.plan_meta[r, g, f])
                if pm == 0:
                    continue
                kind, hb, count, inc = pm & 0xF, (pm >> 4) & 1, (pm >> 16) & 0xFFFF, pm >> 32
                pp, lc, pe = dense.plan_pp[r, g, f], dense.plan_lc[r, g, f], int(dense.plan_epoch[r, g, f])
                tag = abi.CTAG_NONE
                if kind in (abi.PLAN_AE, abi.PLAN_IS):
                    free = ~b & 0xFFFFFFFF
                    if free:
                        tag = (free & -free).bit_length() - 1
                        b |= 1 << tag
                        st.el[tag, g, f] = (pe, int(lc["x"])); st.inc[tag, g, f] = inc
                dcommit = (commit_end - int(lc["y"])) & 0xFFFFFFFFFFFFFFFF
                dprev = 0
                fits = pe == ep[0] and inc == inc_end and count < 64
                if kind == abi.PLAN_AE:
                    dprev = (last_end - int(pp["x"])) & 0xFFFFFFFFFFFFFFFF
                    fits = fits and int(pp["y"]) == term_end and int(lc["x"]) == int(pp["x"]) + count and dprev < 256 and dcommit < 128
                elif kind == abi.PLAN_IS:
                    fits = fits and (int(pp["x"]), int(pp["y"])) == ep and int(lc["x"]) == ep[0] and dcommit < 128
                else:
                    fits = fits and int(pp["x"]) == 0 and int(pp["y"]) == 0 and int(lc["x"]) == 0 and int(lc["y"]) == 0
                pc = kind | (hb << 3) | (tag << 5)
                if fits:
                    pc |= (count << 11) | (dprev << 17) | ((dcommit << 25) if kind in (abi.PLAN_AE, abi.PLAN_IS) else 0)
                else:
                    pc |= 1 << 4
                    esc.append((abi.CESC_PLAN, (r * n + g) * F + f, pm | ((255 if tag == abi.CTAG_NONE else tag) << 8),
                                int(pp["x"]), int(pp["y"]), int(lc["x"]), int(lc["y"]), pe))
                co.plan_c[r, g, f] = pc
            st.bits[g, f] = b
        for r in range(rows):
            rm = int(dense.rep_meta[r, g])
            co.rep_c[r, g] = (rm >> 8) & 0xFF
            if rm & 1:
                co.counts[2] += 1
                esc.append((abi.CESC_REPLY, r * n + g, rm, int(dense.rep_term[r, g]), 0, 0, 0, 0))
            bm = int(dense.ballot_meta[r, g])
            if bm:
                co.counts[1] += 1
                bl = dense.ballot_last[r, g]
                esc.append((abi.CESC_BALLOT, r * n + g, bm, int(dense.ballot_term[r, g]), int(bl["x"]), int(bl["y"]), 0, 0))
    co.counts[0] = len(esc)
    for k, rec in enumerate(esc[:esc_cap]):
        co.esc[k] = rec
    return co


def unpack(ci: compact.CompactInbox, g_term: np.ndarray, st: InFlight) -> abi.Inbox:
    """unpack_kernel + unpack_escapes_kernel: compact inbox -> the dense inbox the step kernel reads."""
    rows, n, F = ci.rows, ci.n, ci.F
    ib = abi.Inbox(rows, n, F)
    ib.flags = abi.INBOX_NO_REQUESTS
    ib.op_cd = None; ib.op_e = None
    if ci.ev_c is None:
        ib.ev_meta = ib.ev_tn = ib.ev_el = None
    else:
        for g in range(n):
            for f in range(F):
                b = int(st.bits[g, f])
                for r in range(rows):
                    w = int(ci.ev_c[r, g, f])
                    kind = w & 0xF
                    if kind in (abi.EV_AE_ACK, abi.EV_IS_ACK):
                        tag = (w >> 8) & 0xFF
                        el, inc = (0, 0), 0
                        if tag < CTAGS:
                            el, inc = (int(st.el[tag, g, f]["x"]), int(st.el[tag, g, f]["y"])), int(st.inc[tag, g, f])
                            b &= ~(1 << tag)
                        ib.ev_meta[r, g, f] = (w & 0x7F) | (inc << 32)
                        ib.ev_tn[r, g, f] = (int(g_term[g]) if (w >> 7) & 1 else 0, int(ci.row_base[r]) + (w >> 16))
                        ib.ev_el[r, g, f] = el
                st.bits[g, f] = b
        for e in ci.esc:
            r, rest = divmod(int(e["slot"]), n * F); g, f = divmod(rest, F)
            ib.ev_meta[r, g, f] = e["ev_meta"]; ib.ev_tn[r, g, f] = (e["term"], e["now_ms"]); ib.ev_el[r, g, f] = (e["epoch_at_send"], e["last_at_send"])
    if ci.op_c is None:
        ib.op_meta = ib.op_nr = ib.op_ab = None
    else:
        c = ci.op_c.astype(np.uint64)
        ib.op_meta[:] = (c & U64(0xF)) | (((c >> U64(4)) & U64(0xFFF)) << U64(16))
        ib.op_nr["x"] = ci.row_base[:, None] + (c >> U64(16)).astype(np.int64)
        if ci.op_unavail is not None:
            ib.op_ab["x"] = ci.op_unavail.astype(np.int64)
        else:
            ib.op_ab = None
    return ib
