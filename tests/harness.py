"""Shared drivers for the parity tests: one script drives the CPU oracle and the CUDA engine
through the identical batch interface and compares outboxes and exported state byte for byte."""
from __future__ import annotations

import ctypes as C

import numpy as np

from rafting_b200 import abi, workload
from rafting_b200.abi import (EV_AE_ACK, EV_IS_ACK, EV_PV_REPLY, EV_RV_REPLY, OUT_CANCELED, OUT_ERROR, OUT_OK,
                              ROLE_CANDIDATE, ROLE_FOLLOWER, ROLE_LEADER)

T0 = workload.T0_MS


def state_bytes(st: abi.GroupState, F: int) -> bytes:
    used = abi.GroupState.followers.offset + C.sizeof(abi.FollowerState) * F
    return bytes(st)[:used]


def assert_states_equal(a, b, gids, F, where=""):
    for gid in gids:
        sa, sb = a.export(gid), b.export(gid)
        if state_bytes(sa, F) != state_bytes(sb, F):
            da, db = sa.as_dict(), sb.as_dict()
            diff = {k: (da[k], db[k]) for k in da if da[k] != db[k]}
            raise AssertionError(f"group {gid} state differs {where}: {diff}")


def assert_outbox_equal(oa: abi.Outbox, ob: abi.Outbox, gids=None, where=""):
    bad = oa.equal(ob, gids)
    if bad:
        name = bad[0]
        a, b = getattr(oa, name), getattr(ob, name)
        if a.shape == b.shape:
            idx = np.argwhere(a != b)[:5].tolist()
        else:
            idx = "shape"
        raise AssertionError(f"outbox column(s) {bad} differ {where}; first diffs of {name} at {idx}")


def init_array(n, terms=None, **common) -> np.ndarray:
    a = np.zeros(n, dtype=abi.GROUP_INIT_DTYPE)
    a["ballot"] = -1
    a["first_index"] = 1
    a["now_ms"] = T0 - 2000
    a["rand_ms"] = 0
    for k, v in common.items():
        a[k] = v
    if terms is not None:
        a["term"] = terms
    return a


def elect_all(sut, w1: workload.WlCfg):
    """Follower -> PreVote round -> Candidate -> Leader for every group, unanimous grants."""
    out = None
    phases = (0, 1, 2) if sut.cfg.pre_vote else (0, 2)
    for ph in phases:
        ib = workload.election_inbox_host(w1, ph, out)
        out = sut.step(ib)
    return out


def run_leader_workload(suts, w: workload.WlCfg, steps: int, compare=True, first_step=0, prevs=None, drop_ab=False):
    """Drives every SUT in `suts` with the SAME stream, generated from the FIRST sut's outbox
    (after checking the others produced the identical outbox).  Returns the last outboxes."""
    prev = prevs
    for k in range(first_step, first_step + steps):
        ib = workload.leader_inbox_host(w, k, prev)
        if drop_ab:
            ib.op_ab = None          # nobody unavailable: the column is simply absent
        outs = [s.step(ib) for s in suts]
        if compare:
            for o in outs[1:]:
                assert_outbox_equal(outs[0], o, where=f"at step {k}")
        prev = outs[0]
    return prev


# ---------------------------------------------------------------------------------------------
# random fuzz over every op / event kind, steered by the reference SUT's exported state
# ---------------------------------------------------------------------------------------------
class Fuzzer:
    def __init__(self, cfg: abi.Cfg, ref, seed=1, rows=3, n=None):
        self.cfg, self.ref = cfg, ref
        self.rng = np.random.default_rng(seed)
        self.rows = rows
        self.n = n or cfg.max_groups
        self.F = cfg.replicas - 1
        self.now = T0
        self.pending = {}      # gid -> list of (lane, kind, inc, epoch, last) replies owed to the SUT

    def _peer(self):
        R = self.cfg.replicas
        s = int(self.rng.integers(0, R - 1))
        return s if s < self.cfg.local_slot else s + 1

    def make(self, last_out: abi.Outbox | None) -> abi.Inbox:
        rng, F = self.rng, self.F
        ib = abi.Inbox(self.rows, self.n, F, ent_cap=self.rows * self.n * 8)
        # harvest reply obligations from the last outbox
        if last_out is not None:
            for r in range(last_out.rows):
                for i in range(self.n):
                    for f in range(F):
                        pm = int(last_out.plan_meta[r, i, f])
                        k = pm & 0xF
                        if k in (abi.PLAN_AE, abi.PLAN_IS):
                            self.pending.setdefault(i, []).append(
                                (f, EV_IS_ACK if k == abi.PLAN_IS else EV_AE_ACK, pm >> 32,
                                 int(last_out.plan_epoch[r, i, f]), int(last_out.plan_lc[r, i, f]["x"])))
                    bm = int(last_out.ballot_meta[r, i])
                    if bm & 0xF:
                        kind = EV_PV_REPLY if (bm & 0xF) == abi.BALLOT_PREVOTE else EV_RV_REPLY
                        for f in range(F):
                            self.pending.setdefault(i, []).append((f, kind, bm >> 32, 0, 0))
        for r in range(self.rows):
            for i in range(self.n):
                self.now += 1
                st = self.ref.export(i)
                term, role = st.current_term, st.role
                u = rng.random()
                if u < 0.25:
                    pass
                elif u < 0.40:
                    ib.submit(r, i, self.now, count=int(rng.integers(1, 4)), unavail=int(rng.integers(0, 4)) if rng.random() < 0.1 else 0)
                elif u < 0.55:
                    ib.timeout(r, i, self.now, rand=int(rng.integers(900, 1801)) if rng.random() < 0.7 else 0,
                               unavail=int(rng.integers(0, 4)) if rng.random() < 0.1 else 0)
                elif u < 0.75:
                    # inbound AppendEntries from a plausible leader
                    t = term + int(rng.choice([-1, 0, 0, 0, 1, 2]))
                    if t < 0:
                        t = 0
                    has = st.last_index >= st.first_index
                    last = st.last_index if has else st.epoch_index
                    mode = rng.random()
                    if mode < 0.6:       # append right after our last entry
                        prev = last
                        prev_t = (st.last_term if has else st.epoch_term)
                    elif mode < 0.8 and last > st.epoch_index + 1:   # overlap: rewrite the tail
                        prev = int(rng.integers(st.epoch_index + 1, last))
                        prev_t = self.ref.log_term(i, prev)
                        if prev_t < 0:
                            prev, prev_t = st.epoch_index, st.epoch_term
                    elif mode < 0.9:     # mismatching prev
                        prev, prev_t = last + int(rng.integers(1, 4)), max(term, 1)
                    else:                # below / at the epoch
                        prev, prev_t = st.epoch_index, st.epoch_term
                    if prev == 0:
                        prev_t = 0
                    n_ent = int(rng.integers(0, 5))
                    base_t = max(prev_t, 1)
                    ent = []
                    for _ in range(n_ent):
                        if rng.random() < 0.2:
                            base_t += 1
                        ent.append(min(base_t, max(t, base_t)))
                    first = None if rng.random() < 0.95 else prev + 1 + int(rng.integers(1, 3))
                    peer = st.current_leader if (st.current_leader >= 0 and rng.random() < 0.9) else self._peer()
                    ib.ae_request(r, i, self.now, peer, t, prev, prev_t, ent, leader_commit=int(rng.integers(0, last + 3)),
                                  first_index=first, rand=int(rng.integers(900, 1801)))
                elif u < 0.85:
                    t = term + int(rng.choice([-1, 0, 1, 1, 2]))
                    li = st.last_index + int(rng.integers(-1, 2)) if st.last_index >= st.first_index else st.epoch_index + int(rng.integers(0, 2))
                    lt = max(0, st.last_term + int(rng.integers(-1, 2)))
                    fn = ib.prevote_request if rng.random() < 0.5 else ib.vote_request
                    fn(r, i, self.now, self._peer(), max(t, 0), max(li, 0), lt, rand=int(rng.integers(900, 1801)))
                elif u < 0.90:
                    t = term + int(rng.choice([-1, 0, 0, 1]))
                    ib.is_request(r, i, self.now, self._peer(), max(t, 0), st.epoch_index + 5, st.epoch_term, bool(rng.integers(0, 2)),
                                  rand=int(rng.integers(900, 1801)))
                else:
                    has = st.last_index >= st.first_index
                    hi = st.last_index if has else st.epoch_index
                    idx = int(rng.integers(max(st.epoch_index - 1, 0), hi + 2))
                    t = self.ref.log_term(i, idx)
                    ib.flush(r, i, self.now, idx, t if t >= 0 else max(st.epoch_term, 1))
                # lane events: answer some owed replies (in order per lane), sometimes forge
                owed = self.pending.get(i, [])
                used = set()
                keep = []
                for (f, kind, inc, ep, la) in owed:
                    if f in used or rng.random() < 0.35:
                        keep.append((f, kind, inc, ep, la))
                        continue
                    used.add(f)
                    v = rng.random()
                    outcome = OUT_OK if v < 0.9 else (OUT_ERROR if v < 0.95 else OUT_CANCELED)
                    rt = term if rng.random() < 0.93 else term + int(rng.integers(1, 4))
                    succ = rng.random() < 0.85
                    self.now += 1
                    if kind in (EV_AE_ACK, EV_IS_ACK):
                        ib.ack(r, i, f, self.now, inc, rt, succ, ep, la, outcome=outcome, snapshot=kind == EV_IS_ACK)
                    else:
                        ib.vote_reply(r, i, f, self.now, inc, rt, succ, outcome=outcome, pre=kind == EV_PV_REPLY)
                if len(keep) > 4 * F:
                    keep = keep[-4 * F:]
                self.pending[i] = keep
        return ib
