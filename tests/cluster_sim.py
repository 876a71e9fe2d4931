"""A closed-loop Raft cluster made of R independent engines (BASELINE config #1: the reference's 3-node
file-append cluster, `R/raft1-3.xml` + `T/cluster/TestNode1-3.java`, generalised to G groups).

Every node is one system under test (the CPU oracle or the CUDA engine) with its own `local_slot`; this module
plays what stays in Java around the engine: the Netty links (FIFO per link, loss, partitions),
the `Async` time-outs, the payload store behind `RaftLog` and the file-append `RaftMachine`
(`T/cluster/cmd/FileMachine.java`: apply = append the command to a file).  Nothing here knows Raft: it only turns
outbox records into the inbox events of the peer they are addressed to, exactly as the tables in
INTEGRATION.md §3/§4 say.  The reference's stated check for this configuration is "the three files are identical"
(`README.md:28-33`); `check()` asserts that plus the Raft safety properties that imply it.

Test infrastructure: used by tests/test_cluster_cpu.py (oracle) and tests/test_cluster_gpu.py (engine vs oracle)."""
from __future__ import annotations

from collections import defaultdict, deque

import numpy as np

from rafting_b200 import abi

T0 = 1_700_000_000_000
TICK_MS = 10
RPC_TIMEOUT_TICKS = 12          # Async time-out of an RPC whose request or reply was lost
ROWS = 4                        # row 0 is the timer-sweep row (its op slot is the implied TIMEOUT), ops use rows 1..3


def _splitmix(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


class Node:
    def __init__(self, sut, slot: int, G: int, R: int):
        self.sut, self.slot, self.G, self.R, self.F = sut, slot, G, R, R - 1
        self.queue = [deque() for _ in range(G)]          # per group: events in arrival order
        self.store = [dict() for _ in range(G)]           # payload store: index -> (term, payload)
        self.applied = [0] * G
        self.file = [[] for _ in range(G)]                # the FileMachine's file, one per group
        self.snap = None                                  # group columns of the previous step
        self.inc_term = [dict() for _ in range(G)]        # incarnation -> term of that role object
        self.snapshot = [None] * G                        # SnapshotArchive: (index, term, file prefix) of the last checkpoint
        self.install = [None] * G                         # pending snapshot installation (download + restore)

    def lane_of(self, slot: int) -> int:
        return slot if slot < self.slot else slot - 1

    def slot_of(self, lane: int) -> int:
        return lane if lane < self.slot else lane + 1


class Cluster:
    def __init__(self, make_sut, G: int = 8, R: int = 3, seed: int = 1, drop_ppm: int = 0, submit_ppm: int = 300_000,
                 heartbeat_ms: int = 50, election_ms: int = 300, compact_every: int = 0, pre_vote: bool = True,
                 guard_candidate_votes: bool = False, cfg_flags: int = 0, shadow_native: bool = False):
        # shadow_native: every step, the C dispatch (rafting_outbox_to_requests) and the C placement (rafting_request_to_inbox)
        # of include/rafting_ingest.h run beside this file's Python pump and must produce the same records / op columns
        self.shadow_native = shadow_native
        # with shadow_native and neither compaction nor the vote guard (both are host-side decisions made while placing), the C
        # inbox builder (rafting_builder_*) receives every queue item this file queues and must build the identical inbox
        self.shadow_builder = shadow_native and not compact_every and not guard_candidate_votes
        # see _vote_request_is_unsafe: the reference's Candidate grants votes without the log check
        self.guard_candidate_votes = guard_candidate_votes
        self.compact_every = compact_every       # RaftRoutine.compactLog: checkpoint + RaftLog.flush every N applied entries
        self.G, self.R, self.seed, self.drop_ppm, self.submit_ppm = G, R, seed, drop_ppm, submit_ppm
        self.nodes = []
        self.cfgs = []
        self.on_outbox = None                    # hook(node, outbox): the pump's durability barrier, before any reply leaves
        for k in range(R):
            cfg = abi.make_cfg(replicas=R, local_slot=k, max_groups=G, max_rows=ROWS, entry_pool_cap=ROWS * G * 64,
                               heartbeat_ms=heartbeat_ms, election_ms=election_ms, timer_seed=0xC0FFEE + 7919 * k,
                               pre_vote=pre_vote, flags=cfg_flags)
            self.cfgs.append(cfg)
            sut = make_sut(cfg)
            init = np.zeros(G, dtype=abi.GROUP_INIT_DTYPE)
            init["ballot"] = -1; init["first_index"] = 1; init["now_ms"] = T0
            sut.open_bulk(0, init)
            self.nodes.append(Node(sut, k, G, R))
            if shadow_native:
                from rafting_b200 import ingest
                self.nodes[-1].dispatch = ingest.Dispatch(G, R - 1, k)
                self.nodes[-1].builder = ingest.Builder(G, R - 1) if self.shadow_builder else None
        self.tick = 0
        self.inflight = []                       # (deliver_tick, seq, dst_slot, gid, item)
        self.link_clock = defaultdict(int)       # FIFO per (src, dst)
        self.seq = 0
        self.cut = set()                         # isolated node slots (partition)
        self.leaders_by_term = [dict() for _ in range(G)]
        self.submitted = 0
        self.errors = []                         # per-event error codes seen (besides NotLeader / NotReady)
        self.counts = defaultdict(int)

    # ---- the network ---------------------------------------------------------------------------
    def _rand(self, *key) -> int:
        h = self.seed
        for k in key:
            h = _splitmix(h ^ (int(k) & 0xFFFFFFFFFFFFFFFF))
        return h

    def _lost(self, src, dst, *key) -> bool:
        if src in self.cut or dst in self.cut:
            return True
        return self._rand(0xD409, src, dst, *key) % 1_000_000 < self.drop_ppm

    def _send(self, src, dst, gid, item, extra_delay=0):
        t = max(self.tick + 1 + extra_delay, self.link_clock[(src, dst)])
        self.link_clock[(src, dst)] = t
        self.seq += 1
        self.inflight.append((t, self.seq, dst, gid, item))

    def _timeout_event(self, node, gid, lane, kind, inc, epoch=0, last=0):
        """The Async of a lost RPC completes with an error after its time-out (Async.java:239-254)."""
        self.seq += 1
        self.inflight.append((self.tick + RPC_TIMEOUT_TICKS, self.seq, node.slot, gid,
                              ("ev", lane, dict(kind=kind, outcome=abi.OUT_ERROR, success=False, inc=inc, term=0,
                                                epoch=epoch, last=last))))

    # ---- one tick ------------------------------------------------------------------------------
    def step(self, submit: bool = True):
        now = T0 + TICK_MS * (self.tick + 1)
        due = [m for m in self.inflight if m[0] <= self.tick]
        self.inflight = [m for m in self.inflight if m[0] > self.tick]
        for _, _, dst, gid, item in sorted(due, key=lambda m: (m[0], m[1])):
            self.nodes[dst].queue[gid].append(item)
            if self.shadow_native and self.shadow_builder:
                self._push_native(self.nodes[dst], gid, item)
        outs = []
        for nd in self.nodes:
            outs.append(self._step_node(nd, now, submit))
        for nd, (ob, placed) in zip(self.nodes, outs):
            if self.on_outbox:
                self.on_outbox(nd, ob)             # persist-before-reply (RaftMember.java:20-26)
            self._dispatch(nd, ob, placed, now)
        self.tick += 1

    def _step_node(self, nd: Node, now: int, submit: bool):
        G, F = self.G, nd.F
        ib = abi.Inbox(ROWS, G, F, ent_cap=ROWS * G * 64, sweep=True)
        ib.row_now[0] = now                                            # fires every due election / keepAlive timer
        placed = {}
        ib2 = None
        nd.placed_native = []
        if self.shadow_native:
            ib2 = abi.Inbox(ROWS, G, F, ent_cap=ROWS * G * 64, sweep=True)
            ib2.row_now[0] = now
        for g in range(G):
            q = nd.queue[g]
            # a client command goes to the node that believes it leads the group (RaftStub.submit)
            if submit and nd.snap is not None and (int(nd.snap.role_word[g]) & 3) == abi.ROLE_LEADER \
                    and self._rand(0x5B, nd.slot, g, self.tick) % 1_000_000 < self.submit_ppm:
                q.appendleft(("op", dict(kind=abi.OP_SUBMIT, count=1 + self._rand(0x5C, nd.slot, g, self.tick) % 3)))
                if self.shadow_native and self.shadow_builder:
                    nd.builder.push_submit(g, q[0][1]["count"])
            if self.compact_every and nd.snap is not None and nd.install[g] is None:
                done = nd.snapshot[g][0] if nd.snapshot[g] else 0
                if nd.applied[g] - done >= self.compact_every:
                    idx = nd.applied[g]; term = nd.store[g][idx][0]
                    nd.snapshot[g] = (idx, term, list(nd.file[g][:idx]))           # RaftMachine.checkpoint
                    q.appendleft(("op", dict(kind=abi.OP_FLUSH, index=idx, term=term)))
                    self.counts["compactions"] += 1
            cursor = 0                                                 # position r*(F+1) + (0 | 1+lane); (0,op) is the sweep
            while q:
                it = q[0]
                if it[0] == "op" and it[1]["kind"] == abi.OP_IS_REQUEST and "result" not in it[1]:
                    q.popleft()
                    q.extendleft(reversed(self._install_snapshot(nd, g, it[1])))
                    continue
                if it[0] == "op":
                    r = cursor // (F + 1) + 1
                    pos = r * (F + 1)
                else:
                    lane = it[1]
                    r = cursor // (F + 1)
                    pos = r * (F + 1) + 1 + lane
                    if pos <= cursor:
                        r += 1; pos += F + 1
                if r >= ROWS:
                    break                                              # stays queued for the next step
                q.popleft()
                if it[0] == "op" and self.guard_candidate_votes and self._vote_request_is_unsafe(nd, g, it[1]):
                    src = self.nodes[it[1]["src"]]                     # dropped on the wire: the asker sees a time-out
                    self._timeout_event(src, g, src.lane_of(nd.slot),
                                        abi.EV_PV_REPLY if it[1]["kind"] == abi.OP_PREVOTE_REQ else abi.EV_RV_REPLY, it[1]["inc"])
                    self.counts["unsafe_vote_requests_dropped"] += 1
                    continue
                cursor = pos
                if it[0] == "op":
                    self._place_op(nd, ib, r, g, now, it[1])
                    placed[(r, g)] = it[1]
                    if ib2 is not None:
                        self._place_op_native(nd, ib2, r, g, now, it[1])
                else:
                    e = it[2]
                    if e["kind"] in (abi.EV_AE_ACK, abi.EV_IS_ACK):
                        ib.ack(r, g, lane, now, e["inc"], e["term"], e["success"], e["epoch"], e["last"],
                               outcome=e["outcome"], snapshot=e["kind"] == abi.EV_IS_ACK)
                    else:
                        ib.vote_reply(r, g, lane, now, e["inc"], e["term"], e["success"], outcome=e["outcome"],
                                      pre=e["kind"] == abi.EV_PV_REPLY)
        if self.shadow_native and self.shadow_builder:
            ib3 = abi.Inbox(ROWS, G, F, ent_cap=ROWS * G * 64, sweep=True)
            placed3, rows3 = nd.builder.build(now, ib3)
            for col in ("row_now", "op_meta", "op_nr", "op_ab", "op_cd", "op_e", "ev_meta", "ev_tn", "ev_el"):
                assert np.array_equal(getattr(ib, col), getattr(ib3, col)), f"native builder differs in {col} (node {nd.slot}, tick {self.tick})"
            assert ib.ent_count == ib3.ent_count and np.array_equal(ib.ent_terms[:ib.ent_count], ib3.ent_terms[:ib3.ent_count])
            assert len(nd.builder) == sum(len(q) for q in nd.queue), "the two sets of queues hold different leftovers"
            assert sorted((int(r), int(q["gid"]), int(q["kind"])) for q, r in zip(placed3, rows3)) == \
                sorted((r, g, op["kind"]) for (r, g), op in placed.items())
            self.counts["native_builds_checked"] += 1
            self.counts["native_build_items"] += int((ib.op_meta != 0).sum()) + int(((ib.ev_meta & np.uint64(0xF)) != 0).sum())
        if ib2 is not None:
            for col in ("op_meta", "op_nr", "op_ab", "op_cd", "op_e"):
                assert np.array_equal(getattr(ib, col), getattr(ib2, col)), f"native placement differs in {col} (node {nd.slot}, tick {self.tick})"
            assert ib.ent_count == ib2.ent_count and np.array_equal(ib.ent_terms[:ib.ent_count], ib2.ent_terms[:ib2.ent_count])
            self.counts["native_placements_checked"] += len(placed)
        ob = nd.sut.step(ib)
        return ob, placed

    def _push_native(self, nd, gid, item):
        """A queue item of this file -> the C builder's queue: requests as rafting_req_rec_t (+ entry terms), lane events as
        rafting_batch_rec_t."""
        from rafting_b200 import ingest
        if item[0] == "op":
            op = item[1]
            k = op["kind"]
            assert k in (abi.OP_AE_REQUEST, abi.OP_PREVOTE_REQ, abi.OP_VOTE_REQ), k
            rec = np.zeros(1, dtype=ingest.REQ_REC)[0]
            rec["gid"], rec["kind"], rec["src_slot"], rec["dst_slot"], rec["incarnation"], rec["term"] = gid, k, op["src"], nd.slot, op["inc"], op["term"]
            terms = []
            if k == abi.OP_AE_REQUEST:
                terms = [t for t, _ in op["entries"]]
                rec["a"], rec["b"], rec["commit"], rec["count"] = op["prev_index"], op["prev_term"], op["leader_commit"], len(terms)
                rec["epoch"], rec["last"] = op.get("epoch", 0), op.get("last", 0)
            else:
                rec["a"], rec["b"] = op["last_index"], op["last_term"]
            nd.builder.push_request(rec, terms)
        else:
            _, lane, e = item
            rec = np.zeros(1, dtype=ingest.BATCH_REC)[0]
            rec["gid"], rec["kind"], rec["lane"], rec["flags"] = gid, e["kind"], lane, e["outcome"] | (4 if e["success"] else 0)
            rec["incarnation"], rec["term"], rec["epoch_at_send"], rec["last_at_send"] = e["inc"], e["term"], e["epoch"], e["last"]
            nd.builder.push_reply(rec)

    def _place_op_native(self, nd, ib2, r, g, now, op):
        """The same op through rafting_request_to_inbox (requests) — submits / flushes are not requests and stay in Python."""
        from rafting_b200 import ingest
        k = op["kind"]
        if k in (abi.OP_SUBMIT, abi.OP_FLUSH):
            return self._place_op(nd, ib2, r, g, now, op)
        rec = np.zeros(1, dtype=ingest.REQ_REC)[0]
        rec["gid"], rec["kind"], rec["src_slot"], rec["dst_slot"], rec["incarnation"], rec["term"] = g, k, op["src"], nd.slot, op["inc"], op["term"]
        terms = []
        if k == abi.OP_AE_REQUEST:
            terms = [t for t, _ in op["entries"]]
            rec["a"], rec["b"], rec["commit"], rec["count"] = op["prev_index"], op["prev_term"], op["leader_commit"], len(terms)
        elif k == abi.OP_IS_REQUEST:
            rec["a"], rec["b"] = op["index"], op["index_term"]
        else:
            rec["a"], rec["b"] = op["last_index"], op["last_term"]
        rec["epoch"], rec["last"] = op.get("epoch", 0), op.get("last", 0)
        rc = ingest.request_to_inbox(rec, terms, r, now, bool(op.get("result", False)), ib2)
        assert rc == 0, (rc, op)
        nd.placed_native.append((r, rec.copy()))

    @staticmethod
    def _vote_request_is_unsafe(nd, g, op):
        """UPSTREAM FLAW, mirrored faithfully by oracle and engine: `Candidate.requestVote` (Candidate.java:49-72, and
        `Candidate.preVote`, which delegates to it, :44-46) answers a request of a higher term with
        `switchTo(Follower, term, candidateId)` — it votes for the asker WITHOUT `logUpToDate`, which only
        `Follower.requestVote` checks (Follower.java:108-127).  A node that is itself mid-election can thereby elect a peer
        whose log misses committed entries (leader completeness breaks, the state machines diverge; reproduced by
        tests/test_cluster_cpu.py::test_upstream_candidate_votes_without_log_check).  With this guard the SIMULATED NETWORK
        drops exactly those requests — asker's log behind the receiver's, receiver not a leader — to show that nothing
        else stands between the restated protocol and Raft's safety properties."""
        if op["kind"] not in (abi.OP_VOTE_REQ, abi.OP_PREVOTE_REQ) or nd.snap is None:
            return False
        if (int(nd.snap.role_word[g]) & 3) == abi.ROLE_LEADER:
            return False
        mi, mt = int(nd.snap.last_entry[g]["x"]), int(nd.snap.last_entry[g]["y"])
        return not (op["last_term"] > mt or (op["last_term"] == mt and op["last_index"] >= mi))

    def _place_op(self, nd, ib, r, g, now, op):
        k = op["kind"]
        if k == abi.OP_SUBMIT:
            ib.submit(r, g, now, count=op["count"])
        elif k == abi.OP_AE_REQUEST:
            ib.ae_request(r, g, now, op["src"], op["term"], op["prev_index"], op["prev_term"],
                          [t for t, _ in op["entries"]], op["leader_commit"])
        elif k == abi.OP_PREVOTE_REQ:
            ib.prevote_request(r, g, now, op["src"], op["term"], op["last_index"], op["last_term"])
        elif k == abi.OP_VOTE_REQ:
            ib.vote_request(r, g, now, op["src"], op["term"], op["last_index"], op["last_term"])
        elif k == abi.OP_IS_REQUEST:
            ib.is_request(r, g, now, op["src"], op["term"], op["index"], op["index_term"], op["result"])
        elif k == abi.OP_FLUSH:
            ib.flush(r, g, now, op["index"], op["term"])
        else:
            raise AssertionError(k)

    def _install_snapshot(self, nd, g, op):
        """Host side of RaftContext.installSnapshot (RaftRoutine.java:408-445): the first request starts the download
        and answers false; once the machine has been restored from the snapshot the next request corrects the log
        epoch (accomplishInstallation -> RaftLog.flush(milestone), :451-475) and answers true."""
        ins = nd.install[g]
        if ins is None and nd.applied[g] >= op["index"]:
            return [("op", dict(op, result=True))]              # the machine is already past that snapshot: nothing to install
        if ins is None:
            snap = self.nodes[op["src"]].snapshot[g]
            nd.install[g] = dict(ready=self.tick + 3, snap=snap)
            return [("op", dict(op, result=False))]
        if self.tick < ins["ready"]:
            return [("op", dict(op, result=False))]
        idx, term, prefix = ins["snap"]
        nd.install[g] = None
        self.counts["snapshots_installed"] += 1
        ok = idx > op["index"] or (idx == op["index"] and term >= op["index_term"])
        return [("op", dict(kind=abi.OP_FLUSH, index=idx, term=term, restore=prefix)), ("op", dict(op, result=ok))]

    # ---- outbox -> messages --------------------------------------------------
    def _dispatch(self, nd: Node, ob: abi.Outbox, placed, now):
        G, F = self.G, nd.F
        prev = nd.snap
        # role objects and their terms: a role object's term is fixed for its lifetime (RaftMember.java:16-26)
        for g in range(G):
            nd.inc_term[g][int(ob.incarnation[g])] = int(ob.current_term[g])
            if (int(ob.role_word[g]) & 3) == abi.ROLE_LEADER:
                t = int(ob.current_term[g])
                who = self.leaders_by_term[g].setdefault(t, nd.slot)
                assert who == nd.slot, f"two leaders in term {t} of group {g}: {who} and {nd.slot}"   # election safety
        # row by row, in the serial order of the step: a plan carries the entries the log held when it was made
        self._sent = [] if self.shadow_native else None              # what this outbox asks the pump to send, in order
        self._replied = [] if self.shadow_native else None           # the replies this outbox makes the pump send back
        for row in range(ROWS):
            self._dispatch_row(nd, ob, {k: v for k, v in placed.items() if k[0] == row}, row, prev)
        if self.shadow_native:
            recs, unknown = nd.dispatch.requests(ob, ROWS)
            got = [(int(q["row"]), int(q["gid"]), int(q["kind"]), int(q["src_slot"]), int(q["dst_slot"]), int(q["incarnation"]),
                    int(q["count"]), int(q["term"]), int(q["a"]), int(q["b"]), int(q["commit"]), int(q["epoch"]), int(q["last"]))
                   for q in recs]
            assert unknown == 0 and got == self._sent, f"native dispatch differs (node {nd.slot}, tick {self.tick}): " \
                f"{[x for x in got if x not in self._sent][:3]} vs {[x for x in self._sent if x not in got][:3]}"
            self.counts["native_requests_checked"] += len(got)
            from rafting_b200 import ingest
            order = sorted(range(len(nd.placed_native)), key=lambda i: (nd.placed_native[i][0], int(nd.placed_native[i][1]["gid"])))
            reps = ingest.outbox_to_replies(ob, G, nd.slot, np.array([nd.placed_native[i][1] for i in order], dtype=ingest.REQ_REC),
                                            [nd.placed_native[i][0] for i in order]) if order else []
            got = [(int(q["gid"]), int(q["kind"]), int(q["lane"]), int(q["flags"]), int(q["incarnation"]), int(q["term"]),
                    int(q["epoch_at_send"]), int(q["last_at_send"])) for q in reps]
            assert got == self._replied, f"native replies differ (node {nd.slot}, tick {self.tick}): {got[:3]} vs {self._replied[:3]}"
            self.counts["native_replies_checked"] += len(got)
        # apply committed commands to the file machine (RaftRoutine.commitState -> applyCommand)
        if self.shadow_native:
            from rafting_b200 import ingest
            applied = np.array(nd.applied, dtype=np.int64)
            got = [(int(a["gid"]), int(a["first"]), int(a["last"])) for a in ingest.apply_ranges(ob, applied)]
            want = [(g, nd.applied[g] + 1, int(ob.commit_index[g])) for g in range(G) if int(ob.commit_index[g]) > nd.applied[g]]
            assert got == want, f"native apply ranges differ (node {nd.slot}, tick {self.tick}): {got[:3]} vs {want[:3]}"
            self.counts["native_apply_ranges_checked"] += len(got)
        for g in range(G):
            c = int(ob.commit_index[g])
            while nd.applied[g] < c:
                nd.applied[g] += 1
                nd.file[g].append(nd.store[g][nd.applied[g]])
        nd.snap = ob

    def _dispatch_row(self, nd: Node, ob: abi.Outbox, placed, row, prev):
        F = nd.F
        # replies to inbound requests, and the outcome of submits
        for (r, g), op in placed.items():
            m = int(ob.rep_meta[r, g]); err = (m >> 8) & 0xFF
            if op["kind"] == abi.OP_SUBMIT:
                if err == 0:
                    # submits (and compaction flushes) lead the step: appended right after the log end of the previous snapshot
                    last = int(prev.last_entry[g]["x"]); term = int(prev.current_term[g])
                    for j in range(op["count"]):
                        self.submitted += 1
                        nd.store[g][last + 1 + j] = (term, f"g{g}:n{nd.slot}:c{self.submitted}")
                    self.counts["submit_ok"] += 1
                elif err not in (24, 25):
                    self.errors.append(("submit", nd.slot, g, err))
                else:
                    self.counts["submit_refused"] += 1
                continue
            if op["kind"] == abi.OP_FLUSH:
                if err:
                    self.errors.append(("flush", nd.slot, g, err))
                    continue
                st = nd.store[g]                                   # RocksLog.flush: deleteRange is end-exclusive
                if st and op["index"] > max(st):
                    st.clear()
                else:
                    for i in [i for i in st if i < op["index"]]:
                        del st[i]
                if "restore" in op:                                # RaftMachine.recover(snapshot)
                    nd.file[g] = list(op["restore"]); nd.applied[g] = op["index"]
                    nd.snapshot[g] = (op["index"], op["term"], list(op["restore"]))
                continue
            src = self.nodes[op["src"]]
            lane = src.lane_of(nd.slot)
            if err == 3 and op["kind"] == abi.OP_AE_REQUEST:
                # upstream quirk, mirrored: a follower whose commitIndex is ahead of min(leaderCommit, last.index)
                # (a new leader that has not yet learnt how far the old one committed) hits "rollback is not
                # allowed" (RocksLog.java:100-103 via Follower.java:80); the event loop logs it, no reply is sent
                self.counts["commit_rollback"] += 1
            elif err:
                self.errors.append(("request", nd.slot, g, op["kind"], err))
            valid, success = m & 1, (m >> 1) & 1
            ekind = {abi.OP_AE_REQUEST: abi.EV_AE_ACK, abi.OP_PREVOTE_REQ: abi.EV_PV_REPLY,
                     abi.OP_VOTE_REQ: abi.EV_RV_REPLY, abi.OP_IS_REQUEST: abi.EV_IS_ACK}[op["kind"]]
            # payload side of the append: on success — and on the one throw site that comes AFTER RaftLog.append inside
            # Follower.appendEntries, the commit rollback assertion (Follower.java:68-80): the entries are in the log
            # although no reply leaves
            if op["kind"] == abi.OP_AE_REQUEST and ((valid and success) or err == 3):
                self._store_entries(nd, g, op["prev_index"], op["entries"])
                self.counts["ae_ok"] += 1
            if op["kind"] in (abi.OP_VOTE_REQ, abi.OP_PREVOTE_REQ) and valid and success and prev is not None and \
                    self._vote_request_is_unsafe(nd, g, op):
                self.counts["votes_granted_to_a_stale_log"] += 1       # the upstream flaw at work (see _vote_request_is_unsafe)
            if valid and self._replied is not None:
                self._replied.append((g, ekind, lane, abi.OUT_OK | (4 if success else 0), op["inc"], int(ob.rep_term[r, g]),
                                      op.get("epoch", 0), op.get("last", 0)))
            if not valid or self._lost(nd.slot, src.slot, g, self.tick, 1):
                self._timeout_event(src, g, lane, ekind, op["inc"], op.get("epoch", 0), op.get("last", 0))
                continue
            self._send(nd.slot, src.slot, g, ("ev", lane, dict(kind=ekind, outcome=abi.OUT_OK, success=bool(success),
                                                               inc=op["inc"], term=int(ob.rep_term[r, g]),
                                                               epoch=op.get("epoch", 0), last=op.get("last", 0))))
        # outbound AppendEntries (Leader.replicateLog)
        pk = (ob.plan_meta[row] & np.uint64(0xF)).astype(np.int64)
        for g, f in np.argwhere((pk == abi.PLAN_AE) | (pk == abi.PLAN_IS)):
            r, g, f = row, int(g), int(f)
            pm = int(ob.plan_meta[r, g, f]); inc = pm >> 32; count = (pm >> 16) & 0xFFFF
            prev_index, prev_term = int(ob.plan_pp[r, g, f]["x"]), int(ob.plan_pp[r, g, f]["y"])
            last, commit = int(ob.plan_lc[r, g, f]["x"]), int(ob.plan_lc[r, g, f]["y"])
            epoch = int(ob.plan_epoch[r, g, f])
            dst = nd.slot_of(f)
            term = nd.inc_term[g].get(inc)
            assert term is not None
            if self._sent is not None:
                is_plan = pk[g, f] == abi.PLAN_IS
                self._sent.append((r, g, abi.OP_IS_REQUEST if is_plan else abi.OP_AE_REQUEST, nd.slot, dst, inc, 0 if is_plan else count,
                                   term, prev_index, prev_term, 0 if is_plan else commit, epoch, last))
            if pk[g, f] == abi.PLAN_IS:                            # (epoch.index, epoch.term) — Leader.java:172
                self.counts["is_sent"] += 1
                if self._lost(nd.slot, dst, g, self.tick, 0):
                    self._timeout_event(nd, g, f, abi.EV_IS_ACK, inc, epoch, last)
                    continue
                self._send(nd.slot, dst, g, ("op", dict(kind=abi.OP_IS_REQUEST, src=nd.slot, term=term, index=prev_index,
                                                        index_term=prev_term, inc=inc, epoch=epoch, last=last)))
                continue
            entries = [nd.store[g][prev_index + 1 + j] for j in range(count)]
            self.counts["ae_sent"] += 1
            if self._lost(nd.slot, dst, g, self.tick, 0):
                self._timeout_event(nd, g, f, abi.EV_AE_ACK, inc, epoch, last)
                continue
            self._send(nd.slot, dst, g, ("op", dict(kind=abi.OP_AE_REQUEST, src=nd.slot, term=term, prev_index=prev_index,
                                                    prev_term=prev_term, entries=entries, leader_commit=commit, inc=inc,
                                                    epoch=epoch, last=last)))
        # PreVote / RequestVote broadcasts
        bk = (ob.ballot_meta[row] & np.uint64(0xF)).astype(np.int64)
        for (g,) in np.argwhere(bk != 0):
            r, g = row, int(g)
            pre = bk[g] == abi.BALLOT_PREVOTE
            inc = int(ob.ballot_meta[r, g]) >> 32
            self.counts["prevote" if pre else "vote"] += 1
            for f in range(F):
                dst = nd.slot_of(f)
                if self._sent is not None:
                    self._sent.append((r, g, abi.OP_PREVOTE_REQ if pre else abi.OP_VOTE_REQ, nd.slot, dst, inc, 0,
                                       int(ob.ballot_term[r, g]), int(ob.ballot_last[r, g]["x"]), int(ob.ballot_last[r, g]["y"]), 0, 0, 0))
                if self._lost(nd.slot, dst, g, self.tick, 2 + f):
                    self._timeout_event(nd, g, f, abi.EV_PV_REPLY if pre else abi.EV_RV_REPLY, inc)
                    continue
                self._send(nd.slot, dst, g, ("op", dict(kind=abi.OP_PREVOTE_REQ if pre else abi.OP_VOTE_REQ, src=nd.slot,
                                                        term=int(ob.ballot_term[r, g]), inc=inc,
                                                        last_index=int(ob.ballot_last[r, g]["x"]),
                                                        last_term=int(ob.ballot_last[r, g]["y"]))))
    @staticmethod
    def _store_entries(nd, g, prev_index, entries):
        """Payload side of RocksLog.conflict / truncate / append (RocksLog.java:169-225)."""
        st = nd.store[g]
        for j, (term, payload) in enumerate(entries):
            idx = prev_index + 1 + j
            have = st.get(idx)
            if have is not None and have[0] == term:
                continue
            if have is not None:
                for i in [i for i in st if i >= idx]:
                    del st[i]
            st[idx] = (term, payload)

    def restart(self, slot, make_sut, restore):
        """Crash + restart of one node (ContextManager.buildContext, ContextManager.java:57-106): the engine state is
        gone; what survives is what the reference keeps on disk — the StableLock record (`restore(gid)` -> term, ballot),
        the payload log (RocksDB: nd.store) and the machine's file + snapshot archive.  commitIndex restarts at 0
        (RocksLog.java:50, volatile) and is re-learnt from the leader."""
        nd = self.nodes[slot]
        sut = make_sut(self.cfgs[slot])
        now = T0 + TICK_MS * (self.tick + 1)
        for g in range(self.G):
            st = restore(g)
            ms = nd.snapshot[g]
            keys = sorted(nd.store[g])
            init = np.zeros(1, dtype=abi.GROUP_INIT_DTYPE)
            init["term"] = st.term; init["ballot"] = st.ballot; init["now_ms"] = now
            init["epoch_index"] = ms[0] if ms else 0; init["epoch_term"] = ms[1] if ms else 0
            init["first_index"] = keys[0] if keys else 1
            init["last_index"] = keys[-1] if keys else 0
            init["last_term"] = nd.store[g][keys[-1]][0] if keys else 0
            sut.open_bulk(g, init)
            runs = []
            for i in keys:
                t = nd.store[g][i][0]
                if not runs or runs[-1][1] != t:
                    runs.append((i, t))
            if len(runs) > 1:
                sut.load_runs(g, runs)
            nd.queue[g].clear()
            if self.shadow_native and self.shadow_builder:
                nd.builder.clear_group(g)
            nd.install[g] = None
        nd.sut, nd.snap = sut, None
        nd.inc_term = [dict() for _ in range(self.G)]
        if self.shadow_native:
            from rafting_b200 import ingest
            nd.dispatch = ingest.Dispatch(self.G, nd.F, slot)
        self.inflight = [m for m in self.inflight if m[2] != slot]     # nothing addressed to the dead process survives
        self.counts["restarts"] += 1

    # ---- scenario helpers ------------------------------------------------------------------------
    def run(self, ticks, submit=True):
        for _ in range(ticks):
            self.step(submit)

    def leader_of(self, g):
        best = None
        for nd in self.nodes:
            if nd.snap is not None and (int(nd.snap.role_word[g]) & 3) == abi.ROLE_LEADER:
                t = int(nd.snap.current_term[g])
                if best is None or t > best[0]:
                    best = (t, nd.slot)
        return None if best is None else best[1]

    # ---- checks ------------------------------------------------------------------------------------
    def check(self, converged: bool):
        assert not self.errors, self.errors[:5]
        for g in range(self.G):
            files = [nd.file[g] for nd in self.nodes]
            # state machine safety: every pair of files agrees on its common prefix
            for a in files:
                for b in files:
                    n = min(len(a), len(b))
                    assert a[:n] == b[:n], f"group {g}: applied logs diverge"
            for nd in self.nodes:
                st = nd.sut.export(g)
                # the engine's term table and the payload store describe the same log
                lo, hi = st.epoch_index + 1, st.last_index
                if nd.store[g]:
                    assert max(nd.store[g]) >= hi
                for i in range(max(lo, hi - 40), hi + 1):
                    assert nd.sut.log_term(g, i) == nd.store[g][i][0], (g, nd.slot, i)
                assert st.commit_index <= nd.applied[g]        # equal unless a snapshot moved the machine ahead of the log
            if converged:
                assert all(f == files[0] for f in files), f"group {g}: the files differ"
