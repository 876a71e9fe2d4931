"""CPU restatement of the PAYLOAD side of the reference's RaftLog, for the entry-buffer parity tests.
TEST INFRASTRUCTURE ONLY (same rule as raft_oracle.c).  PARITY UNPINNED BY UPSTREAM: the reference has no
test of RocksLog either.

RocksLog keeps one RocksDB per context: key = 8-byte BE index, value = 8-byte BE term || Kryo(cmd)
(M/command/storage/RocksLog.java:82-89).  Restated as a dict; the operations follow
  newEntry / append : put                         RocksLog.java:82-89,169-196
  get / batch       : point read / multiGet       RocksLog.java:122-166   (null outside the stored keys)
  truncate          : deleteRange(index, last+1)  RocksLog.java:219-225
  flush             : deleteRange(epoch, index)   RocksLog.java:228-242   (end-exclusive: `index` survives)
"""
from __future__ import annotations


class PayloadLog:
    def __init__(self):
        self.kv: dict[int, tuple[int, bytes]] = {}

    def put(self, index: int, term: int, payload: bytes):
        self.kv[index] = (term, payload)

    def truncate(self, index: int):
        for k in [k for k in self.kv if k >= index]:
            del self.kv[k]

    def flush(self, old_epoch: int, index: int):
        for k in [k for k in self.kv if old_epoch <= k < index]:
            del self.kv[k]

    def batch(self, first: int, n: int):
        out = []
        for i in range(first, first + n):
            if i not in self.kv:
                break
            out.append((i, self.kv[i][0], self.kv[i][1]))
        return out
