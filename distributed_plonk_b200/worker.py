"""Host-side mirror of the reference worker's RPC surface (src/worker.rs:125-439,
src/hello_world.capnp:15-52): same method names, argument meaning and payload formats
(`List(Data)` = list of byte chunks cut at 2^28 bytes; raw in-memory structs, utils.rs:27-43),
with every arkworks call replaced by the C ABI (include/dplonk.h).  The Cap'n Proto transport
itself stays in the Rust worker; tests and bench.py drive this class directly.
"""
from __future__ import annotations

import numpy as np

from ._binding import Context, DpError

CHUNK = 1 << 28  # dispatcher.rs:61-63


def chunks(buf, size: int = CHUNK):
    """`v.chunks(1 << 28)` of the dispatcher stubs: the List(Data) form of a byte string."""
    if isinstance(buf, np.ndarray):
        b = memoryview(np.ascontiguousarray(buf).view(np.uint8).reshape(-1))
    else:
        b = memoryview(buf)
    return [b[i:i + size] for i in range(0, len(b), size)] or [b[0:0]]


def concat(chunk_list) -> np.ndarray:
    """`extend_from_slice` over the Data chunks (worker.rs:136-141, 172-175, 244-247)."""
    parts = [np.frombuffer(c, dtype=np.uint8) for c in chunk_list]
    return parts[0] if len(parts) == 1 else np.concatenate(parts)


class PlonkSlave:
    """PlonkImpl: plonk_slave::Server + plonk_peer::Server for one GPU."""

    def __init__(self, cdll, me: int = 0, n_workers: int = 1, device: int = 0):
        self.me, self.n_workers = me, n_workers
        self.ctx = Context(cdll, device, me, n_workers)
        self._dims = {}      # task id -> (r, c, n_cols)
        self._log = (0, 0)

    # init @0 (bases :List(Data), domainSize, quotDomainSize)            worker.rs:126-157
    def init(self, bases, domain_size: int, quot_domain_size: int):
        self.ctx.init(concat(bases), domain_size, quot_domain_size)
        lg = lambda n: max(0, (n - 1).bit_length())
        self._log = (lg(domain_size), lg(quot_domain_size))

    # varMsm @1 (workload: MsmWorkload, scalars :List(Data)) -> (result: Data)   worker.rs:159-185
    def var_msm(self, workload, scalars) -> bytes:
        start, end = workload
        return self.ctx.msm(start, end, concat(scalars)).tobytes()

    # fftInit @2                                                          worker.rs:187-233
    def fft_init(self, task_id: int, workloads, is_quot: bool, is_inv: bool, is_coset: bool):
        self.ctx.fft_init(task_id, workloads, is_quot, is_inv, is_coset)
        L = self._log[1 if is_quot else 0]
        r = 1 << (L >> 1)
        mine = workloads[self.me]
        self._dims[task_id] = (r, (1 << L) // r, mine[3] - mine[2])

    # fft1 @3 (id, i, v :List(Data))                                      worker.rs:235-278
    def fft1(self, task_id: int, i: int, v):
        self.ctx.fft1(task_id, i, concat(v))

    # fft2Prepare @4 (id)   + the worker<->worker fftExchange             worker.rs:280-345, 412-438
    def fft2_prepare(self, task_id: int, exchange=None):
        """n_workers == 1: local.  Otherwise `exchange(send_ptr, recv_ptr, block_elems)` must move
        block q of send to rank q's recv block `me` (an all-to-all; see parallel.py)."""
        if self.n_workers == 1 or self.ctx.peer_ready():
            # peers attached: the row kernel stores into the owners' arenas over NVLink; the caller
            # (the dispatcher's join over all fft2Prepare replies) is the barrier before fft2
            self.ctx.fft2_prepare(task_id)
            return
        if exchange is None:
            raise DpError(-5, "fft2_prepare on a multi-worker task needs an exchange callable")
        send, recv, n = self.ctx.fft_exchange_begin(task_id)
        exchange(send, recv, n)
        self.ctx.fft_exchange_end(task_id)

    # fft2 @5 (id) -> (v :List(Data))   one Data per local column          worker.rs:347-381
    def fft2(self, task_id: int):
        cols = self.fft2_array(task_id)
        return [cols[k].tobytes() for k in range(cols.shape[0])]

    def fft2_array(self, task_id: int) -> np.ndarray:
        r, _, n_cols = self._dims.pop(task_id)
        return self.ctx.fft2(task_id, n_cols, r)

    # round1 @6 (w :List(Data)) -> (c :Data)                               worker.rs:383-408
    def round1(self, w, blind=None) -> bytes:
        """blind: two secret uniformly random Fr (tests inject them); None = drawn by the library from getrandom(2)"""
        return self.ctx.round1(concat(w), blind).tobytes()

    def close(self):
        self.ctx.close()
