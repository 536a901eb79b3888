"""How a worker-side host keeps one GPU busy with a proof's transforms and commitments.

Host mirror of the dispatcher's concurrency (`join_all` over FFT tasks and commitments,
src/dispatcher2.rs:294-306, 316-321, 382-414, 526-532), expressed over the C ABI with HOST buffers:
every transform is fft_init + fft1 (asynchronous copy-in) + fft2_prepare (asynchronous kernels) and
later fft2 (copy-out, blocks for that task only); every commitment is either part of a blocking
dp_msm_batch (serial schedule) or a dp_msm_submit / dp_msm_collect pair (overlapped schedule).

Two schedules over the same work:
  * serial      the commitments of each prover round as one batch, then the transforms with a few
                tasks of look-ahead (copy-in, kernels and copy-out of neighbouring tasks overlap);
  * overlapped  a commitment is queued after every second transform.  Transforms are bound by the
                PCIe copies (32 B in and out per element for ~10 ns of kernel time), commitments by
                the multiplier (tens of ms of kernels per 128 MiB of scalars), so the commitments'
                kernels run while the copy engines move the transforms around them.
Used by bench.py's end-to-end leg and by the tests (emulator on CPU, real library on the GPU).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Sequence


@dataclass
class Transform:
    """one Prover::fft share of this worker: n_rows rows in (host pointer), its columns out"""
    in_ptr: int
    out_ptr: int
    out_bytes: int
    workloads: Sequence          # FftWorkload of every worker
    n_rows: int
    is_quot: bool
    is_inv: bool
    is_coset: bool
    row_len: Optional[int] = None   # leading entries handed in per row (None: whole rows of c); in_ptr is then compact


@dataclass
class Commitment:
    """one varMsm request: bases [start, end) against n scalars at a host pointer"""
    start: int
    end: int
    scalars_ptr: int
    n: int


class Runner:
    def __init__(self, ctx, exchange: Optional[Callable] = None, first_id: int = 1):
        self.ctx, self.lib, self.exchange = ctx, ctx.lib, exchange
        self.next_id = first_id

    def _id(self) -> int:
        self.next_id += 1
        return self.next_id

    def submit(self, t: Transform) -> int:
        """fft_init + fft1 over all local rows (async H2D) + fft2_prepare (async kernels)"""
        ctx, tid = self.ctx, self._id()
        ctx.fft_init(tid, t.workloads, t.is_quot, t.is_inv, t.is_coset)
        if t.row_len is None:
            ctx._ck(self.lib.dp_fft1_rows(ctx.h, tid, 0, t.n_rows, t.in_ptr))
        else:
            ctx._ck(self.lib.dp_fft1_rows_short(ctx.h, tid, 0, t.n_rows, t.in_ptr, t.row_len))
        if self.exchange is None:
            ctx.fft2_prepare(tid)
        else:   # several workers: one all-to-all on the task's own send / receive buffers
            ordered = getattr(self.exchange, "stream_ordered", False)   # enqueued on the compute stream: nothing to wait for
            s, r, blk = ctx.fft_exchange_begin_async(tid) if ordered else ctx.fft_exchange_begin(tid)
            self.exchange(s, r, blk)
            ctx.fft_exchange_end(tid)
        return tid

    def collect(self, tid: int, t: Transform):
        """fft2: copy-out of the task's columns; blocks for this task only"""
        self.ctx._ck(self.lib.dp_fft2(self.ctx.h, tid, t.out_ptr, t.out_bytes))

    def run_serial(self, transforms: Sequence[Transform], commitment: Commitment, rounds: Sequence[int], lookahead: int = 2,
                   on_fft: Optional[Callable] = None, on_msm: Optional[Callable] = None):
        for cnt in rounds:       # one varMsm batch per prover round
            outs = self.ctx.msm_batch([(commitment.start, commitment.end, commitment.scalars_ptr, commitment.n)] * cnt)
            if on_msm:
                for o in outs:
                    on_msm(o)
        pending = []
        for t in transforms:
            pending.append((self.submit(t), t))
            if len(pending) > lookahead:
                self._finish_fft(pending.pop(0), on_fft)
        while pending:
            self._finish_fft(pending.pop(0), on_fft)

    def _finish_fft(self, entry, on_fft):
        self.collect(*entry)
        if on_fft:
            on_fft(entry[1])

    def run_overlapped(self, transforms: Sequence[Transform], commitment: Commitment, n_commitments: int, lookahead: int = 4,
                       on_fft: Optional[Callable] = None, on_msm: Optional[Callable] = None):
        items, left = [], n_commitments
        for j, t in enumerate(transforms):
            items.append(t)
            if j % 2 == 1 and left:
                items.append(None)
                left -= 1
        items += [None] * left
        pending = []

        def finish(entry):
            if entry[0] == "fft":
                self._finish_fft(entry[1:], on_fft)
            else:
                out = self.ctx.msm_collect(entry[1])
                if on_msm:
                    on_msm(out)

        for it in items:
            if it is None:
                mid = self._id()
                self.ctx.msm_submit(mid, commitment.start, commitment.end, commitment.scalars_ptr, commitment.n)
                pending.append(("msm", mid))
            else:
                pending.append(("fft", self.submit(it), it))
            if len(pending) > lookahead:
                finish(pending.pop(0))
        while pending:
            finish(pending.pop(0))


def pick_schedule(runner: Runner, transforms: Sequence[Transform], commitment: Commitment, rounds: Sequence[int],
                  checksum: Callable, timed: Callable, all_agree: Callable = lambda ok: ok, allow_overlap: bool = True):
    """Decide which schedule the end-to-end measurement times.  The overlapped schedule is used only if,
    on this machine, it reproduces the serial one bit for bit - every commitment and `checksum(t)` of every
    transform's output, in order - and is not slower in a one-step trial.
      checksum(transform) -> hashable digest of the transform's host output buffer
      timed(step) -> seconds for one call of step() (the caller's clock, max over ranks)
      all_agree(ok) -> ok on every rank (identity for one worker)
    Returns (step, description): step() runs one pass of the chosen schedule."""
    n_msm = sum(rounds)

    def serial(check=None):
        runner.run_serial(transforms, commitment, rounds, 2, *_observers(check, checksum))

    def overlapped(check=None):
        runner.run_overlapped(transforms, commitment, n_msm, 4, *_observers(check, checksum))

    if not allow_overlap:
        return serial, "serial"
    same, why = False, "gave different results"
    try:
        ref, got = {"msm": [], "fft": []}, {"msm": [], "fft": []}
        serial(ref)
        overlapped(got)
        same = sorted(ref["msm"]) == sorted(got["msm"]) and ref["fft"] == got["fft"] and len(got["msm"]) == n_msm
    except Exception as exc:   # the serial schedule stays available whatever happened to the other one
        why = f"failed: {str(exc)[:120]}"
        runner.ctx.sync()
    if not all_agree(same):
        return serial, f"serial (overlapped schedule {why}: disabled)"
    t_ser, t_ovl = timed(serial), timed(overlapped)
    if t_ovl <= t_ser:
        return overlapped, "overlapped (commitments queued between transforms; verified against the serial schedule)"
    return serial, f"serial (overlapped schedule verified but slower in a one-step trial: {t_ovl * 1e3:.0f} vs {t_ser * 1e3:.0f} ms)"


def _observers(check, checksum):
    if check is None:
        return None, None
    return (lambda t: check["fft"].append(checksum(t))), (lambda o: check["msm"].append(o.tobytes()))
