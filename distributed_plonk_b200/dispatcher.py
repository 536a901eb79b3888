"""Host-side mirror of the dispatcher's use of the hot path (src/dispatcher2.rs:731-787 Prover::fft,
834-893 Prover::commit_polynomial; stubs 945-1086), over in-process PlonkSlave objects instead of
Cap'n Proto connections.  Index conventions follow the reference line by line."""
from __future__ import annotations

import numpy as np

from .worker import chunks


def fft_workloads(domain_log: int, n_workers: int):
    """FftWorkload per worker: equal row / column blocks (dispatcher2.rs:1143-1156)."""
    r = 1 << (domain_log >> 1)
    c = (1 << domain_log) // r
    W = n_workers
    return [(p * r // W, (p + 1) * r // W, p * c // W, (p + 1) * c // W) for p in range(W)]


def dispatcher_rows(coeffs: np.ndarray, domain_log: int) -> np.ndarray:
    """dispatcher2.rs:746,754: resize to the domain, chunk by r, transpose -> [r][c][4]."""
    N = 1 << domain_log
    r = 1 << (domain_log >> 1)
    c = N // r
    x = np.zeros((N, 4), dtype=np.uint64)
    x[: coeffs.shape[0]] = coeffs
    return np.ascontiguousarray(x.reshape(c, r, 4).transpose(1, 0, 2))


def assemble(cols: np.ndarray) -> np.ndarray:
    """dispatcher2.rs:780-786: u[col] = vec(r); transpose(u).concat()  ->  out[j*c + i] = col_i[j]."""
    c, r = cols.shape[0], cols.shape[1]
    return np.ascontiguousarray(cols.transpose(1, 0, 2)).reshape(r * c, 4)


def fft(workers, domain_log: int, coeffs: np.ndarray, is_quot: bool, is_inv: bool, is_coset: bool,
        task_id: int, exchange=None) -> np.ndarray:
    """Prover::fft (dispatcher2.rs:731-787) against in-process workers.  Returns the transformed
    vector (N x 4 u64, raw Fr)."""
    W = len(workers)
    r = 1 << (domain_log >> 1)
    c = (1 << domain_log) // r
    wl = fft_workloads(domain_log, W)
    rows = dispatcher_rows(coeffs, domain_log)
    for w in workers:
        w.fft_init(task_id, wl, is_quot, is_inv, is_coset)
    for p, w in enumerate(workers):                       # one fft1 RPC per row (dispatcher2.rs:756-766)
        for j in range(wl[p][1] - wl[p][0]):
            w.fft1(task_id, j, chunks(rows[wl[p][0] + j]))
    for w in workers:
        w.fft2_prepare(task_id, exchange)
    u = [None] * c
    for p, w in enumerate(workers):
        for j, v in enumerate(w.fft2(task_id)):           # dispatcher2.rs:780-784
            u[p * c // W + j] = np.frombuffer(v, dtype=np.uint64).reshape(r, 4)
    return assemble(np.stack(u))


def commit_polynomial(workers, n_bases: int, scalars_canonical: np.ndarray) -> list:
    """Prover::commit_polynomial (dispatcher2.rs:834-893) with the tested contract of
    dispatcher.rs:213-229: every worker holds all bases and gets the global index range of its
    chunk.  Returns the per-worker 144-byte partials (the dispatcher sums them, 887-890)."""
    W = len(workers)
    plain = np.zeros((n_bases, 4), dtype=np.uint64)
    plain[: scalars_canonical.shape[0]] = scalars_canonical
    out = []
    for i, w in enumerate(workers):
        lo, hi = i * n_bases // W, (i + 1) * n_bases // W
        out.append(w.var_msm((lo, hi), chunks(plain[lo:hi])))
    return out
