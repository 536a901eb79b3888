"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle on the same
seeded inputs - bit-exact Fr arrays for the NTT pieces, identical affine point (and normalised
raw bytes) for the MSM - plus size-independent properties at BASELINE.json's full sizes."""
import os

import numpy as np
import pytest
import torch

from distributed_plonk_b200 import dispatcher as disp
from distributed_plonk_b200 import parallel
from distributed_plonk_b200._binding import Context, DpError
from distributed_plonk_b200.worker import PlonkSlave, chunks
from tests import common

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz")


def device_copy(dst, src, n):
    parallel.as_tensor(dst, n, True).copy_(parallel.as_tensor(src, n, True))
    torch.cuda.synchronize()


@pytest.fixture(scope="module")
def bases(orc):
    return orc.gen_bases(5, (1 << 16) + 32, 2048, True)


@pytest.fixture(scope="module")
def ctx(gpu_lib, orc, bases):
    c = Context(gpu_lib, 0, 0, 1)
    c.init(bases, 1 << 12, 1 << 15)            # BASELINE config 0 sizes (n = 2^12, quotient 2^15)
    yield c
    c.close()


def test_golden_vectors_on_gpu(orc, gpu_lib):
    g = np.load(GOLDEN)
    c = Context(gpu_lib, 0, 0, 1)
    c.init(g["msm_bases"], 1 << 6, 1 << 9)
    for inv in (0, 1):
        for cos in (0, 1):
            assert np.array_equal(c.ntt(g["ntt_in"], 6, bool(inv), bool(cos)), g[f"ntt_out_{inv}{cos}"])
    out = c.msm(0, int(g["msm_n"]), g["msm_scalars"])
    assert np.array_equal(orc.normalize(out), g["msm_out_affine"])
    c.close()


@pytest.mark.parametrize("log_n", [0, 1, 2, 5, 8, 11, 12, 13, 16, 18, 19, 20])
def test_whole_ntt(orc, ctx, log_n):
    common.check_whole_ntt(orc, ctx, log_n, 40 + log_n)
    if log_n >= 5:
        common.check_whole_ntt(orc, ctx, log_n, 90 + log_n, n_in=(1 << log_n) // 8 + 3)


@pytest.mark.parametrize("limits,logs", [((3, 2), (4, 5, 6)), ((4, 3), (7, 9)), ((6, 5), (12, 15))])
def test_whole_ntt_forced_multi_pass(orc, ctx, limits, logs):
    ctx.debug_set_limits(limits[0], limits[1], 0)
    try:
        for log_n in logs:
            common.check_whole_ntt(orc, ctx, log_n, 50 + log_n)
    finally:
        ctx.debug_set_limits(11, 9, 0)


@pytest.mark.parametrize("W,logn,logq,limits", [(1, 12, 15, (11, 9)), (1, 6, 9, (2, 2)), (2, 12, 15, (11, 9)),
                                                (4, 8, 11, (3, 2)), (8, 12, 15, (4, 3)), (1, 11, 13, (11, 9))])
def test_distributed_fft_like_reference_test_fft(orc, gpu_lib, W, logn, logq, limits):
    """dispatcher.rs:246-350 (n = 2^11, quotient 2^13 there) with W workers as W contexts on one GPU"""
    workers = [PlonkSlave(gpu_lib, p, W) for p in range(W)]
    try:
        for w in workers:
            w.init([b""], 1 << logn, 1 << logq)
            w.ctx.debug_set_limits(limits[0], limits[1], 0)
        common.check_distributed_fft(orc, workers, logn, False, 3, device_copy)
        common.check_distributed_fft(orc, workers, logq, True, 4, device_copy)
        common.check_distributed_fft(orc, workers, logq, True, 5, device_copy, n_in=(1 << logq) // 8)
    finally:
        for w in workers:
            w.close()


@pytest.mark.parametrize("W,logn,logq,limits", [(2, 12, 15, (11, 9)), (4, 8, 11, (3, 2))])
def test_distributed_fft_fused_peer_exchange(orc, gpu_lib, W, logn, logq, limits):
    """W contexts on ONE GPU attached through CUDA IPC is not possible (IPC handles cannot be opened
    in the exporting process), so on a single device the peer arenas are exercised across processes
    in tests/test_gpu_multi.py; here only the API's error behaviour is checked."""
    w = PlonkSlave(gpu_lib, 0, W)
    try:
        w.init([b""], 1 << logn, 1 << logq)
        assert not w.ctx.peer_ready()
        h = w.ctx.peer_arena_create(2 * (1 << logq) * 32 // W)
        assert len(h) == 64 and not w.ctx.peer_ready()
        with pytest.raises(DpError):
            w.ctx.peer_arena_create(1 << 20)          # second arena
        with pytest.raises(DpError):
            w.ctx.peer_attach(W, h)                   # peer index out of range
    finally:
        w.close()


@pytest.mark.parametrize("logq", [20, 23])
def test_distributed_fft_large(orc, gpu_lib, logq):
    """2^20: single-pass rows/cols (r = c = 2^10); 2^23: r = 2^11, c = 2^12 -> both phases split"""
    c = Context(gpu_lib, 0, 0, 1)
    try:
        c.init(np.zeros(0, dtype=np.uint8), 1 << 10, 1 << logq)
        N = 1 << logq
        r = 1 << (logq >> 1)
        cc = N // r
        for k, (inv, cos) in enumerate([(False, True), (True, True), (True, False)]):
            x = orc.gen_fr(200 + k, N)
            rows = disp.dispatcher_rows(x, logq)
            c.fft_init(k, disp.fft_workloads(logq, 1), True, inv, cos)
            c.fft1_rows(k, 0, rows, r)
            c.fft2_prepare(k)
            got = disp.assemble(c.fft2(k, cc, r))
            assert np.array_equal(got, orc.fft(x, inv, cos)), f"2^{logq} inv={inv} coset={cos}"
    finally:
        c.close()


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 1000, 1 << 12, (1 << 12) + 32, (1 << 16) + 32])
def test_msm_vs_oracle(orc, ctx, bases, n):
    common.check_msm(orc, ctx, bases, n, 300 + (n % 97))


def test_msm_windowed_path_vs_precomputed(orc, ctx, bases):
    """the default path uses the precomputed window multiples (SRS of 2^16+32 bases); mode 1 forces
    the per-window bucket sets: both must give the oracle's point"""
    n = (1 << 16) + 32
    ctx.debug_set_limits(11, 9, 1)
    try:
        common.check_msm(orc, ctx, bases, n, 777, which=("uniform", "witness-like"))
    finally:
        ctx.debug_set_limits(11, 9, 0)
    common.check_msm(orc, ctx, bases, n, 777, which=("uniform", "witness-like"))


@pytest.mark.parametrize("c", [4, 5, 9, 12, 15, 16, 17, 18])
def test_msm_every_window_geometry(orc, ctx, bases, c):
    ctx.debug_set_limits(11, 9, c)
    try:
        common.check_msm(orc, ctx, bases, 5000, 400 + c, which=("uniform", "witness-like", "all r-1"))
    finally:
        ctx.debug_set_limits(11, 9, 0)


def test_msm_dev_batch(orc, ctx, bases):
    """a round's commitments issued together (dp_msm_dev_batch): device pointers, tails overlapped"""
    n = (1 << 16) + 32
    scs = [orc.gen_fr(900 + k, n, False) for k in range(5)]
    dev = [torch.from_numpy(s.view(np.int64)).cuda() for s in scs]
    outs = [torch.zeros(18, dtype=torch.int64, device="cuda") for _ in range(5)]
    ranges = [(0, n), (0, n), (5, 60000), (0, n), (n - 100, n)]
    ctx.msm_dev_batch([(lo, hi, dev[k].data_ptr(), hi - lo, outs[k].data_ptr()) for k, (lo, hi) in enumerate(ranges)])
    for k, (lo, hi) in enumerate(ranges):
        got = outs[k].cpu().numpy().view(np.uint8)
        common.assert_point_eq(orc, got, orc.msm(bases[lo:hi], scs[k][: hi - lo]), f"batch job {k}")
    outs = ctx.msm_batch([(lo, hi, scs[k], hi - lo) for k, (lo, hi) in enumerate(ranges)])
    for k, (lo, hi) in enumerate(ranges):
        common.assert_point_eq(orc, outs[k], orc.msm(bases[lo:hi], scs[k][: hi - lo]), f"host batch job {k}")


def test_msm_edges(orc, ctx, bases):
    sc = orc.gen_fr(9, 600, False)
    assert orc.normalize(ctx.msm(10, 10, sc[:0]))[96] == 1
    common.assert_point_eq(orc, ctx.msm(100, 333, sc[:233]), orc.msm(bases[100:333], sc[:233]), "sub-range")
    common.assert_point_eq(orc, ctx.msm(0, 600, sc[:50]), orc.msm(bases[:50], sc[:50]), "truncate to scalars")
    common.assert_point_eq(orc, ctx.msm(3, 4, sc[:1]), orc.msm(bases[3:4], sc[:1]), "infinity base")
    with pytest.raises(DpError):
        ctx.msm(0, bases.shape[0] + 1, sc)
    s2 = np.zeros((2049, 4), dtype=np.uint64)       # bases 0 and 2048 are the same point
    s2[0] = common.u256(5)
    s2[2048] = common.u256(common.R_MOD - 5)
    assert orc.normalize(ctx.msm(0, 2049, s2))[96] == 1
    s2[2048] = common.u256(5)
    common.assert_point_eq(orc, ctx.msm(0, 2049, s2), orc.msm(bases[:2049], s2), "same point twice")


def test_sharded_msm_like_reference_test_msm(orc, gpu_lib, bases):
    """dispatcher.rs:177-244: full bases on every worker, global index ranges, partials summed"""
    n = 1 << 14
    workers = [PlonkSlave(gpu_lib, p, 4) for p in range(4)]
    try:
        for w in workers:
            w.init(chunks(bases[:n]), 1 << 4, 1 << 7)
        common.check_sharded_msm(orc, workers, bases, n, 77)
    finally:
        for w in workers:
            w.close()


def test_commit_and_round1(orc, ctx, bases):
    co = orc.gen_fr(21, 3000, True)
    common.assert_point_eq(orc, ctx.commit(co), orc.commit(bases, co), "commit_polynomial")
    n = 1 << 12
    evals = orc.gen_fr(22, n, True)
    blind = orc.gen_fr(23, 2, True)
    got = ctx.round1(evals, blind)
    poly = orc.fft(evals, True, False)
    L = orc.lib()
    wire = np.zeros((n + 2, 4), dtype=np.uint64)
    wire[:n] = poly
    for k in range(2):
        L.orc_fr_sub(wire[k].ctypes.data, blind[k].ctypes.data, wire[k].ctypes.data)
        wire[n + k] = blind[k]
    assert np.array_equal(ctx.get_wire(), wire)
    common.assert_point_eq(orc, got, orc.commit(bases, wire), "round1 commitment")
    # internally drawn blinders: commitment must still open to get_wire()
    got2 = ctx.round1(evals, None)
    common.assert_point_eq(orc, got2, orc.commit(bases, ctx.get_wire()), "round1 (internal blinders)")


@pytest.mark.parametrize("n", [1, 2, 1000, 1 << 12, (1 << 16) + 3])
def test_perm_product(orc, ctx, n):
    """next row §8(f)-3: round-2 grand product (dispatcher2.rs:329-345), 5 wire types"""
    common.check_perm_product(orc, ctx, n, 5, 500 + (n % 89))
    if n == 1 << 12:   # device-resident variant
        w, i_, s_ = (np.stack([orc.gen_fr(700 + 10 * k + t, n) for t in range(5)]) for k in range(3))
        beta, gamma = orc.gen_fr(790, 1)[0], orc.gen_fr(791, 1)[0]
        dev = [torch.from_numpy(a.view(np.int64)).cuda() for a in (w, i_, s_)]
        out = torch.empty((n, 4), dtype=torch.int64, device="cuda")
        ctx.perm_product_dev(dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(), 5, n, beta, gamma, out.data_ptr())
        assert np.array_equal(out.cpu().numpy().view(np.uint64), orc.perm_product(w, i_, s_, beta, gamma))


# ---------------------------------------------------------------- full BASELINE sizes: properties
def test_full_size_msm_2p20_vs_oracle_and_linearity(orc, gpu_lib):
    """MSM at 2^20+32 against the oracle, then linearity msm(s)+msm(t) == msm(s+t mod r) and the
    sharded sum at the same size."""
    n = (1 << 20) + 32
    bases = orc.gen_bases(8, n, 2048, True)
    c = Context(gpu_lib, 0, 0, 1)
    try:
        c.init(bases, 1 << 10, 1 << 13)
        s = orc.gen_fr(501, n, False)
        t = orc.gen_fr(502, n, False)
        ms = c.msm(0, n, s)
        common.assert_point_eq(orc, ms, orc.msm(bases, s), "msm 2^20+32")
        mt = c.msm(0, n, t)
        # linearity on a sparse subset (every 4099th scalar, the rest zero): s + t mod r in Python ints
        R = common.R_MOD
        idx = list(range(0, n, 4099))
        sub_s = np.zeros_like(s)
        sub_t = np.zeros_like(t)
        sub_st = np.zeros_like(s)
        for i in idx:
            a = int.from_bytes(s[i].tobytes(), "little")
            b = int.from_bytes(t[i].tobytes(), "little")
            sub_s[i], sub_t[i], sub_st[i] = s[i], t[i], common.u256((a + b) % R)
        lhs = orc.g1_add(c.msm(0, n, sub_s), c.msm(0, n, sub_t))
        assert np.array_equal(orc.normalize(lhs), orc.normalize(c.msm(0, n, sub_st)))
        # index-range shards sum to the whole
        acc = c.msm(0, n // 2, s[: n // 2])
        acc = orc.g1_add(acc, c.msm(n // 2, n, s[n // 2:]))
        assert np.array_equal(orc.normalize(acc), orc.normalize(ms))
        assert not np.array_equal(orc.normalize(ms), orc.normalize(mt))
    finally:
        c.close()


def test_full_size_msm_2p22_vs_oracle(orc, gpu_lib):
    """BASELINE's headline size: MSM over 2^22+32 DISTINCT bases (the synthetic SRS bench.py uses,
    generated on the device; points checked on-curve by the oracle's own arithmetic when it adds
    them) against the oracle's Pippenger, uniform and witness-like scalars."""
    n = (1 << 22) + 32
    c = Context(gpu_lib, 0, 0, 1)
    try:
        bases = c.gen_bases(0xD15791B07E5EED, n)
        bases[3, :] = 0
        bases[3, 96] = 1                       # infinity at index 3 (dispatcher2.rs:1100-1101)
        c.init(bases, 1 << 10, 1 << 13)
        for name, sc in common.scalar_sets(orc, n, 4242).items():
            if name in ("uniform", "witness-like"):
                common.assert_point_eq(orc, c.msm(0, n, sc), orc.msm(bases, sc), f"msm 2^22+32 {name}")
    finally:
        c.close()


def test_full_size_ntt_2p25_roundtrip_and_spot_checks(orc, gpu_lib):
    """quotient-domain size of the 2^22-gate config: coset NTT of an n-coefficient polynomial on
    the 8n domain through the worker path; spot-check outputs by O(N) Horner evaluation, then the
    inverse transform must give the input back bit for bit."""
    logq = 25
    N, n = 1 << logq, 1 << 22
    c = Context(gpu_lib, 0, 0, 1)
    try:
        c.init(np.zeros(0, dtype=np.uint8), n, N)
        r = 1 << (logq >> 1)
        cc = N // r
        x = np.zeros((N, 4), dtype=np.uint64)
        x[:n] = orc.gen_fr(600, n)
        c.fft_init(1, disp.fft_workloads(logq, 1), True, False, True)
        c.fft1_rows(1, 0, disp.dispatcher_rows(x, logq), r)
        c.fft2_prepare(1)
        y = disp.assemble(c.fft2(1, cc, r))
        for k in (0, 1, 12345, N // 2 + 7, N - 1):
            assert np.array_equal(y[k], orc.ntt_output_at(x, k, False, True)), f"X[{k}]"
        c.fft_init(2, disp.fft_workloads(logq, 1), True, True, True)
        c.fft1_rows(2, 0, disp.dispatcher_rows(y, logq), r)
        c.fft2_prepare(2)
        back = disp.assemble(c.fft2(2, cc, r))
        assert np.array_equal(back, x)
        # whole-domain entry point agrees with the 2-D path
        assert np.array_equal(c.ntt(x[:n], logq, False, True), y)
    finally:
        c.close()
