"""Kernel-LOGIC tests on a CPU-only box: the library's .cu sources compiled against the CUDA
execution-model emulator in tests/emul (barriers, shared memory, atomics and carry flags emulated;
the real PTX paths are exercised by the `-m gpu` tests).  Same checks as tests/test_gpu_parity.py,
at sizes the emulator finishes in seconds."""
import ctypes as C
import os

import numpy as np
import pytest

from distributed_plonk_b200._binding import Context, DpError
from distributed_plonk_b200 import dispatcher as disp
from distributed_plonk_b200.worker import PlonkSlave, chunks
from tests import common


def host_copy(dst, src, n):
    C.memmove(dst, src, n)


FULL = os.environ.get("DP_TEST_FULL", "0") == "1"   # the larger shapes of a few tests: DP_TEST_FULL=1


@pytest.fixture(scope="module")
def ctx(emul_lib, orc):
    c = Context(emul_lib, 0, 0, 1)
    c.init(orc.gen_bases(5, 600, 64, True), 1 << 6, 1 << 9)
    yield c
    c.close()


def test_whole_ntt_single_pass(orc, ctx):
    ctx.debug_set_limits(11, 9, 0)
    for log_n in (0, 1, 2, 3, 6, 9, 11):
        common.check_whole_ntt(orc, ctx, log_n, 40 + log_n)
    common.check_whole_ntt(orc, ctx, 8, 77, n_in=100)          # zero-padded short input


@pytest.mark.parametrize("limits,logs", [((3, 2), (4, 5, 6)), ((4, 3), (7, 9)), ((11, 9), (12,))])
def test_whole_ntt_multi_pass_plans(orc, ctx, limits, logs):
    ctx.debug_set_limits(limits[0], limits[1], 0)
    for log_n in logs:
        common.check_whole_ntt(orc, ctx, log_n, 50 + log_n)
    ctx.debug_set_limits(11, 9, 0)


@pytest.mark.parametrize("W,logn,logq,limits", [(1, 6, 9, (11, 9)), (1, 6, 9, (2, 2)), (2, 6, 9, (11, 9)),
                                                (4, 6, 9, (2, 2)), (2, 8, 7, (3, 2))])
def test_distributed_fft_like_reference_test_fft(orc, emul_lib, W, logn, logq, limits):
    workers = [PlonkSlave(emul_lib, p, W) for p in range(W)]
    for w in workers:
        w.init([b""], 1 << logn, 1 << logq)
        w.ctx.debug_set_limits(limits[0], limits[1], 0)
    common.check_distributed_fft(orc, workers, logn, False, 3, host_copy)
    common.check_distributed_fft(orc, workers, logq, True, 4, host_copy)
    common.check_distributed_fft(orc, workers, logq, True, 5, host_copy, n_in=(1 << logq) // 8)  # n coeffs on the 8n domain
    if W > 1:   # stream-ordered begin (dp_fft_exchange_begin_async) + the compute-stream handle
        common.check_distributed_fft(orc, workers, logq, True, 6, host_copy, async_begin=True)
        assert isinstance(workers[0].ctx.compute_stream(), int)
    for w in workers:
        w.close()


def test_fft1_short_and_long_rows(orc, emul_lib):
    w = PlonkSlave(emul_lib, 0, 1)
    w.init([b""], 1 << 6, 1 << 9)
    common.check_fft1_row_lengths(orc, w, 9, True, 60)
    common.check_fft1_row_lengths(orc, w, 6, False, 61)
    w.close()


@pytest.mark.parametrize("W,limits", [(2, (11, 9)), (4, (2, 2))])
def test_distributed_fft_fused_peer_exchange(orc, emul_lib, W, limits):
    """dp_peer_arena_create / dp_peer_attach: the row kernel stores into the owners' arenas (the
    emulator's IPC handle is the pointer itself); two slots alternate across consecutive tasks"""
    workers = [PlonkSlave(emul_lib, p, W) for p in range(W)]
    for w in workers:
        w.init([b""], 1 << 6, 1 << 9)
        w.ctx.debug_set_limits(limits[0], limits[1], 0)
    common.attach_in_process(workers, 2 * (1 << 9) * 32 // W)
    common.check_distributed_fft(orc, workers, 6, False, 3, host_copy)
    common.check_distributed_fft(orc, workers, 9, True, 4, host_copy)
    for w in workers:
        w.close()


@pytest.mark.parametrize("logn,logq", [(0, 0), (0, 3), (1, 4), (2, 5)])
def test_distributed_fft_tiny_domains(orc, emul_lib, logn, logq):
    w = PlonkSlave(emul_lib, 0, 1)
    w.init([b""], 1 << logn, 1 << logq)
    common.check_distributed_fft(orc, [w], logn, False, 3, host_copy)
    common.check_distributed_fft(orc, [w], logq, True, 4, host_copy)
    w.close()


def test_msm_distributions_and_geometries(orc, ctx):
    bases = orc.gen_bases(5, 600, 64, True)
    ctx.debug_set_limits(11, 9, 0)
    common.check_msm(orc, ctx, bases, 600 if FULL else 300, 21)
    for c in (4, 7, 13):
        ctx.debug_set_limits(11, 9, c)
        common.check_msm(orc, ctx, bases, 300 if FULL else 150, 30 + c, which=("uniform", "witness-like"))
    ctx.debug_set_limits(11, 9, 0)


def test_msm_precomputed_window_multiples(orc, emul_lib):
    """SRS of >= 2^11 bases: dp_init builds the table 2^(c*w) * P_i and the MSM runs over one shared
    bucket set; small sub-ranges and the forced modes still take the per-window path."""
    n = 2048
    bases = orc.gen_bases(11, n, 64, True)
    c = Context(emul_lib, 0, 0, 1)
    c.init(bases, 1 << 4, 1 << 7)
    common.check_msm(orc, c, bases, n, 61, which=("uniform", "witness-like", "all one"))
    sc = orc.gen_fr(62, n, False)
    common.assert_point_eq(orc, c.msm(500, 1700, sc[:1200]), orc.msm(bases[500:1700], sc[:1200]), "pre sub-range")
    common.assert_point_eq(orc, c.msm(7, 57, sc[:50]), orc.msm(bases[7:57], sc[:50]), "small range -> windowed")
    c.debug_set_limits(11, 9, 1)
    common.assert_point_eq(orc, c.msm(0, n, sc), orc.msm(bases, sc), "forced windowed")
    c.close()
    # worker 1 of 2: the table covers only its MsmWorkload shard [n, 2n)
    bases2 = orc.gen_bases(12, 2 * n, 64, True)
    c = Context(emul_lib, 0, 1, 2)
    c.init(bases2, 1 << 4, 1 << 7)
    common.assert_point_eq(orc, c.msm(n, 2 * n, sc), orc.msm(bases2[n:], sc), "own shard (table)")
    common.assert_point_eq(orc, c.msm(0, n, sc), orc.msm(bases2[:n], sc), "other shard (per-window path)")
    common.assert_point_eq(orc, c.msm(n - 5, 2 * n, sc), orc.msm(bases2[n - 5:], sc), "straddling range")
    c.close()


@pytest.mark.parametrize("levels", [1, 2, 3])
def test_msm_batched_affine_levels(orc, emul_lib, monkeypatch, levels):
    """the batched-affine tree levels in front of the XYZZ chunks (DP_MSM_AFFINE): every scalar distribution, both
    bucket layouts (per window / one shared set over the window-multiple table), the additions that are not chords -
    holes, infinity, P + P, P + (-P) - and a batch"""
    monkeypatch.setenv("DP_MSM_AFFINE", str(levels))
    monkeypatch.setenv("DP_MSM_AFFINE_MIN", "0")
    bases = orc.gen_bases(5, 600, 64, True)                  # 64 distinct points tiled: equal points meet in the buckets
    c = Context(emul_lib, 0, 0, 1)
    c.init(bases, 1 << 4, 1 << 7)
    assert c.msm_tuning()["levels"] == levels and c.msm_tuning()["equal"] == -1       # forced: nothing was tuned
    l0 = c.launch_count()
    which = None if levels == 2 else ("witness-like", "all r-1")
    common.check_msm(orc, c, bases, 200, 21, which=which)
    assert c.launch_count() - l0 == (5 if levels == 2 else 2) * (11 + 3 * levels + 2)
    for cbits in (4, 13) if levels == 2 else (7,):
        c.debug_set_limits(11, 9, cbits)
        common.check_msm(orc, c, bases, 150, 30 + cbits, which=("witness-like",))
    c.debug_set_limits(11, 9, 0)
    sc = orc.gen_fr(9, 600, False)
    common.assert_point_eq(orc, c.msm(3, 4, sc[:1]), orc.msm(bases[3:4], sc[:1]), "infinity base")
    s2 = np.zeros((65, 4), dtype=np.uint64)
    s2[0] = common.u256(5)
    s2[64] = common.u256(common.R_MOD - 5)
    assert orc.normalize(c.msm(0, 65, s2))[96] == 1           # P + (-P)
    s2[64] = common.u256(5)
    common.assert_point_eq(orc, c.msm(0, 65, s2), orc.msm(bases[:65], s2), "same point twice")
    if levels == 2:
        scs = [np.ascontiguousarray(orc.gen_fr(70 + k, 300, False)) for k in range(3)]
        outs = c.msm_batch([(0, 300, scs[0], 300), (100, 250, scs[1], 150), (0, 0, scs[2], 0)])
        for k, (lo, hi) in enumerate([(0, 300), (100, 250), (0, 0)]):
            common.assert_point_eq(orc, outs[k], orc.msm(bases[lo:hi], scs[k][: hi - lo]), f"host batch job {k}")
    c.close()
    # (the shared bucket set over the window-multiple table: test_msm_tuning_at_init)


def test_msm_tuning_at_init(orc, emul_lib, monkeypatch):
    """dp_init times the plain pipeline against two tree levels over the context's own table and keeps the levels only
    if both give the same 144 bytes; whatever it chose, MSMs agree with the oracle"""
    monkeypatch.delenv("DP_MSM_AFFINE", raising=False)
    monkeypatch.delenv("DP_MSM_TUNE", raising=False)
    monkeypatch.setenv("DP_MSM_AFFINE_MIN", "0")               # (by default only SRS with >= 2^22 digits per MSM are tuned)
    n = 2048
    bases = orc.gen_bases(11, n, 64, True)
    c = Context(emul_lib, 0, 0, 1)
    c.init(bases, 1 << 4, 1 << 7)
    t = c.msm_tuning()                                         # default: the plain pipeline against two levels
    assert t["equal"] == 1 and t["plain_ms"] > 0 and t["affine_ms"] > 0 and t["levels"] in (0, 2)
    assert t["ms_by_levels"][1] == 0 and t["ms_by_levels"][3] == 0 and t["ms_by_levels"][2] > 0
    monkeypatch.setenv("DP_MSM_TUNE", "2")                     # the wider search bench.py's probe runs: 1, 2 and 3 levels
    c3 = Context(emul_lib, 0, 0, 1)
    c3.init(bases, 1 << 4, 1 << 7)
    t3 = c3.msm_tuning()
    assert t3["equal"] == 1 and all(v > 0 for v in t3["ms_by_levels"]) and t3["levels"] in (0, 1, 2, 3)
    c3.close()
    monkeypatch.delenv("DP_MSM_TUNE")
    common.check_msm(orc, c, bases, n, 63, which=("uniform",))
    monkeypatch.setenv("DP_MSM_AFFINE", "2")                   # two levels over the table's shared bucket set, whatever the tuning chose
    c2 = Context(emul_lib, 0, 0, 1)
    c2.init(bases, 1 << 4, 1 << 7)
    common.check_msm(orc, c2, bases, n, 61, which=("witness-like",))
    c2.close()
    monkeypatch.delenv("DP_MSM_AFFINE")
    c.init(orc.gen_bases(5, 100, 64, True), 1 << 4, 1 << 7)   # a small SRS has no table: nothing to tune, plain pipeline
    assert c.msm_tuning() == {"plain_ms": 0.0, "affine_ms": 0.0, "levels": 0, "equal": -1}
    c.close()
    monkeypatch.setenv("DP_MSM_TUNE", "0")                     # switched off: dp_init never tunes and never selects the levels
    c = Context(emul_lib, 0, 0, 1)
    c.init(bases, 1 << 4, 1 << 7)
    assert c.msm_tuning() == {"plain_ms": 0.0, "affine_ms": 0.0, "levels": 0, "equal": -1}
    c.close()


def test_msm_probe_is_a_guarded_opt_in(monkeypatch):
    """tune.probe runs dp_init's tuning in a child process; whatever happens there, the caller only opts in on
    "identical and faster" (here the child cannot even create a context: no GPU)"""
    from distributed_plonk_b200 import tune
    assert tune.choose({"plain_ms": 23.0, "affine_ms": 21.0, "levels": 2, "equal": 1}) == 2
    assert tune.choose({"plain_ms": 23.0, "affine_ms": 24.0, "levels": 0, "equal": 1}) == 0      # slower: the tuning said 0
    assert tune.choose({"plain_ms": 23.0, "affine_ms": 21.0, "levels": 2, "equal": 0}) == 0      # different results
    assert tune.choose({"plain_ms": 0.0, "affine_ms": 0.0, "levels": 0, "equal": -1}) == 0       # not tuned
    assert tune.choose({"error": "probe exited with -11"}) == 0
    import torch
    if not torch.cuda.is_available():
        res = tune.probe(0, 0, 1, 12, timeout=120)
        assert "error" in res and tune.choose(res) == 0
    # what the parent makes of a child's output: the last line is the result, anything before it is ignored
    import subprocess
    import types
    seen = {}

    def fake_run(cmd, env=None, **kw):
        seen["cmd"], seen["tune"], seen["forced"] = cmd, env.get("DP_MSM_TUNE"), "DP_MSM_AFFINE" in env
        return types.SimpleNamespace(returncode=0, stderr="", stdout='some warning\n{"plain_ms": 23.5, "affine_ms": 22.6, "levels": 2, "equal": 1}\n')

    monkeypatch.setenv("DP_MSM_AFFINE", "0")
    monkeypatch.setattr(subprocess, "run", fake_run)
    res = tune.probe(3, 5, 8, 22)
    assert res["levels"] == 2 and res["equal"] == 1 and "probe_seconds" in res and tune.choose(res) == 2
    assert seen["cmd"][-4:] == ["3", "5", "8", "22"] and seen["tune"] == "2" and not seen["forced"]
    monkeypatch.setattr(subprocess, "run", lambda *a, **k: types.SimpleNamespace(returncode=-11, stderr="Segmentation fault", stdout=""))
    assert "error" in tune.probe(0, 0, 1, 22)
    monkeypatch.setattr(subprocess, "run", lambda *a, **k: types.SimpleNamespace(returncode=0, stderr="", stdout="no json here"))
    assert "error" in tune.probe(0, 0, 1, 22)


def test_msm_dev_batch(orc, ctx):
    """dp_msm_dev_batch: three MSMs in flight (tails on their own stream), device pointers in/out
    (host memory under the emulator)"""
    bases = orc.gen_bases(5, 600, 64, True)
    scs = [np.ascontiguousarray(orc.gen_fr(70 + k, 600, False)) for k in range(3)]
    outs = [np.zeros(144, dtype=np.uint8) for _ in range(3)]
    ranges = [(0, 600), (100, 400), (0, 0)]
    ctx.msm_dev_batch([(lo, hi, scs[k].ctypes.data, hi - lo, outs[k].ctypes.data) for k, (lo, hi) in enumerate(ranges)])
    for k, (lo, hi) in enumerate(ranges):
        common.assert_point_eq(orc, outs[k], orc.msm(bases[lo:hi], scs[k][: hi - lo]), f"batch job {k}")
    # host-buffer variant (dp_msm_batch): copy-in of job k+1 under the kernels of job k
    outs = ctx.msm_batch([(lo, hi, scs[k], hi - lo) for k, (lo, hi) in enumerate(ranges)])
    for k, (lo, hi) in enumerate(ranges):
        common.assert_point_eq(orc, outs[k], orc.msm(bases[lo:hi], scs[k][: hi - lo]), f"host batch job {k}")


def test_msm_edges(orc, ctx):
    bases = orc.gen_bases(5, 600, 64, True)
    sc = orc.gen_fr(9, 600, False)
    ident = orc.normalize(ctx.msm(10, 10, sc[:0]))
    assert ident[96] == 1                                                       # empty range -> identity
    common.assert_point_eq(orc, ctx.msm(100, 333, sc[:233]), orc.msm(bases[100:333], sc[:233]), "sub-range")
    common.assert_point_eq(orc, ctx.msm(0, 600, sc[:50]), orc.msm(bases[:50], sc[:50]), "truncate to scalars")
    common.assert_point_eq(orc, ctx.msm(0, 40, sc), orc.msm(bases[:40], sc[:40]), "truncate to bases")
    common.assert_point_eq(orc, ctx.msm(3, 4, sc[:1]), orc.msm(bases[3:4], sc[:1]), "infinity base")
    with pytest.raises(DpError) as e:
        ctx.msm(0, 601, sc)
    assert e.value.code == -1
    # P + (-P) and repeated points (doubling path): bases 0 and 64 are the same point
    s2 = np.zeros((65, 4), dtype=np.uint64)
    s2[0] = common.u256(5)
    s2[64] = common.u256(common.R_MOD - 5)
    assert orc.normalize(ctx.msm(0, 65, s2))[96] == 1
    s2[64] = common.u256(5)
    common.assert_point_eq(orc, ctx.msm(0, 65, s2), orc.msm(bases[:65], s2), "same point twice")


def test_commit_and_round1(orc, ctx):
    bases = orc.gen_bases(5, 600, 64, True)
    co = orc.gen_fr(21, 300, True)
    common.assert_point_eq(orc, ctx.commit(co), orc.commit(bases, co), "commit_polynomial")
    # round1 (worker.rs:383-408) with injected blinders
    n = 1 << 6
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


def test_perm_product(orc, ctx):
    for n, t in ((1, 1), (2, 5), (7, 5), (256, 5), (1025, 3), (2500, 5)):
        common.check_perm_product(orc, ctx, n, t, 300 + n)
    # zero denominator -> DP_E_ARG (the reference's `a / b` panics)
    n = 8
    w = np.stack([orc.gen_fr(1, n)])
    sg = np.stack([orc.gen_fr(2, n)])
    beta = orc.gen_fr(3, 1)[0]
    L = orc.lib()
    t = np.zeros(4, dtype=np.uint64)
    L.orc_fr_mul(beta.ctypes.data, sg[0, 3].ctypes.data, t.ctypes.data)       # beta * sigma[3]
    gamma = np.zeros(4, dtype=np.uint64)
    zero = np.zeros(4, dtype=np.uint64)
    L.orc_fr_add(w[0, 3].ctypes.data, t.ctypes.data, t.ctypes.data)           # w + beta*sigma
    L.orc_fr_sub(zero.ctypes.data, t.ctypes.data, gamma.ctypes.data)          # gamma = -(w + beta*sigma)
    with pytest.raises(DpError) as e:
        ctx.perm_product(w, w, sg, beta, gamma)
    assert e.value.code == -1


def test_rounds_3_to_5(orc, ctx):
    common.check_quotient(orc, ctx, 1 << 6, 1 << 9, 900)
    common.check_poly_ops(orc, ctx, (1, 2, 7, 8, 9, 255, 2047, 2048, 2049, 6145, 10000), 910)
    # the zero polynomial and the constant polynomial
    pt = orc.gen_fr(3, 1)[0]
    assert not ctx.poly_eval(np.zeros((0, 4), dtype=np.uint64), pt).any()
    q, rem = ctx.poly_div_linear(orc.gen_fr(4, 1), pt)
    assert q.shape[0] == 0 and np.array_equal(rem, orc.gen_fr(4, 1)[0])
    with pytest.raises(DpError) as e:
        ctx.poly_lincomb([orc.gen_fr(1, 4)] * 33, orc.gen_fr(2, 33))
    assert e.value.code == -1


def test_satisfied_circuit_divides_exactly(orc, ctx):
    common.check_satisfied_circuit(orc, ctx, 6, 1200)


def test_published_vector(orc, emul_lib):
    common.check_published_vector(orc, lambda: Context(emul_lib, 0, 0, 1))


def test_compressed_srs_ingest(orc, emul_lib):
    common.check_compressed_init(orc, lambda: Context(emul_lib, 0, 0, 1), 40, 1300)


def test_async_msm(orc, emul_lib):
    bases = orc.gen_bases(5, 600, 64, True)
    c = Context(emul_lib, 0, 0, 1)
    c.init(bases, 1 << 6, 1 << 9)
    common.check_async_msm(orc, c, bases, 300, 1500)
    c.msm_submit(7, 0, 100, orc.gen_fr(1, 100, False))      # a job still pending at re-init / close is dropped
    c.init(bases, 1 << 6, 1 << 9)
    with pytest.raises(DpError):
        c.msm_collect(7)
    c.msm_submit(8, 0, 100, orc.gen_fr(1, 100, False))
    c.close()


def test_resident_rounds(orc, emul_lib):
    bases = orc.gen_bases(5, 80, 64, True)
    c = Context(emul_lib, 0, 0, 1)
    c.init(bases, 1 << 6, 1 << 9)
    common.check_resident_rounds(orc, c, bases, 6, 1800)
    c.close()


def test_host_schedules(orc, emul_lib):
    bases = orc.gen_bases(5, 200, 64, True)
    c = Context(emul_lib, 0, 0, 1)
    c.init(bases, 1 << 6, 1 << 9)
    common.check_schedule(orc, c, bases, 6, 9, 1600)
    c.close()


def test_kzg_opening_identity(orc, emul_lib):
    common.check_kzg_opening(orc, lambda: Context(emul_lib, 0, 0, 1), 100, 1400)


def test_golden_rounds(orc, emul_lib):
    def make(n, m):
        c = Context(emul_lib, 0, 0, 1)
        c.init(orc.gen_bases(5, 40, 8, False), n, m)
        return c
    common.check_golden_rounds(make)


def test_rounds_quotient_domain_ratios(orc, emul_lib):
    """quotient / gate domain ratios other than 8, and a quotient domain smaller than one block"""
    for n, m in ((4, 32), (8, 16), (16, 16), (2, 32)):
        c = Context(emul_lib, 0, 0, 1)
        c.init(orc.gen_bases(5, 40, 8, False), n, m)
        common.check_quotient(orc, c, n, m, 950 + n + m)
        c.close()


def test_quotient_inverse_table_cache(orc, emul_lib):
    """the cached 1/(x_i - 1) table (quotient_kernel<true>) serves repeated proofs on one domain and is rebuilt
    when dp_init changes the quotient domain under the same context"""
    c = Context(emul_lib, 0, 0, 1)
    bases = orc.gen_bases(5, 40, 8, False)
    for n, m, seeds in ((4, 32, (961, 962, 963)), (64, 512, (964, 965)), (8, 16, (966,)), (4, 32, (967,))):
        c.init(bases, n, m)
        for seed in seeds:
            common.check_quotient(orc, c, n, m, seed)
    c.close()


def test_error_behaviour(orc, emul_lib):
    c = Context(emul_lib, 0, 0, 1)
    with pytest.raises(DpError) as e:
        c.msm(0, 0, np.zeros((0, 4), dtype=np.uint64))
    assert e.value.code == -2                      # before init
    c.init(np.zeros(0, dtype=np.uint8), 1 << 4, 1 << 7)
    wl = [(0, 4, 0, 4)]
    c.fft_init(1, wl, False, False, False)
    c.fft1(1, 0, np.ones((4, 4), dtype=np.uint64))
    c.fft_init(1, wl, False, False, False)         # same id again: the task is replaced (fft_tasks.insert, worker.rs:215)
    c.fft1(1, 0, np.zeros((3, 4), dtype=np.uint64))       # a short row is zero-extended (fft_in_place resizes)
    with pytest.raises(DpError) as e:
        c.fft1(1, 4, np.zeros((4, 4), dtype=np.uint64))   # row index outside the worker's range
    assert e.value.code == -1
    with pytest.raises(DpError) as e:
        c.fft2_prepare(1)                          # rows missing
    assert e.value.code == -2
    with pytest.raises(DpError):
        c.fft2(7, 4, 4)                            # unknown task
    with pytest.raises(DpError):
        c.fft_init(2, [(0, 3, 0, 4)], False, False, False)   # not the equal split
    c.close()
    with pytest.raises(DpError):
        Context(emul_lib, 0, 2, 2)                 # me >= n_workers


def test_round1_default_blinders_come_from_the_os(orc, emul_lib):
    """blind = NULL: the two blinders are drawn from getrandom(2) (the reference uses ThreadRng, worker.rs:400):
    different between calls and between contexts, canonical residues, and the commitment still opens to
    the blinded polynomial"""
    bases = orc.gen_bases(5, 80, 64, True)
    n = 1 << 6
    evals = orc.gen_fr(22, n, True)
    poly = orc.fft(evals, True, False)
    L = orc.lib()
    seen = set()
    for _ in range(2):
        c = Context(emul_lib, 0, 0, 1)
        c.init(bases, n, 1 << 9)
        for _ in range(2):
            got = c.round1(evals, None)
            wire = c.get_wire()
            b = wire[n:n + 2]
            for k in range(2):
                v = sum(int(b[k, i]) << (64 * i) for i in range(4))
                assert v < common.R_MOD
                seen.add(v)
                t = np.zeros(4, dtype=np.uint64)
                L.orc_fr_sub(poly[k].ctypes.data, b[k].ctypes.data, t.ctypes.data)
                assert np.array_equal(wire[k], t)
            common.assert_point_eq(orc, got, orc.commit(bases, wire), "round1 commitment (library blinders)")
        c.close()
    assert len(seen) == 8                                      # 2 contexts x 2 calls x 2 blinders, all distinct


def test_fused_exchange_many_slots(orc, emul_lib):
    """an arena of five receive slots: five tasks between fft2_prepare and fft2 (the reference dispatcher keeps up to 26
    in flight), collected out of order, a sixth refused"""
    W, L, S = 2, 9, 5
    workers = [PlonkSlave(emul_lib, p, W) for p in range(W)]
    for w in workers:
        w.init([b""], 1 << 6, 1 << L)
    common.attach_in_process(workers, S * (1 << L) * 32 // W)
    wl = disp.fft_workloads(L, W)
    xs = [orc.gen_fr(740 + t, 1 << L) for t in range(S + 1)]
    for t in range(S + 1):
        rows = disp.dispatcher_rows(xs[t], L)
        for p, w in enumerate(workers):
            w.fft_init(t, wl, True, True, True)
            w.ctx.fft1_rows(t, 0, np.ascontiguousarray(rows[wl[p][0]:wl[p][1]]), wl[p][1] - wl[p][0])
    for t in range(S):
        for w in workers:
            w.fft2_prepare(t)
    for w in workers:
        with pytest.raises(DpError) as e:
            w.fft2_prepare(S)
        assert e.value.code == -2
    for t in (3, 0, 4, 1, 2):
        got = disp.assemble(np.concatenate([w.fft2_array(t) for w in workers], axis=0))
        assert np.array_equal(got, orc.fft(xs[t], True, True)), f"task {t}"
    for w in workers:
        w.fft2_prepare(S)
    got = disp.assemble(np.concatenate([w.fft2_array(S) for w in workers], axis=0))
    assert np.array_equal(got, orc.fft(xs[S], True, True))
    for w in workers:
        w.close()


def test_fused_exchange_in_flight_limit(orc, emul_lib):
    """two receive slots per arena: a third fft2_prepare before any fft2 is refused (DP_E_STATE) without
    consuming a slot; after an fft2 it goes through and every task still yields the right transform"""
    W, L = 2, 9
    workers = [PlonkSlave(emul_lib, p, W) for p in range(W)]
    for w in workers:
        w.init([b""], 1 << 6, 1 << L)
    common.attach_in_process(workers, 2 * (1 << L) * 32 // W)
    wl = disp.fft_workloads(L, W)
    xs = [orc.gen_fr(700 + t, 1 << L) for t in range(3)]
    for t in range(3):
        rows = disp.dispatcher_rows(xs[t], L)
        for p, w in enumerate(workers):
            w.fft_init(t, wl, True, False, True)
            for j in range(wl[p][1] - wl[p][0]):
                w.fft1(t, j, chunks(rows[wl[p][0] + j]))
    for t in range(2):
        for w in workers:
            w.fft2_prepare(t)
    for w in workers:
        with pytest.raises(DpError) as e:
            w.fft2_prepare(2)
        assert e.value.code == -2
    outs = {}
    outs[0] = [w.fft2_array(0) for w in workers]              # frees slot 0 on every worker
    for w in workers:
        w.fft2_prepare(2)
    for t in (1, 2):
        outs[t] = [w.fft2_array(t) for w in workers]
    for t in range(3):
        got = disp.assemble(np.concatenate(outs[t], axis=0))
        assert np.array_equal(got, orc.fft(xs[t], False, True)), f"task {t}"
    for w in workers:
        w.close()


def test_exchange_end_twice_and_in_place_division_are_refused(orc, emul_lib):
    W, L = 2, 9
    workers = [PlonkSlave(emul_lib, p, W) for p in range(W)]
    for w in workers:
        w.init([b""], 1 << 6, 1 << L)
        w.ctx.debug_set_limits(3, 2, 0)                        # two-pass column plan: works in place on recv
    x = orc.gen_fr(720, 1 << L)
    wl = disp.fft_workloads(L, W)
    rows = disp.dispatcher_rows(x, L)
    for p, w in enumerate(workers):
        w.fft_init(5, wl, True, False, False)
        for j in range(wl[p][1] - wl[p][0]):
            w.fft1(5, j, chunks(rows[wl[p][0] + j]))
    for dst, src, n in common.local_exchange([w.ctx for w in workers], 5):
        host_copy(dst, src, n)
    for w in workers:
        with pytest.raises(DpError) as e:
            w.ctx.fft_exchange_end(5)                          # the column phase was queued once already
        assert e.value.code == -2
    got = disp.assemble(np.concatenate([w.fft2_array(5) for w in workers], axis=0))
    assert np.array_equal(got, orc.fft(x, False, False))
    # dp_poly_div_linear_dev: quotient over the dividend -> DP_E_ARG
    c = workers[0].ctx
    p = np.ascontiguousarray(orc.gen_fr(721, 5000))
    pt = orc.gen_fr(722, 1)[0]
    with pytest.raises(DpError) as e:
        c.poly_div_linear(p.ctypes.data, pt, 5000, p.ctypes.data)
    assert e.value.code == -1
    for w in workers:
        w.close()


class _HostBuf:
    """device-pointer stand-in under the emulator: host memory"""

    def __init__(self, a):
        self.a = np.ascontiguousarray(a)
        self.ptr = self.a.ctypes.data

    def read(self):
        return self.a


@pytest.mark.parametrize("W,limits", [(1, (11, 9)), (1, (3, 2)), (2, (3, 2)), (2, (11, 9))])
def test_short_rows_skip_the_zero_stages(orc, emul_lib, W, limits):
    """n coefficients on the 8n domain and other cuts: power-of-two row lengths drop butterfly stages (single-pass and
    two-pass row plans), other lengths read a zero-filled tail; through dp_fft1_rows_short and per-row dp_fft1"""
    workers = [PlonkSlave(emul_lib, p, W) for p in range(W)]
    for w in workers:
        w.init([b""], 1 << 6, 1 << 9)
        w.ctx.debug_set_limits(limits[0], limits[1], 0)
    c_q = 1 << 5                                                   # quotient domain 2^9 = 16 x 32
    for k, row_len in enumerate((c_q // 8, 1, c_q // 2, 3, c_q // 8 + 1, c_q)):
        common.check_short_rows(orc, workers, 9, True, row_len, 40 + k, host_copy)
    common.check_short_rows(orc, workers, 6, False, 2, 50, host_copy)
    common.check_short_rows(orc, workers, 9, True, c_q // 8, 51, host_copy, per_row=True)
    if W == 1:
        for valid in (c_q // 8, 1, 5, c_q):
            common.check_dev_valid_cols_hint(orc, workers[0].ctx, 9, True, valid, 60 + valid, _HostBuf)
    for w in workers:
        w.close()


@pytest.mark.parametrize("limits,domains", [((3, 2), (6, 5)), ((3, 3), (6, 5)), ((2, 2), (4, 6))])
def test_whole_ntt_fused_coset_tables_and_short_inputs(orc, emul_lib, limits, domains):
    """whole-domain transforms of the two domains given to dp_init: coset scaling from the two factor tables (2- and
    3-pass splits), inputs shorter than the domain (the first pass drops the zero-input stages), resident padded form"""
    c = Context(emul_lib, 0, 0, 1)
    c.debug_set_limits(limits[0], limits[1], 0)       # before init: the factor tables follow the pass split
    c.init(np.zeros(0, dtype=np.uint8), 1 << domains[0], 1 << domains[1])
    for log_n in domains:
        N = 1 << log_n
        for n_in in (None, max(1, N // 8), 3, 1, N // 2 + 1):
            common.check_whole_ntt(orc, c, log_n, 70 + log_n, n_in=n_in)
        # dp_ntt_dev_padded on a "device" buffer (host memory under the emulator)
        n_in = max(1, N // 8)
        x = orc.gen_fr(90 + log_n, n_in)
        for inv, cos in common.FLAG_COMBOS:
            buf = np.zeros((N, 4), dtype=np.uint64)
            buf[:n_in] = x
            ref = orc.fft(buf, inv, cos)
            c.ntt_dev_padded(buf.ctypes.data, n_in, log_n, inv, cos)
            assert np.array_equal(buf, ref), f"ntt_dev_padded log_n={log_n} inv={inv} coset={cos}"
    common.check_whole_ntt(orc, c, 3, 77)              # a size that is neither domain: scaling kernel fallback
    c.close()


@pytest.mark.parametrize("logn,logq", [(6, 9), (8, 11)] if FULL else [(6, 9)])
def test_single_worker_three_pass_plan(orc, emul_lib, logn, logq):
    """n_workers == 1: the transform as three passes over digit groups of the element index instead of two row and two
    column passes (plan_single_worker3) - same rows in, same columns out, every flag combination, full and short rows,
    the fft1 / fft2 task path and the resident dp_fft_dev path; domains too small for the split fall back"""
    w = PlonkSlave(emul_lib, 0, 1)
    w.init([b""], 1 << logn, 1 << logq)
    w.ctx.debug_set_three_pass(9)
    common.check_distributed_fft(orc, [w], logq, True, 21, host_copy)
    common.check_distributed_fft(orc, [w], logn, False, 22, host_copy)
    common.check_distributed_fft(orc, [w], logq, True, 23, host_copy, n_in=(1 << logq) // 8)
    c_q = (1 << logq) >> (logq >> 1)
    for k, row_len in enumerate((c_q // 8, 1, 3, c_q // 2 + 1)):
        common.check_short_rows(orc, [w], logq, True, row_len, 30 + k)
    for valid in (c_q // 8, 5, c_q):
        common.check_dev_valid_cols_hint(orc, w.ctx, logq, True, valid, 40 + valid, _HostBuf)
    w.ctx.debug_set_three_pass(0)                          # off: the 2-D plan gives the same bytes
    common.check_distributed_fft(orc, [w], logq, True, 24, host_copy)
    w.close()
