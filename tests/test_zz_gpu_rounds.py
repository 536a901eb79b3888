"""GPU parity of the "next" row SURVEY.md §8(f)-1: rounds 3-5 of Prover::prove (quotient evaluations,
evaluation at a point, linear combinations, division by X - z) through the C ABI against the
oracle's sequential restatement of src/dispatcher2.rs:363-690, plus the size-independent property
(a satisfied circuit's quotient divides exactly).  The file sorts last on purpose: these kernels were
added after the round's GPU budget was spent, so under `pytest -x` they run after every test that
has already been seen green on hardware."""
import numpy as np
import pytest
import torch

from distributed_plonk_b200._binding import Context
from tests import common

pytestmark = pytest.mark.gpu


def dev(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def host(t: torch.Tensor) -> np.ndarray:
    return t.cpu().numpy().view(np.uint64)


@pytest.fixture(scope="module")
def ctx(gpu_lib, orc):
    c = Context(gpu_lib, 0, 0, 1)
    c.init(orc.gen_bases(5, 64, 16, False), 1 << 12, 1 << 15)   # BASELINE config 0 sizes
    yield c
    c.close()


def test_golden_rounds_on_gpu(orc, gpu_lib):
    def make(n, m):
        c = Context(gpu_lib, 0, 0, 1)
        c.init(orc.gen_bases(5, 40, 8, False), n, m)
        return c
    common.check_golden_rounds(make)


def test_quotient_evals(orc, ctx):
    common.check_quotient(orc, ctx, 1 << 12, 1 << 15, 2000)


@pytest.mark.parametrize("n,m", [(4, 32), (16, 16), (2, 32), (64, 128), (1 << 10, 1 << 13)])
def test_quotient_domain_shapes(orc, gpu_lib, n, m):
    """ratios 1, 2, 8, 16 and quotient domains smaller than one thread block"""
    c = Context(gpu_lib, 0, 0, 1)
    c.init(orc.gen_bases(5, 40, 8, False), n, m)
    common.check_quotient(orc, c, n, m, 2100 + n + m)
    c.close()


def test_quotient_evals_device_resident(orc, ctx):
    m, n = 1 << 15, 1 << 12
    sel = [orc.gen_fr(2200 + i, m) for i in range(13)]
    sig = [orc.gen_fr(2220 + i, m) for i in range(5)]
    w = [orc.gen_fr(2230 + i, m) for i in range(5)]
    z, pi, k = orc.gen_fr(2240, m), orc.gen_fr(2241, m), orc.gen_fr(2242, 5)
    al, be, ga = (orc.gen_fr(2243 + i, 1)[0] for i in range(3))
    d = {name: [dev(v) for v in arrs] for name, arrs in (("sel", sel), ("sig", sig), ("w", w), ("zp", [z, pi]))}
    out = torch.empty((m, 4), dtype=torch.int64, device="cuda")
    ctx.quotient_evals_dev([t.data_ptr() for t in d["sel"]], [t.data_ptr() for t in d["sig"]], [t.data_ptr() for t in d["w"]],
                           d["zp"][0].data_ptr(), d["zp"][1].data_ptr(), k, al, be, ga, out.data_ptr())
    ref = orc.quotient_evals(np.stack(sel), np.stack(sig), np.stack(w), z, pi, k, al, be, ga, n)
    assert np.array_equal(host(out), ref)


def test_poly_ops(orc, ctx):
    # 1 .. 3 levels of the chunk recursion (2048 coefficients per block): 2^22 + 5 needs three
    common.check_poly_ops(orc, ctx, (1, 2, 7, 8, 9, 255, 2047, 2048, 2049, 6145, 100003, (1 << 22) + 5), 2300)
    pt = orc.gen_fr(3, 1)[0]
    assert not ctx.poly_eval(np.zeros((0, 4), dtype=np.uint64), pt).any()
    q, rem = ctx.poly_div_linear(orc.gen_fr(4, 1), pt)
    assert q.shape[0] == 0 and np.array_equal(rem, orc.gen_fr(4, 1)[0])


def test_poly_ops_device_resident(orc, ctx):
    n = (1 << 16) + 11
    c, pt = orc.gen_fr(2400, n), orc.gen_fr(2401, 1)[0]
    cd = dev(c)
    assert np.array_equal(ctx.poly_eval(cd.data_ptr(), pt, n), orc.poly_eval(c, pt))
    qd = torch.empty((n - 1, 4), dtype=torch.int64, device="cuda")
    _, rem = ctx.poly_div_linear(cd.data_ptr(), pt, n, qd.data_ptr())
    assert np.array_equal(rem, orc.poly_eval(c, pt))
    assert np.array_equal(host(qd), orc.poly_div_linear(c, pt))
    # q(X) (X - z) + rem == p(X) at a fresh point t: the defining identity, through the library only
    t = orc.gen_fr(2402, 1)[0]
    L = orc.lib()
    lhs, tz = np.zeros(4, dtype=np.uint64), np.zeros(4, dtype=np.uint64)
    L.orc_fr_sub(t.ctypes.data, pt.ctypes.data, tz.ctypes.data)
    qt = ctx.poly_eval(qd.data_ptr(), t, n - 1)
    L.orc_fr_mul(qt.ctypes.data, tz.ctypes.data, lhs.ctypes.data)
    L.orc_fr_add(lhs.ctypes.data, rem.ctypes.data, lhs.ctypes.data)
    assert np.array_equal(lhs, ctx.poly_eval(cd.data_ptr(), t, n))
    # linear combination of device-resident polynomials of different lengths
    lens = [n, 5, 70000, 1]
    polys = [orc.gen_fr(2410 + i, ln) for i, ln in enumerate(lens)]
    cf = orc.gen_fr(2420, len(lens))
    pd = [dev(p) for p in polys]
    od = torch.empty((70001, 4), dtype=torch.int64, device="cuda")
    ctx.poly_lincomb([t_.data_ptr() for t_ in pd], cf, out_len=70001, lens=lens, out_ptr=od.data_ptr())
    assert np.array_equal(host(od), orc.poly_lincomb(polys, cf, 70001))


@pytest.mark.parametrize("log_n", [6, 12, 16])
def test_satisfied_circuit_divides_exactly(orc, gpu_lib, log_n):
    """rounds 2-3 end to end on the GPU at growing sizes: exact division by Z_H iff the witness is valid"""
    c = Context(gpu_lib, 0, 0, 1)
    c.init(orc.gen_bases(5, 40, 8, False), 1 << log_n, 8 << log_n)
    common.check_satisfied_circuit(orc, c, log_n, 2500 + log_n)
    c.close()


@pytest.mark.parametrize("n", [40, 5000])
def test_compressed_srs_ingest(orc, gpu_lib, n):
    """"next" row §8(f)-4: ark-serialize compressed points decompressed (and subgroup-checked) on the GPU"""
    common.check_compressed_init(orc, lambda: Context(gpu_lib, 0, 0, 1), n, 2600 + n)


def test_kzg_opening_identity(orc, gpu_lib):
    """round 5 end to end over an SRS with a known trapdoor: (tau - z) commit(q) + p(z) G == commit(p)"""
    common.check_kzg_opening(orc, lambda: Context(gpu_lib, 0, 0, 1), 3000, 2700)


def test_async_msm(orc, gpu_lib):
    """dp_msm_submit / dp_msm_collect (commitments queued behind transforms) == blocking dp_msm"""
    n = 5000
    bases = orc.gen_bases(5, n, 2048, True)
    c = Context(gpu_lib, 0, 0, 1)
    c.init(bases, 1 << 6, 1 << 9)
    common.check_async_msm(orc, c, bases, n, 2800)
    c.close()


def test_host_schedules(orc, gpu_lib):
    """serial and overlapped end-to-end schedules of bench.py (schedule.py) against the oracle, host buffers"""
    bases = orc.gen_bases(5, 3000, 2048, True)
    c = Context(gpu_lib, 0, 0, 1)
    c.init(bases, 1 << 12, 1 << 15)
    common.check_schedule(orc, c, bases, 12, 15, 2900)
    c.close()


@pytest.mark.parametrize("log_n", [6, 12])
def test_resident_rounds(orc, gpu_lib, log_n):
    """rounds 2-5 on worker-resident polynomials (dp_poly_* + *_dev entries + dp_ntt_dev + dp_commit_dev)"""
    bases = orc.gen_bases(5, (1 << log_n) + 32, 2048, True)
    c = Context(gpu_lib, 0, 0, 1)
    c.init(bases, 1 << log_n, 8 << log_n)
    common.check_resident_rounds(orc, c, bases, log_n, 3000 + log_n)
    c.close()


def test_published_vector_on_gpu(orc, gpu_lib):
    """2*G1 as published (EIP-2537 vector): the one absolute value on this path that exists outside the reference"""
    common.check_published_vector(orc, lambda: Context(gpu_lib, 0, 0, 1))


@pytest.mark.parametrize("mode", ["schedule", "schedule_stream_ordered"])
def test_host_schedules_two_ranks_nccl(orc, tmp_path, mode):
    """the serial / overlapped host schedules with the exchange as one NCCL all-to-all per transform (2 GPUs):
    host-synchronised (make_exchange) and enqueued on the compute stream (make_stream_ordered_exchange)"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    import distributed_plonk_b200 as dp
    from tests.test_distributed_cpu import _free_port, _worker
    dp.load()
    orc.build()
    mp.spawn(_worker, args=(2, _free_port(), dp.library_path(), str(tmp_path), "nccl", mode), nprocs=2, join=True)
    for r in range(2):
        assert (tmp_path / f"rank{r}.txt").read_text() == "ok"


@pytest.mark.parametrize("world", [2, 4])
def test_device_side_barrier_transform(orc, tmp_path, world):
    """dp_fft_dev_p2p: rows -> peer stores -> p2p_barrier_kernel -> columns on one stream (no host barrier)"""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    import distributed_plonk_b200 as dp
    from tests.test_distributed_cpu import _free_port, _worker
    dp.load()
    orc.build()
    mp.spawn(_worker, args=(world, _free_port(), dp.library_path(), str(tmp_path), "nccl", "p2p_barrier"), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"rank{r}.txt").read_text() == "ok"


def test_cpp_host_mirror_rounds_on_gpu(orc, tmp_path):
    """the C++ mirror's Promise-style varMsm and round 4-5 bodies against the real library"""
    import distributed_plonk_b200 as dp
    from tests.test_host_mirror import build_cli, run_case
    dp.load()
    exe = build_cli(dp.library_path(), str(tmp_path / "host_mirror_gpu"))
    run_case(orc, exe, tmp_path, (1 << 12) + 32, 12, 15, 0b101, 1 << 12, rounds=True)


def test_fft1_short_and_long_rows(orc, gpu_lib):
    """fft1 rows shorter / longer than c behave like the reference's `fft_in_place` resize"""
    from distributed_plonk_b200.worker import PlonkSlave
    w = PlonkSlave(gpu_lib, 0, 1)
    w.init([b""], 1 << 10, 1 << 13)
    common.check_fft1_row_lengths(orc, w, 13, True, 3100)
    common.check_fft1_row_lengths(orc, w, 10, False, 3101)
    w.close()
