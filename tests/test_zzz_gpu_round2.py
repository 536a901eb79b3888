"""GPU parity of what round 2 added: rows handed in short (the zero-input butterfly stages dropped), whole-domain
transforms with the fused coset factor tables and padded inputs, the resident prover (rounds 1-5 on device
polynomials), library-drawn blinders, the asynchronous device-barrier transform stream (>= 2 GPUs)."""
import numpy as np
import pytest
import torch

from distributed_plonk_b200._binding import Context
from distributed_plonk_b200.worker import PlonkSlave
from tests import common
from tests.test_gpu_parity import device_copy

pytestmark = pytest.mark.gpu


class _DevBuf:
    def __init__(self, a):
        self.t = torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
        self.ptr = self.t.data_ptr()

    def read(self):
        torch.cuda.synchronize()
        return self.t.cpu().numpy().view(np.uint64)


@pytest.mark.parametrize("logn,logq", [(12, 15), (20, 23)])
def test_short_rows(orc, gpu_lib, logn, logq):
    """single-pass (c <= 2^11) and two-pass (c = 2^12) row plans; n coefficients on the 8n domain and other cuts"""
    w = PlonkSlave(gpu_lib, 0, 1)
    w.init([b""], 1 << logn, 1 << logq)
    c_q = (1 << logq) >> (logq >> 1)
    lens = (c_q // 8, 1, 3, c_q // 8 + 1) if logq <= 15 else (c_q // 8, c_q // 2 + 5)
    for k, row_len in enumerate(lens):
        common.check_short_rows(orc, [w], logq, True, row_len, 400 + k)
    common.check_short_rows(orc, [w], logn, False, 2, 410)
    if logq <= 15:
        common.check_short_rows(orc, [w], logq, True, c_q // 8, 411, per_row=True)
        for valid in (c_q // 8, 5, c_q):
            common.check_dev_valid_cols_hint(orc, w.ctx, logq, True, valid, 420 + valid, _DevBuf)
    else:
        common.check_dev_valid_cols_hint(orc, w.ctx, logq, True, c_q // 8, 430, _DevBuf)
    w.close()


@pytest.mark.parametrize("logn,logq", [(12, 15), (20, 23)])
def test_whole_ntt_coset_tables_and_padded_inputs(orc, gpu_lib, logn, logq):
    c = Context(gpu_lib, 0, 0, 1)
    c.init(np.zeros(0, dtype=np.uint8), 1 << logn, 1 << logq)
    for log_n in (logn, logq):
        N = 1 << log_n
        for n_in in ((None, N // 8, 3, N // 2 + 1) if log_n <= 15 else (None, N // 8)):
            common.check_whole_ntt(orc, c, log_n, 500 + log_n, n_in=n_in)
        n_in = N // 8
        x = orc.gen_fr(520 + log_n, n_in)
        for inv, cos in common.FLAG_COMBOS:
            buf = np.zeros((N, 4), dtype=np.uint64)
            buf[:n_in] = x
            ref = orc.fft(buf, inv, cos)
            d = _DevBuf(buf)
            c.ntt_dev_padded(d.ptr, n_in, log_n, inv, cos)
            assert np.array_equal(d.read(), ref), f"ntt_dev_padded log_n={log_n} inv={inv} coset={cos}"
    c.close()


@pytest.mark.parametrize("log_n", [6, 10])
def test_resident_prover(orc, gpu_lib, log_n):
    from tests.test_resident import check_resident_prover
    n = 1 << log_n
    bases = orc.gen_bases(5, n + 32, 64, True)
    c = Context(gpu_lib, 0, 0, 1)
    c.init(bases, n, 8 * n)
    check_resident_prover(orc, c, bases, log_n, 2300 + log_n, "cuda")
    c.close()


def test_round1_library_blinders(orc, gpu_lib):
    bases = orc.gen_bases(5, 80, 64, True)
    n = 1 << 6
    evals = orc.gen_fr(22, n, True)
    seen = set()
    for _ in range(2):
        c = Context(gpu_lib, 0, 0, 1)
        c.init(bases, n, 1 << 9)
        got = c.round1(evals, None)
        wire = c.get_wire()
        common.assert_point_eq(orc, got, orc.commit(bases, wire), "round1 commitment (library blinders)")
        for k in range(2):
            v = sum(int(wire[n + k, i]) << (64 * i) for i in range(4))
            assert 0 < v < common.R_MOD
            seen.add(v)
        c.close()
    assert len(seen) == 4


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 4, 8])
def test_device_barrier_async_stream(orc, tmp_path, world):
    """dp_fft_dev_p2p_async: six transforms queued back to back on every rank with no host synchronisation in
    between (the two arena slots are recycled behind the device-side barriers), then the in-flight limit of the
    fft2_prepare form of the fused exchange"""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    import distributed_plonk_b200 as dp
    from tests.test_distributed_cpu import _free_port, _worker
    dp.load()
    orc.build()
    mp.spawn(_worker, args=(world, _free_port(), dp.library_path(), str(tmp_path), "nccl", "p2p_async"), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"rank{r}.txt").read_text() == "ok"
