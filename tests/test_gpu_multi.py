"""N > 1 on real GPUs (runs only when the box exposes >= 2 devices): the same two-rank scenario as
tests/test_distributed_cpu.py, over NCCL / NVLink with the real library."""
import pytest
import torch
import torch.multiprocessing as mp

from tests.test_distributed_cpu import _free_port, _worker

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_fft_and_msm(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, box has {torch.cuda.device_count()}")
    import distributed_plonk_b200 as dp
    from oracle import loader
    loader.build()
    dp.load()
    mp.spawn(_worker, args=(world, _free_port(), dp.library_path(), str(tmp_path), "nccl"), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"rank{r}.txt").read_text() == "ok"
