"""The batched-affine MSM levels (DP_MSM_AFFINE / DP_MSM_TUNE) on the GPU.  Last file of the suite on purpose: these
kernels were written after the round's GPU budget was spent (their inner loops ran on hardware in microbenchmark form
and, as dp_init's default tuning at 2^20 and 2^22 points, with the round's last GPU seconds: profiles/r02i_msm_tuning.txt),
so nothing else in the suite runs after the forced-level cases and the wider search exercised here."""
import numpy as np
import pytest

from distributed_plonk_b200._binding import Context
from tests import common

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bases(orc):
    return orc.gen_bases(5, (1 << 16) + 32, 2048, True)


@pytest.mark.parametrize("levels", [2, 1, 3])
def test_forced_levels_vs_oracle(orc, gpu_lib, bases, monkeypatch, levels):
    monkeypatch.setenv("DP_MSM_AFFINE", str(levels))
    monkeypatch.setenv("DP_MSM_AFFINE_MIN", "0")
    c = Context(gpu_lib, 0, 0, 1)
    c.init(bases, 1 << 12, 1 << 15)
    assert c.msm_tuning()["levels"] == levels
    for n in (1, 33, 1000, (1 << 12) + 32, (1 << 16) + 32) if levels == 2 else (1000, (1 << 16) + 32):
        common.check_msm(orc, c, bases, n, 4000 + n)          # uniform, witness-like, all r-1, all zero, all one
    for cbits in (5, 12, 17):                                  # per-window bucket sets
        c.debug_set_limits(11, 9, cbits)
        common.check_msm(orc, c, bases, 3000, 4100 + cbits, which=("uniform", "witness-like"))
    c.debug_set_limits(11, 9, 0)
    s2 = np.zeros((2049, 4), dtype=np.uint64)                  # bases 0 and 2048 are the same point (2048 distinct, tiled)
    s2[0] = common.u256(5)
    s2[2048] = common.u256(common.R_MOD - 5)
    assert orc.normalize(c.msm(0, 2049, s2))[96] == 1          # P + (-P)
    s2[2048] = common.u256(5)
    common.assert_point_eq(orc, c.msm(0, 2049, s2), orc.msm(bases[:2049], s2), "same point twice")
    sc = orc.gen_fr(4200, 1 << 16, False)
    outs = c.msm_batch([(0, 1 << 16, sc, 1 << 16), (100, 40000, sc, 39900), (0, 0, sc, 0)])
    for k, (lo, hi) in enumerate([(0, 1 << 16), (100, 40000), (0, 0)]):
        common.assert_point_eq(orc, outs[k], orc.msm(bases[lo:hi], sc[: hi - lo]), f"batch job {k}")
    c.close()


def test_tuning_at_init_agrees(orc, gpu_lib, monkeypatch):
    """dp_init times the plain pipeline and 1, 2, 3 tree levels over the context's own table (DP_MSM_TUNE=2: the wider
    search; the default compares plain and two levels); all must give the same 144 bytes"""
    monkeypatch.delenv("DP_MSM_AFFINE", raising=False)
    monkeypatch.setenv("DP_MSM_TUNE", "2")
    n = (1 << 20) + 32
    c = Context(gpu_lib, 0, 0, 1)
    b = c.gen_bases(77, n)
    c.init(b, 1 << 20, 1 << 23)
    t = c.msm_tuning()
    print("msm tuning at 2^20:", t)
    assert t["equal"] == 1, f"the two MSM pipelines disagree: {t}"
    assert t["plain_ms"] > 0 and t["affine_ms"] > 0 and t["levels"] in (0, 1, 2, 3)
    sc = orc.gen_fr(4300, n, False)
    sc[::5] = 0
    common.assert_point_eq(orc, c.msm(0, n, sc), orc.msm(b, sc), "2^20 MSM through the tuned pipeline")
    c.close()


def test_probe_in_a_child_process(gpu_lib):
    from distributed_plonk_b200 import tune
    res = tune.probe(0, 0, 1, 18)
    print("probe at 2^18:", res)
    assert "error" not in res and res["equal"] in (1, -1)
    assert tune.choose(res) in (0, 1, 2, 3)
