"""Pins the oracle: public constants, tier-0 (Python ints, by definition) vs tier-1 (C, ark-faithful),
and the committed golden vectors.  CPU only."""
import os

import numpy as np

from oracle.py import bls12_381 as B

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz")


def ints(a):
    return B.fr_vec_from_bytes(np.ascontiguousarray(a).tobytes())


def test_public_constants():
    # BLS12-381 parameters as published (and as hard-coded in ark-bls12-381 0.3.0)
    assert B.FR_MOD.bit_length() == 255 and B.FQ_MOD.bit_length() == 381
    x = -0xD201000000010000                      # the BLS parameter
    assert B.FR_MOD == x**4 - x**2 + 1
    assert B.FQ_MOD == (x - 1) ** 2 * B.FR_MOD // 3 + x
    assert B.g1_is_on_curve(B.G1_GEN)
    assert B.g1_mul(B.G1_GEN, B.FR_MOD) is None   # generator has order r
    assert (B.FR_MOD - 1) % (1 << 32) == 0 and ((B.FR_MOD - 1) >> 32) % 2 == 1
    assert B.FR_TWO_ADIC_ROOT == 0x16A2A19EDFE81F20D09B681922C813B4B63683508C2280B93829971F439F0D2B
    assert pow(B.FR_TWO_ADIC_ROOT, 1 << 31, B.FR_MOD) == B.FR_MOD - 1
    # ark's Montgomery constants
    assert B.FR_R.to_bytes(32, "little") == bytes.fromhex(
        "feffffff01000000024803 00fab78458f54fbcecef4f8c996f05c5ac59b12418".replace(" ", ""))
    assert pow(7, (B.FR_MOD - 1) // 2, B.FR_MOD) == B.FR_MOD - 1   # 7 is a non-residue


def test_field_ops_c_vs_python(orc):
    rng = np.random.default_rng(1)
    L = orc.lib()
    for name, mod, nl, R in (("fr", B.FR_MOD, 4, B.FR_R), ("fq", B.FQ_MOD, 6, B.FQ_R)):
        vals = [int.from_bytes(rng.bytes(nl * 8), "little") % mod for _ in range(300)]
        vals[:5] = [0, 1, mod - 1, mod - 2, 2]
        rinv = pow(R, -1, mod)
        for a, b in zip(vals, vals[::-1]):
            A = np.frombuffer(a.to_bytes(nl * 8, "little"), dtype=np.uint64).copy()
            Bb = np.frombuffer(b.to_bytes(nl * 8, "little"), dtype=np.uint64).copy()
            O = np.zeros(nl, dtype=np.uint64)
            for op, f in (("mul", a * b * rinv % mod), ("add", (a + b) % mod), ("sub", (a - b) % mod)):
                getattr(L, f"orc_{name}_{op}")(A.ctypes.data, Bb.ctypes.data, O.ctypes.data)
                assert int.from_bytes(O.tobytes(), "little") == f


def test_ntt_tier1_vs_definition(orc):
    for log_n in (0, 1, 2, 5, 7):
        n = 1 << log_n
        x = orc.gen_fr(100 + log_n, n)
        xi, d = ints(x), B.Domain(n)
        for inv, cos, f in ((0, 0, B.fft), (1, 0, B.ifft), (0, 1, B.coset_fft), (1, 1, B.coset_ifft)):
            assert ints(orc.fft(x, bool(inv), bool(cos))) == f(d, xi)


def test_2d_pipeline_equals_plain_ntt(orc):
    """playground.rs:82-103 restated: the 2-D decomposition == Radix2EvaluationDomain, incl. padding"""
    for log_n in (4, 7, 9):
        n = 1 << log_n
        x = orc.gen_fr(7 + log_n, n)
        for inv in (False, True):
            for cos in (False, True):
                ref = orc.fft(x, inv, cos)
                for W in (1, 2, 4):
                    for as_written in (True, False):
                        assert np.array_equal(orc.distributed_fft(x, n, inv, cos, W, as_written), ref)
    # tier-0 model of the exchange indexing agrees too
    x = orc.gen_fr(3, 64)
    for inv in (False, True):
        for cos in (False, True):
            assert B.distributed_fft(B.Domain(64), ints(x), inv, cos, 4) == ints(orc.fft(x, inv, cos))
    # zero-padding invariance + round trip (playground.rs:100-102)
    short = orc.gen_fr(5, 32)
    pad = np.zeros((64, 4), dtype=np.uint64)
    pad[:32] = short
    assert np.array_equal(orc.distributed_fft(short, 64, False, True), orc.fft(pad, False, True))
    assert np.array_equal(orc.fft(orc.fft(pad, False, True), True, True), pad)


def test_fft_helpers_vs_definition(orc):
    d = B.Domain(1 << 7)
    r, c = d.split()
    for inv in (False, True):
        for cos in (False, True):
            v = orc.gen_fr(11, c)
            for i in (0, 3, r - 1):
                assert ints(orc.fft1_helper(v, i, cos, inv, d.size, True)) == B.fft1_helper(ints(v), i, cos, inv, d)
            v = orc.gen_fr(12, r)
            for i in (0, 5, c - 1):
                assert ints(orc.fft2_helper(v, i, cos, inv, d.size, True)) == B.fft2_helper(ints(v), i, cos, inv, d)


def test_msm_tier1_vs_double_and_add(orc):
    n = 48
    bases = orc.gen_bases(5, n, 16, True)
    pts = [B.g1_affine_from_bytes(bases[i].tobytes()) for i in range(n)]
    assert pts[3] is None and all(B.g1_is_on_curve(p) for p in pts)
    assert pts[0] == pts[16]                                   # tiled by doubling
    sc = orc.gen_fr(7, n, False)
    sc[1] = 0
    sc[2] = [1, 0, 0, 0]
    sc[4] = np.frombuffer((B.FR_MOD - 1).to_bytes(32, "little"), dtype=np.uint64)
    got = B.g1_affine_from_bytes(orc.normalize(orc.msm(bases, sc)).tobytes())
    ss = [int.from_bytes(sc[i].tobytes(), "little") for i in range(n)]
    assert got == B.msm_naive(pts, ss)
    # ark window rule
    assert [int(orc.lib().orc_msm_window_c(k)) for k in (1, 31, 32, 1 << 12, (1 << 20) + 32, (1 << 22) + 32)] == [3, 3, 5, 10, 16, 17]
    # into_repr / commit
    co = orc.gen_fr(9, 20, True)
    rep = orc.into_repr(co)
    assert [int.from_bytes(rep[i].tobytes(), "little") for i in range(20)] == ints(co)
    got = B.g1_affine_from_bytes(orc.normalize(orc.commit(bases, co)).tobytes())
    assert got == B.msm_naive(pts, ints(co))


def test_golden_vectors(orc):
    """fixtures written by tests/golden/make_golden.py (tier-1 outputs, spot-verified by tier-0 there)"""
    g = np.load(GOLDEN)
    x = orc.gen_fr(int(g["ntt_seed"]), int(g["ntt_n"]))
    assert np.array_equal(x, g["ntt_in"])
    for inv in (0, 1):
        for cos in (0, 1):
            assert np.array_equal(orc.fft(x, bool(inv), bool(cos)), g[f"ntt_out_{inv}{cos}"])
    bases = orc.gen_bases(int(g["msm_seed"]), int(g["msm_n"]), 16, True)
    assert np.array_equal(bases, g["msm_bases"])
    assert np.array_equal(orc.normalize(orc.msm(bases, g["msm_scalars"])), g["msm_out_affine"])
