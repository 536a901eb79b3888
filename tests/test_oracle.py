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


# Published known-answer vector: 2*G1 of BLS12-381 (EIP-2537 test vector "bls_g1add_(g1+g1=2*g1)", also the
# doubled generator in the zkcrypto/bls12_381 test-suite).  The only absolute value for this path that exists
# outside the (vector-less) reference; everything else is pinned by definition-level recomputation.
TWO_G1 = (0x0572CBEA904D67468808C8EB50A9450C9721DB309128012543902D0AC358A62AE28F75BB8F1C7C42C39A8C5529BF0F4E,
          0x166A9D8CABC673A322FDA673779D8E3822BA3ECB8670E461F73BB9021D5FD76A4C56D9D4CD16BD1BBA86881979749D28)


# EIP-2537 test vector "bls_g1add_(g1+p1)": the second test point of that suite and its sum with the generator
P1_EIP2537 = (0x112B98340EEE2777CC3C14163DEA3EC97977AC3DC5C70DA32E6E87578F44912E902CCEF9EFE28D4A78B8999DFBCA9426,
              0x186B28D92356C4DFEC4B5201AD099DBDEDE3781F8998DDF929B4CD7756192185CA7B8F4EF7088F813270AC3D48868A21)
G1_PLUS_P1 = (0x0A40300CE2DEC9888B60690E9A41D3004FDA4886854573974FAB73B046D3147BA5B7A5BDE85279FFEDE1B45B3918D82D,
              0x06D3D887E9F53B9EC4EB6CEDF5607226754B07C01ACE7834F57F3E7315FAEFB739E59018E22C492006190FBA4A870025)


def test_published_addition_vector(orc):
    assert B.g1_is_on_curve(P1_EIP2537) and B.g1_mul(P1_EIP2537, B.FR_MOD) is None
    assert B.g1_add(B.G1_GEN, P1_EIP2537) == G1_PLUS_P1
    bases = np.stack([np.frombuffer(B.g1_affine_to_bytes(pt), dtype=np.uint8) for pt in (B.G1_GEN, P1_EIP2537)])
    one = np.array([1, 0, 0, 0], dtype=np.uint64)
    assert B.g1_affine_from_bytes(orc.normalize(orc.msm(bases, np.stack([one, one]))).tobytes()) == G1_PLUS_P1


def test_published_doubling_vector(orc):
    assert B.g1_add(B.G1_GEN, B.G1_GEN) == TWO_G1 and B.g1_mul(B.G1_GEN, 2) == TWO_G1
    gen = np.zeros(104, dtype=np.uint8)
    orc.lib().orc_g1_generator(gen.ctypes.data)
    two = np.array([2, 0, 0, 0], dtype=np.uint64)
    assert B.g1_affine_from_bytes(orc.g1_mul(gen, two).tobytes()) == TWO_G1
    # through the Pippenger restatement as well: msm([G, G], [1, 1]) and msm([G], [2])
    one = np.array([1, 0, 0, 0], dtype=np.uint64)
    assert B.g1_affine_from_bytes(orc.normalize(orc.msm(np.stack([gen, gen]), np.stack([one, one]))).tobytes()) == TWO_G1
    assert B.g1_affine_from_bytes(orc.normalize(orc.msm(gen[None], two[None])).tobytes()) == TWO_G1


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


def _mont(vals):
    """list of ints -> [len, 4] uint64 Montgomery Fr"""
    return np.frombuffer(B.fr_vec_to_bytes(vals), dtype=np.uint64).reshape(-1, 4).copy()


def satisfying_circuit(orc, n: int, seed: int):
    """A random TurboPlonk instance whose gates and copy constraints HOLD (GATE_WIDTH 4, 5 wire types):
    random a..d and selectors, q_o = 1, e = the gate output; a non-trivial wire permutation between
    equal-valued cells.  Returns polynomials (coefficient form, ints) and the challenges."""
    g = B.SplitMix64(seed)
    dom, P = B.Domain(n), B.FR_MOD
    w = [[g.fr() for _ in range(n)] for _ in range(4)]
    # copy constraints: cell (i, j) == cell (i2, j2) for a few pairs; sigma swaps them
    cells = [(i, j) for i in range(5) for j in range(n)]
    pairs = [((0, 1), (2, 3)), ((1, 0), (3, n - 2)), ((0, 2), (0, 5))]
    for (i1, j1), (i2, j2) in pairs:
        w[i2][j2] = w[i1][j1]
    sel = [[g.fr() for _ in range(n)] for _ in range(13)]
    sel[10] = [1] * n                                    # q_o
    sel[11] = [0] * n                                    # q_c (public input = 0)
    e = []
    for j in range(n):
        a, b, c, d = (w[i][j] for i in range(4))
        e.append((sel[0][j] * a + sel[1][j] * b + sel[2][j] * c + sel[3][j] * d + sel[4][j] * a * b + sel[5][j] * c * d
                  + sel[6][j] * a**5 + sel[7][j] * b**5 + sel[8][j] * c**5 + sel[9][j] * d**5) % P)
    sel[12] = [0] * n                                    # q_ecc would make e implicit; keep it off
    w.append(e)
    # permutation: extended identity k_i * omega^j; sigma = identity with the paired cells swapped
    k = [1] + [g.fr() for _ in range(4)]
    ident = [[k[i] * pow(dom.group_gen, j, P) % P for j in range(n)] for i in range(5)]
    sigma = [row[:] for row in ident]
    for (i1, j1), (i2, j2) in pairs:
        sigma[i1][j1], sigma[i2][j2] = sigma[i2][j2], sigma[i1][j1]
    beta, gamma, alpha = g.fr(), g.fr(), g.fr()
    return dict(n=n, w=w, sel=sel, k=k, ident=ident, sigma=sigma, beta=beta, gamma=gamma, alpha=alpha, cells=cells)


def test_rounds_tier1_vs_definition(orc):
    """quotient evaluations, evaluate, linear combination and division by (X - z): C oracle == Python ints"""
    g = B.SplitMix64(77)
    n, m = 4, 32
    sel = [[g.fr() for _ in range(m)] for _ in range(13)]
    sig = [[g.fr() for _ in range(m)] for _ in range(5)]
    w = [[g.fr() for _ in range(m)] for _ in range(5)]
    z, pi, k = [g.fr() for _ in range(m)], [g.fr() for _ in range(m)], [g.fr() for _ in range(5)]
    al, be, ga = g.fr(), g.fr(), g.fr()
    got = orc.quotient_evals(np.stack([_mont(v) for v in sel]), np.stack([_mont(v) for v in sig]), np.stack([_mont(v) for v in w]),
                             _mont(z), _mont(pi), _mont(k), _mont([al]), _mont([be]), _mont([ga]), n)
    assert ints(got) == B.quotient_evals(sel, sig, w, z, pi, k, al, be, ga, n)
    for ln in (1, 2, 3, 17, 64):
        c, pt = [g.fr() for _ in range(ln)], g.fr()
        assert ints(orc.poly_eval(_mont(c), _mont([pt]))[None]) == [B.poly_eval(c, pt)]
        assert ints(orc.poly_div_linear(_mont(c), _mont([pt]))) == B.poly_div_linear(c, pt)
    polys = [[g.fr() for _ in range(ln)] for ln in (5, 9, 1, 9, 3)]
    cf = [g.fr() for _ in polys]
    assert ints(orc.poly_lincomb([_mont(p) for p in polys], _mont(cf))) == B.poly_lincomb(polys, cf)


def test_quotient_of_a_satisfied_circuit_is_a_polynomial(orc):
    """The algebraic end-to-end check the reference gets from its verifier (test_plonk): for a witness that
    satisfies every gate and copy constraint, the round-3 evaluations interpolate to a polynomial of degree
    <= 5(n+1)+2 - with no blinding even lower - i.e. the division by Z_H is exact (dispatcher2.rs:506-517)."""
    n = 16
    m = 8 * n
    c = satisfying_circuit(orc, n, 5)
    P = B.FR_MOD
    z = B.perm_product(c["w"], c["ident"], c["sigma"], c["beta"], c["gamma"])
    assert ints(orc.perm_product(np.stack([_mont(v) for v in c["w"]]), np.stack([_mont(v) for v in c["ident"]]),
                                 np.stack([_mont(v) for v in c["sigma"]]), _mont([c["beta"]]), _mont([c["gamma"]]))) == z
    # the grand product closes: z[n-1] * a[n-1] / b[n-1] == 1
    a = b = 1
    for i in range(5):
        a = a * (c["w"][i][n - 1] + c["gamma"] + c["beta"] * c["ident"][i][n - 1]) % P
        b = b * (c["w"][i][n - 1] + c["gamma"] + c["beta"] * c["sigma"][i][n - 1]) % P
    assert z[n - 1] * a % P == b

    def to_coset(evals):     # evaluations over H -> coefficients -> evaluations over g*H_m (lines 386-388)
        coeffs = orc.fft(_mont(evals), True, False)
        pad = np.zeros((m, 4), dtype=np.uint64)
        pad[:n] = coeffs
        return orc.fft(pad, False, True)

    sel = np.stack([to_coset(v) for v in c["sel"]])
    sig = np.stack([to_coset(v) for v in c["sigma"]])
    w = np.stack([to_coset(v) for v in c["w"]])
    q = orc.quotient_evals(sel, sig, w, to_coset(z), to_coset([0] * n), _mont(c["k"]), _mont([c["alpha"]]), _mont([c["beta"]]),
                           _mont([c["gamma"]]), n)
    coeffs = ints(orc.fft(q, True, True))
    deg = max(j for j, v in enumerate(coeffs) if v)
    # unblinded degrees: gate 5(n-1)+ (n-1) ... bounded by the permutation term 6(n-1) - n
    assert deg <= 5 * (n + 1) + 2 and deg >= 4 * n, deg
    # and the same evaluations with ONE wire value corrupted do not divide: high coefficients appear
    c["w"][0][3] = (c["w"][0][3] + 1) % P
    w_bad = np.stack([to_coset(v) for v in c["w"]])
    q_bad = orc.quotient_evals(sel, sig, w_bad, to_coset(z), to_coset([0] * n), _mont(c["k"]), _mont([c["alpha"]]),
                               _mont([c["beta"]]), _mont([c["gamma"]]), n)
    assert max(j for j, v in enumerate(ints(orc.fft(q_bad, True, True))) if v) > 7 * n


def test_golden_rounds(orc):
    from tests import common
    common.check_golden_rounds(None, impl=orc)


def test_point_encoding_tier1_vs_tier0(orc):
    """ark-serialize compressed G1 ("next" row 4): C oracle == Python ints; round trip; error cases"""
    bases = orc.gen_bases(11, 24, 24, True)
    comp = orc.g1_compress(bases)
    for k in range(bases.shape[0]):
        pt = B.g1_affine_from_bytes(bases[k].tobytes())
        assert comp[k].tobytes() == B.g1_compress(pt)
        assert B.g1_decompress(comp[k].tobytes()) == pt
    back, rcs = orc.g1_decompress(comp)
    assert rcs == [0] * len(rcs) and np.array_equal(back, bases)
    # the generator: y is the smaller root, so the encoding is x little-endian with no flag bits
    gen = np.zeros(104, dtype=np.uint8)
    orc.lib().orc_g1_generator(gen.ctypes.data)
    assert orc.g1_compress(gen[None])[0].tobytes() == B.G1_GEN[0].to_bytes(48, "little")
    neg = B.g1_compress(B.g1_neg(B.G1_GEN))
    assert neg[47] >> 7 == 1 and neg[:47] == B.G1_GEN[0].to_bytes(48, "little")[:47]
    # errors: x >= p, both flags, x with no point, a curve point outside the subgroup
    bad_x = np.frombuffer((B.FQ_MOD + 1).to_bytes(48, "little"), dtype=np.uint8)
    both = comp[0].copy()
    both[47] |= 0xC0
    outside = orc.g1_point_outside_subgroup()
    assert orc.g1_decompress(np.stack([bad_x, both, outside]))[1] == [-1, -2, -4]
    assert orc.g1_decompress(outside[None], check_subgroup=False)[1] == [0]
    no_point = next(x for x in range(1, 50) if pow((x**3 + 4) % B.FQ_MOD, (B.FQ_MOD - 1) // 2, B.FQ_MOD) != 1)
    assert orc.g1_decompress(np.frombuffer(no_point.to_bytes(48, "little"), dtype=np.uint8)[None])[1] == [-3]
    for raw, exc in ((bad_x, "canonical"), (both, "flags"), (outside, "subgroup")):
        try:
            B.g1_decompress(raw.tobytes())
            assert False
        except ValueError as e:
            assert exc in str(e)
