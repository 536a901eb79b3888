"""Writes tests/golden/golden_v1.npz: seeded inputs and oracle outputs for the hot path.

The reference (Rust / arkworks) holds no golden vectors and cannot run here, so these are produced
by the tier-1 C oracle and every value is re-derived by the independent tier-0 Python-integer
oracle before being written (O(N^2) DFT, double-and-add).  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader as L  # noqa: E402
from oracle.py import bls12_381 as B  # noqa: E402


def ints(a):
    return B.fr_vec_from_bytes(np.ascontiguousarray(a).tobytes())


def main():
    out = {}
    seed, n = 20260922, 64
    x = L.gen_fr(seed, n)
    out.update(ntt_seed=seed, ntt_n=n, ntt_in=x)
    d = B.Domain(n)
    for inv, cos, f in ((0, 0, B.fft), (1, 0, B.ifft), (0, 1, B.coset_fft), (1, 1, B.coset_ifft)):
        y = L.fft(x, bool(inv), bool(cos))
        assert ints(y) == f(d, ints(x))
        out[f"ntt_out_{inv}{cos}"] = y
    mseed, mn = 381, 96
    bases = L.gen_bases(mseed, mn, 16, True)
    sc = L.gen_fr(mseed + 1, mn, False)
    sc[::5] = 0
    sc[1::7] = [1, 0, 0, 0]
    aff = L.normalize(L.msm(bases, sc))
    pts = [B.g1_affine_from_bytes(bases[i].tobytes()) for i in range(mn)]
    ss = [int.from_bytes(sc[i].tobytes(), "little") for i in range(mn)]
    assert B.g1_affine_from_bytes(aff.tobytes()) == B.msm_naive(pts, ss)
    out.update(msm_seed=mseed, msm_n=mn, msm_bases=bases, msm_scalars=sc, msm_out_affine=aff)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
