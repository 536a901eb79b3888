"""Writes tests/golden/golden_rounds_v1.npz: seeded inputs and outputs of the round 2-5 arithmetic
(permutation product, quotient evaluations, evaluate, division by X - z, linear combination;
src/dispatcher2.rs:329-345, 363-504, 535-690).

Produced by the tier-1 C oracle; every value is re-derived by the tier-0 Python-integer restatement
(oracle/py/bls12_381.py, by definition) before being written.  Run from the repo root:
    python tests/golden/make_golden_rounds.py
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
    seed, n, m = 20260923, 8, 64
    out = dict(seed=seed, n=n, m=m)
    sel = np.stack([L.gen_fr(seed + i, m) for i in range(13)])
    sig = np.stack([L.gen_fr(seed + 20 + i, m) for i in range(5)])
    w = np.stack([L.gen_fr(seed + 30 + i, m) for i in range(5)])
    z, pi, k = L.gen_fr(seed + 40, m), L.gen_fr(seed + 41, m), L.gen_fr(seed + 42, 5)
    al, be, ga = (L.gen_fr(seed + 43 + i, 1)[0] for i in range(3))
    q = L.quotient_evals(sel, sig, w, z, pi, k, al, be, ga, n)
    assert ints(q) == B.quotient_evals([ints(v) for v in sel], [ints(v) for v in sig], [ints(v) for v in w], ints(z), ints(pi),
                                       ints(k), ints(al[None])[0], ints(be[None])[0], ints(ga[None])[0], n)
    out.update(q_sel=sel, q_sig=sig, q_w=w, q_z=z, q_pi=pi, q_k=k, q_alpha=al, q_beta=be, q_gamma=ga, q_out=q)
    # permutation product over n rows
    pw, pid, psg = (np.stack([L.gen_fr(seed + 100 + 10 * j + i, n) for i in range(5)]) for j in range(3))
    pz = L.perm_product(pw, pid, psg, be, ga)
    assert ints(pz) == B.perm_product([ints(v) for v in pw], [ints(v) for v in pid], [ints(v) for v in psg], ints(be[None])[0], ints(ga[None])[0])
    out.update(p_w=pw, p_id=pid, p_sigma=psg, p_out=pz)
    # evaluate / divide / combine
    c, pt = L.gen_fr(seed + 200, 37), L.gen_fr(seed + 201, 1)[0]
    ev, dv = L.poly_eval(c, pt), L.poly_div_linear(c, pt)
    assert ints(ev[None]) == [B.poly_eval(ints(c), ints(pt[None])[0])]
    assert ints(dv) == B.poly_div_linear(ints(c), ints(pt[None])[0])
    out.update(e_coeffs=c, e_point=pt, e_out=ev, d_out=dv)
    polys = [L.gen_fr(seed + 210 + i, ln) for i, ln in enumerate((37, 5, 12))]
    cf = L.gen_fr(seed + 220, 3)
    lc = L.poly_lincomb(polys, cf)
    assert ints(lc) == B.poly_lincomb([ints(p) for p in polys], ints(cf))
    out.update(l_p0=polys[0], l_p1=polys[1], l_p2=polys[2], l_coeffs=cf, l_out=lc)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_rounds_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
