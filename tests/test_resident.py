"""distributed_plonk_b200/resident.py (rounds 1-5 on worker-resident polynomials) against the oracle's
sequential restatement of src/dispatcher2.rs:294-690, every commitment and every evaluation.  CPU: the kernel-logic
emulator with torch CPU tensors as the "device" buffers; the same check runs on the GPU from tests/test_zz_gpu_rounds.py."""
import numpy as np
import pytest
import torch

from distributed_plonk_b200._binding import Context
from distributed_plonk_b200.resident import N_SEL, N_WIRE, NumpyField, ResidentProver
from tests import common


def check_resident_prover(orc, ctx, bases, log_n, seed, device):
    n, m = 1 << log_n, 8 << log_n
    F = NumpyField(log_n)
    pr = ResidentProver(ctx, torch, log_n, device, F)
    sel = [orc.gen_fr(seed + i, n) for i in range(N_SEL)]
    sig = [orc.gen_fr(seed + 20 + i, n) for i in range(N_WIRE)]
    sig_ev = [orc.fft(s, False, False) for s in sig]
    id_ev = [orc.gen_fr(seed + 30 + i, n) for i in range(N_WIRE)]
    k = np.stack([F.from_u64(v) for v in (1, 7, 13, 17, 23)])
    pr.load_key(sel, sig, sig_ev, id_ev, k)
    wires = [orc.gen_fr(seed + 40 + i, n) for i in range(N_WIRE)]
    pub = orc.gen_fr(seed + 50, n)
    ch = {name: orc.gen_fr(seed + 60 + j, 1)[0] for j, name in enumerate(("beta", "gamma", "alpha", "zeta", "v"))}
    w_host = torch.as_tensor(np.concatenate(wires).view(np.int64))
    p_host = torch.as_tensor(pub.view(np.int64))
    if device != "cpu":
        w_host, p_host = w_host.pin_memory(), p_host.pin_memory()
    for rep in range(2 if device != "cpu" else 1):               # twice on hardware: nothing of proof k may leak into proof k+1
        com, ev = pr.prove(w_host, p_host, ch)
    # ---- the dispatcher's sequential computation (oracle)
    want_com, pad = [], lambda c: np.concatenate([c, np.zeros((m - c.shape[0], 4), dtype=np.uint64)])
    w_coef = [orc.fft(w, True, False) for w in wires]
    want_com += [orc.commit(bases, c) for c in w_coef]
    z_ev = orc.perm_product(np.stack(wires), np.stack(id_ev), np.stack(sig_ev), ch["beta"], ch["gamma"])
    z = orc.fft(z_ev, True, False)
    want_com.append(orc.commit(bases, z))
    pub_coef = orc.fft(pub, True, False)
    cos = [orc.fft(pad(c), False, True) for c in sel + sig + w_coef + [z, pub_coef]]
    q_ev = orc.quotient_evals(np.stack(cos[:13]), np.stack(cos[13:18]), np.stack(cos[18:23]), cos[23], cos[24], k, ch["alpha"], ch["beta"], ch["gamma"], n)
    quot = orc.fft(q_ev, True, True)
    chunk = n + 2
    chunks = [quot[j * chunk:(j + 1) * chunk] for j in range(N_WIRE)]
    want_com += [orc.commit(bases, c) for c in chunks]
    zeta, zeta_w = ch["zeta"], F.mul(ch["zeta"], F.omega)
    w_ev = [orc.poly_eval(c, zeta) for c in w_coef]
    s_ev = [orc.poly_eval(c, zeta) for c in sig[:-1]]
    z_next = orc.poly_eval(z, zeta_w)
    for got, want in zip(ev, w_ev + s_ev + [z_next]):
        assert np.array_equal(got, want), "round-4 evaluation"
    # round 5 with Python integers (dispatcher2.rs:557-633)
    D, E = F._dec, F._enc
    a, b, c, d, e = (D(x) for x in w_ev)
    al, be, ga, ze, v = (D(ch[x]) for x in ("alpha", "beta", "gamma", "zeta", "v"))
    R = F.R_MOD
    vanish = (pow(ze, n, R) - 1) % R
    lag1 = vanish * pow(n * (ze - 1) % R, -1, R) % R
    cz = al
    for wv, kk in zip((a, b, c, d, e), (1, 7, 13, 17, 23)):
        cz = cz * (wv + be * kk * ze + ga) % R
    cz = (cz + al * al * lag1) % R
    cs = al * be * D(z_next) % R
    for wv, sv in zip((a, b, c, d), (D(x) for x in s_ev)):
        cs = cs * (wv + be * sv + ga) % R
    zn2 = (vanish + 1) * ze * ze % R
    coeffs = [a, b, c, d, a * b, c * d, pow(a, 5, R), pow(b, 5, R), pow(c, 5, R), pow(d, 5, R), -e, 1, a * b * c * d * e, cz, -cs]
    coeffs += [-vanish * pow(zn2, j, R) for j in range(N_WIRE)]
    lin = orc.poly_lincomb(sel + [z, sig[-1]] + chunks, np.stack([E(x) for x in coeffs]), chunk)
    batch = orc.poly_lincomb([lin] + w_coef + sig[:-1], np.stack([E(pow(v, j, R)) for j in range(2 * N_WIRE)]), chunk)
    want_com.append(orc.commit(bases, orc.poly_div_linear(batch, zeta)))
    want_com.append(orc.commit(bases, orc.poly_div_linear(z, zeta_w)))
    assert len(com) == len(want_com) == 13
    for j, (got, want) in enumerate(zip(com, want_com)):
        common.assert_point_eq(orc, got, want, f"commitment {j} of the resident proof")


def test_resident_prover_matches_the_dispatchers_arithmetic(orc, emul_lib):
    bases = orc.gen_bases(5, 80, 64, True)
    c = Context(emul_lib, 0, 0, 1)
    c.init(bases, 1 << 6, 1 << 9)
    check_resident_prover(orc, c, bases, 6, 2100, "cpu")
    c.close()
