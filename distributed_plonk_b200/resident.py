"""One TurboPlonk proof's polynomial arithmetic with every polynomial RESIDENT on the worker.

What the reference's dispatcher does between its worker calls (src/dispatcher2.rs:192-713) - and ships
every polynomial over the wire for, twice per transform - restated over the worker-resident entries of
the C ABI (the schema already declares round3*/round4*/round5* RPCs for exactly this,
src/hello_world.capnp:26-44, the worker implements none of them).  The witness (wire and public-input
evaluations) goes in once; 13 commitments and 10 evaluations come back:

  round 1  5 x (iNTT(n) -> commit)                                        dispatcher2.rs:294-321
  round 2  grand product -> iNTT(n) -> commit                             329-361
  round 3  25 x coset-NTT(8n) of n coefficients, quotient evaluations,    363-532
           coset-iNTT(8n), 5 commitments of the (n+2)-coefficient chunks
  round 4  10 evaluations at zeta / zeta*omega                            535-555
  round 5  linearisation and batch polynomials, two divisions by          557-690
           (X - point), 2 commitments

Out of scope here, as everywhere in this repo: the Fiat-Shamir transcript (challenges are inputs) and the
blinding scalars (two extra coefficients per wire; `dp_round1` does them for the RPC the worker has).
The proving key (13 selector + 5 sigma polynomials in coefficient form, sigma / identity permutation
evaluations) stays resident across proofs, as `State` keeps the bases (worker.rs:42-59).

Buffers are torch tensors (int64 [count, 4] = raw Fr) on the worker's device - or CPU tensors when the
library under test is the kernel-logic emulator, whose "device" memory is host memory.
"""
from __future__ import annotations

import numpy as np

N_SEL, N_WIRE = 13, 5


class ResidentProver:
    def __init__(self, ctx, torch, log_n: int, device: str, field):
        """field: helpers over raw Montgomery Fr as np.uint64[4] - mul(a,b), add(a,b), sub(a,b), inv(a), from_u64(v),
        pow_u64(a, e), omega (the generator of the n-point domain); tests and the bench pass the oracle's / numpy ones:
        the handful of scalar challenge products of rounds 4-5 are host-side glue, not hot-path work"""
        self.ctx, self.torch, self.F = ctx, torch, field
        self.log_n, self.n, self.m = log_n, 1 << log_n, 8 << log_n
        self.dev = device
        n, m = self.n, self.m

        def buf(count):
            return torch.zeros((count, 4), dtype=torch.int64, device=device)

        self.sel_coef = [buf(n) for _ in range(N_SEL)]
        self.sig_coef = [buf(n) for _ in range(N_WIRE)]
        self.sig_eval = buf(N_WIRE * n)
        self.id_eval = buf(N_WIRE * n)
        self.wire_eval = buf(N_WIRE * n)
        self.pub = buf(n)
        self.wire_coef = [buf(n) for _ in range(N_WIRE)]
        self.z = buf(n)
        self.big = [buf(m) for _ in range(N_SEL + 2 * N_WIRE + 2)]   # coset evaluations: 13 sel, 5 sigma, 5 wires, z, pub
        self.quot = buf(m)
        self.lin = buf(n + 2)
        self.batch = buf(n + 2)
        self.wit = [buf(n + 2), buf(n + 2)]

    # ---- setup: the proving key, once
    def load_key(self, sel_coef, sig_coef, sig_eval, id_eval, k):
        """selector / sigma polynomials in coefficient form ([n,4] each), sigma and identity permutation evaluations
        ([5][n,4]), the coset representatives k[5]"""
        t = self.torch
        for dst, src in zip(self.sel_coef + self.sig_coef, list(sel_coef) + list(sig_coef)):
            dst.copy_(t.as_tensor(np.ascontiguousarray(src).view(np.int64)))
        self.sig_eval.copy_(t.as_tensor(np.concatenate(sig_eval).view(np.int64)))
        self.id_eval.copy_(t.as_tensor(np.concatenate(id_eval).view(np.int64)))
        self.k = np.ascontiguousarray(k, dtype=np.uint64)

    def _sync(self):
        if self.dev != "cpu":
            self.torch.cuda.current_stream().synchronize()

    # ---- one proof
    def prove(self, wire_evals_host, pub_host, ch):
        """wire_evals_host: torch tensor [5n,4] (pinned host memory on a GPU), pub_host [n,4]; ch: dict of the
        challenges beta, gamma, alpha, zeta, v as raw Fr.  Returns (commitments: list of 13 x 144 B, evals: list)"""
        ctx, n, m, log_n, F = self.ctx, self.n, self.m, self.log_n, self.F
        log_m = log_n + 3
        com, P = [], lambda t: t.data_ptr()
        # witness in: the only bulk host->device traffic of the proof
        self.wire_eval.copy_(wire_evals_host, non_blocking=True)
        self.pub.copy_(pub_host, non_blocking=True)
        for i in range(N_WIRE):
            self.wire_coef[i].copy_(self.wire_eval[i * n:(i + 1) * n])
        self._sync()                                           # torch's stream -> the library's streams
        # round 1
        for i in range(N_WIRE):
            ctx.ntt_dev(P(self.wire_coef[i]), log_n, True, False)
            com.append(ctx.commit_dev(P(self.wire_coef[i]), n))
        # round 2
        ctx.perm_product_dev(P(self.wire_eval), P(self.id_eval), P(self.sig_eval), N_WIRE, n, ch["beta"], ch["gamma"], P(self.z))
        ctx.ntt_dev(P(self.z), log_n, True, False)
        com.append(ctx.commit_dev(P(self.z), n))
        ctx.ntt_dev(P(self.pub), log_n, True, False)
        # round 3: 25 coset evaluations on the 8n domain; only the n coefficients at the head of each buffer are read
        srcs = self.sel_coef + self.sig_coef + self.wire_coef + [self.z, self.pub]
        for dst, src in zip(self.big, srcs):
            dst[:n].copy_(src)
        self._sync()
        for dst in self.big:
            ctx.ntt_dev_padded(P(dst), n, log_m, False, True, wait=False)
        b = [P(t) for t in self.big]
        ctx.quotient_evals_dev(b[:13], b[13:18], b[18:23], b[23], b[24], self.k, ch["alpha"], ch["beta"], ch["gamma"], P(self.quot))
        ctx.ntt_dev(P(self.quot), log_m, True, True)
        chunk = n + 2
        for j in range(N_WIRE):
            com.append(ctx.commit_dev(P(self.quot) + 32 * j * chunk, chunk))
        # round 4
        zeta = ch["zeta"]
        zeta_w = F.mul(zeta, F.omega)
        w_ev = [ctx.poly_eval(P(self.wire_coef[i]), zeta, n) for i in range(N_WIRE)]
        s_ev = [ctx.poly_eval(P(self.sig_coef[i]), zeta, n) for i in range(N_WIRE - 1)]
        z_next = ctx.poly_eval(P(self.z), zeta_w, n)
        # round 5: the scalar coefficients are host glue (a few dozen field operations), the polynomials stay put
        a, bb, c, d, e = w_ev
        ab, cd = F.mul(a, bb), F.mul(c, d)
        p5 = lambda x: F.mul(F.mul(F.mul(x, x), F.mul(x, x)), x)
        neg = lambda x: F.sub(F.from_u64(0), x)
        one = F.from_u64(1)
        vanish = F.sub(F.pow_u64(zeta, n), one)
        lag1 = F.mul(vanish, F.inv(F.mul(F.from_u64(n), F.sub(zeta, one))))
        cz = ch["alpha"]
        for wv, kk in zip(w_ev, self.k):
            cz = F.mul(cz, F.add(F.add(wv, F.mul(F.mul(ch["beta"], kk), zeta)), ch["gamma"]))
        cz = F.add(cz, F.mul(F.mul(ch["alpha"], ch["alpha"]), lag1))
        cs = F.mul(F.mul(ch["alpha"], ch["beta"]), z_next)
        for wv, sv in zip(w_ev[:-1], s_ev):
            cs = F.mul(cs, F.add(F.add(wv, F.mul(ch["beta"], sv)), ch["gamma"]))
        cs = neg(cs)
        zn2 = F.mul(F.add(vanish, one), F.mul(zeta, zeta))
        qc, cur = [], neg(vanish)
        for _ in range(N_WIRE):
            qc.append(cur)
            cur = F.mul(cur, zn2)
        coeffs = [a, bb, c, d, ab, cd, p5(a), p5(bb), p5(c), p5(d), neg(e), one, F.mul(F.mul(ab, cd), e), cz, cs] + qc
        polys = [P(t) for t in self.sel_coef] + [P(self.z), P(self.sig_coef[N_WIRE - 1])] + [P(self.quot) + 32 * j * chunk for j in range(N_WIRE)]
        lens = [n] * (N_SEL + 2) + [chunk] * N_WIRE
        ctx.poly_lincomb(polys, np.stack(coeffs), out_len=chunk, lens=lens, out_ptr=P(self.lin))
        vs, cur = [], one
        for _ in range(1 + N_WIRE + N_WIRE - 1):
            vs.append(cur)
            cur = F.mul(cur, ch["v"])
        polys = [P(self.lin)] + [P(t) for t in self.wire_coef] + [P(t) for t in self.sig_coef[:-1]]
        ctx.poly_lincomb(polys, np.stack(vs), out_len=chunk, lens=[chunk] + [n] * (2 * N_WIRE - 1), out_ptr=P(self.batch))
        ctx.poly_div_linear(P(self.batch), zeta, chunk, P(self.wit[0]))
        com.append(ctx.commit_dev(P(self.wit[0]), chunk - 1))
        ctx.poly_div_linear(P(self.z), zeta_w, n, P(self.wit[1]))
        com.append(ctx.commit_dev(P(self.wit[1]), n - 1))
        return com, w_ev + s_ev + [z_next]


class NumpyField:
    """raw Montgomery Fr scalars (np.uint64[4]) through Python integers - the few dozen challenge products of rounds 4-5"""
    R_MOD = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    R = (1 << 256) % R_MOD

    def __init__(self, log_n: int):
        root = pow(7, (self.R_MOD - 1) >> 32, self.R_MOD)
        self.omega = self._enc(pow(root, 1 << (32 - log_n), self.R_MOD))

    @classmethod
    def _dec(cls, a) -> int:
        v = sum(int(a[i]) << (64 * i) for i in range(4))
        return v * pow(cls.R, -1, cls.R_MOD) % cls.R_MOD

    @classmethod
    def _enc(cls, v: int) -> np.ndarray:
        v = v % cls.R_MOD * cls.R % cls.R_MOD
        return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)

    def mul(self, a, b):
        return self._enc(self._dec(a) * self._dec(b))

    def add(self, a, b):
        return self._enc(self._dec(a) + self._dec(b))

    def sub(self, a, b):
        return self._enc(self._dec(a) - self._dec(b))

    def inv(self, a):
        return self._enc(pow(self._dec(a), -1, self.R_MOD))

    def from_u64(self, v):
        return self._enc(v)

    def pow_u64(self, a, e):
        return self._enc(pow(self._dec(a), e, self.R_MOD))


def bench_leg(ctx, torch, log_n: int, rand_fr, timed, steps: int = 3):
    """bench.py's e2e_resident: proofs/s of ResidentProver.prove on synthetic data (random polynomials: the
    arithmetic is data-independent; the quotient is then not a polynomial of the expected degree, which changes
    nothing about the work), witness copied from pinned host memory inside the timed region"""
    if log_n > 22:
        return {"skipped": f"2^{log_n}: 25 resident coset evaluations need {25 * (8 << log_n) * 32 / 2**30:.0f} GiB"}
    n = 1 << log_n
    F = NumpyField(log_n)
    pr = ResidentProver(ctx, torch, log_n, "cuda", F)
    for t in pr.sel_coef + pr.sig_coef:
        t.copy_(rand_fr(n))
    pr.sig_eval.copy_(rand_fr(N_WIRE * n))
    pr.id_eval.copy_(rand_fr(N_WIRE * n))
    pr.k = np.stack([F.from_u64(v) for v in (1, 7, 13, 17, 23)])
    wires = rand_fr(N_WIRE * n).cpu().pin_memory()
    pub = rand_fr(n).cpu().pin_memory()
    ch = {name: F.from_u64(v) for name, v in (("beta", 0xB17A), ("gamma", 0x6A33A), ("alpha", 0xA1FA), ("zeta", 0x2E7A), ("v", 0x55))}
    out = {}

    def step():
        out["r"] = pr.prove(wires, pub, ch)

    dt, _ = timed(step, steps, 1, False)
    com, ev = out["r"]
    return {"value": steps / dt, "unit": "proofs/s", "ms_per_step": dt / steps * 1e3, "steps": steps,
            "h2d_bytes_per_step": int((N_WIRE + 1) * n * 32), "d2h_bytes_per_step": int(len(com) * 144 + len(ev) * 32),
            "what": ("rounds 1-5 of one proof on worker-resident polynomials (distributed_plonk_b200/resident.py): witness in once, "
                     "13 commitments + 10 evaluations out; includes the round-2 grand product, the quotient evaluations, the round-4 "
                     "evaluations and the round-5 folds / divisions that the 33-transform + 13-MSM schedule of `value` leaves to the dispatcher"),
            "timing": "host clock between barrier+synchronize"}
