"""Parity checks shared by the emulator (CPU) and the GPU test files: every check drives the C ABI
exactly like the reference's own tests drive the workers (dispatcher.rs:177-350, dispatcher2.rs:
1088-1216) and compares with the oracle."""
import ctypes as C

import os

import numpy as np

from distributed_plonk_b200 import dispatcher as disp
from distributed_plonk_b200._binding import Context, DpError
from distributed_plonk_b200.worker import PlonkSlave, chunks

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
FLAG_COMBOS = [(inv, cos) for inv in (False, True) for cos in (False, True)]


def u256(v: int) -> np.ndarray:
    return np.frombuffer(int(v).to_bytes(32, "little"), dtype=np.uint64).copy()


def check_whole_ntt(orc, ctx: Context, log_n: int, seed: int, n_in=None):
    N = 1 << log_n
    n_in = N if n_in is None else n_in
    x = orc.gen_fr(seed, n_in)
    padded = np.zeros((N, 4), dtype=np.uint64)
    padded[:n_in] = x
    for inv, cos in FLAG_COMBOS:
        got = ctx.ntt(x, log_n, inv, cos)
        ref = orc.fft(padded, inv, cos)
        assert np.array_equal(got, ref), f"dp_ntt log_n={log_n} inv={inv} coset={cos}"


def local_exchange(ctxs, tid, async_begin=False):
    """in-process all-to-all over the split API (single process holding every worker).  async_begin: the
    stream-ordered form (dp_fft_exchange_begin_async) - the copies here are host-side, so the streams are
    drained explicitly before them, which is exactly the wait the blocking form does internally"""
    W = len(ctxs)
    if async_begin:
        bufs = [c.fft_exchange_begin_async(tid) for c in ctxs]
        for c in ctxs:
            c.sync()
    else:
        bufs = [c.fft_exchange_begin(tid) for c in ctxs]
    import_cuda = None
    for p in range(W):
        for q in range(W):
            n = bufs[p][2] * 32
            yield bufs[q][1] + p * n, bufs[p][0] + q * n, n
    for c in ctxs:
        c.fft_exchange_end(tid)


def attach_in_process(workers, arena_bytes):
    """single process holding every worker: swap the arena handles directly"""
    handles = [w.ctx.peer_arena_create(arena_bytes) for w in workers]
    for p, w in enumerate(workers):
        for q, h in enumerate(handles):
            if q != p:
                w.ctx.peer_attach(q, h)
        assert w.ctx.peer_ready()


def check_distributed_fft(orc, workers, domain_log, is_quot, seed, copy_fn, n_in=None, async_begin=False):
    """test_fft (dispatcher.rs:246-350): all flag combos through fft_init / fft1 / fft2_prepare /
    fft2 must equal Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}."""
    N = 1 << domain_log
    n_in = N if n_in is None else n_in
    W = len(workers)
    for k, (inv, cos) in enumerate(FLAG_COMBOS):
        x = orc.gen_fr(seed + k, n_in)
        padded = np.zeros((N, 4), dtype=np.uint64)
        padded[:n_in] = x
        tid = 0xF00D0000 + seed * 16 + k

        def exchange(send, recv, n, _tid=tid):
            raise AssertionError("single-process test drives the split API directly")

        if W == 1 or workers[0].ctx.peer_ready():
            # one worker, or fused peer-memory exchange: the plain dispatcher sequence is enough
            # (all fft2_prepare calls return before the first fft2 = the barrier)
            got = disp.fft(workers, domain_log, x, is_quot, inv, cos, tid)
        else:
            # drive the stubs by hand so the exchange can be done in-process
            wl = disp.fft_workloads(domain_log, W)
            rows = disp.dispatcher_rows(x, domain_log)
            for w in workers:
                w.fft_init(tid, wl, is_quot, inv, cos)
            for p, w in enumerate(workers):
                for j in range(wl[p][1] - wl[p][0]):
                    w.fft1(tid, j, chunks(rows[wl[p][0] + j]))
            for dst, src, n in local_exchange([w.ctx for w in workers], tid, async_begin):
                copy_fn(dst, src, n)
            cols = np.concatenate([w.fft2_array(tid) for w in workers], axis=0)
            got = disp.assemble(cols)
        ref = orc.fft(padded, inv, cos)
        assert np.array_equal(got, ref), f"distributed fft L={domain_log} W={W} inv={inv} coset={cos}"


def scalar_sets(orc, n, seed):
    """input distributions of SURVEY §8d: (A) uniform, (B) witness-like, (C) all r-1, plus edges"""
    a = orc.gen_fr(seed, n, False)
    b = orc.gen_fr(seed + 1, n, False)
    b[::2] = 0
    b[1::10] = u256(1)
    if n > 7:
        b[5] = u256(R_MOD - 1)
        b[7] = u256(2)
    c = np.tile(u256(R_MOD - 1), (n, 1))
    z = np.zeros((n, 4), dtype=np.uint64)
    o = np.tile(u256(1), (n, 1))
    return {"uniform": a, "witness-like": b, "all r-1": c, "all zero": z, "all one": o}


def assert_point_eq(orc, got144, ref144, what):
    g, r = orc.normalize(got144), orc.normalize(ref144)
    assert np.array_equal(g, r), f"{what}: affine mismatch"
    # the library promises a normalised representative: identical raw bytes after normalisation
    norm = np.zeros(144, dtype=np.uint8)
    norm[:96] = g[:96]
    if g[96]:
        norm[96:144] = 0
    else:
        from oracle.py import bls12_381 as B
        norm[96:144] = np.frombuffer(B.fq_to_mont_bytes(1), dtype=np.uint8)
    assert np.array_equal(np.asarray(got144, dtype=np.uint8), norm), f"{what}: output is not the normalised Jacobian"


def check_msm(orc, ctx: Context, bases, n, seed, which=None):
    for name, sc in scalar_sets(orc, n, seed).items():
        if which and name not in which:
            continue
        got = ctx.msm(0, n, sc)
        ref = orc.msm(bases[:n], sc)
        assert_point_eq(orc, got, ref, f"msm n={n} {name}")


def check_sharded_msm(orc, workers, bases, n, seed):
    """test_msm (dispatcher.rs:177-244): sum of the workers' partials == one multi_scalar_mul"""
    sc = orc.gen_fr(seed, n, False)
    parts = disp.commit_polynomial(workers, n, sc)
    acc = np.frombuffer(parts[0], dtype=np.uint8)
    for p in parts[1:]:
        acc = orc.g1_add(acc, np.frombuffer(p, dtype=np.uint8))
    ref = orc.msm(bases[:n], sc)
    assert np.array_equal(orc.normalize(acc), orc.normalize(ref))


def check_perm_product(orc, ctx: Context, n: int, n_types: int, seed: int):
    """round-2 grand product: GPU scans + one inversion vs the dispatcher's row-by-row division"""
    w = np.stack([orc.gen_fr(seed + i, n) for i in range(n_types)])
    idp = np.stack([orc.gen_fr(seed + 20 + i, n) for i in range(n_types)])
    sg = np.stack([orc.gen_fr(seed + 40 + i, n) for i in range(n_types)])
    beta, gamma = orc.gen_fr(seed + 60, 1)[0], orc.gen_fr(seed + 61, 1)[0]
    got = ctx.perm_product(w, idp, sg, beta, gamma)
    assert np.array_equal(got, orc.perm_product(w, idp, sg, beta, gamma)), f"perm product n={n}"
    # a valid permutation (sigma = id re-ordered) closes the cycle: z[n-1] * a[n-1]/b[n-1] = 1
    return got


def check_quotient(orc, ctx: Context, n: int, m: int, seed: int):
    """round 3: one fused kernel vs the dispatcher's per-point loop (dispatcher2.rs:434-504)"""
    sel = [orc.gen_fr(seed + i, m) for i in range(13)]
    sig = [orc.gen_fr(seed + 20 + i, m) for i in range(5)]
    w = [orc.gen_fr(seed + 30 + i, m) for i in range(5)]
    z, pi = orc.gen_fr(seed + 40, m), orc.gen_fr(seed + 41, m)
    k = orc.gen_fr(seed + 42, 5)
    al, be, ga = (orc.gen_fr(seed + 43 + i, 1)[0] for i in range(3))
    got = ctx.quotient_evals(sel, sig, w, z, pi, k, al, be, ga)
    ref = orc.quotient_evals(np.stack(sel), np.stack(sig), np.stack(w), z, pi, k, al, be, ga, n)
    assert np.array_equal(got, ref), f"quotient evaluations n={n} m={m}"


def check_poly_ops(orc, ctx: Context, sizes, seed: int):
    """rounds 4-5: evaluate, divide by (X - z), linear combination vs the sequential restatement"""
    for n in sizes:
        c, pt = orc.gen_fr(seed + n % 1000, n), orc.gen_fr(seed + 1, 1)[0]
        ev = orc.poly_eval(c, pt)
        assert np.array_equal(ctx.poly_eval(c, pt), ev), f"poly_eval n={n}"
        q, rem = ctx.poly_div_linear(c, pt)
        assert np.array_equal(rem, ev), f"poly_div_linear remainder n={n}"
        assert np.array_equal(q, orc.poly_div_linear(c, pt)), f"poly_div_linear n={n}"
    # special points: 0, 1, -1
    c = orc.gen_fr(seed + 5, 300)
    zero = np.zeros(4, dtype=np.uint64)
    for pt in (zero, _fr_one(orc), _fr_neg_one(orc)):
        assert np.array_equal(ctx.poly_eval(c, pt), orc.poly_eval(c, pt))
        assert np.array_equal(ctx.poly_div_linear(c, pt)[0], orc.poly_div_linear(c, pt))
    lens = [5, 2049, 1, 2049, 300, 0]
    polys = [orc.gen_fr(seed + 70 + i, ln) for i, ln in enumerate(lens)]
    cf = orc.gen_fr(seed + 80, len(lens))
    assert np.array_equal(ctx.poly_lincomb(polys, cf), orc.poly_lincomb(polys, cf))
    assert np.array_equal(ctx.poly_lincomb(polys, cf, out_len=100), orc.poly_lincomb(polys, cf)[:100])
    assert np.array_equal(ctx.poly_lincomb(polys, cf, out_len=3000)[2049:], np.zeros((3000 - 2049, 4), dtype=np.uint64))


def _fr_one(orc):
    """Montgomery 1 = R mod r"""
    a = np.array([1, 0, 0, 0], dtype=np.uint64)
    return orc.from_repr(a[None])[0]


def _fr_neg_one(orc):
    out, zero = np.zeros(4, dtype=np.uint64), np.zeros(4, dtype=np.uint64)
    one = _fr_one(orc)
    orc.lib().orc_fr_sub(zero.ctypes.data, one.ctypes.data, out.ctypes.data)
    return out


def make_satisfied_instance(orc, ctx: Context, log_n: int, seed: int):
    """A random TurboPlonk instance (GATE_WIDTH 4, five wire types) whose gates and copy constraints
    hold, as evaluations over H: wires w[5], selectors sel[13], public input, identity / sigma
    permutation values, challenges, and the round-2 product z (computed by the library)."""
    n = 1 << log_n
    one = _fr_one(orc)
    V = orc.vec_op
    w = [orc.gen_fr(seed + i, n) for i in range(4)]
    # copy constraints between equal-valued cells of different wire types / rows
    pairs = [((0, 1), (2, 3)), ((1, 0), (3, n - 2)), ((0, 2), (0, n // 2 + 1)), ((3, 7 % n), (1, n - 1))]
    for (i1, j1), (i2, j2) in pairs:
        w[i2][j2] = w[i1][j1]
    sel = [orc.gen_fr(seed + 10 + i, n) for i in range(13)]
    pub = np.zeros((n, 4), dtype=np.uint64)
    pub[:3] = orc.gen_fr(seed + 30, 3)                          # a few public inputs
    a, b, c, d = w
    ab, cd = V("mul", a, b), V("mul", c, d)
    p5 = lambda v: V("mul", V("mul", V("mul", v, v), V("mul", v, v)), v)
    rest = V("add", sel[11], pub)
    for q, v in ((sel[0], a), (sel[1], b), (sel[2], c), (sel[3], d), (sel[4], ab), (sel[5], cd),
                 (sel[6], p5(a)), (sel[7], p5(b)), (sel[8], p5(c)), (sel[9], p5(d))):
        rest = V("add", rest, V("mul", q, v))
    # q_c + pi + ... + q_ecc ab cd e - q_o e = 0   =>   e = rest / (q_o - q_ecc ab cd)
    e = V("mul", rest, V("inv", V("sub", sel[10], V("mul", sel[12], V("mul", ab, cd)))))
    w = w + [e]
    k = orc.gen_fr(seed + 40, 5)
    k[0] = one
    x_poly = np.zeros((n, 4), dtype=np.uint64)
    x_poly[1 % n] = one
    omegas = ctx.ntt(x_poly, log_n, False, False)              # omega^j = evaluations of X over H
    ident = [ctx.poly_lincomb([omegas], k[i][None]) for i in range(5)]
    sigma = [v.copy() for v in ident]
    for (i1, j1), (i2, j2) in pairs:
        sigma[i1][j1], sigma[i2][j2] = sigma[i2][j2].copy(), sigma[i1][j1].copy()
    beta, gamma, alpha = (orc.gen_fr(seed + 50 + i, 1)[0] for i in range(3))
    z = ctx.perm_product(np.stack(w), np.stack(ident), np.stack(sigma), beta, gamma)
    return dict(n=n, w=w, sel=sel, pub=pub, k=k, ident=ident, sigma=sigma, beta=beta, gamma=gamma, alpha=alpha, z=z)


def check_satisfied_circuit(orc, ctx: Context, log_n: int, seed: int):
    """Size-independent property of rounds 2-3 (what the reference gets from its verifier, test_plonk):
    for a witness that satisfies every gate and every copy constraint the quotient evaluations
    interpolate to a polynomial of degree <= 5(n+1)+2, i.e. the division by Z_H is exact; with ONE
    wire value corrupted it is not.  The instance is built with the oracle's elementwise vector ops;
    perm product, iNTT, coset NTT, quotient kernel and coset iNTT all run through the library."""
    n, log_m = 1 << log_n, log_n + 3
    c = make_satisfied_instance(orc, ctx, log_n, seed)
    w, sel, sigma, z, pub, k = c["w"], c["sel"], c["sigma"], c["z"], c["pub"], c["k"]
    alpha, beta, gamma = c["alpha"], c["beta"], c["gamma"]

    def to_coset(evals):                                       # dispatcher2.rs:381-432
        return ctx.ntt(ctx.ntt(evals, log_n, True, False), log_m, False, True)

    fixed = ([to_coset(v) for v in sel], [to_coset(v) for v in sigma], to_coset(z), to_coset(pub))

    def degree(wires):
        q = ctx.quotient_evals(fixed[0], fixed[1], [to_coset(v) for v in wires], fixed[2], fixed[3], k, alpha, beta, gamma)
        coeffs = ctx.ntt(q, log_m, True, True)                 # dispatcher2.rs:507
        nz = np.nonzero(coeffs.any(axis=1))[0]
        return int(nz[-1]) if nz.size else -1

    deg = degree(w)
    assert 4 * n <= deg <= 5 * (n + 1) + 2, f"quotient degree {deg} for n={n}"
    bad = [v.copy() for v in w]
    bad[1][n // 3] = orc.gen_fr(seed + 60, 1)[0]
    assert degree(bad) > 7 * n


def check_resident_rounds(orc, ctx: Context, bases, log_n: int, seed: int):
    """Rounds 2-5 with every polynomial RESIDENT on the worker (dp_poly_* store + the *_dev entries +
    dp_ntt_dev + dp_commit_dev): each polynomial crosses PCIe once on the way in; only commitments,
    evaluations and the final witness commitments come back.  Checked against the host-buffer entries /
    the oracle at every step, and by the protocol's own invariants (exact division, opening identity)."""
    n, m, log_m = 1 << log_n, 8 << log_n, log_n + 3
    c = make_satisfied_instance(orc, ctx, log_n, seed)
    alpha, beta, gamma, k = c["alpha"], c["beta"], c["gamma"], c["k"]
    ids = {}

    def put(name, evals):
        ids[name] = len(ids) + 100
        return ctx.poly_put(ids[name], evals, m)

    # round 1-2: evaluations over H in, coefficient form by an in-place iNTT of the first n entries
    names = [f"sel{i}" for i in range(13)] + [f"sig{i}" for i in range(5)] + [f"w{i}" for i in range(5)] + ["z", "pub"]
    evals = c["sel"] + c["sigma"] + c["w"] + [c["z"], c["pub"]]
    ptr = {}
    for name, ev in zip(names, evals):
        ptr[name] = put(name, ev)
        ctx.ntt_dev(ptr[name], log_n, True, False)
    coeff = {name: ctx.poly_get(ids[name], n) for name in ("w0", "w4", "z", "sig1")}
    assert np.array_equal(coeff["w0"], orc.fft(c["w"][0], True, False))
    assert not ctx.poly_get(ids["w0"], m - n, n).any(), "zero padding of a resident polynomial"
    com_w0 = ctx.commit_dev(ptr["w0"], n)                                     # wire commitment from device memory
    assert_point_eq(orc, com_w0, orc.commit(bases, coeff["w0"]), "commit_dev")
    # round 4 evaluations before the buffers are transformed in place
    zeta = orc.gen_fr(seed + 70, 1)[0]
    ev_w4 = ctx.poly_eval(ptr["w4"], zeta, n)
    assert np.array_equal(ev_w4, orc.poly_eval(coeff["w4"], zeta))
    # round 5 witness of z at zeta, straight into another resident polynomial, then committed
    wit = ctx.poly_put(999, np.zeros((1, 4), dtype=np.uint64), n)
    _, rem = ctx.poly_div_linear(ptr["z"], zeta, n, wit)
    assert np.array_equal(rem, orc.poly_eval(coeff["z"], zeta))
    assert np.array_equal(ctx.poly_get(999, n - 1), orc.poly_div_linear(coeff["z"], zeta))
    assert_point_eq(orc, ctx.commit_dev(wit, n - 1), orc.commit(bases, orc.poly_div_linear(coeff["z"], zeta)), "witness commitment")
    # round 3: coset evaluations in place, quotient kernel, coset iNTT in place, split commitments
    for name in names:
        ctx.ntt_dev(ptr[name], log_m, False, True)
    q_ptr = ctx.poly_put(998, np.zeros((1, 4), dtype=np.uint64), m)
    ctx.quotient_evals_dev([ptr[f"sel{i}"] for i in range(13)], [ptr[f"sig{i}"] for i in range(5)], [ptr[f"w{i}"] for i in range(5)],
                           ptr["z"], ptr["pub"], k, alpha, beta, gamma, q_ptr)
    host_q = ctx.quotient_evals([ctx.poly_get(ids[f"sel{i}"]) for i in range(13)], [ctx.poly_get(ids[f"sig{i}"]) for i in range(5)],
                                [ctx.poly_get(ids[f"w{i}"]) for i in range(5)], ctx.poly_get(ids["z"]), ctx.poly_get(ids["pub"]),
                                k, alpha, beta, gamma)
    assert np.array_equal(ctx.poly_get(998), host_q), "resident quotient == host-buffer quotient"
    ctx.ntt_dev(q_ptr, log_m, True, True)
    quot = ctx.poly_get(998)
    nz = np.nonzero(quot.any(axis=1))[0]
    assert 4 * n <= int(nz[-1]) <= 5 * (n + 1) + 2, "exact division by Z_H"
    for j in range(0, int(nz[-1]) + 1, n + 2):                                # split_quot_polys (dispatcher2.rs:509-524)
        ln = min(n + 2, m - j)
        if n + 2 <= bases.shape[0]:
            assert_point_eq(orc, ctx.commit_dev(q_ptr + 32 * j, ln), orc.commit(bases, quot[j:j + ln]), f"split quotient commitment at {j}")
    for pid in list(ids.values()) + [998, 999]:
        ctx.poly_free(pid)
    try:
        ctx.poly_ptr(998)
        raise AssertionError("freed polynomial still known")
    except DpError as e:
        assert e.code == -1


GOLDEN_ROUNDS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_rounds_v1.npz")


def check_golden_rounds(make_ctx, impl=None):
    """tests/golden/golden_rounds_v1.npz (written by make_golden_rounds.py, tier-0 verified there) against
    `impl` (the oracle loader) or against the library (make_ctx(n, m) -> Context)"""
    g = np.load(GOLDEN_ROUNDS)
    n, m = int(g["n"]), int(g["m"])
    if impl is not None:
        assert np.array_equal(impl.quotient_evals(g["q_sel"], g["q_sig"], g["q_w"], g["q_z"], g["q_pi"], g["q_k"], g["q_alpha"],
                                                  g["q_beta"], g["q_gamma"], n), g["q_out"])
        assert np.array_equal(impl.perm_product(g["p_w"], g["p_id"], g["p_sigma"], g["q_beta"], g["q_gamma"]), g["p_out"])
        assert np.array_equal(impl.poly_eval(g["e_coeffs"], g["e_point"]), g["e_out"])
        assert np.array_equal(impl.poly_div_linear(g["e_coeffs"], g["e_point"]), g["d_out"])
        assert np.array_equal(impl.poly_lincomb([g["l_p0"], g["l_p1"], g["l_p2"]], g["l_coeffs"]), g["l_out"])
        return
    ctx = make_ctx(n, m)
    assert np.array_equal(ctx.quotient_evals(list(g["q_sel"]), list(g["q_sig"]), list(g["q_w"]), g["q_z"], g["q_pi"], g["q_k"],
                                             g["q_alpha"], g["q_beta"], g["q_gamma"]), g["q_out"])
    assert np.array_equal(ctx.perm_product(g["p_w"], g["p_id"], g["p_sigma"], g["q_beta"], g["q_gamma"]), g["p_out"])
    assert np.array_equal(ctx.poly_eval(g["e_coeffs"], g["e_point"]), g["e_out"])
    q, rem = ctx.poly_div_linear(g["e_coeffs"], g["e_point"])
    assert np.array_equal(q, g["d_out"]) and np.array_equal(rem, g["e_out"])
    assert np.array_equal(ctx.poly_lincomb([g["l_p0"], g["l_p1"], g["l_p2"]], g["l_coeffs"]), g["l_out"])
    ctx.close()


def check_compressed_init(orc, make_ctx, n: int, seed: int):
    """"next" row 4: SRS ingest from ark-serialize compressed points == ingest of the raw structs"""
    bases = orc.gen_bases(seed, n, min(n, 64), True)
    comp = orc.g1_compress(bases)
    ctx = make_ctx()
    for check in (False, True):
        ctx.init_compressed(comp, 1 << 4, 1 << 7, check)
        assert np.array_equal(ctx.get_bases(0, n), bases), f"decompressed bases differ (check_subgroup={check})"
    sc = orc.gen_fr(seed + 1, n, False)
    assert_point_eq(orc, ctx.msm(0, n, sc), orc.msm(bases, sc), "msm over decompressed bases")
    # rejected encodings name the first bad index and leave the context uninitialised
    bad = comp.copy()
    bad[n // 2] = np.frombuffer((B_FQ_MOD + 1).to_bytes(48, "little"), dtype=np.uint8)
    both = comp.copy()
    both[3, 47] |= 0xC0
    nosq = comp.copy()
    no_point = next(x for x in range(1, 50) if pow((x**3 + 4) % B_FQ_MOD, (B_FQ_MOD - 1) // 2, B_FQ_MOD) != 1)
    nosq[n - 1] = np.frombuffer(no_point.to_bytes(48, "little"), dtype=np.uint8)
    nosq[1] = nosq[n - 1]
    outside = comp.copy()
    outside[5] = orc.g1_point_outside_subgroup()
    for arr, idx, word in ((bad, n // 2, "canonical"), (both, 3, "flag"), (nosq, 1, "square"), (outside, 5, "subgroup")):
        try:
            ctx.init_compressed(arr, 1 << 4, 1 << 7, True)
            raise AssertionError(f"accepted an invalid encoding ({word})")
        except DpError as e:
            assert e.code == -1 and f"point {idx} " in str(e) and word in str(e), str(e)
        try:
            ctx.msm(0, 1, sc[:1])
            raise AssertionError("context usable after a failed init")
        except DpError as e:
            assert e.code == -2
    ctx.init_compressed(outside, 1 << 4, 1 << 7, False)       # the unchecked variant accepts it (deserialize_unchecked)
    ctx.close()


B_FQ_MOD = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB


def check_kzg_opening(orc, make_ctx, n: int, seed: int):
    """Protocol-level property of round 5 (what the reference's verifier checks with a pairing in
    test_plonk): over an SRS with a KNOWN trapdoor tau, commit(p) = p(tau) G, so for the witness
    q = (p - p(z)) / (X - z) computed by the library   (tau - z) commit(q) + p(z) G == commit(p).
    MSM (dp_commit), evaluation and division kernels all take part; the oracle only does group algebra."""
    from oracle.py import bls12_381 as B
    tau = orc.gen_fr(seed, 1, False)[0]
    srs = orc.gen_srs(tau, n)
    ctx = make_ctx()
    ctx.init(srs, 1 << 4, 1 << 7)
    p, z = orc.gen_fr(seed + 1, n), orc.gen_fr(seed + 2, 1)[0]
    q, rem = ctx.poly_div_linear(p, z)
    assert np.array_equal(rem, ctx.poly_eval(p, z))
    c_p, c_q = ctx.commit(p), ctx.commit(q)
    tau_i = int.from_bytes(tau.tobytes(), "little")
    z_i, rem_i = B.fr_from_mont_bytes(z.tobytes()), B.fr_from_mont_bytes(rem.tobytes())
    as_k = lambda v: np.frombuffer((v % B.FR_MOD).to_bytes(32, "little"), dtype=np.uint64)
    gen = np.zeros(104, dtype=np.uint8)
    orc.lib().orc_g1_generator(gen.ctypes.data)
    lhs = orc.g1_add(orc.affine_to_jacobian(orc.g1_mul(orc.normalize(c_q), as_k(tau_i - z_i))),
                     orc.affine_to_jacobian(orc.g1_mul(gen, as_k(rem_i))))
    assert np.array_equal(orc.normalize(lhs), orc.normalize(c_p)), "KZG opening identity"
    # and commit(p) really is p(tau) G
    p_tau = B.fr_from_mont_bytes(ctx.poly_eval(p, orc.from_repr(tau[None])[0]).tobytes())
    assert np.array_equal(orc.g1_mul(gen, as_k(p_tau)), orc.normalize(c_p))
    ctx.close()


def check_async_msm(orc, ctx: Context, bases, n: int, seed: int):
    """dp_msm_submit / dp_msm_collect: several commitments in flight, interleaved with a transform and
    collected out of order, equal the blocking dp_msm; error behaviour of the job table"""
    sets = [orc.gen_fr(seed + k, n - 3 * k, False) for k in range(4)]
    sets[2][::3] = 0
    ranges = [(0, n), (3, n), (0, n - 6), (5, n - 4)]
    for k, (sc, (lo, hi)) in enumerate(zip(sets, ranges)):
        ctx.msm_submit(100 + k, lo, hi, sc)
        if k == 1:   # a whole transform goes through the context while jobs are pending
            x = orc.gen_fr(seed + 9, 64)
            assert np.array_equal(ctx.ntt(x, 6, False, True), orc.fft(x, False, True))
    ctx.msm_submit(200, 7, 7, np.zeros((0, 4), dtype=np.uint64))          # empty range -> identity
    for k in (2, 0, 3, 1):
        sc, (lo, hi) = sets[k], ranges[k]
        m = min(hi - lo, sc.shape[0])
        assert_point_eq(orc, ctx.msm_collect(100 + k), orc.msm(bases[lo:lo + m], sc[:m]), f"async msm job {k}")
    assert np.array_equal(orc.normalize(ctx.msm_collect(200))[96:97], np.array([1], dtype=np.uint8))
    # job table errors
    ctx.msm_submit(1, 0, n, sets[0])
    for call, code in ((lambda: ctx.msm_submit(1, 0, n, sets[0]), -2), (lambda: ctx.msm_collect(999), -2),
                       (lambda: ctx.msm_submit(2, 0, n + 10**6, sets[0]), -1)):
        try:
            call()
            raise AssertionError("expected a DpError")
        except DpError as e:
            assert e.code == code, str(e)
    assert_point_eq(orc, ctx.msm_collect(1), orc.msm(bases[:n], sets[0]), "job 1 after the rejected calls")
    # a scalar >= 2^255 (not a canonical Fr): whatever the blocking call does with it - an error when the
    # top window overflows, the plain 256-bit multiple otherwise - the asynchronous one does too
    bad = sets[0].copy()
    bad[5] = np.uint64(0xFFFFFFFFFFFFFFFF)
    outcome = []
    for call in (lambda: ctx.msm(0, n, bad), lambda: (ctx.msm_submit(3, 0, n, bad), ctx.msm_collect(3))[1]):
        try:
            outcome.append(orc.normalize(call()).tobytes())
        except DpError as e:
            outcome.append(e.code)
    assert outcome[0] == outcome[1], outcome
    # 64 jobs may be pending, the 65th is refused; all of them still complete (empty jobs: the identity)
    empty = np.zeros((0, 4), dtype=np.uint64)
    for k in range(63):
        ctx.msm_submit(1000 + k, 7, 7, empty)
    ctx.msm_submit(1063, 0, 8, sets[0][:8])
    try:
        ctx.msm_submit(2000, 0, 8, sets[0][:8])
        raise AssertionError("65 pending jobs accepted")
    except DpError as e:
        assert e.code == -2
    assert_point_eq(orc, ctx.msm_collect(1063), orc.msm(bases[:8], sets[0][:8]), "the 64th job")
    for k in range(63):
        assert orc.normalize(ctx.msm_collect(1000 + k))[96] == 1


def check_schedule(orc, ctx: Context, bases, log_n: int, log_m: int, seed: int, rank: int = 0, world: int = 1, exchange=None, gather=None):
    """distributed_plonk_b200/schedule.py: the serial and the overlapped host schedules (transforms with
    look-ahead, commitments batched or queued between transforms) give every transform and every
    commitment of the job list exactly as the oracle does.  Distinct inputs per transform.
    gather(array) -> concatenation over ranks (identity for one worker)."""
    from distributed_plonk_b200 import schedule
    gather = gather or (lambda a: a)
    specs = [(log_n, False, True, False), (log_m, True, False, True), (log_m, True, False, True), (log_m, True, True, True),
             (log_n, False, False, False), (log_m, True, False, True), (log_n, False, True, True)]
    transforms, expect, keep = [], {}, []
    for k, (L, is_quot, inv, cos) in enumerate(specs):
        x = orc.gen_fr(seed + k, 1 << L)
        wl = disp.fft_workloads(L, world)
        r, c = 1 << (L >> 1), (1 << L) >> (L >> 1)
        rows = np.ascontiguousarray(disp.dispatcher_rows(x, L)[wl[rank][0]:wl[rank][1]])
        out = np.zeros(((wl[rank][3] - wl[rank][2]) * r, 4), dtype=np.uint64)
        keep += [rows, out]
        t = schedule.Transform(rows.ctypes.data, out.ctypes.data, out.nbytes, wl, wl[rank][1] - wl[rank][0], is_quot, inv, cos)
        transforms.append(t)
        expect[id(t)] = (out, orc.fft(x, inv, cos), c, r)
    n_b = bases.shape[0]
    lo, hi = rank * n_b // world, (rank + 1) * n_b // world
    sc = np.ascontiguousarray(orc.gen_fr(seed + 50, n_b, False)[lo:hi])
    com = schedule.Commitment(lo, hi, sc.ctypes.data, hi - lo)
    ref_msm = orc.normalize(orc.msm(bases[lo:hi], sc))
    seen = {"fft": 0, "msm": 0}

    def on_fft(t):
        out, ref, c, r = expect[id(t)]
        cols = gather(out.copy()).reshape(c, r, 4)
        assert np.array_equal(disp.assemble(cols), ref), "transform result"
        out[:] = 0
        seen["fft"] += 1

    def on_msm(o):
        assert np.array_equal(orc.normalize(o), ref_msm), "commitment result"
        seen["msm"] += 1

    run = schedule.Runner(ctx, exchange, first_id=7000)
    run.run_serial(transforms, com, (2, 1), 2, on_fft, on_msm)
    assert seen == {"fft": len(specs), "msm": 3}
    run.run_overlapped(transforms, com, 5, 4, on_fft, on_msm)          # more commitments than pairs of transforms
    assert seen == {"fft": 2 * len(specs), "msm": 8}
    run.run_overlapped(transforms[:2], com, 0, 1, on_fft, on_msm)
    assert seen == {"fft": 2 * len(specs) + 2, "msm": 8}
    # bench.py's decision procedure: verified -> overlapped unless the trial says it is slower; a schedule
    # that does not reproduce the serial results (here: a corrupting checksum) or raises -> serial
    out_of = {t.out_ptr: expect[id(t)][0] for t in transforms}
    digest = lambda t: out_of[t.out_ptr].tobytes()
    agree = (lambda ok: ok) if world == 1 else None
    if world == 1:
        clock = iter([2.0, 1.0, 1.0, 2.0])
        step, why = schedule.pick_schedule(run, transforms, com, (1,), digest, lambda f: next(clock), agree)
        assert why.startswith("overlapped"), why
        step()
        step, why = schedule.pick_schedule(run, transforms, com, (1,), digest, lambda f: next(clock), agree)
        assert why.startswith("serial (overlapped schedule verified but slower"), why
        flip = iter(range(10**6))
        step, why = schedule.pick_schedule(run, transforms, com, (1,), lambda t: next(flip), lambda f: 1.0, agree)
        assert "different results" in why, why
        step, why = schedule.pick_schedule(run, transforms, com, (1,), digest, lambda f: 1.0, agree, allow_overlap=False)
        assert why == "serial"

        class Boom(schedule.Runner):
            def run_overlapped(self, *a, **k):
                raise DpError(-4, "simulated failure")
        step, why = schedule.pick_schedule(Boom(ctx, exchange, first_id=9000), transforms, com, (1,), digest, lambda f: 1.0, agree)
        assert "failed" in why and "simulated" in why, why
        step()


TWO_G1_PUBLISHED = (0x0572CBEA904D67468808C8EB50A9450C9721DB309128012543902D0AC358A62AE28F75BB8F1C7C42C39A8C5529BF0F4E,
                    0x166A9D8CABC673A322FDA673779D8E3822BA3ECB8670E461F73BB9021D5FD76A4C56D9D4CD16BD1BBA86881979749D28)


P1_EIP2537 = (0x112B98340EEE2777CC3C14163DEA3EC97977AC3DC5C70DA32E6E87578F44912E902CCEF9EFE28D4A78B8999DFBCA9426,
              0x186B28D92356C4DFEC4B5201AD099DBDEDE3781F8998DDF929B4CD7756192185CA7B8F4EF7088F813270AC3D48868A21)
G1_PLUS_P1_PUBLISHED = (0x0A40300CE2DEC9888B60690E9A41D3004FDA4886854573974FAB73B046D3147BA5B7A5BDE85279FFEDE1B45B3918D82D,
                        0x06D3D887E9F53B9EC4EB6CEDF5607226754B07C01ACE7834F57F3E7315FAEFB739E59018E22C492006190FBA4A870025)


def check_published_vector(orc, make_ctx):
    """The library against the one PUBLISHED absolute value on this path: 2*G1 (EIP-2537 "bls_g1add_(g1+g1=2*g1)").
    MSMs over copies of the generator must land on it whatever the bucket geometry."""
    from oracle.py import bls12_381 as B
    gen = np.zeros(104, dtype=np.uint8)
    orc.lib().orc_g1_generator(gen.ctypes.data)
    ctx = make_ctx()
    # G1 + P1 (EIP-2537 "bls_g1add_(g1+p1)") as an MSM with unit scalars, alone and inside a longer SRS
    p1 = np.frombuffer(B.g1_affine_to_bytes(P1_EIP2537), dtype=np.uint8)
    for n in (2, 2500):
        bases = np.stack([gen, p1] + [gen] * (n - 2))
        ctx.init(bases, 1 << 4, 1 << 7)
        sc = np.zeros((n, 4), dtype=np.uint64)
        sc[0, 0] = sc[1, 0] = 1
        assert B.g1_affine_from_bytes(orc.normalize(ctx.msm(0, n, sc)).tobytes()) == G1_PLUS_P1_PUBLISHED, f"G+P1, n={n}"
    for n in (1, 2, 40, 3000):
        ctx.init(np.stack([gen] * n), 1 << 4, 1 << 7)
        sc = np.zeros((n, 4), dtype=np.uint64)
        sc[0, 0] = 2
        assert B.g1_affine_from_bytes(orc.normalize(ctx.msm(0, n, sc)).tobytes()) == TWO_G1_PUBLISHED, f"2*G, n={n}"
        if n >= 2:
            sc[:] = 0
            sc[0, 0] = sc[n - 1, 0] = 1
            assert B.g1_affine_from_bytes(orc.normalize(ctx.msm(0, n, sc)).tobytes()) == TWO_G1_PUBLISHED, f"G+G, n={n}"
            # r-1 copies of ... : (r + 2) * G = 2 * G  via scalars summing to r + 2 over identical bases
            sc[:] = 0
            sc[0] = np.frombuffer((B.FR_MOD - 1).to_bytes(32, "little"), dtype=np.uint64)
            sc[1, 0] = 3
            assert B.g1_affine_from_bytes(orc.normalize(ctx.msm(0, n, sc)).tobytes()) == TWO_G1_PUBLISHED, f"(r-1)G+3G, n={n}"
    ctx.close()


def check_fft1_row_lengths(orc, worker: PlonkSlave, domain_log: int, is_quot: bool, seed: int):
    """fft1 with rows shorter than c (the zero tail of a padded polynomial left out) or longer than c:
    `c_domain.fft_in_place(&mut v)` resizes v, so the result is that of the zero-extended / cut row"""
    N = 1 << domain_log
    r = 1 << (domain_log >> 1)
    c = N // r
    x = orc.gen_fr(seed, N // 8 if N >= 8 else N)              # n coefficients on an 8n domain
    padded = np.zeros((N, 4), dtype=np.uint64)
    padded[: x.shape[0]] = x
    rows = disp.dispatcher_rows(x, domain_log)                  # [r][c], columns >= c/8 are zero
    keep = max(1, c // 8)
    assert not rows[:, keep:].any()
    wl = disp.fft_workloads(domain_log, 1)
    for tid, (inv, cos) in enumerate(FLAG_COMBOS):
        if tid == 0:             # an abandoned task under the same id is simply replaced (fft_tasks.insert)
            worker.fft_init(9100, wl, is_quot, not inv, cos)
            worker.fft1(9100, 0, chunks(orc.gen_fr(seed + 99, c)))
        worker.fft_init(9100 + tid, wl, is_quot, inv, cos)
        for j in range(r):
            if j % 3 == 0:       # full row
                worker.fft1(9100 + tid, j, chunks(rows[j]))
            elif j % 3 == 1:     # only the non-zero prefix
                worker.fft1(9100 + tid, j, chunks(np.ascontiguousarray(rows[j, :keep])))
            else:                # a row with junk beyond c: cut
                worker.fft1(9100 + tid, j, chunks(np.concatenate([rows[j], orc.gen_fr(seed + j, 3)])))
        worker.fft2_prepare(9100 + tid)
        got = disp.assemble(worker.fft2_array(9100 + tid))
        assert np.array_equal(got, orc.fft(padded, inv, cos)), f"short / long rows, inv={inv} coset={cos}"


def check_short_rows(orc, workers, domain_log: int, is_quot: bool, row_len: int, seed: int, copy_fn=None, per_row=False):
    """Rows handed in cut after `row_len` entries (dp_fft1_rows_short, or dp_fft1 per row when per_row): the transform
    is that of the zero-extended rows (Radix2EvaluationDomain::fft_in_place resizes, worker.rs:81-85), whether the
    row kernel drops the zero-input butterfly stages (power-of-two counts) or reads a zero-filled tail."""
    N = 1 << domain_log
    r = 1 << (domain_log >> 1)
    c = N // r
    W = len(workers)
    wl = disp.fft_workloads(domain_log, W)
    for k, (inv, cos) in enumerate(FLAG_COMBOS):
        rows = np.zeros((r, c, 4), dtype=np.uint64)
        rows[:, :row_len] = orc.gen_fr(seed + k, r * row_len).reshape(r, row_len, 4)
        x = np.ascontiguousarray(rows.transpose(1, 0, 2)).reshape(N, 4)          # rows[i][j] = x[i + r*j]
        tid = 0x5A0000 + seed * 16 + k
        for w in workers:
            w.fft_init(tid, wl, is_quot, inv, cos)
        for p, w in enumerate(workers):
            lo, hi = wl[p][0], wl[p][1]
            short = np.ascontiguousarray(rows[lo:hi, :row_len])
            if per_row:
                for j in range(hi - lo):
                    w.ctx.fft1(tid, j, short[j])
            else:
                half = (hi - lo) // 2                                              # two calls: the bookkeeping is per row
                if half:
                    w.ctx.fft1_rows_short(tid, 0, np.ascontiguousarray(short[:half]), half, row_len)
                w.ctx.fft1_rows_short(tid, half, np.ascontiguousarray(short[half:]), hi - lo - half, row_len)
        if W == 1 or workers[0].ctx.peer_ready():
            for w in workers:
                w.fft2_prepare(tid)
        else:
            for dst, src, nb in local_exchange([w.ctx for w in workers], tid):
                copy_fn(dst, src, nb)
        got = disp.assemble(np.concatenate([w.fft2_array(tid) for w in workers], axis=0))
        assert np.array_equal(got, orc.fft(x, inv, cos)), f"short rows L={domain_log} W={W} len={row_len} inv={inv} coset={cos}"


def check_dev_valid_cols_hint(orc, ctx: Context, domain_log: int, is_quot: bool, valid: int, seed: int, as_ptr):
    """dp_fft_dev with the promise that every row is zero from column `valid` on (W = 1)"""
    N = 1 << domain_log
    r = 1 << (domain_log >> 1)
    c = N // r
    rows = np.zeros((r, c, 4), dtype=np.uint64)
    rows[:, :valid] = orc.gen_fr(seed, r * valid).reshape(r, valid, 4)
    x = np.ascontiguousarray(rows.transpose(1, 0, 2)).reshape(N, 4)
    ctx.fft_dev_hint_valid_cols(is_quot, valid)
    try:
        for inv, cos in FLAG_COMBOS:
            src, dst = as_ptr(rows), as_ptr(np.zeros((c, r, 4), dtype=np.uint64))
            ctx.fft_dev(src.ptr, dst.ptr, is_quot, inv, cos)
            got = disp.assemble(dst.read().reshape(c, r, 4))
            # inverse transforms ignore the promise (their inputs are evaluations); the rows here are zero-padded anyway
            assert np.array_equal(got, orc.fft(x, inv, cos)), f"valid-cols hint L={domain_log} valid={valid} inv={inv} coset={cos}"
    finally:
        ctx.fft_dev_hint_valid_cols(is_quot, 0)
