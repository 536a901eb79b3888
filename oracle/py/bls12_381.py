"""Tier-0 oracle: BLS12-381 Fr / Fq / G1 in plain Python integers.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (distributed_plonk_b200/,
include/, bench.py's GPU arm) may import this; only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs use oracle/ as the checker.

PARITY UNPINNED by reference fixtures: the reference (/root/reference, Rust, arkworks 0.3.0)
holds no golden vectors / KATs and cannot be compiled here (no rustc).
The published absolute values that exist for this path - 2*G1 and G1 + P1, EIP-2537 test vectors
"bls_g1add_(g1+g1=2*g1)" and "bls_g1add_(g1+p1)" - are checked in tests/test_oracle.py (both tiers)
and against the library.
Beyond it values are pinned by mathematics: NTT outputs and the MSM group element are unique,
and this file computes them by definition (O(N^2) DFT, double-and-add) so it is independent
of the fast C restatement in oracle/c/ that it cross-checks.  The public constants below
(moduli, generator, two-adic root) are checked against the published BLS12-381 /
arkworks values in tests/test_oracle.py.

Reference call sites this restates (file:line into /root/reference):
  * Fr/G1 types, raw in-memory layouts ........ src/utils.rs:27-43
  * Radix2EvaluationDomain::new / fft / ifft .. src/worker.rs:143-154, 81-85, 104-108
  * 2-D decomposition model ................... src/playground.rs:21-80
  * fft1_helper / fft2_helper ................. src/worker.rs:66-115
  * exchange scatter .......................... src/worker.rs:327-330, 432-435
  * dispatcher row/col <-> flat mapping ....... src/dispatcher2.rs:731-787
  * VariableBaseMSM::multi_scalar_mul ......... src/worker.rs:117-123, 177-182
  * rounds 2-5 arithmetic of Prover::prove .... src/dispatcher2.rs:329-345, 363-504, 535-690
"""
from __future__ import annotations

# ----------------------------------------------------------------------------- constants
FR_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
FQ_MOD = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
FR_R = (1 << 256) % FR_MOD          # Montgomery radix, ark-ff Fp256 (4 x u64)
FQ_R = (1 << 384) % FQ_MOD          # Montgomery radix, ark-ff Fp384 (6 x u64)
FR_GENERATOR = 7                    # ark-bls12-381 FrParameters::GENERATOR
FR_TWO_ADICITY = 32
FR_TWO_ADIC_ROOT = pow(FR_GENERATOR, (FR_MOD - 1) >> FR_TWO_ADICITY, FR_MOD)
G1_B = 4
G1_GEN = (
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
)

FR_BYTES = 32
FQ_BYTES = 48
G1_AFFINE_BYTES = 104      # GroupAffine{x:Fq,y:Fq,infinity:bool} padded (SURVEY §8a layouts)
G1_PROJ_BYTES = 144        # GroupProjective{x,y,z:Fq} Jacobian


# ----------------------------------------------------------------------------- byte layouts
def fr_to_mont_bytes(v: int) -> bytes:
    """Fr value -> 32 B raw ark-ff Fp256 (Montgomery, 4 x u64 LE)."""
    return ((v % FR_MOD) * FR_R % FR_MOD).to_bytes(32, "little")


def fr_from_mont_bytes(b: bytes) -> int:
    return int.from_bytes(b, "little") * pow(FR_R, -1, FR_MOD) % FR_MOD


def fr_vec_to_bytes(vs) -> bytes:
    return b"".join(fr_to_mont_bytes(v) for v in vs)


def fr_vec_from_bytes(b: bytes):
    rinv = pow(FR_R, -1, FR_MOD)
    return [int.from_bytes(b[i:i + 32], "little") * rinv % FR_MOD for i in range(0, len(b), 32)]


def bigint256_to_bytes(v: int) -> bytes:
    """canonical (non-Montgomery) BigInteger256 as produced by Fr::into_repr."""
    return int(v).to_bytes(32, "little")


def fq_to_mont_bytes(v: int) -> bytes:
    return ((v % FQ_MOD) * FQ_R % FQ_MOD).to_bytes(48, "little")


def fq_from_mont_bytes(b: bytes) -> int:
    return int.from_bytes(b, "little") * pow(FQ_R, -1, FQ_MOD) % FQ_MOD


def g1_affine_to_bytes(pt) -> bytes:
    """pt = None (infinity) or (x, y).  ark identity = (0, 1, infinity=true)."""
    if pt is None:
        return fq_to_mont_bytes(0) + fq_to_mont_bytes(1) + b"\x01" + b"\x00" * 7
    return fq_to_mont_bytes(pt[0]) + fq_to_mont_bytes(pt[1]) + b"\x00" * 8


def g1_affine_from_bytes(b: bytes):
    if b[96] != 0:
        return None
    return (fq_from_mont_bytes(b[0:48]), fq_from_mont_bytes(b[48:96]))


def g1_jacobian_from_bytes(b: bytes):
    """144 B raw GroupProjective -> affine tuple / None."""
    X = fq_from_mont_bytes(b[0:48])
    Y = fq_from_mont_bytes(b[48:96])
    Z = fq_from_mont_bytes(b[96:144])
    if Z == 0:
        return None
    zi = pow(Z, -1, FQ_MOD)
    return (X * zi * zi % FQ_MOD, Y * zi * zi * zi % FQ_MOD)


def g1_jacobian_to_bytes(pt) -> bytes:
    """normalised Jacobian (Z = 1) or the ark identity (0, 1, 0)."""
    if pt is None:
        return fq_to_mont_bytes(0) + fq_to_mont_bytes(1) + fq_to_mont_bytes(0)
    return fq_to_mont_bytes(pt[0]) + fq_to_mont_bytes(pt[1]) + fq_to_mont_bytes(1)


# ----------------------------------------------------------------------------- G1 (affine, by definition)
def g1_is_on_curve(pt) -> bool:
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - G1_B) % FQ_MOD == 0


def g1_neg(pt):
    return None if pt is None else (pt[0], (-pt[1]) % FQ_MOD)


def g1_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % FQ_MOD == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, FQ_MOD) % FQ_MOD
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, FQ_MOD) % FQ_MOD
    x3 = (lam * lam - x1 - x2) % FQ_MOD
    y3 = (lam * (x1 - x3) - y1) % FQ_MOD
    return (x3, y3)


def g1_mul(pt, k: int):
    acc = None
    add = pt
    while k:
        if k & 1:
            acc = g1_add(acc, add)
        add = g1_add(add, add)
        k >>= 1
    return acc


def msm_naive(bases, scalars):
    """sum_i scalars[i] * bases[i], truncating to min(len) like ark's multi_scalar_mul."""
    acc = None
    for b, s in zip(bases, scalars):
        acc = g1_add(acc, g1_mul(b, s))
    return acc


# ----------------------------------------------------------------------------- domains / NTT (by definition)
def log2_ceil(n: int) -> int:
    return 0 if n <= 1 else (n - 1).bit_length()


class Domain:
    """Radix2EvaluationDomain::<Fr>::new(k) (ark-poly 0.3.0; worker.rs:143-154)."""

    def __init__(self, min_size: int):
        self.log_size = log2_ceil(min_size)
        self.size = 1 << self.log_size
        assert self.log_size <= FR_TWO_ADICITY
        self.group_gen = pow(FR_TWO_ADIC_ROOT, 1 << (FR_TWO_ADICITY - self.log_size), FR_MOD)
        self.group_gen_inv = pow(self.group_gen, -1, FR_MOD)
        self.size_inv = pow(self.size, -1, FR_MOD)

    def split(self):
        """(r, c) rule of worker.rs:144-147 / playground.rs:22-23."""
        r = 1 << (self.log_size >> 1)
        return r, self.size // r


def dft(x, omega):
    n = len(x)
    pw = [1] * n
    for i in range(1, n):
        pw[i] = pw[i - 1] * omega % FR_MOD
    return [sum(x[j] * pw[(j * k) % n] for j in range(n)) % FR_MOD for k in range(n)]


def fft(dom: Domain, x):
    x = list(x) + [0] * (dom.size - len(x))
    return dft(x, dom.group_gen)


def ifft(dom: Domain, x):
    x = list(x) + [0] * (dom.size - len(x))
    return [v * dom.size_inv % FR_MOD for v in dft(x, dom.group_gen_inv)]


def coset_fft(dom: Domain, x):
    g = FR_GENERATOR
    return fft(dom, [v * pow(g, j, FR_MOD) % FR_MOD for j, v in enumerate(x)])


def coset_ifft(dom: Domain, x):
    gi = pow(FR_GENERATOR, -1, FR_MOD)
    return [v * pow(gi, j, FR_MOD) % FR_MOD for j, v in enumerate(ifft(dom, x))]


# ----------------------------------------------------------------------------- the worker's 2-D pieces
def fft1_helper(v, i, is_coset, is_inv, dom: Domain):
    """worker.rs:66-94 on global row index i (c elements in, c out)."""
    r, c = dom.split()
    cdom = Domain(c)
    v = list(v)
    assert len(v) == c
    if is_coset and not is_inv:
        v = [u * pow(FR_GENERATOR, i + j * r, FR_MOD) % FR_MOD for j, u in enumerate(v)]
    v = ifft(cdom, v) if is_inv else fft(cdom, v)
    w = dom.group_gen_inv if is_inv else dom.group_gen
    return [u * pow(w, i * j, FR_MOD) % FR_MOD for j, u in enumerate(v)]


def fft2_helper(v, i, is_coset, is_inv, dom: Domain):
    """worker.rs:96-115 on global column index i (r elements in, r out)."""
    r, c = dom.split()
    rdom = Domain(r)
    v = list(v)
    assert len(v) == r
    v = ifft(rdom, v) if is_inv else fft(rdom, v)
    if is_coset and is_inv:
        gi = pow(FR_GENERATOR, -1, FR_MOD)
        v = [u * pow(gi, i + j * c, FR_MOD) % FR_MOD for j, u in enumerate(v)]
    return v


def dispatcher_rows(dom: Domain, coeffs):
    """dispatcher2.rs:746,754: zero-pad, chunk by r, transpose -> r rows of length c."""
    r, c = dom.split()
    x = list(coeffs) + [0] * (dom.size - len(coeffs))
    return [[x[b + a * r] for a in range(c)] for b in range(r)]


def distributed_fft(dom: Domain, coeffs, is_inv, is_coset, n_workers=1):
    """Full 4-RPC pipeline of dispatcher2.rs:731-787 + worker.rs:235-381, 412-438."""
    r, c = dom.split()
    rows = dispatcher_rows(dom, coeffs)
    W = n_workers
    # fft1 on every worker's row block
    rows = [fft1_helper(rows[i], i, is_coset, is_inv, dom) for i in range(r)]
    # exchange: worker p sends rows[p-block][cols of q] flattened row-major; q scatters
    cols = [[0] * r for _ in range(c)]
    for p in range(W):
        rs, re = p * r // W, (p + 1) * r // W
        for q in range(W):
            cs, ce = q * c // W, (q + 1) * c // W
            flat = [rows[i][k] for i in range(rs, re) for k in range(cs, ce)]
            ncols = ce - cs
            for t, val in enumerate(flat):
                cols[cs + t % ncols][rs + t // ncols] = val
    cols = [fft2_helper(cols[k], k, is_coset, is_inv, dom) for k in range(c)]
    # dispatcher2.rs:780-786: transpose(u).concat()  -> out[j*c + i] = col_i[j]
    return [cols[i][j] for j in range(r) for i in range(c)]


# ----------------------------------------------------------------------------- canonical point encoding
def g1_compress(pt) -> bytes:
    """ark-serialize 0.3.0 compressed GroupAffine (48 B): x little-endian, bit 7 of the last byte =
    (y > -y), bit 6 = infinity.  pt = (x, y) ints or None."""
    if pt is None:
        return bytes(47) + bytes([1 << 6])
    x, y = pt
    b = bytearray(x.to_bytes(48, "little"))
    if y > FQ_MOD - y:
        b[47] |= 1 << 7
    return bytes(b)


def g1_decompress(b: bytes, check_subgroup: bool = True):
    """inverse of g1_compress; raises ValueError like ark's deserialize returns Err"""
    positive, infinity = b[47] >> 7 & 1, b[47] >> 6 & 1
    if positive and infinity:
        raise ValueError("bad flags")
    if infinity:
        return None
    x = int.from_bytes(b[:47] + bytes([b[47] & 0x3F]), "little")
    if x >= FQ_MOD:
        raise ValueError("x not canonical")
    rhs = (x * x * x + G1_B) % FQ_MOD
    y = pow(rhs, (FQ_MOD + 1) // 4, FQ_MOD)
    if y * y % FQ_MOD != rhs:
        raise ValueError("not on the curve")
    if (y > FQ_MOD - y) != bool(positive):
        y = FQ_MOD - y
    if check_subgroup and g1_mul((x, y), FR_MOD) is not None:
        raise ValueError("not in the subgroup")
    return (x, y)


# ----------------------------------------------------------------------------- rounds 2-5 of Prover::prove
def perm_product(wires, idp, sigma, beta, gamma):
    """dispatcher2.rs:329-345: z[0] = 1, z[j+1] = z[j] * a_j / b_j.  wires/idp/sigma: [types][n] ints."""
    n = len(wires[0])
    z = [1]
    for j in range(n - 1):
        a = b = 1
        for w, i_, s_ in zip(wires, idp, sigma):
            a = a * (w[j] + gamma + beta * i_[j]) % FR_MOD
            b = b * (w[j] + gamma + beta * s_[j]) % FR_MOD
        z.append(z[-1] * a * pow(b, -1, FR_MOD) % FR_MOD)
    return z


def quotient_evals(sel, sig, w, z, pi, k, alpha, beta, gamma, n):
    """dispatcher2.rs:363-504 by definition.  sel [13][m], sig / w [5][m], z / pi [m], k [5] (ints)."""
    m = len(z)
    ratio = m // n
    omega = Domain(m).group_gen
    x = [FR_GENERATOR * pow(omega, i, FR_MOD) % FR_MOD for i in range(m)]
    zh_inv = [pow(pow(x[i], n, FR_MOD) - 1, -1, FR_MOD) for i in range(ratio)]
    a2n = alpha * alpha * pow(n, -1, FR_MOD) % FR_MOD
    out = []
    for i in range(m):
        a, b, c, d, e = (w[j][i] for j in range(5))
        gate = (sel[11][i] + pi[i] + sel[0][i] * a + sel[1][i] * b + sel[2][i] * c + sel[3][i] * d
                + sel[4][i] * a * b + sel[5][i] * c * d + sel[12][i] * a * b * c * d * e
                + sel[6][i] * a**5 + sel[7][i] * b**5 + sel[8][i] * c**5 + sel[9][i] * d**5 - sel[10][i] * e)
        acc1, acc2 = z[i], z[(i + ratio) % m]
        for j in range(5):
            acc1 = acc1 * (w[j][i] + gamma + k[j] * x[i] * beta) % FR_MOD
            acc2 = acc2 * (w[j][i] + gamma + sig[j][i] * beta) % FR_MOD
        t3 = a2n * (z[i] - 1) * pow(x[i] - 1, -1, FR_MOD)
        out.append((zh_inv[i % ratio] * (gate + alpha * (acc1 - acc2)) + t3) % FR_MOD)
    return out


def poly_eval(coeffs, point):
    """DensePolynomial::evaluate (dispatcher2.rs:535-548), by definition"""
    return sum(c * pow(point, j, FR_MOD) for j, c in enumerate(coeffs)) % FR_MOD


def poly_div_linear(coeffs, point):
    """quotient q of p = q * (X - point) + p(point)  (dispatcher2.rs:651-666), by definition:
    q_j = sum_{k > j} p_k * point^(k-j-1)"""
    n = len(coeffs)
    return [sum(coeffs[k] * pow(point, k - j - 1, FR_MOD) for k in range(j + 1, n)) % FR_MOD for j in range(n - 1)]


def poly_lincomb(polys, coeffs):
    n = max(len(p) for p in polys)
    return [sum(c * p[j] for p, c in zip(polys, coeffs) if j < len(p)) % FR_MOD for j in range(n)]


# ----------------------------------------------------------------------------- seeded PRNG shared with the C oracle
class SplitMix64:
    """SplitMix64; the one PRNG every oracle / test / bench generator shares."""

    MASK = (1 << 64) - 1

    def __init__(self, seed: int):
        self.s = seed & self.MASK

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & self.MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & self.MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & self.MASK
        return z ^ (z >> 31)

    def fr(self) -> int:
        """uniform Fr by rejection on 255 bits (like Fr::rand)."""
        while True:
            limbs = [self.next() for _ in range(4)]
            v = limbs[0] | (limbs[1] << 64) | (limbs[2] << 128) | (limbs[3] << 192)
            v &= (1 << 255) - 1
            if v < FR_MOD:
                return v
