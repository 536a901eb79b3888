"""The limb-level algorithms that run on the GPU (distributed_plonk_b200/csrc/{ptx_arith,field,g1}.cuh),
compiled for the host with the PTX carry flag emulated, against Python integers: Montgomery product
/ add / sub / inverse of Fr and Fq on random and adversarial operands, the XYZZ group law including
its special cases, and (hypothesis) algebraic identities.  CPU only."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle.py import bls12_381 as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emul", "emul_field.cpp")
OUT = os.path.join(ROOT, "tests", "emul", "_build", "libemul_field.so")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
FIELDS = {"fr": (B.FR_MOD, 4, B.FR_R), "fq": (B.FQ_MOD, 6, B.FQ_R)}


@pytest.fixture(scope="module")
def E():
    deps = [SRC] + [os.path.join(ROOT, "distributed_plonk_b200", "csrc", f) for f in ("ptx_arith.cuh", "field.cuh", "g1.cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call([CXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", SRC, "-o", OUT])
    return C.CDLL(OUT)


def arr(vals, nl):
    return np.frombuffer(b"".join(int(v).to_bytes(nl * 8, "little") for v in vals), dtype=np.uint64).reshape(-1, nl).copy()


def ints(a):
    return [int.from_bytes(a[i].tobytes(), "little") for i in range(a.shape[0])]


def binop(E, name, op, xs, ys, nl):
    A, Bm = arr(xs, nl), arr(ys, nl)
    O = np.empty_like(A)
    getattr(E, f"emu_{name}_{op}")(A.ctypes.data_as(C.c_void_p), Bm.ctypes.data_as(C.c_void_p), O.ctypes.data_as(C.c_void_p),
                                    C.c_uint64(len(xs)))
    return ints(O)


@pytest.mark.parametrize("name", ["fr", "fq"])
def test_field_ops_random_and_edges(E, name):
    mod, nl, R = FIELDS[name]
    rng = random.Random(7)
    edge = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, (1 << (nl * 64 - 3)) % mod, R % mod, (mod - R) % mod,
            (1 << 32) - 1, (1 << 64) - 1, mod >> 1, (mod >> 32) << 32]
    xs = edge + [rng.randrange(mod) for _ in range(4000)]
    ys = list(reversed(edge)) + [rng.randrange(mod) for _ in range(4000)]
    rinv = pow(R, -1, mod)
    assert binop(E, name, "mul", xs, ys, nl) == [x * y * rinv % mod for x, y in zip(xs, ys)]
    # the dedicated squaring (symmetric partial products + separate reduction) against Python and against the product
    A, O = arr(xs + ys, nl), np.empty((len(xs) + len(ys), nl), dtype=np.uint64)
    getattr(E, f"emu_{name}_sqr")(A.ctypes.data_as(C.c_void_p), O.ctypes.data_as(C.c_void_p), C.c_uint64(A.shape[0]))
    assert ints(O) == [x * x * rinv % mod for x in xs + ys]
    assert binop(E, name, "add", xs, ys, nl) == [(x + y) % mod for x, y in zip(xs, ys)]
    assert binop(E, name, "sub", xs, ys, nl) == [(x - y) % mod for x, y in zip(xs, ys)]
    for x in (1, 2, mod - 1, rng.randrange(1, mod)):
        a, o = arr([x], nl), np.zeros((1, nl), dtype=np.uint64)
        getattr(E, f"emu_{name}_inverse")(a.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p))
        assert ints(o)[0] == pow(x * rinv % mod, -1, mod) * R % mod
    # the single-thread inversion of msm_final (binary extended Euclid): same value as Fermat's, edge operands included
    for x in [1, 2, 3, mod - 1, mod - 2, R % mod, (mod - R) % mod, (1 << 32), (1 << (nl * 64 - 3)) % mod, mod >> 1] + [rng.randrange(1, mod) for _ in range(300)]:
        a, o = arr([x], nl), np.zeros((1, nl), dtype=np.uint64)
        getattr(E, f"emu_{name}_inverse_vartime")(a.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p))
        assert ints(o)[0] == pow(x * rinv % mod, -1, mod) * R % mod, hex(x)


@settings(max_examples=200, deadline=None)
@given(st.integers(0, B.FQ_MOD - 1), st.integers(0, B.FQ_MOD - 1), st.integers(0, B.FQ_MOD - 1))
def test_fq_ring_identities(E, a, b, c):
    """(a+b)*c == a*c + b*c and (a-b)+b == a on the emulated limb code"""
    ab = binop(E, "fq", "add", [a], [b], 6)[0]
    lhs = binop(E, "fq", "mul", [ab], [c], 6)[0]
    rhs = binop(E, "fq", "add", binop(E, "fq", "mul", [a], [c], 6), binop(E, "fq", "mul", [b], [c], 6), 6)[0]
    assert lhs == rhs
    assert binop(E, "fq", "add", binop(E, "fq", "sub", [a], [b], 6), [b], 6)[0] == a


def aff96(pt):
    return b"\0" * 96 if pt is None else B.fq_to_mont_bytes(pt[0]) + B.fq_to_mont_bytes(pt[1])


def from96(b):
    return None if int.from_bytes(b, "little") == 0 else (B.fq_from_mont_bytes(b[:48]), B.fq_from_mont_bytes(b[48:96]))


def test_g1_xyzz_group_law(E):
    rng = random.Random(11)
    G = B.G1_GEN
    for k in [0, 1, 2, 3, B.FR_MOD - 1, B.FR_MOD, rng.randrange(B.FR_MOD)]:
        out = C.create_string_buffer(96)
        ka = (C.c_uint64 * 4)(*[(k >> (64 * i)) & (2 ** 64 - 1) for i in range(4)])
        E.emu_g1_mul(C.create_string_buffer(aff96(G)), ka, out)
        assert from96(out.raw) == B.g1_mul(G, k)
    pts = [B.g1_mul(G, rng.randrange(1, B.FR_MOD)) for _ in range(9)]
    pts[3] = None                      # infinity operand
    pts[5] = pts[1]                    # doubling through the mixed add
    pts[7] = B.g1_neg(pts[0])          # P + (-P)
    pts[8] = pts[2]                    # doubling through the full add after the merge
    exp = None
    for p in pts:
        exp = B.g1_add(exp, p)
    out = C.create_string_buffer(96)
    E.emu_g1_sum(C.create_string_buffer(b"".join(aff96(p) for p in pts)), C.c_uint64(len(pts)), out)
    assert from96(out.raw) == exp
    E.emu_g1_add_xyzz_self(C.create_string_buffer(aff96(pts[1])), out)
    assert from96(out.raw) == B.g1_add(pts[1], pts[1])
