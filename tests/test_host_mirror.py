"""The C++ host mirror of the reference's worker / dispatcher (distributed_plonk_b200/host/
plonk_worker.hpp): compiled with g++ and linked against the library - the kernel-logic emulator
build on CPU, the real CUDA library on the GPU box - then checked against the oracle."""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "host_mirror_cli.cpp")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def build_cli(lib_path: str, out: str) -> str:
    d, name = os.path.dirname(lib_path), os.path.basename(lib_path)
    subprocess.check_call([CXX, "-O2", "-std=c++17", SRC, "-o", out, f"-L{d}", f"-l:{name}", f"-Wl,-rpath,{d}", "-pthread"])
    return out


def run_case(orc, exe, tmp_path, n_bases, log_n, log_q, flags, n_coeffs, rounds=True):
    bases = orc.gen_bases(31, n_bases, 32, True)
    sc = orc.gen_fr(32, n_bases, False)
    co = orc.gen_fr(33, n_coeffs, True)
    req, rep = tmp_path / "req.bin", tmp_path / "rep.bin"
    with open(req, "wb") as f:
        f.write(struct.pack("<5Q", n_bases, log_n, log_q, n_coeffs, flags))
        f.write(bases.tobytes())
        f.write(sc.tobytes())
        f.write(co.tobytes())
    subprocess.check_call([exe, str(req), str(rep)] + (["rounds"] if rounds else []))
    raw = np.fromfile(rep, dtype=np.uint8)
    L = log_q if flags & 1 else log_n
    part, rest = raw[:144], raw[144:].view(np.uint64).reshape(-1, 4)
    out = rest[:1 << L]
    assert np.array_equal(orc.normalize(part), orc.normalize(orc.msm(bases, sc)))
    pad = np.zeros((1 << L, 4), dtype=np.uint64)
    pad[:n_coeffs] = co
    assert np.array_equal(out, orc.fft(pad, bool(flags & 2), bool(flags & 4)))
    if not rounds:
        assert rest.shape[0] == 1 << L
        return
    pz, q, lc = rest[1 << L], rest[(1 << L) + 1:(1 << L) + n_coeffs], rest[(1 << L) + n_coeffs:]
    # rounds 4-5 through the C++ mirror: p(z), p / (X - z), z p + z p[:n/2]   (z = p[0])
    assert np.array_equal(pz, orc.poly_eval(co, co[0]))
    assert np.array_equal(q, orc.poly_div_linear(co, co[0]))
    assert np.array_equal(lc, orc.poly_lincomb([co, co[:n_coeffs // 2]], np.stack([co[0], co[0]])))


def test_cpp_host_mirror_on_emulator(orc, tmp_path):
    from tests.emul import build as emul_build
    exe = build_cli(emul_build.build(), str(tmp_path / "host_mirror_emul"))
    run_case(orc, exe, tmp_path, 200, 6, 9, 0b101, 64)     # coset NTT of n coeffs on the 8n domain
    run_case(orc, exe, tmp_path, 64, 6, 9, 0b010, 64)      # iNTT on the gate domain


@pytest.mark.gpu
def test_cpp_host_mirror_on_gpu(orc, tmp_path):
    import distributed_plonk_b200 as dp
    dp.load()
    exe = build_cli(dp.library_path(), str(tmp_path / "host_mirror_gpu"))
    run_case(orc, exe, tmp_path, (1 << 12) + 32, 12, 15, 0b101, 1 << 12, rounds=False)
    run_case(orc, exe, tmp_path, 1000, 12, 15, 0b110, 1 << 12, rounds=False)
