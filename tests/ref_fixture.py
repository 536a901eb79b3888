"""`ref_v1.bin`: golden vectors written by the REFERENCE's own code (rust/dump_fixtures.rs; format in rust/README.md).
Reader, checker and - so that the checker itself is tested while no Rust toolchain exists - a writer that produces a
file of the same format from the oracle."""
from __future__ import annotations

import struct

import numpy as np

MAGIC = b"DPREFv1\0"
LAYOUT, NTT, FFT1, FFT2, MSM, COMMIT, DIST_FFT, COMPRESSED = range(1, 9)
FLAGS = [(False, False), (True, False), (False, True), (True, True)]


def read(path):
    """-> list of (tag, (a, b, c, d), [blob bytes])"""
    data = open(path, "rb").read()
    assert data[:8] == MAGIC, "not a DPREFv1 file"
    (count,), off, out = struct.unpack_from("<I", data, 8), 12, []
    for _ in range(count):
        tag, a, b, c, d, nb = struct.unpack_from("<IQQQQI", data, off)
        off += 40
        blobs = []
        for _ in range(nb):
            (ln,) = struct.unpack_from("<Q", data, off)
            blobs.append(data[off + 8:off + 8 + ln])
            off += 8 + ln
        out.append((tag, (a, b, c, d), blobs))
    assert off == len(data), "trailing bytes"
    return out


def write(path, records):
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<I", len(records)))
        for tag, p, blobs in records:
            f.write(struct.pack("<IQQQQI", tag, *p, len(blobs)))
            for b in blobs:
                b = bytes(b)
                f.write(struct.pack("<Q", len(b)) + b)


def fr(blob) -> np.ndarray:
    return np.frombuffer(blob, dtype=np.uint64).reshape(-1, 4).copy()


def pts(blob, size) -> np.ndarray:
    return np.frombuffer(blob, dtype=np.uint8).reshape(-1, size).copy()


class OracleImpl:
    """the functions a record exercises, answered by the CPU oracle"""

    def __init__(self, orc):
        self.orc = orc

    def ntt(self, x, log, inv, coset):
        pad = np.zeros((1 << log, 4), dtype=np.uint64)
        pad[: x.shape[0]] = x
        return self.orc.fft(pad, inv, coset)

    def dist_fft(self, x, log, inv, coset):
        return self.orc.distributed_fft(x, 1 << log, inv, coset, 1, True)

    def fft1(self, row, i, log, inv, coset):
        return self.orc.fft1_helper(row, i, coset, inv, 1 << log, True)

    def fft2(self, col, i, log, inv, coset):
        return self.orc.fft2_helper(col, i, coset, inv, 1 << log, True)

    def msm(self, bases, a, b, scalars):
        return self.orc.normalize(self.orc.msm(bases[a:b], scalars))

    def commit(self, bases, coeffs):
        return self.orc.normalize(self.orc.commit(bases, coeffs))

    def decompress(self, comp):
        out, rcs = self.orc.g1_decompress(comp, True)
        assert not any(rcs)
        return out


class LibraryImpl:
    """... by the library behind the C ABI (the CUDA build, or the kernel-logic emulator on a CPU box)"""

    def __init__(self, orc, lib, device=0):
        self.orc, self.lib, self.device = orc, lib, device

    def _ctx(self, bases, n_log, q_log):
        from distributed_plonk_b200._binding import Context
        c = Context(self.lib, self.device, 0, 1)
        c.init(bases, 1 << n_log, 1 << q_log)
        return c

    def ntt(self, x, log, inv, coset):
        c = self._ctx(np.zeros(0, dtype=np.uint8), log, log)
        try:
            return c.ntt(x, log, inv, coset)
        finally:
            c.close()

    def dist_fft(self, x, log, inv, coset):
        from distributed_plonk_b200 import dispatcher as disp
        from distributed_plonk_b200.worker import PlonkSlave
        w = PlonkSlave(self.lib, 0, 1, device=self.device)
        try:
            w.init([b""], 1 << log, 1 << log)
            return disp.fft([w], log, x, False, inv, coset, task_id=77)
        finally:
            w.close()

    fft1 = fft2 = None      # single rows / columns are internal to the kernels: covered through dist_fft

    def msm(self, bases, a, b, scalars):
        c = self._ctx(bases, 2, 2)
        try:
            return self.orc.normalize(c.msm(a, b, scalars))
        finally:
            c.close()

    def commit(self, bases, coeffs):
        c = self._ctx(bases, 2, 2)
        try:
            return self.orc.normalize(c.commit(coeffs))
        finally:
            c.close()

    def decompress(self, comp):
        from distributed_plonk_b200._binding import Context
        c = Context(self.lib, self.device, 0, 1)
        try:
            c.init_compressed(comp, 4, 4, True)
            return c.get_bases(0, comp.shape[0])
        finally:
            c.close()


def check(records, impl, orc, max_log=99):
    """every record of the file against impl, byte for byte; returns how many were compared"""
    n = 0
    for tag, (a, b, c, d), blobs in records:
        if tag == LAYOUT:
            assert (a, b, c, d) == (32, 104, 144, 32), f"struct sizes {(a, b, c, d)}: the raw layouts of utils.rs:27-43 differ from 32/104/144/32"
            one, seven = fr(blobs[0])[0], fr(blobs[1])[0]
            assert np.array_equal(orc.from_repr(np.array([[1, 0, 0, 0]], dtype=np.uint64))[0], one), "Fr::one() bytes (Montgomery R mod r)"
            assert np.array_equal(orc.from_repr(np.array([[7, 0, 0, 0]], dtype=np.uint64))[0], seven)
            assert np.array_equal(np.frombuffer(blobs[5], dtype=np.uint64), np.array([7, 0, 0, 0], dtype=np.uint64)), "BigInteger256 is canonical"
            gen, inf = pts(blobs[2], 104)[0], pts(blobs[3], 104)[0]
            assert np.array_equal(gen[:96], orc.g1_generator()[:96]) and gen[96] == 0, "G1Affine: x at 0, y at 48, infinity flag at 96"
            assert inf[96] == 1 and np.array_equal(orc.affine_to_jacobian(inf), pts(blobs[4], 144)[0]), "identity encodings"
        elif tag in (NTT, DIST_FFT):
            if a > max_log:
                continue
            f = impl.ntt if tag == NTT else impl.dist_fft
            x, want = fr(blobs[0]), fr(blobs[1])
            assert x.shape[0] == d
            assert np.array_equal(f(x, a, bool(b), bool(c)), want), f"record tag {tag} log={a} inv={b} coset={c} n_in={d}"
        elif tag in (FFT1, FFT2):
            f = impl.fft1 if tag == FFT1 else impl.fft2
            if f is None or a > max_log:
                continue
            assert np.array_equal(f(fr(blobs[0]), b, a, bool(c), bool(d)), fr(blobs[1])), f"helper record tag {tag} log={a} i={b}"
        elif tag == MSM:
            bases, sc = pts(blobs[0], 104), fr(blobs[1])
            got = impl.msm(bases, a, b, sc)
            assert np.array_equal(got, orc.normalize(pts(blobs[2], 144)[0])), f"MSM [{a},{b}) vs the raw GroupProjective"
            assert np.array_equal(got, pts(blobs[3], 104)[0]), f"MSM [{a},{b}) vs into_affine()"
        elif tag == COMMIT:
            got = impl.commit(pts(blobs[0], 104), fr(blobs[1]))
            assert np.array_equal(got, pts(blobs[2], 104)[0]), "commit_polynomial"
        elif tag == COMPRESSED:
            raw, comp = pts(blobs[0], 104), pts(blobs[1], 48)
            assert np.array_equal(orc.g1_compress(raw), comp), "ark-serialize compressed encoding"
            got = impl.decompress(comp)
            assert np.array_equal(orc.g1_compress(got), comp) and np.array_equal(got[:, 96], raw[:, 96]), "decompression"
        else:
            raise AssertionError(f"unknown record tag {tag}")
        n += 1
    return n


def make_from_oracle(orc, small=True):
    """the record set of rust/dump_fixtures.rs with the ORACLE as the producer (inputs differ: ChaCha20 is Rust-side)"""
    O, rec, seed = OracleImpl(orc), [], [1000]

    def rnd_fr(n):
        seed[0] += 1
        return orc.gen_fr(seed[0], n)

    inf = np.zeros(104, dtype=np.uint8)
    inf[96] = 1
    inf[48:96] = orc.affine_to_jacobian(inf)[48:96]            # GroupAffine::zero() = (0, 1, true)
    rec.append((LAYOUT, (32, 104, 144, 32), [
        orc.from_repr(np.array([[1, 0, 0, 0]], dtype=np.uint64)).tobytes(), orc.from_repr(np.array([[7, 0, 0, 0]], dtype=np.uint64)).tobytes(),
        orc.g1_generator().tobytes(), inf.tobytes(), orc.affine_to_jacobian(inf).tobytes(), np.array([7, 0, 0, 0], dtype=np.uint64).tobytes()]))
    for log in ((0, 1, 3, 6, 9) if small else (0, 1, 3, 6, 9, 11, 12, 15, 16)):
        for inv, cos in FLAGS:
            for n_in in (1 << log, max(1, (1 << log) // 8)):
                x = rnd_fr(n_in)
                rec.append((NTT, (log, int(inv), int(cos), n_in), [x.tobytes(), O.ntt(x, log, inv, cos).tobytes()]))
                if log >= 3:
                    rec.append((DIST_FFT, (log, int(inv), int(cos), n_in), [x.tobytes(), O.dist_fft(x, log, inv, cos).tobytes()]))
    for log in (6, 9):
        r = 1 << (log >> 1)
        c = (1 << log) // r
        for inv, cos in FLAGS:
            for i in (0, 1, r - 1):
                row = rnd_fr(c)
                rec.append((FFT1, (log, i, int(inv), int(cos)), [row.tobytes(), O.fft1(row, i, log, inv, cos).tobytes()]))
            for i in (0, 2, c - 1):
                col = rnd_fr(r)
                rec.append((FFT2, (log, i, int(inv), int(cos)), [col.tobytes(), O.fft2(col, i, log, inv, cos).tobytes()]))
    for n in (1, 33, 600):
        bases = orc.gen_bases(seed[0] + n, n, min(n, 64), n > 3)
        seed[0] += 1
        sc = orc.gen_fr(seed[0], n, False)
        for a, b in ((0, n), (n // 3, n - n // 4)):
            res = orc.msm(bases[a:b], sc[: b - a])
            rec.append((MSM, (a, b, 0, 0), [bases.tobytes(), sc[: b - a].tobytes(), res.tobytes(), orc.normalize(res).tobytes()]))
        co = rnd_fr(max(1, n - n // 5))
        rec.append((COMMIT, (co.shape[0], 0, 0, 0), [bases.tobytes(), co.tobytes(), orc.normalize(orc.commit(bases, co)).tobytes()]))
    raw = orc.gen_bases(4242, 40, 40, True)
    rec.append((COMPRESSED, (40, 0, 0, 0), [raw.tobytes(), orc.g1_compress(raw).tobytes()]))
    return rec
