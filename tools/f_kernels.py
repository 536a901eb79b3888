"""Drives the kernels of the "next" rows at bench size so that ncu can capture them, and prints their
CUDA-event times:  round-2 grand product (perm_*), quotient evaluations (both variants), p(z) / division by (X - z) /
linear combination (poly_*), SRS decompression (g1_decompress_kernel).
  python tools/f_kernels.py [log_n, default 22]
  ncu --set full --clock-control none -k regex:'perm_|poly_|quotient_kernel|g1_decompress' -c 24 -f -o gpurun_out/f python tools/f_kernels.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import distributed_plonk_b200 as dp  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n, m = 1 << log_n, 8 << log_n
lib = dp.load()
gen = torch.Generator(device="cuda")
gen.manual_seed(7)


def rand_fr(count):
    t = torch.randint(-(1 << 63), (1 << 63) - 1, (count, 4), dtype=torch.int64, device="cuda", generator=gen)
    t[:, 3] &= (1 << 61) - 1          # < 2^253 < r: valid Montgomery residues
    return t


ch = np.array([[3, 1, 4, 1], [5, 9, 2, 6], [5, 3, 5, 8], [9, 7, 9, 3], [2, 3, 8, 4], [6, 2, 6, 4], [3, 3, 8, 3], [2, 7, 9, 5]], dtype=np.uint64)
for table in (1, 0):               # the cached-table variant first (what a worker runs), then the product-tree variant
    os.environ["DP_QUOT_TABLE"] = str(table)
    ctx = dp.Context(lib, 0, 0, 1)
    bases = ctx.gen_bases(5, 64)
    ctx.init(bases, n, m)
    arrs = [rand_fr(m) for _ in range(25)]
    qo = torch.empty((m, 4), dtype=torch.int64, device="cuda")
    ptr = [t.data_ptr() for t in arrs]
    for rep in range(2):
        ctx.quotient_evals_dev(ptr[:13], ptr[13:18], ptr[18:23], ptr[23], ptr[24], ch[:5], ch[5], ch[6], ch[7], qo.data_ptr())
        print(f"quotient_evals 2^{log_n + 3} points, table={table}, call {rep}: {ctx.last_timing()[0]:.3f} ms ({ctx.last_timing()[1]} launches)", flush=True)
    if table == 1:
        for rep in range(2):
            ctx.poly_eval(ptr[0], ch[5], n + 3)
        print(f"poly_eval n+3: {ctx.last_timing()[0]:.3f} ms", flush=True)
        for rep in range(2):
            ctx.poly_div_linear(ptr[0], ch[5], n + 3, qo.data_ptr())
        print(f"poly_div_linear n+3: {ctx.last_timing()[0]:.3f} ms", flush=True)
        for rep in range(2):
            ctx.poly_lincomb(ptr[:12], ch[[0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3]], out_len=n + 3, lens=[n + 3] * 12, out_ptr=qo.data_ptr())
        print(f"poly_lincomb 12 x (n+3): {ctx.last_timing()[0]:.3f} ms", flush=True)
        wt = [rand_fr(5 * n) for _ in range(3)]
        zt = torch.empty((n, 4), dtype=torch.int64, device="cuda")
        for rep in range(2):
            ctx.perm_product_dev(wt[0].data_ptr(), wt[1].data_ptr(), wt[2].data_ptr(), 5, n, ch[0], ch[1], zt.data_ptr())
        print(f"perm_product 5 x n: {ctx.last_timing()[0]:.3f} ms", flush=True)
        del wt, zt
    del arrs, qo
    ctx.close()
    torch.cuda.empty_cache()

# SRS ingest: 2^14 compressed points (ark-serialize 0.3.0: canonical x little-endian, bit 7 of byte 47 = y > -y,
# bit 6 = infinity), produced here from the library's own synthetic bases with Python integers
P_MOD = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R_INV = pow(1 << 384, -1, P_MOD)
ctx = dp.Context(lib, 0, 0, 1)
k = 1 << 14
raw = np.ascontiguousarray(ctx.gen_bases(11, k)).reshape(k, 104)
comp = np.zeros((k, 48), dtype=np.uint8)
for i in range(k):
    x = int.from_bytes(raw[i, :48].tobytes(), "little") * R_INV % P_MOD
    y = int.from_bytes(raw[i, 48:96].tobytes(), "little") * R_INV % P_MOD
    b = bytearray(x.to_bytes(48, "little"))
    if y > P_MOD - y:
        b[47] |= 0x80
    comp[i] = np.frombuffer(bytes(b), dtype=np.uint8)
for rep in range(2):
    ctx.init_compressed(comp, 1 << 10, 1 << 13, True)
print(f"init_compressed 2^14 points with subgroup check: {ctx.last_timing()[0]:.3f} ms (whole dp_init)", flush=True)
back = np.ascontiguousarray(ctx.get_bases(0, k)).reshape(k, 104)
print("decompressed == original:", bool(np.array_equal(back[:, :96], raw[:, :96])), flush=True)
ctx.close()
print("done")
