"""Exploratory timing on a GPU box (not the bench): prints device-side ms of the main entry points."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import distributed_plonk_b200 as dp
from oracle import loader as orc

lib = dp.load()
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = (1 << logn) + 32
t0 = time.time()
ctx = dp.Context(lib, 0, 0, 1)
bases = ctx.gen_bases(5, n)
print("gen bases (gpu, distinct)", round(time.time() - t0, 2), flush=True)
t0 = time.time()
ctx.init(bases, 1 << logn, 1 << (logn + 3))
print("init ms", round((time.time() - t0) * 1e3, 1), ctx.last_timing(), flush=True)

def dev(arr):
    return torch.from_numpy(arr.view(np.int64)).cuda()

for L in (logn, logn + 3):
    x = dev(orc.gen_fr(1, 1 << L))
    y = torch.empty_like(x)
    for (inv, cos) in ((0, 0), (0, 1), (1, 1)):
        for rep in range(3):
            ctx.fft_dev(x.data_ptr(), y.data_ptr(), L == logn + 3, bool(inv), bool(cos))
        ms, nl = ctx.last_timing()
        print(f"fft_dev 2^{L} inv={inv} coset={cos}: {ms:.3f} ms, {nl} launches, {64*(1<<L)*nl/ms/1e6:.0f} GB/s algorithmic", flush=True)
    z = x.clone()
    for rep in range(3):
        ctx.ntt_dev(z.data_ptr(), L, False, False)
    ms, nl = ctx.last_timing()
    print(f"ntt_dev 2^{L}: {ms:.3f} ms, {nl} launches", flush=True)
    del x, y, z

out = torch.zeros(18, dtype=torch.int64, device="cuda")
for kind in ("uniform", "witness"):
    sc = orc.gen_fr(3, n, False)
    if kind == "witness":
        sc[::2] = 0
        sc[1::10] = [1, 0, 0, 0]
    scd = dev(sc)
    for m in (n, (1 << (logn - 2)) + 32):
        for rep in range(3):
            ctx.msm_dev(0, m, scd.data_ptr(), m, out.data_ptr())
        ms, nl = ctx.last_timing()
        print(f"msm_dev {kind} n={m}: {ms:.3f} ms, {nl} launches, sort/accumulate/reduce ms = {[round(v,3) for v in ctx.msm_breakdown()]}", flush=True)
torch.cuda.synchronize()
print("done")
