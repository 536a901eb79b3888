"""Full-size check of the 2-D transform on ONE GPU for a given domain: the single-worker path (dp_fft_dev) and the
sharded path with W workers held in one process (dp_fft_dev_rows -> in-process all-to-all -> dp_fft_dev_cols), both
spot-checked against the oracle's O(N) Horner evaluation.  python tools/debug_fft_shapes.py LOG_N W"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import distributed_plonk_b200 as dp  # noqa: E402
from distributed_plonk_b200 import parallel  # noqa: E402
from oracle import loader as orc  # noqa: E402

L, W = int(sys.argv[1]), int(sys.argv[2])
orc.build()
orc.set_num_threads(os.cpu_count() or 1)
lib = dp.load()
N = 1 << L
r = 1 << (L >> 1)
c = N // r
gen = torch.Generator(device="cuda")
gen.manual_seed(1234)
x = torch.randint(-(1 << 63), (1 << 63) - 1, (N, 4), dtype=torch.int64, device="cuda", generator=gen)
x[:, 3] &= (1 << 62) - 1
x_host = x.cpu().numpy().view(np.uint64)


def rows_of(p, Wn):
    rows = r // Wn
    return x.view(c, r, 4)[:, p * rows:(p + 1) * rows, :].permute(1, 0, 2).contiguous().view(-1, 4)


def check(name, outs, Wn, inv, coset):
    cols = c // Wn
    ks, vals = [], []
    for p in range(Wn):
        for k2, k1 in ((0, 1), (cols - 1, r - 1), (cols // 2, (7919 * (p + 1)) % r), (cols // 3, (104729 * (p + 3)) % r)):
            ks.append(p * cols + k2 + c * k1)
            vals.append(outs[p][k2 * r + k1].cpu().numpy().view(np.uint64))
    want = orc.ntt_outputs_at(x_host, N, np.array(ks, dtype=np.uint64), inv, coset)
    bad = [i for i in range(len(ks)) if not np.array_equal(vals[i], want[i])]
    print(f"{name} L={L} inv={inv} coset={coset}: {len(ks) - len(bad)}/{len(ks)} positions ok" + (f"  BAD ranks/positions {[(i // 4, i % 4) for i in bad]}" if bad else ""), flush=True)


for inv, coset in ((True, False), (False, False), (True, True)):
    if W == 1 or os.environ.get("ALSO_SINGLE"):
        ctx = dp.Context(lib, 0, 0, 1)
        ctx.init(np.zeros(0, dtype=np.uint8), N, 1 << 10)
        out = torch.empty((N, 4), dtype=torch.int64, device="cuda")
        rows0 = rows_of(0, 1)
        torch.cuda.synchronize()
        ctx.fft_dev(rows0.data_ptr(), out.data_ptr(), False, inv, coset)
        check("single worker", [out], 1, inv, coset)
        ctx.close()
        del out
    if W > 1:
        ctxs = [dp.Context(lib, 0, p, W) for p in range(W)]
        bufs, ins = [], []
        for p, cx in enumerate(ctxs):
            cx.init(np.zeros(0, dtype=np.uint8), N, 1 << 10)
            ins.append(rows_of(p, W))
            torch.cuda.synchronize()
            bufs.append(cx.fft_dev_rows(ins[-1].data_ptr(), False, inv, coset))
        blk = bufs[0][2] * 32
        for p in range(W):
            for q in range(W):
                parallel.as_tensor(bufs[q][1] + p * blk, blk, True).copy_(parallel.as_tensor(bufs[p][0] + q * blk, blk, True))
        torch.cuda.synchronize()
        outs = []
        for p, cx in enumerate(ctxs):
            o = torch.empty(((c // W) * r, 4), dtype=torch.int64, device="cuda")
            cx.fft_dev_cols(o.data_ptr())
            outs.append(o)
        check(f"{W} workers in process", outs, W, inv, coset)
        for cx in ctxs:
            cx.close()
        del outs, ins
    torch.cuda.empty_cache()
