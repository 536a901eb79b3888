"""MSM of ONE worker's shard at W = 1, 2, 4, 8 on a single GPU (a context created as worker 0 of W holds the window
table of its own shard only): per-phase device times and the batch throughput, to see what limits multi-GPU scaling.
  python tools/msm_shard_profile.py [log_n]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import distributed_plonk_b200 as dp  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n, nb = 1 << log_n, (1 << log_n) + 32
lib = dp.load()
for W in (1, 2, 4, 8):
    ctx = dp.Context(lib, 0, 0, W)
    bases = torch.empty((nb, 104), dtype=torch.uint8, device="cuda")
    ctx.gen_bases_into(0xD15791B07E5EED, nb, bases.data_ptr())
    torch.cuda.synchronize()
    ctx.init_ptr(bases.data_ptr(), nb, n, 8 * n)
    del bases
    lo, hi = 0, nb // W
    sc = [torch.randint(-(1 << 63), (1 << 63) - 1, (hi - lo, 4), dtype=torch.int64, device="cuda") for _ in range(3)]
    for s in sc:
        s[:, 3] &= (1 << 62) - 1
    outs = [torch.zeros(18, dtype=torch.int64, device="cuda") for _ in range(5)]
    torch.cuda.synchronize()
    for _ in range(3):
        ctx.msm_dev(lo, hi, sc[0].data_ptr(), hi - lo, outs[0].data_ptr())
    one, br = [], []
    for k in range(6):
        ctx.msm_dev(lo, hi, sc[k % 3].data_ptr(), hi - lo, outs[0].data_ptr())
        one.append(ctx.last_timing()[0])
        br.append(ctx.msm_breakdown())
    batch = []
    for _ in range(4):
        ctx.msm_dev_batch([(lo, hi, sc[j % 3].data_ptr(), hi - lo, outs[j].data_ptr()) for j in range(5)])
        batch.append(ctx.last_timing()[0] / 5)
    b = np.median(np.array(br), axis=0)
    print(f"W={W} shard={hi - lo}: one MSM {np.median(one):.3f} ms (sort {b[0]:.3f} accumulate {b[1]:.3f} tail {b[2]:.3f}); "
          f"batch of 5: {np.median(batch):.3f} ms per MSM; ideal share of the W=1 accumulate: see W=1 line / {W}", flush=True)
    ctx.close()
    del sc, outs
    torch.cuda.empty_cache()
