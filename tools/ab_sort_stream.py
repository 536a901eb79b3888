"""A/B of the MSM digit sort on its own stream (DP_MSM_SORT_STREAM=1, default) against the sort queued in front of its
accumulation on the compute stream (=0): batches of 5 MSMs, the full 2^22+32 set on one worker and one worker's shard of 8.
  python tools/ab_sort_stream.py [log_n]"""
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
results = {}
for W in (1, 8):
    for own in (1, 0, 1, 0):
        os.environ["DP_MSM_SORT_STREAM"] = str(own)
        ctx = dp.Context(lib, 0, 0, W)
        bases = torch.empty((nb, 104), dtype=torch.uint8, device="cuda")
        ctx.gen_bases_into(0xD15791B07E5EED, nb, bases.data_ptr())
        torch.cuda.synchronize()
        ctx.init_ptr(bases.data_ptr(), nb, n, 8 * n)
        del bases
        lo, hi = 0, nb // W
        gen = torch.Generator(device="cuda")
        gen.manual_seed(1234 + W)                  # the same scalars for both settings: the results must then be identical
        sc = [torch.randint(-(1 << 63), (1 << 63) - 1, (hi - lo, 4), dtype=torch.int64, device="cuda", generator=gen) for _ in range(3)]
        for s in sc:
            s[:, 3] &= (1 << 62) - 1
        outs = [torch.zeros(18, dtype=torch.int64, device="cuda") for _ in range(5)]
        torch.cuda.synchronize()
        jobs = [(lo, hi, sc[j % 3].data_ptr(), hi - lo, outs[j].data_ptr()) for j in range(5)]
        for _ in range(2):
            ctx.msm_dev_batch(jobs)
        t = []
        for _ in range(5):
            ctx.msm_dev_batch(jobs)
            t.append(ctx.last_timing()[0] / 5)
        res = [o.cpu().numpy().tobytes() for o in outs]
        key = (W, own)
        results.setdefault(key, []).append(float(np.median(t)))
        results.setdefault(("out", W), []).append(res)
        print(f"W={W} shard={hi - lo} sort stream own={own}: batch of 5: {np.median(t):.3f} ms per MSM (min {min(t):.3f})", flush=True)
        ctx.close()
        del sc, outs
        torch.cuda.empty_cache()
for W in (1, 8):
    outs = results[("out", W)]
    print(f"W={W}: results identical across the four runs: {all(o == outs[0] for o in outs)}; "
          f"own stream {np.mean(results[(W, 1)]):.3f} ms vs compute stream {np.mean(results[(W, 0)]):.3f} ms per MSM")
