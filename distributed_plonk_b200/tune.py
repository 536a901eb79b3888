"""MSM pipeline tuning in a child process.

`dp_init` times the plain MSM pipeline against batched-affine tree levels over the context's own window table and keeps
the levels only if they give the identical 144 bytes faster (csrc/dplonk.cu: msm_tune).  By default it compares the two
pipelines that have run on hardware (plain, two levels); the wider search (one, two, three levels: DP_MSM_TUNE=2) is run by
`probe()` in a CHILD process, so that whatever goes wrong there - a crash included - cannot touch the caller's CUDA
context; the caller then forces the answer (DP_MSM_AFFINE).  bench.py does this once per run.

  python -m distributed_plonk_b200.tune DEVICE ME N_WORKERS LOG_N     ->  one JSON line
"""
from __future__ import annotations

import json
import os
import subprocess
import sys

GEN_SEED = 0xD15791B07E5EED


def _child(device: int, me: int, n_workers: int, log_n: int) -> dict:
    from . import Context, load
    lib = load()
    ctx = Context(lib, device, me, n_workers)
    n, nb = 1 << log_n, (1 << log_n) + 32
    cap = (nb * 104 + 31) // 32                      # a device buffer from the library itself: no torch in this process
    ptr = ctx.poly_put(1, __import__("numpy").zeros((0, 4), dtype="uint64"), capacity=cap)
    ctx.gen_bases_into(GEN_SEED, nb, ptr)
    ctx.init_ptr(ptr, nb, n, 8 * n)
    out = ctx.msm_tuning()
    ctx.close()
    return out


def probe(device: int, me: int, n_workers: int, log_n: int, timeout: float = 150.0) -> dict:
    """dp_init's MSM tuning for worker `me` of `n_workers` on GPU `device`, run in a child process.
    Returns its {"plain_ms", "affine_ms", "levels", "equal"} or {"error": ...}."""
    env = dict(os.environ, DP_MSM_TUNE="2")
    env.pop("DP_MSM_AFFINE", None)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
    import time
    t0 = time.perf_counter()
    try:
        r = subprocess.run([sys.executable, "-m", "distributed_plonk_b200.tune", str(device), str(me), str(n_workers), str(log_n)],
                           env=env, capture_output=True, text=True, timeout=timeout, cwd=root)
    except (subprocess.TimeoutExpired, OSError) as exc:
        return {"error": f"probe did not finish: {exc}"[:200]}
    if r.returncode != 0:
        return {"error": f"probe exited with {r.returncode}: {(r.stderr or r.stdout).strip()[-200:]}"}
    try:
        return dict(json.loads(r.stdout.strip().splitlines()[-1]), probe_seconds=round(time.perf_counter() - t0, 2))
    except (ValueError, IndexError):
        return {"error": f"probe printed no result: {r.stdout.strip()[-200:]}"}


def choose(result: dict) -> int:
    """levels to force (DP_MSM_AFFINE) given a probe result: the tuned choice when the two pipelines agreed, else 0"""
    if result.get("equal") == 1 and result.get("levels") in (1, 2, 3):
        return int(result["levels"])
    return 0


if __name__ == "__main__":
    print(json.dumps(_child(*(int(a) for a in sys.argv[1:5]))), flush=True)
