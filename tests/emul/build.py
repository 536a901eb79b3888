"""Build the CPU kernel-logic emulator of the library (TEST INFRASTRUCTURE ONLY; see cuda_emul.h).

Compiles distributed_plonk_b200/csrc/dplonk.cu with g++ -DDP_EMUL into
tests/emul/_build/libdplonk_emul.so.  Never loaded by the package itself.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "distributed_plonk_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libdplonk_emul.so")
OUT_ASYNC = os.path.join(HERE, "_build", "libdplonk_emul_async.so")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def build(force=False, async_streams=False):
    """async_streams: streams are worker threads, events are real, operations get random delays
    (cuda_emul.h, DP_EMUL_ASYNC) - catches missing stream / event dependencies on a CPU"""
    out = OUT_ASYNC if async_streams else OUT
    return _build(out, ["-DDP_EMUL_ASYNC"] if async_streams else [], force)


def _build(OUT, extra, force):
    deps = [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.join(HERE, "cuda_emul.h"),
                                                             os.path.join(ROOT, "include", "dplonk.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    import fcntl
    lock = open(OUT + ".lock", "w")                  # one builder at a time; released when the process drops the file
    fcntl.flock(lock, fcntl.LOCK_EX)
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    tmp = f"{OUT}.tmp.{os.getpid()}"
    cmd = [CXX, "-O2", "-g", "-std=c++20", "-fPIC", "-shared", "-pthread", "-DDP_EMUL", *extra, "-Wall", "-Wno-unknown-pragmas",
           "-x", "c++", os.path.join(SRC, "dplonk.cu"), "-o", tmp]
    subprocess.check_call(cmd)
    os.replace(tmp, OUT)          # atomic: concurrent builders / loaders never see a half-written library
    return OUT


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, async_streams="--async" in sys.argv))
