"""Stream / event dependencies of the host-side pipeline, checked on a CPU.

tests/emul/build.py --async builds the kernel-logic emulator with REAL asynchronous streams (every
stream is a worker thread, events carry record / completion generations, every operation gets a random
delay).  DP_EMUL_SLOW="i:us" additionally makes one of the context's five streams (0 compute, 1 copy-in,
2 copy-out, 3 MSM tail, 4 MSM digit sort) pathologically slow, so anything that should have waited for it and does not is
practically certain to run too early and produce a wrong result.  (Removing any one of the library's
cudaStreamWaitEvent calls makes these tests fail; the pipeline-heavy emulator tests are re-run under each
of the five adversarial schedules.)  TEST INFRASTRUCTURE ONLY, like the rest of tests/emul."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the pipeline-heavy tests (three copy / compute streams, MSM head / tail streams, deferred zero-fill of short rows);
# DP_TEST_FULL=1 adds the slower transform / resident-round tests, which use the same stream dependencies
SELECT = "host_schedules or async_msm or commit_and_round1 or msm_dev_batch or fused_peer_exchange or (short_rows and limits0)"
if os.environ.get("DP_TEST_FULL", "0") == "1":
    SELECT += " or resident_rounds or like_reference_test_fft or short_rows"


@pytest.mark.timeout(1500)
def test_pipeline_under_adversarial_stream_schedules():
    from tests.emul import build as emul_build
    emul_build.build(async_streams=True)            # build once, before the children race for it
    procs = []
    # the five streams with the default MSM pipeline, then the optional sort stream (DP_MSM_SORT_STREAM=1) with the two
    # stream itself made slow
    configs = [(0, "0"), (1, "0"), (2, "0"), (3, "0"), (4, "1")]
    if os.environ.get("DP_TEST_FULL", "0") != "1":
        configs = [(0, "0"), (1, "0"), (3, "0"), (4, "1")]      # (the copy-out stream only feeds dp_fft2's blocking read)
    for slow, sort_stream in configs:
        env = dict(os.environ, DP_TEST_EMUL_ASYNC="1", DP_EMUL_SLOW=f"{slow}:1500", DP_MSM_SORT_STREAM=sort_stream)
        procs.append(subprocess.Popen(
            [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_emul_kernels.py"), "-q", "-x", "-p", "no:cacheprovider",
             "-k", SELECT], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for slow, p in enumerate(procs):
        out, _ = p.communicate()
        assert p.returncode == 0, f"adversarial schedule {slow}:\n{out[-3000:]}"
        assert " passed" in out and "failed" not in out
