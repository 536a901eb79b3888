"""Short randomised run of tools/fuzz_emul.py (fixed seed, ~12 s): random API sequences on the
kernel-logic emulator against the oracle.  Longer runs: `python tools/fuzz_emul.py <seed> <seconds>`."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fuzz_emulator_short():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_emul.py"), "20260922", "12"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "fails 0" in r.stdout
