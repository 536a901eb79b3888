"""Reference-side pin (rust/README.md): when `tests/golden/ref_v1.bin` - written by the reference's own arkworks build
through rust/dump_fixtures.rs - is present, the oracle (always), the emulator build and (with -m gpu) the CUDA
library must reproduce every record byte for byte.  Without the file those tests skip; the reader / checker are
exercised either way on a file of the same format that the oracle writes."""
import os

import pytest

from tests import ref_fixture as rf

REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_v1.bin")
needs_ref = pytest.mark.skipif(not os.path.exists(REF), reason="tests/golden/ref_v1.bin absent: produce it with rust/dump_fixtures.rs (needs cargo)")


def test_format_round_trip_and_checker(orc, emul_lib, tmp_path):
    recs = rf.make_from_oracle(orc)
    path = str(tmp_path / "oracle_made.bin")
    rf.write(path, recs)
    back = rf.read(path)
    assert len(back) == len(recs) and all(a[0] == b[0] and tuple(a[1]) == tuple(b[1]) for a, b in zip(back, recs))
    assert rf.check(back, rf.OracleImpl(orc), orc) == len(recs)
    n = rf.check(back, rf.LibraryImpl(orc, emul_lib), orc, max_log=6)     # the library (emulator build) on the same records
    assert n >= len(recs) - 90
    # a flipped bit in an expected output must be caught
    tag, p, blobs = back[5]
    bad = bytearray(blobs[1])
    bad[0] ^= 1
    back[5] = (tag, p, [blobs[0], bytes(bad)])
    with pytest.raises(AssertionError):
        rf.check(back, rf.OracleImpl(orc), orc)


@needs_ref
def test_reference_fixture_pins_the_oracle(orc):
    assert rf.check(rf.read(REF), rf.OracleImpl(orc), orc) > 0


@needs_ref
def test_reference_fixture_pins_the_emulated_kernels(orc, emul_lib):
    assert rf.check(rf.read(REF), rf.LibraryImpl(orc, emul_lib), orc, max_log=12) > 0


@pytest.mark.gpu
@needs_ref
def test_reference_fixture_pins_the_cuda_library(orc, gpu_lib):
    assert rf.check(rf.read(REF), rf.LibraryImpl(orc, gpu_lib), orc) > 0


@pytest.mark.gpu
def test_oracle_made_fixture_on_the_cuda_library(orc, gpu_lib, tmp_path):
    recs = rf.make_from_oracle(orc, small=False)
    assert rf.check(recs, rf.LibraryImpl(orc, gpu_lib), orc) > 100
