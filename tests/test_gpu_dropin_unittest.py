"""Runs the reference's UNCHANGED unit test (UnitTest/main.cpp: Demo.Sort32 / Demo.SortKeyValue / Demo.Scan)
compiled against this repo's drop-in headers (include/Adl, include/Tahoe) and linked with libb200rs.so.

The binary is built in the container by `make unittest` (it needs /root/reference/UnitTest/main.cpp and the
vendored gtest, which never enter the repo) into oracle/_ref/UnitTest64 and travels to the GPU box from there.
Its CPU side (Tahoe::RadixSort::sort, the reference's own object code) is the checker; every device result goes
through Pprims -> the C ABI -> the CUDA kernels.  Demo.Scan includes n = 1048576, which the reference's GPU
path cannot do (Pprims.cpp:132-138).
"""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "UnitTest64")


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/UnitTest64 is built only where /root/reference exists (make unittest)")
def test_unchanged_reference_unittest_passes():
    r = subprocess.run([BIN], cwd=os.path.dirname(BIN), capture_output=True, text=True, timeout=600)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert "[  PASSED  ] 3 tests." in r.stdout, tail
    for name in ("Demo.Sort32", "Demo.SortKeyValue", "Demo.Scan"):
        assert f"[       OK ] {name}" in r.stdout, tail
    assert "1024.0K elems" in r.stdout  # the 1M-element cases ran
