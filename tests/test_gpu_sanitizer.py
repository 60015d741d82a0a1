"""compute-sanitizer memcheck over small instances of every entry point (tools/sanitize_small.py, which also checks every
result against numpy).  racecheck and synccheck of the same program are recorded in profiles/r2_final_compute_sanitizer.txt."""
import os
import shutil
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_memcheck_small_instances():
    tool = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(tool):
        pytest.skip("compute-sanitizer not installed")
    r = subprocess.run([tool, "--tool", "memcheck", "--print-limit", "10", sys.executable, os.path.join(ROOT, "tools", "sanitize_small.py")],
                       cwd=ROOT, capture_output=True, text=True, timeout=900)
    out = r.stdout + r.stderr
    if "ERROR SUMMARY" not in out:
        pytest.skip("compute-sanitizer could not attach here: " + out[-300:])
    assert "sanitize_small: all results correct" in out, out[-2000:]
    assert "ERROR SUMMARY: 0 errors" in out, out[-3000:]
