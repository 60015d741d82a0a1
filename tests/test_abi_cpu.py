"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/b200rs.h declares, and refuses to work without a GPU instead of falling back."""
import ctypes
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from oclradixsort_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.run(["make", "-s", "-C", ROOT], check=True)
    return _lib.lib()


def test_every_declared_symbol_is_exported_and_bound(L):
    from oclradixsort_b200 import _lib
    declared = _lib.declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(L, name), f"libb200rs.so does not export {name}"
        assert name in _lib.SIGNATURES, f"{name} declared in b200rs.h but not bound in _lib.py"
    assert set(_lib.SIGNATURES) == set(declared)


def test_version_and_error_strings(L):
    assert L.b200rs_version() == 100
    assert L.b200rs_error_string(0) == b"ok"
    assert b"temp" in L.b200rs_error_string(-2)


def test_size_queries_need_no_gpu_but_compute_fails_loudly(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-box behaviour")
    n = ctypes.c_int(-1)
    assert L.b200rs_device_count(ctypes.byref(n)) != 0 and n.value == 0
    dev = ctypes.c_void_p()
    assert L.b200rs_device_create(0, ctypes.byref(dev)) != 0 and not dev.value  # no device => error, not a CPU device
    import oclradixsort_b200 as ob
    with pytest.raises(Exception):
        ob.DeviceUtils.allocate(ob.TYPE_CL)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package or include/ may reference it."""
    bad = []
    for base in ("oclradixsort_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".h", ".cpp", ".inl", ".cuh")):
                    text = open(os.path.join(dirpath, fn), errors="ignore").read()
                    if "pyoracle" in text or "radixsort_oracle" in text or "liboracle" in text:
                        bad.append(os.path.join(dirpath, fn))
    assert not bad, bad


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/b200rs.h must compile as C99 on its own (what a cgo / JNI / ctypes-cffi binding sees)."""
    src = tmp_path / "t.c"
    src.write_text('#include "b200rs.h"\nint main(void) { b200rs_pair p = {1u, 2u}; b200rs_profile_entry e; (void)e; return (int)(p.key + p.value) - 3 + (B200RS_OK); }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-fsyntax-only", str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_dropin_headers_compile_as_cxx11(tmp_path):
    """include/Adl + include/Tahoe are what a C++ user of the reference includes: a TU that only includes them must compile
    with the reference's language level (premake4.lua builds with the compiler default of its day; -std=c++11 here)."""
    src = tmp_path / "t.cpp"
    src.write_text('#include <Adl/Adl.h>\n#include <Tahoe/ParallelPrimitives/Pprims.h>\n#include <Tahoe/Algorithm/Sort/RadixSort.h>\n'
                   'char adl::s_cacheDirectory[128];\nint main() { adl::Stopwatch sw; Tahoe::SortData d(1u, 2u); return (int)d.m_key - 1 + sw.getNIntervals() + 1; }\n')
    r = subprocess.run(["g++", "-std=c++11", "-Wall", "-I", os.path.join(ROOT, "include"), "-fsyntax-only", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
