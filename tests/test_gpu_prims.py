"""GPU parity of the element-wise primitives Pprims::copy / Pprims::fill (SURVEY.md section 8f row 3;
reference: Pprims.cpp:31-121, PprimsKernels.cl:9-48) against the CPU restatement of their loops.  Bit-exact."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po  # noqa: E402  (the checker)

F4 = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("w", "<f4")])
SIZES = [0, 1, 3, 4, 5, 255, 1024, 4099, 100003, (1 << 22) + 3]


@pytest.fixture(scope="module")
def ctx():
    import oclradixsort_b200 as ob
    d = ob.DeviceUtils.allocate(ob.TYPE_CL)
    p = ob.Pprims()
    yield ob, d, p
    p.release()
    ob.DeviceUtils.deallocate(d)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("dtype", [np.int32, np.uint32])
def test_copy_and_fill_4byte(ctx, n, dtype):
    ob, d, p = ctx
    rng = np.random.default_rng(n + 5)
    cap = n + 9  # elements behind n must stay untouched
    src = rng.integers(0, 2**32, size=cap, dtype=np.uint64).astype(np.uint32).view(dtype)
    old = rng.integers(0, 2**32, size=cap, dtype=np.uint64).astype(np.uint32).view(dtype)
    bs, bd = ob.Buffer(d, cap, dtype), ob.Buffer(d, cap, dtype)
    bs.write(src)
    bd.write(old)
    p.copy(d, bd, bs, n)
    d.waitForCompletion()
    assert np.array_equal(bd.read(), po.copy_elems(old, src, n))
    value = dtype(-7) if dtype == np.int32 else dtype(0xDEADBEEF)
    p.fill(d, bd, value, n)
    d.waitForCompletion()
    assert np.array_equal(bd.read(), po.fill_elems(po.copy_elems(old, src, n), value, n))
    bs.release(); bd.release()


@pytest.mark.parametrize("n", [0, 1, 2, 1023, 65537])
def test_copy_and_fill_float4(ctx, n):
    ob, d, p = ctx
    rng = np.random.default_rng(n)
    cap = n + 3
    src = rng.standard_normal((cap, 4)).astype(np.float32).view(F4).reshape(cap)
    old = rng.standard_normal((cap, 4)).astype(np.float32).view(F4).reshape(cap)
    bs, bd = ob.Buffer(d, cap, F4), ob.Buffer(d, cap, F4)
    bs.write(src)
    bd.write(old)
    p.copy(d, bd, bs, n)
    d.waitForCompletion()
    want = po.copy_elems(old, src, n)
    assert bd.read().tobytes() == want.tobytes()
    value = np.array((1.5, -2.25, 3.0e-9, float("inf")), dtype=F4)[()]
    p.fill(d, bd, (1.5, -2.25, 3.0e-9, float("inf")), n)
    d.waitForCompletion()
    assert bd.read().tobytes() == po.fill_elems(want, value, n).tobytes()
    bs.release(); bd.release()


@pytest.mark.parametrize("dst_off,src_off", [(0, 0), (1, 1), (3, 3), (1, 2), (0, 3), (2, 0)])
def test_copy_unaligned_subbuffers(ctx, dst_off, src_off):
    """Pointers that are only 4-byte aligned: equal misalignment takes the 128-bit body with scalar ends, different
    misalignment the element-wise kernel; bytes outside [0, n) are never touched."""
    ob, d, p = ctx
    from oclradixsort_b200._lib import check, lib
    n, cap = 70001, 70016
    rng = np.random.default_rng(dst_off * 7 + src_off)
    src = rng.integers(0, 2**32, size=cap, dtype=np.uint64).astype(np.uint32)
    old = rng.integers(0, 2**32, size=cap, dtype=np.uint64).astype(np.uint32)
    bs, bd = ob.Buffer(d, cap, np.uint32), ob.Buffer(d, cap, np.uint32)
    bs.write(src)
    bd.write(old)
    check(lib().b200rs_copy_u32(d.handle, ctypes.c_void_p(bd.m_ptr + 4 * dst_off), ctypes.c_void_p(bs.m_ptr + 4 * src_off), n), "b200rs_copy_u32")
    d.waitForCompletion()
    want = old.copy()
    want[dst_off:dst_off + n] = src[src_off:src_off + n]
    assert np.array_equal(bd.read(), want)
    check(lib().b200rs_fill_u32(d.handle, ctypes.c_void_p(bd.m_ptr + 4 * dst_off), 0x01020304, n), "b200rs_fill_u32")
    d.waitForCompletion()
    want[dst_off:dst_off + n] = 0x01020304
    assert np.array_equal(bd.read(), want)
    bs.release(); bd.release()


def test_prims_reject_bad_arguments(ctx):
    ob, d, p = ctx
    from oclradixsort_b200._lib import lib
    b = ob.Buffer(d, 64, np.uint32)
    words = (ctypes.c_uint32 * 4)(1, 2, 3, 4)
    assert lib().b200rs_copy_u32(d.handle, ctypes.c_void_p(b.m_ptr + 2), ctypes.c_void_p(b.m_ptr), 4) == -1   # 2-byte aligned
    assert lib().b200rs_copy_u128(d.handle, ctypes.c_void_p(b.m_ptr + 4), ctypes.c_void_p(b.m_ptr), 1) == -1  # float4 needs 16
    assert lib().b200rs_fill_u128(d.handle, ctypes.c_void_p(b.m_ptr + 8), words, 1) == -1
    assert lib().b200rs_fill_u32(None, ctypes.c_void_p(b.m_ptr), 0, 1) == -1
    assert lib().b200rs_copy_u32(d.handle, None, None, 0) == 0  # n == 0 is a no-op
    b.release()


def test_cpp_dropin_prims_stopwatch_profile_csv(tmp_path):
    """The C++ side of the same rows: tests/cpp/prims_dropin.cpp (uArray copy/fill, adl::Stopwatch on device events,
    Device::writeProfileCsv), built by `make prims_test` against include/ + libb200rs.so."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tools", "_build", "prims_dropin")
    if not os.path.exists(exe):
        subprocess.run(["make", "-s", "-C", root, "prims_test"], check=True)
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "PRIMS DROPIN OK" in r.stdout, (r.stdout + r.stderr)[-2000:]
    rows = open(tmp_path / "gpurun_out_profile_test.csv").read().strip().splitlines()
    assert len(rows) >= 8 and rows[0].startswith('"digit_histogram_keys"'), rows[:3]
    assert all(len(r.split(",")) == 5 for r in rows)


def test_buffer_map_unmap_pinned_ring():
    """adl.Buffer.getHostPtr / returnHostPtr (the reference caller's sequence, UnitTest/main.cpp:118-139) through the pinned
    staging ring: a fresh buffer is mapped without a copy, unmap does not block, two buffers filled back to back keep their
    contents, the sorted result comes back through a second mapping."""
    import oclradixsort_b200 as ob
    from oracle import pyoracle as po
    d = ob.DeviceUtils.allocate(ob.TYPE_CL)
    p = ob.Pprims()
    n = 300007
    rng = np.random.default_rng(2)
    keys = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    a, b = ob.Buffer(d, n, np.uint32), ob.Buffer(d, n, np.uint32)
    m = a.getHostPtr(n)          # never written: no device -> host copy, the view can be filled at once
    m[:] = keys
    a.returnHostPtr(m)           # returns without waiting
    m2 = b.getHostPtr(n)         # must not be the block the copy above is still reading
    m2[:] = keys[::-1]
    b.returnHostPtr(m2)
    p.radixSort(d, a, n)
    m = a.getHostPtr(n)
    d.waitForCompletion()
    assert np.array_equal(m, po.sort_u32(keys))
    a.returnHostPtr(m)
    assert np.array_equal(b.read(), keys[::-1])
    a.release(); b.release(); p.release()
    ob.DeviceUtils.deallocate(d)
