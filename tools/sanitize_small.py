"""tools/sanitize_small.py -- small instances of every entry point, for compute-sanitizer (memcheck / racecheck / synccheck):
each result is also compared with numpy, so a sanitizer-clean but wrong run still fails."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oclradixsort_b200 as ob

d = ob.DeviceUtils.allocate(ob.TYPE_CL, 0)
p = ob.Pprims()
rng = np.random.default_rng(3)
for n in (1, 257, 8960 * 2 + 3, 6400 * 9 + 1):
    k = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    for bits in (32, 12):
        b = ob.Buffer(d, n, np.uint32); b.write(k); p.radixSort(d, b, n, bits); d.waitForCompletion()
        mask = np.uint32((1 << bits) - 1) if bits < 32 else np.uint32(0xFFFFFFFF)
        want = k[np.argsort(k & mask, kind="stable")]
        assert np.array_equal(b.read(), want), ("keys", n, bits)
        b.release()
    kv = np.empty(n, dtype=ob.PAIR_DTYPE); kv["key"] = k & np.uint32(0xFFFF00FF); kv["value"] = np.arange(n, dtype=np.uint32)
    b = ob.Buffer(d, n, ob.PAIR_DTYPE); b.write(kv); p.radixSort(d, b, n); d.waitForCompletion()
    assert np.array_equal(b.read(), kv[np.argsort(kv["key"], kind="stable")]), ("pairs", n)
    b.release()
    s, o = ob.Buffer(d, n, np.uint32), ob.Buffer(d, n, np.uint32)
    s.write(k); tot = p.scan(d, o, s, n, sumOut=True)
    c = np.cumsum(k.astype(np.uint64)); want = np.concatenate([[0], c[:-1]]).astype(np.uint64) & 0xFFFFFFFF
    assert np.array_equal(o.read().astype(np.uint64), want) and tot == int(c[-1] & 0xFFFFFFFF), ("scan", n)
    p.fill(d, o, 0xABCD1234, n); p.copy(d, s, o, n); d.waitForCompletion()
    assert np.all(s.read() == np.uint32(0xABCD1234)), ("fill/copy", n)
    s.release(); o.release()

# ---- round 2: the kernels the sizes above no longer reach ----
import ctypes
from oclradixsort_b200._lib import check, lib

# multi-kernel LSD chain (n > 2^20: full-size tiles, histogram + four scatter passes), keys and pairs
n = (1 << 20) + 4099
k = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
b = ob.Buffer(d, n, np.uint32); b.write(k); p.radixSort(d, b, n, 32); d.waitForCompletion()
assert np.array_equal(b.read(), np.sort(k)), "keys, LSD chain"
b.release()
kv = np.empty(n, dtype=ob.PAIR_DTYPE); kv["key"] = k & np.uint32(0xFF00FFFF); kv["value"] = np.arange(n, dtype=np.uint32)
b = ob.Buffer(d, n, ob.PAIR_DTYPE); b.write(kv); p.radixSort(d, b, n); d.waitForCompletion()
assert np.array_equal(b.read(), kv[np.argsort(kv["key"], kind="stable")]), "pairs, LSD chain"
b.release()

# key-only MSD pipeline, forced at small n (histogram, plan, two partition passes, bucket kernel incl. its robust route)
fn = lib().b200rs_sort_keys_u32_msd
for n, kind in ((70001, "uniform"), (300007, "uniform"), (40000, "dups")):
    k = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    if kind == "dups":
        k = (k & np.uint32(0xFFFF0000)) | np.uint32(0x1234)  # every bucket holds copies of one value
    b = ob.Buffer(d, n, np.uint32); b.write(k)
    need, used = ctypes.c_size_t(0), ctypes.c_int(-1)
    check(fn(d.handle, None, n, None, ctypes.byref(need), ctypes.byref(used)), "size")
    t = ob.Buffer(d, need.value, np.uint8)
    check(fn(d.handle, ctypes.c_void_p(b.m_ptr), n, ctypes.c_void_p(t.m_ptr), ctypes.byref(need), ctypes.byref(used)), "msd")
    d.waitForCompletion()
    assert used.value == 1 and np.array_equal(b.read(), np.sort(k)), ("msd", n, kind)
    t.release(); b.release()

# exchange kernel (helper warp, two-level look-back, bulk copies) into 16 and 32 local parts
fx = lib().b200rs_exchange_pairs
for n, parts in ((50021, 16), (3584 * 9 + 5, 32), (3000, 3)):
    kv = np.empty((n, 2), dtype=np.uint32)
    kv[:, 0] = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32); kv[:, 1] = np.arange(n, dtype=np.uint32)
    top = kv[:, 0] >> 24
    lut = (np.arange(256) * parts // 256).astype(np.uint8)
    counts = np.bincount(lut[top], minlength=parts)
    starts = np.cumsum(counts + 1) - (counts + 1)  # one pair of gap: odd and even destination phases
    src = ob.Buffer(d, n, ob.PAIR_DTYPE); src.write(kv.view(ob.PAIR_DTYPE).reshape(-1))
    dst = ob.Buffer(d, n + parts + 8, ob.PAIR_DTYPE)
    lut_d = ob.Buffer(d, 256, np.uint8); lut_d.write(lut)
    base = ob.Buffer(d, parts, np.uint64); base.write((np.uint64(dst.m_ptr) + 8 * starts.astype(np.uint64)).astype(np.uint64))
    need = ctypes.c_size_t(0)
    check(fx(d.handle, None, n, 24, 8, None, None, parts, None, None, ctypes.byref(need)), "size")
    t = ob.Buffer(d, need.value, np.uint8)
    check(fx(d.handle, ctypes.c_void_p(src.m_ptr), n, 24, 8, ctypes.c_void_p(lut_d.m_ptr), ctypes.c_void_p(base.m_ptr), parts, None, ctypes.c_void_p(t.m_ptr),
             ctypes.byref(need)), "exchange")
    d.waitForCompletion()
    got = dst.read().view(np.uint32).reshape(-1, 2)
    want = kv[np.argsort(lut[top], kind="stable")]
    at = 0
    for q in range(parts):
        assert np.array_equal(got[starts[q]:starts[q] + counts[q]], want[at:at + counts[q]]), ("exchange", n, parts, q)
        at += counts[q]
    for x in (src, dst, lut_d, base, t): x.release()
p.release(); ob.DeviceUtils.deallocate(d)
print("sanitize_small: all results correct")
