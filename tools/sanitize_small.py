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
p.release(); ob.DeviceUtils.deallocate(d)
print("sanitize_small: all results correct")
