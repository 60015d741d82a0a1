"""tests/golden/make_golden.py -- regenerates tests/golden/reference_hashes.json.

Run in the build container only (needs oracle/_ref, i.e. /root/reference):
    python tests/golden/make_golden.py

Inputs are the reference unit test's own (UnitTest/main.cpp:105-109,122,144,152,183: srand(123),
glibc rand() through getRandom); outputs come from the UNMODIFIED reference compiled into
oracle/_ref/libref_oclradixsort.so, through its Adl Host-backend route
(DeviceUtils::allocate(TYPE_HOST) -> Pprims::radixSort -> RadixSort::sort).  The scan has no Host
path in the reference (Pprims.cpp:124-127), so its expected output is the serial running sum the
test itself checks against (UnitTest/main.cpp:193-199), computed here with numpy.
Hash = FNV-1a-64 of the raw little-endian bytes, the same table as SURVEY.md section 8c.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import pyoracle as po  # noqa: E402

SIZES = [1024 << k for k in range(11)]  # main.cpp:105


def kv_sizes():
    """main.cpp:144 does `testSize += 13` on the LOOP VARIABLE: 1037, 2087, 4187, ... 1075187."""
    t, out = 1024, []
    while t < 2 * 1024 * 1024:
        t += 13
        out.append(t)
        t *= 2
    return out


def main():
    assert po.have_ref(), "oracle/_ref missing: run `make -C oracle ref` in the build container"
    out = {"hash": "fnv1a64 over raw little-endian bytes", "seed": 123, "sort32": [], "sortkeyvalue": [], "scan": []}
    for n, m in zip(SIZES, kv_sizes()):
        keys = po.gen_sort32(n)
        srt = po.ref_sort_u32(keys, host_backend=True)
        assert np.array_equal(srt, po.ref_sort_u32(keys, host_backend=False))
        out["sort32"].append({"n": n, "in": f"{po.fnv1a64(keys):016x}", "out": f"{po.fnv1a64(srt):016x}",
                              "first": f"{srt[0]:08x}", "last": f"{srt[-1]:08x}"})
        kv = po.gen_keyvalue(m)
        skv = po.ref_sort_pairs(kv, host_backend=True)
        assert np.array_equal(skv, po.ref_sort_pairs(kv, host_backend=False))
        dup = int(np.count_nonzero(skv["key"][1:] == skv["key"][:-1]))
        out["sortkeyvalue"].append({"n": m, "in": f"{po.fnv1a64(kv):016x}", "out": f"{po.fnv1a64(skv):016x}",
                                    "first_key": f"{skv['key'][0]:08x}", "first_value": int(skv["value"][0]),
                                    "adjacent_equal_keys": dup})
        s = po.gen_scan(n)
        c = np.cumsum(s.astype(np.uint64)).astype(np.uint32)
        ex = np.concatenate([[np.uint32(0)], c[:-1]]).astype(np.uint32)
        out["scan"].append({"n": n, "in": f"{po.fnv1a64(s):016x}", "out": f"{po.fnv1a64(ex):016x}", "total": int(c[-1])})
    # Low-entropy fixtures (many duplicate keys => stability is really exercised; the uniform
    # unit-test data above has almost none).  Small enough to commit in full.
    rng = np.random.default_rng(20261017)
    n = 4099
    base = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    fx = {}
    for name, mask in (("mask3f", 0x3f), ("maskffff", 0xffff), ("maskff00ff00", 0xff00ff00), ("allequal", 0)):
        kv = np.empty(n, dtype=po.PAIR_DTYPE)
        kv["key"] = (base & np.uint32(mask)) if mask else np.uint32(0xdeadbeef)
        kv["value"] = np.arange(n, dtype=np.uint32)
        fx[f"kv_{name}_in"] = kv.view(np.uint32).reshape(n, 2)
        fx[f"kv_{name}_out"] = po.ref_sort_pairs(kv, host_backend=True).view(np.uint32).reshape(n, 2)
        k = np.ascontiguousarray(kv["key"])
        fx[f"keys_{name}_in"] = k
        fx[f"keys_{name}_out"] = po.ref_sort_u32(k, host_backend=True)
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "stability_fixtures.npz"), **fx)
    path = os.path.join(os.path.dirname(__file__), "reference_hashes.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
