"""GPU parity tests: the CUDA path, called through the C ABI (via the host-side mirror of the
reference interface), against the CPU oracle on the same seeded inputs.  Bit-exact: integer work.

Reference test being mirrored: UnitTest/main.cpp:105-205 (Demo.Sort32 / Demo.SortKeyValue / Demo.Scan).
"""
import ctypes
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po  # noqa: E402  (the checker, never the thing under test)


@pytest.fixture(scope="module")
def ctx():
    import oclradixsort_b200 as ob
    d = ob.DeviceUtils.allocate(ob.TYPE_CL)
    p = ob.Pprims()
    yield ob, d, p
    p.release()
    ob.DeviceUtils.deallocate(d)


def _keys(kind: str, n: int, seed: int = 1) -> np.ndarray:
    rng = np.random.default_rng(seed + n)
    u = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    if kind == "uniform":
        return u
    if kind == "lowbyte":
        return u & np.uint32(0xFF)
    if kind == "few":
        return (u & np.uint32(0x3)) * np.uint32(0x01010101)
    if kind == "allequal":
        return np.full(n, 0xDEADBEEF, dtype=np.uint32)
    if kind == "allmax":
        return np.full(n, 0xFFFFFFFF, dtype=np.uint32)
    if kind == "sorted":
        return np.sort(u)
    if kind == "reversed":
        return np.sort(u)[::-1].copy()
    if kind == "strided":  # presorted arithmetic progression c*i (BASELINE config 4): every digit equally frequent in every tile
        return (np.arange(n, dtype=np.uint64) * np.uint64(2**32 // max(n, 1))).astype(np.uint32)
    if kind == "strided_reversed":
        return (np.arange(n, dtype=np.uint64) * np.uint64(2**32 // max(n, 1))).astype(np.uint32)[::-1].copy()
    if kind == "and3":
        a = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
        b = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
        return u & a & b
    raise ValueError(kind)


SIZES = [1, 2, 31, 32, 33, 255, 256, 257, 1000, 4095, 4096, 4097, 6399, 6400, 6401, 8191, 8192, 8193, 8959, 8960, 8961, 100003, 1 << 20, (1 << 22) + 7]
KINDS = ["uniform", "lowbyte", "few", "allequal", "allmax", "sorted", "reversed", "and3", "strided", "strided_reversed"]


def _sort_keys(ctx, keys, bits=32):
    ob, d, p = ctx
    buf = ob.Buffer(d, keys.size, np.uint32)
    buf.write(keys)
    p.radixSort(d, buf, keys.size, bits)
    d.waitForCompletion()
    out = buf.read()
    buf.release()
    return out


def _sort_pairs(ctx, kv, bits=32):
    ob, d, p = ctx
    buf = ob.Buffer(d, kv.shape[0], ob.PAIR_DTYPE)
    buf.write(kv)
    p.radixSort(d, buf, kv.shape[0], bits)
    d.waitForCompletion()
    out = buf.read()
    buf.release()
    return out


@pytest.mark.parametrize("n", SIZES)
def test_sort_keys_sizes_uniform(ctx, n):
    k = _keys("uniform", n)
    assert np.array_equal(_sort_keys(ctx, k), po.sort_u32(k))


@pytest.mark.parametrize("kind", KINDS)
def test_sort_keys_distributions(ctx, kind):
    for n in (8192 * 3 + 5, 1 << 18, (1 << 19) + 8960 * 3 + 1):
        k = _keys(kind, n)
        assert np.array_equal(_sort_keys(ctx, k), po.sort_u32(k)), (kind, n)


@pytest.mark.parametrize("bits", [0, 1, 4, 8, 12, 16, 20, 24, 28, 31, 32])
def test_sort_keys_partial_bits_stable_on_ignored_bits(ctx, bits):
    for n in (300007, 700001):  # 2048-element tiles / full-size tiles
        k = _keys("uniform", n)
        assert np.array_equal(_sort_keys(ctx, k, bits), po.sort_u32(k, bits)), (n, bits)


@pytest.mark.parametrize("n", [2, 33, 1000, 4097, 8191, 8192, 8193])
def test_small_inputs_single_cta_path(ctx, n):
    """n <= 8192 takes the one-launch, one-CTA sort (small_sort_kernel); 8193 is the first size of the multi-kernel chain.
    Partial sortBits (stability on the ignored bits), low-entropy keys (stability of pairs) and every tail shape."""
    ob = ctx[0]
    for bits in (32, 20, 8, 3):
        k = _keys("uniform", n, seed=bits)
        assert np.array_equal(_sort_keys(ctx, k, bits), po.sort_u32(k, bits)), (n, bits)
    for kind in ("few", "lowbyte", "allequal", "strided", "and3"):
        kv = np.empty(n, dtype=ob.PAIR_DTYPE)
        kv["key"], kv["value"] = _keys(kind, n), np.arange(n, dtype=np.uint32)
        for bits in (32, 12):
            assert np.array_equal(_sort_pairs(ctx, kv, bits), po.sort_pairs(kv, bits)), (n, kind, bits)


@pytest.mark.parametrize("n", SIZES)
def test_sort_pairs_sizes_uniform(ctx, n):
    ob = ctx[0]
    kv = np.empty(n, dtype=ob.PAIR_DTYPE)
    kv["key"], kv["value"] = _keys("uniform", n), np.arange(n, dtype=np.uint32)
    assert np.array_equal(_sort_pairs(ctx, kv), po.sort_pairs(kv))


@pytest.mark.parametrize("kind", KINDS)
def test_sort_pairs_stability(ctx, kind):
    ob = ctx[0]
    for n in (4096 * 3 + 5, 1 << 18, (1 << 19) + 6400 * 3 + 1):
        kv = np.empty(n, dtype=ob.PAIR_DTYPE)
        kv["key"], kv["value"] = _keys(kind, n), np.arange(n, dtype=np.uint32)
        got = _sort_pairs(ctx, kv)
        assert np.array_equal(got, po.sort_pairs(kv)), (kind, n)


# The scatter pass groups tiles by 8 for its two-level look-back (csrc/b200rs_onesweep2.cuh, LB_GROUP).  Tiles are 2048
# elements up to 2^19 elements (mid-size path) and 8960 keys / 6400 pairs above.  Sizes around whole groups, one tile
# more, one element less, and several groups deep -- for both tile sizes.
GROUP_EDGE_TILES = [7, 8, 9, 16, 17, 41]


@pytest.mark.parametrize("tiles", GROUP_EDGE_TILES)
def test_sort_lookback_group_boundaries(ctx, tiles):
    ob = ctx[0]
    big = 56  # 56 tiles of either default size are more than 2^19 elements: the full-size tiles are in use from there on
    key_sizes = [tiles * 2048 + d for d in (-1, 0, 1)] + [(big + tiles) * 8960 + d for d in (-1, 0, 1)]
    pair_sizes = [tiles * 2048 + d for d in (-1, 0, 1)] + [(big + 32 + tiles) * 6400 + d for d in (-1, 0, 1)]
    for n in key_sizes:
        k = _keys("and3", n)  # low entropy: long equal-key runs cross tile and group edges
        assert np.array_equal(_sort_keys(ctx, k), po.sort_u32(k)), n
    for n in pair_sizes:
        kv = np.empty(n, dtype=ob.PAIR_DTYPE)
        kv["key"], kv["value"] = _keys("few", n), np.arange(n, dtype=np.uint32)
        assert np.array_equal(_sort_pairs(ctx, kv), po.sort_pairs(kv)), n


@pytest.mark.parametrize("mask", [0x000000FF, 0x0000FF00, 0x00FF0000, 0xFF000000, 0x00FFFF00, 0xFF0000FF, 0x00FF00FF, 0xFFFF0000, 0x0])
def test_sort_skipped_passes(ctx, mask):
    """A pass whose digit is the same for every element moves nothing: the device detects it (digit_start_kernel) and
    that pass's kernel only copies its tiles; 0, 1, 2, 3 or 4 real passes then run."""
    ob = ctx[0]
    for n in (150001, 650001):  # 2048-element tiles / full-size tiles (identity passes copy whole tiles of either size)
        k = (_keys("uniform", n) & np.uint32(mask)) | np.uint32(0x5A5A5A5A & ~mask)
        assert np.array_equal(_sort_keys(ctx, k), po.sort_u32(k)), (n, hex(mask))
        for bits in (16, 24):
            assert np.array_equal(_sort_keys(ctx, k, bits), po.sort_u32(k, bits)), (n, hex(mask), bits)
        kv = np.empty(n, dtype=ob.PAIR_DTYPE)
        kv["key"], kv["value"] = k, np.arange(n, dtype=np.uint32)[::-1]
        assert np.array_equal(_sort_pairs(ctx, kv), po.sort_pairs(kv)), (n, hex(mask))


@pytest.mark.parametrize("bits", [4, 12, 16, 24])
def test_sort_pairs_partial_bits(ctx, bits):
    ob = ctx[0]
    for n in (200003, 600011):  # 2048-element tiles / full-size tiles
        kv = np.empty(n, dtype=ob.PAIR_DTYPE)
        kv["key"], kv["value"] = _keys("uniform", n), np.arange(n, dtype=np.uint32)[::-1]
        assert np.array_equal(_sort_pairs(ctx, kv, bits), po.sort_pairs(kv, bits)), (n, bits)


def test_reference_unit_test_vectors(ctx, golden_dir):
    """The reference's own sweep (UnitTest/main.cpp:105-205) against hashes of the reference's output."""
    ob, d, p = ctx
    with open(os.path.join(golden_dir, "reference_hashes.json")) as f:
        g = json.load(f)
    for e in g["sort32"]:
        out = _sort_keys(ctx, po.gen_sort32(e["n"]))
        assert f"{po.fnv1a64(out):016x}" == e["out"], ("sort32", e["n"])
    for e in g["sortkeyvalue"]:
        out = _sort_pairs(ctx, po.gen_keyvalue(e["n"]))
        assert f"{po.fnv1a64(out):016x}" == e["out"], ("kv", e["n"])
    for e in g["scan"]:
        s = po.gen_scan(e["n"])
        src, dst = ob.Buffer(d, s.size, np.int32), ob.Buffer(d, s.size, np.int32)
        src.write(s)
        total = p.scan(d, dst, src, s.size, sumOut=True)
        out = dst.read()
        assert f"{po.fnv1a64(out):016x}" == e["out"], ("scan", e["n"])  # includes n = 1048576, which the reference cannot do
        assert total == e["total"]
        src.release(); dst.release()


def test_stability_fixtures_from_reference(ctx, golden_dir):
    ob = ctx[0]
    fx = np.load(os.path.join(golden_dir, "stability_fixtures.npz"))
    for name in sorted({k[:-3] for k in fx.files if k.endswith("_in")}):
        if name.startswith("kv_"):
            got = _sort_pairs(ctx, np.ascontiguousarray(fx[name + "_in"]).view(ob.PAIR_DTYPE).reshape(-1))
            assert np.array_equal(got.view(np.uint32).reshape(-1, 2), fx[name + "_out"]), name
        else:
            assert np.array_equal(_sort_keys(ctx, fx[name + "_in"]), fx[name + "_out"]), name


SCAN_SIZES = [1, 2, 31, 32, 33, 511, 512, 513, 4095, 4096, 4097, 100003, (1 << 20) - 1, 1 << 20, (1 << 20) + 1, (1 << 23) + 11]


@pytest.mark.parametrize("n", SCAN_SIZES)
def test_scan_sizes_with_wraparound(ctx, n):
    ob, d, p = ctx
    rng = np.random.default_rng(n)
    s = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)  # full range => wraps mod 2^32
    src, dst = ob.Buffer(d, n, np.uint32), ob.Buffer(d, n, np.uint32)
    src.write(s)
    total = p.scan(d, dst, src, n, sumOut=True)
    want, want_total = po.scan_u32(s)
    assert np.array_equal(dst.read(), want)
    assert total == want_total
    # in place (dst == src), no total requested
    assert p.scan(d, src, src, n) is None
    d.waitForCompletion()
    assert np.array_equal(src.read(), want)
    src.release(); dst.release()


def test_unaligned_device_pointers(ctx):
    """Pointers that are only element-aligned (sub-buffer at +1 element) take the non-vector paths."""
    ob, d, p = ctx
    n = 50001
    k = _keys("uniform", n)
    big = ob.Buffer(d, n + 4, np.uint32)
    view = ob.Buffer(d, n, np.uint32, ptr=big.m_ptr + 4)
    view.write(k)
    p.radixSort(d, view, n)
    d.waitForCompletion()
    assert np.array_equal(view.read(), po.sort_u32(k))
    view.write(k)
    dst = ob.Buffer(d, n + 4, np.uint32)
    dview = ob.Buffer(d, n, np.uint32, ptr=dst.m_ptr + 4)
    p.scan(d, dview, view, n)
    d.waitForCompletion()
    assert np.array_equal(dview.read(), po.scan_u32(k)[0])
    kvbig = ob.Buffer(d, n + 1, ob.PAIR_DTYPE)
    kview = ob.Buffer(d, n, ob.PAIR_DTYPE, ptr=kvbig.m_ptr + 8)
    kv = np.empty(n, dtype=ob.PAIR_DTYPE)
    kv["key"], kv["value"] = k & np.uint32(0xFFFF), np.arange(n, dtype=np.uint32)
    kview.write(kv)
    p.radixSort(d, kview, n)
    d.waitForCompletion()
    assert np.array_equal(kview.read(), po.sort_pairs(kv))
    big.release(); dst.release(); kvbig.release()


def test_host_buffer_entry_points(ctx):
    ob, d, p = ctx
    from oclradixsort_b200._lib import check, lib
    n = 123457
    k = _keys("uniform", n)
    got = k.copy()
    check(lib().b200rs_sort_keys_u32_host(d.handle, ctypes.c_void_p(got.ctypes.data), n, 32), "sort_keys_host")
    assert np.array_equal(got, po.sort_u32(k))
    kv = np.empty(n, dtype=ob.PAIR_DTYPE)
    kv["key"], kv["value"] = k & np.uint32(0xFFF), np.arange(n, dtype=np.uint32)
    got = kv.copy()
    check(lib().b200rs_sort_pairs_u32_host(d.handle, ctypes.c_void_p(got.ctypes.data), n, 32), "sort_pairs_host")
    assert np.array_equal(got, po.sort_pairs(kv))
    out = np.empty_like(k)
    total = ctypes.c_uint32(0)
    check(lib().b200rs_exclusive_scan_u32_host(d.handle, ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(k.ctypes.data), n,
                                               ctypes.byref(total)), "scan_host")
    want, want_total = po.scan_u32(k)
    assert np.array_equal(out, want) and total.value == want_total
    check(lib().b200rs_device_release_scratch(d.handle), "release_scratch")


def test_host_batch_entry_points(ctx):
    """b200rs_sort_*_host_batch: several host arrays sorted in place by one pipelined call (1, 2, 3 and 5 arrays: the
    double buffering wraps around), pageable and pinned memory, against the oracle."""
    import torch
    ob, d, p = ctx
    from oclradixsort_b200._lib import check, lib
    n = 70001
    for count in (1, 2, 3, 5):
        arrays = [_keys("and3" if i % 2 else "uniform", n, seed=10 + i) for i in range(count)]
        got = [a.copy() for a in arrays]
        ptrs = (ctypes.c_void_p * count)(*[g.ctypes.data for g in got])
        check(lib().b200rs_sort_keys_u32_host_batch(d.handle, ptrs, count, n, 32), "sort_keys_host_batch")
        for a, g in zip(arrays, got):
            assert np.array_equal(g, po.sort_u32(a)), count
    count = 4
    pinned = [torch.empty((n, 2), dtype=torch.int32).pin_memory() for _ in range(count)]
    want = []
    for i, t in enumerate(pinned):
        kv = np.empty(n, dtype=ob.PAIR_DTYPE)
        kv["key"], kv["value"] = _keys("few" if i % 2 else "uniform", n, seed=20 + i), np.arange(n, dtype=np.uint32)[::-1]
        t.numpy().view(np.uint32)[:] = kv.view(np.uint32).reshape(n, 2)
        want.append(po.sort_pairs(kv))
    ptrs = (ctypes.c_void_p * count)(*[t.data_ptr() for t in pinned])
    check(lib().b200rs_sort_pairs_u32_host_batch(d.handle, ptrs, count, n, 32), "sort_pairs_host_batch")
    for t, w in zip(pinned, want):
        assert np.array_equal(t.numpy().view(np.uint32).reshape(n, 2), w.view(np.uint32).reshape(n, 2))
    assert lib().b200rs_sort_pairs_u32_host_batch(d.handle, None, 2, n, 32) != 0  # null array list is an error
    check(lib().b200rs_device_release_scratch(d.handle), "release_scratch")


def test_empty_inputs_and_errors(ctx):
    ob, d, p = ctx
    from oclradixsort_b200._lib import lib
    b = ob.Buffer(d, 16, np.uint32)
    p.radixSort(d, b, 0)
    p.scan(d, b, b, 0)
    d.waitForCompletion()
    need = ctypes.c_size_t(0)
    assert lib().b200rs_sort_keys_u32(d.handle, None, 1000, 33, None, ctypes.byref(need)) == -1  # bad sort_bits
    assert lib().b200rs_sort_keys_u32(d.handle, None, 1000, 32, None, ctypes.byref(need)) == 0 and need.value > 4000
    small = ctypes.c_size_t(16)
    assert lib().b200rs_sort_keys_u32(d.handle, ctypes.c_void_p(b.m_ptr), 1000, 32, ctypes.c_void_p(b.m_ptr), ctypes.byref(small)) == -2
    with pytest.raises(NotImplementedError):
        ob.DeviceUtils.allocate(ob.TYPE_HOST)  # no CPU backend
    b.release()


def test_large_sizes_by_properties(ctx):
    """BASELINE sizes the serial oracle is too slow for in a test: sortedness + multiset checksums +
    stability (values increasing inside equal-key runs), checked on the device with torch."""
    import torch
    ob, d, p = ctx
    n = 1 << 27
    g = torch.Generator(device="cuda").manual_seed(5)
    keys = torch.randint(0, 2**31, (n,), device="cuda", dtype=torch.int64, generator=g).to(torch.int32)
    keys = keys * 2 + torch.randint(0, 2, (n,), device="cuda", dtype=torch.int32, generator=g)  # all 32 bits random
    def multiset_hash(t):
        # order-independent: sums (mod 2^64) of two different non-linear mixes of every element
        x = t.to(torch.int64) & 0xFFFFFFFF
        a = x * -7046029254386353131  # 0x9E3779B97F4A7C15 as int64; products wrap
        a = a ^ (a >> 29)
        b = (x + 0x632BE59B) * -4417276706812531889  # 0xC2B2AE3D27D4EB4F
        b = b ^ (b >> 31)
        return int(a.sum().item()), int((b * b).sum().item())

    hash_in = multiset_hash(keys)
    torch.cuda.synchronize()
    buf = ob.Buffer(d, n, np.uint32, ptr=keys.data_ptr())
    p.radixSort(d, buf, n)
    d.waitForCompletion()
    u = keys.to(torch.int64) & 0xFFFFFFFF
    assert bool((u[1:] >= u[:-1]).all())
    assert multiset_hash(keys) == hash_in  # sorted + same multiset <=> the oracle's output (key-only)
    del u
    # pairs: low-entropy keys, value = index  => stable result is unique
    m = 1 << 26
    kv = torch.empty((m, 2), device="cuda", dtype=torch.int32)
    kv[:, 0] = torch.randint(0, 1 << 12, (m,), device="cuda", dtype=torch.int32, generator=g) * 1048583
    kv[:, 1] = torch.arange(m, device="cuda", dtype=torch.int32)
    torch.cuda.synchronize()
    pb = ob.Buffer(d, m, ob.PAIR_DTYPE, ptr=kv.data_ptr())
    p.radixSort(d, pb, m)
    d.waitForCompletion()
    k64 = kv[:, 0].to(torch.int64) & 0xFFFFFFFF
    v64 = kv[:, 1].to(torch.int64)
    assert bool((k64[1:] >= k64[:-1]).all())
    same = k64[1:] == k64[:-1]
    assert bool((v64[1:][same] > v64[:-1][same]).all())
    assert int(v64.sum().item()) == m * (m - 1) // 2
    # scan at 2^28 against torch.cumsum (mod 2^32)
    s = torch.randint(0, 2**31, (1 << 28,), device="cuda", dtype=torch.int64, generator=g).to(torch.int32)
    want = (torch.cumsum(s.to(torch.int64) & 0xFFFFFFFF, 0) - (s.to(torch.int64) & 0xFFFFFFFF)) & 0xFFFFFFFF
    torch.cuda.synchronize()
    sb = ob.Buffer(d, s.numel(), np.uint32, ptr=s.data_ptr())
    p.scan(d, sb, sb, s.numel())
    d.waitForCompletion()
    assert bool(((s.to(torch.int64) & 0xFFFFFFFF) == want).all())
