"""GPU parity tests of the key-only MSD path (csrc/b200rs_msd.cuh) and bit-exact checks against the oracle at the
BASELINE.json sizes (configs 2, 3, 4): everything through the C ABI, oracle = oracle/pyoracle (restatement of
RadixSort::sort, Tahoe/Algorithm/Sort/RadixSort.cpp:10-104).

b200rs_sort_keys_u32_msd tries the MSD pipeline at any n, so its edge cases (partial tiles, one-digit tiles, bins that
overflow into the dense route, buckets that overflow the 4-bit counters into the robust route, inputs that are not
eligible) are covered at sizes the oracle sorts in milliseconds; b200rs_sort_keys_u32 picks the path by itself."""
import ctypes

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


def _keys(kind: str, n: int, seed: int = 3) -> np.ndarray:
    rng = np.random.default_rng(seed + n)
    u = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    i = np.arange(n, dtype=np.uint64)
    step = np.uint64(max(1, 2**32 // max(n, 1)))
    if kind == "uniform":
        return u
    if kind == "strided":            # presorted, spans the whole key range (BASELINE config 4 "presorted")
        return (i * step).astype(np.uint32)
    if kind == "strided_reversed":
        return (i * step).astype(np.uint32)[::-1].copy()
    if kind == "and3":
        return u & rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32) & rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    if kind == "allequal":
        return np.full(n, 0xDEADBEEF, dtype=np.uint32)
    if kind == "few":
        return (u & np.uint32(0x3)) * np.uint32(0x01010101)
    if kind == "dups_low":           # every top-16 bucket holds copies of ONE value: the 4-bit counters overflow -> robust route
        return (u & np.uint32(0xFFFF0000)) | np.uint32(0x1234)
    if kind == "dups_some":          # 8 neighbouring values folded into one: several equal keys per bucket, nibble-sum limits
        return u & np.uint32(0xFFFFFFF8)
    if kind == "staircase":          # presorted runs of 37 equal keys
        return ((i // np.uint64(37)) * np.uint64(37) * step).astype(np.uint32)
    if kind == "low16":              # one bucket holds everything: never eligible
        return u & np.uint32(0xFFFF)
    if kind == "two_digit_tiles":    # presorted, two first-pass digits per tile: the dense route of the partition kernel
        return np.sort(u & np.uint32(0x01FFFFFF) | ((u >> np.uint32(31)) << np.uint32(31)))
    raise ValueError(kind)


def _sort_msd(ctx, keys, offset_elems=0):
    """b200rs_sort_keys_u32_msd on a device copy of `keys` (optionally at an element offset inside its allocation)."""
    from oclradixsort_b200._lib import check, lib
    ob, d, p = ctx
    n = keys.size
    big = ob.Buffer(d, n + offset_elems + 4, np.uint32)
    view = ob.Buffer(d, n, np.uint32, ptr=big.m_ptr + 4 * offset_elems)
    view.write(keys)
    need, used = ctypes.c_size_t(0), ctypes.c_int(-1)
    fn = lib().b200rs_sort_keys_u32_msd
    check(fn(d.handle, None, n, None, ctypes.byref(need), ctypes.byref(used)), "size query")
    temp = ob.Buffer(d, need.value, np.uint8)
    have = ctypes.c_size_t(need.value)
    check(fn(d.handle, ctypes.c_void_p(view.m_ptr), n, ctypes.c_void_p(temp.m_ptr), ctypes.byref(have), ctypes.byref(used)), "b200rs_sort_keys_u32_msd")
    d.waitForCompletion()
    out = view.read()
    temp.release(); big.release()
    return out, used.value


MSD_SIZES = [2, 3, 31, 257, 4097, 12287, 12288, 12289, 24577, 65537, 300007, (1 << 20) + 5, (1 << 22) + 3]
MSD_KINDS = ["uniform", "strided", "strided_reversed", "and3", "allequal", "few", "dups_low", "dups_some", "staircase", "low16", "two_digit_tiles"]


@pytest.mark.parametrize("kind", MSD_KINDS)
def test_msd_sizes_and_distributions(ctx, kind):
    for n in MSD_SIZES:
        k = _keys(kind, n)
        got, used = _sort_msd(ctx, k)
        assert np.array_equal(got, po.sort_u32(k)), (kind, n, used)
        if kind in ("uniform", "strided", "strided_reversed", "dups_low", "dups_some", "staircase"):
            assert used == 1, (kind, n)       # every top-16 bucket is small: the MSD pipeline must have run
        if kind in ("allequal", "low16") and n > 12288:
            assert used == 0, (kind, n)       # one bucket holds everything: the LSD path must have run


def test_msd_unaligned_input_takes_the_lsd_path(ctx):
    n = 70001
    k = _keys("uniform", n)
    got, used = _sort_msd(ctx, k, offset_elems=1)  # 4-byte aligned only
    assert used == 0 and np.array_equal(got, po.sort_u32(k))
    got, used = _sort_msd(ctx, k, offset_elems=4)  # 16-byte aligned, not 128
    assert used == 1 and np.array_equal(got, po.sort_u32(k))


def test_msd_bucket_sizes_at_the_shape_limits(ctx):
    """Top-16 buckets just below / above the capacities of the counting kernel's shapes (4608, 6144, 9216, 12288 keys minus
    the 31-key alignment allowance) and above the largest (not eligible)."""
    rng = np.random.default_rng(11)
    for bucket in (4577, 4578, 6113, 6114, 9185, 9186, 12257, 12258, 20000):
        low = rng.integers(0, 2**16, size=bucket, dtype=np.uint64).astype(np.uint32)
        filler = rng.integers(0, 2**32, size=100000, dtype=np.uint64).astype(np.uint32) & np.uint32(0x7FFFFFFF)
        k = np.concatenate([np.uint32(0xABCD0000) | low, filler])
        rng.shuffle(k)
        got, used = _sort_msd(ctx, k)
        assert np.array_equal(got, po.sort_u32(k)), bucket
        assert used == (1 if bucket <= 12257 else 0), bucket


def test_msd_equal_keys_inside_a_bucket(ctx):
    """15 equal keys fit a 4-bit counter; 16 and more send the bucket to the robust route (stable LSD sort in shared memory)."""
    rng = np.random.default_rng(12)
    for copies in (15, 16, 17, 255, 3000, 9000):
        hot = np.full(copies, 0x77771234, dtype=np.uint32)
        other = rng.integers(0, 2**32, size=200000, dtype=np.uint64).astype(np.uint32)
        same_bucket = np.uint32(0x77770000) | rng.integers(0, 2**16, size=500, dtype=np.uint64).astype(np.uint32)
        k = np.concatenate([hot, other, same_bucket])
        rng.shuffle(k)
        got, used = _sort_msd(ctx, k)
        assert used == 1 and np.array_equal(got, po.sort_u32(k)), copies


# ---- bit-exact against the oracle at the BASELINE.json sizes ("next round" 3b) -------------------------

def _device_sort(ctx, arr, dtype, bits=32):
    ob, d, p = ctx
    buf = ob.Buffer(d, arr.shape[0], dtype)
    buf.write(arr)
    p.radixSort(d, buf, arr.shape[0], bits)
    d.waitForCompletion()
    out = buf.read()
    buf.release()
    return out


@pytest.mark.parametrize("log2n", [24, 26])
def test_pairs_bit_exact_at_config2_sizes(ctx, log2n):
    ob = ctx[0]
    n = 1 << log2n
    kv = np.empty(n, dtype=ob.PAIR_DTYPE)
    kv["key"], kv["value"] = _keys("uniform", n), np.arange(n, dtype=np.uint32)
    assert np.array_equal(_device_sort(ctx, kv, ob.PAIR_DTYPE), po.sort_pairs(kv))


@pytest.mark.parametrize("log2n", [26, 28])
def test_keys_bit_exact_at_config4_sizes(ctx, log2n):
    n = 1 << log2n
    k = _keys("uniform", n)
    assert np.array_equal(_device_sort(ctx, k, np.uint32), po.sort_u32(k))  # 2^28: the MSD path (picked automatically)


@pytest.mark.parametrize("kind", ["and3", "few", "strided", "strided_reversed", "allequal"])
def test_keys_distributions_bit_exact_at_2p26(ctx, kind):
    n = 1 << 26
    k = _keys(kind, n)
    assert np.array_equal(_device_sort(ctx, k, np.uint32), po.sort_u32(k)), kind
    got, used = _sort_msd(ctx, k)
    assert np.array_equal(got, po.sort_u32(k)), (kind, used)


def test_keys_sortbits16_stability_at_2p26(ctx):
    """Partial sortBits: only the low 16 bits are ordered, the high 16 ride along like a payload and must keep input order."""
    n = 1 << 26
    k = _keys("uniform", n)
    assert np.array_equal(_device_sort(ctx, k, np.uint32, 16), po.sort_u32(k, 16))


def test_scan_bit_exact_at_2p28(ctx):
    ob, d, p = ctx
    n = 1 << 28
    s = _keys("uniform", n)  # full-range inputs: the sums wrap mod 2^32
    src = ob.Buffer(d, n, np.uint32)
    src.write(s)
    total = p.scan(d, src, src, n, sumOut=True)
    want, want_total = po.scan_u32(s)
    assert total == want_total and np.array_equal(src.read(), want)
    src.release()
