"""oracle/pyoracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes/numpy front-end to oracle/liboracle.so (the C restatement of the reference's Host path,
oracle/radixsort_oracle.c) and, when it was built in the container, to
oracle/_ref/libref_oclradixsort.so (the UNMODIFIED reference compiled from /root/reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing under oclradixsort_b200/ does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libref_oclradixsort.so")

PAIR_DTYPE = np.dtype([("key", "<u4"), ("value", "<u4")])  # RadixSort.h:10-21 layout

_lib = None
_ref = None


def build(with_ref: bool | None = None) -> None:
    """Compile liboracle.so (always) and oracle/_ref (only where /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    if with_ref is None:
        with_ref = os.path.isdir("/root/reference")
    if with_ref:
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_ORACLE_SO):
            build(with_ref=False)
        L = ctypes.CDLL(_ORACLE_SO)
        vp, sz, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        L.oracle_sort_u32.argtypes = [vp, sz, ci]
        L.oracle_sort_u32.restype = ci
        L.oracle_sort_pairs.argtypes = [vp, sz, ci]
        L.oracle_sort_pairs.restype = ci
        L.oracle_scan_u32.argtypes = [vp, vp, sz, vp]
        L.oracle_scan_u32.restype = None
        for name in ("oracle_gen_sort32", "oracle_gen_keyvalue", "oracle_gen_scan"):
            getattr(L, name).argtypes = [vp, sz, ctypes.c_uint]
            getattr(L, name).restype = None
        L.oracle_fnv1a64.argtypes = [vp, sz]
        L.oracle_fnv1a64.restype = ctypes.c_uint64
        _lib = L
    return _lib


def have_ref() -> bool:
    return os.path.exists(_REF_SO)


def ref() -> ctypes.CDLL:
    """The unmodified reference (oracle/_ref); raises if it was not built."""
    global _ref
    if _ref is None:
        L = ctypes.CDLL(_REF_SO)
        for name in ("ref_radixsort_u32", "ref_radixsort_pairs", "ref_hostbackend_sort_u32",
                     "ref_hostbackend_sort_pairs"):
            getattr(L, name).argtypes = [ctypes.c_void_p, ctypes.c_int]
            getattr(L, name).restype = None
        for name in ("ref_hostbackend_sort_u32_timed", "ref_hostbackend_sort_pairs_timed"):
            getattr(L, name).argtypes = [ctypes.c_void_p, ctypes.c_int]
            getattr(L, name).restype = ctypes.c_double
        _ref = L
    return _ref


def _ptr(a: np.ndarray) -> ctypes.c_void_p:
    assert a.flags["C_CONTIGUOUS"]
    return ctypes.c_void_p(a.ctypes.data)


# ---- the restated oracle -------------------------------------------------------------------

def sort_u32(keys: np.ndarray, sort_bits: int = 32) -> np.ndarray:
    out = np.ascontiguousarray(keys, dtype=np.uint32).copy()
    assert lib().oracle_sort_u32(_ptr(out), out.size, sort_bits) == 0
    return out


def sort_pairs(pairs: np.ndarray, sort_bits: int = 32) -> np.ndarray:
    """pairs: structured PAIR_DTYPE array, or uint32 array of shape (n, 2) = (key, value)."""
    out = np.ascontiguousarray(pairs).copy()
    assert out.dtype == PAIR_DTYPE or (out.dtype == np.uint32 and out.ndim == 2 and out.shape[1] == 2)
    n = out.shape[0]
    assert lib().oracle_sort_pairs(_ptr(out), n, sort_bits) == 0
    return out


def scan_u32(src: np.ndarray) -> tuple[np.ndarray, int]:
    s = np.ascontiguousarray(src).view(np.uint32)
    dst = np.empty_like(s)
    total = ctypes.c_uint32(0)
    lib().oracle_scan_u32(_ptr(dst), _ptr(s), s.size, ctypes.byref(total))
    return dst, int(total.value)


def gen_sort32(n: int, seed: int = 123) -> np.ndarray:
    out = np.empty(n, dtype=np.uint32)
    lib().oracle_gen_sort32(_ptr(out), n, seed)
    return out


def gen_keyvalue(n: int, seed: int = 123) -> np.ndarray:
    out = np.empty(n, dtype=PAIR_DTYPE)
    lib().oracle_gen_keyvalue(_ptr(out), n, seed)
    return out


def gen_scan(n: int, seed: int = 123) -> np.ndarray:
    out = np.empty(n, dtype=np.int32)
    lib().oracle_gen_scan(_ptr(out), n, seed)
    return out


def fnv1a64(a: np.ndarray) -> int:
    a = np.ascontiguousarray(a)
    return int(lib().oracle_fnv1a64(_ptr(a), a.nbytes))


# ---- the unmodified reference (only where oracle/_ref was built) -----------------------------

def ref_sort_u32(keys: np.ndarray, host_backend: bool = False) -> np.ndarray:
    out = np.ascontiguousarray(keys, dtype=np.uint32).copy()
    fn = ref().ref_hostbackend_sort_u32 if host_backend else ref().ref_radixsort_u32
    fn(_ptr(out), out.size)
    return out


def ref_sort_pairs(pairs: np.ndarray, host_backend: bool = False) -> np.ndarray:
    out = np.ascontiguousarray(pairs).copy()
    fn = ref().ref_hostbackend_sort_pairs if host_backend else ref().ref_radixsort_pairs
    fn(_ptr(out), out.shape[0])
    return out


def ref_time_hostbackend(data: np.ndarray, pairs: bool) -> float:
    """Seconds the UNMODIFIED reference's Host-backend Pprims::radixSort takes on `data` (sorted in place)."""
    assert data.flags["C_CONTIGUOUS"]
    fn = ref().ref_hostbackend_sort_pairs_timed if pairs else ref().ref_hostbackend_sort_u32_timed
    return float(fn(_ptr(data), data.shape[0]))


def copy_elems(dst: np.ndarray, src: np.ndarray, n: int) -> np.ndarray:
    """Pprims::copy's CPU loop `for i<n: dst[i] = src[i]` (Pprims.cpp:34-38, :51-55); returns the new dst."""
    out = dst.copy()
    out[:n] = src[:n]
    return out


def fill_elems(dst: np.ndarray, value, n: int) -> np.ndarray:
    """Pprims::fill's CPU loop `for i<n: dst[i] = src` (Pprims.cpp:69-73, :86-90, :103-107); returns the new dst."""
    out = dst.copy()
    out[:n] = value
    return out
