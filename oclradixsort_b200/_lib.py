"""ctypes binding of libb200rs.so (C ABI declared in include/b200rs.h).

The library is the product: if it is missing or fails to load this module raises -- there is no
Python/CPU fallback for any entry point.
"""
from __future__ import annotations

import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200RS_LIB: developer tools only (tools/sweep.py, tools/msd_probe.py) -- points the binding at the experiments build
# (make experiments -> tools/_build/libb200rs_exp.so).  The library itself reads no environment variable.
LIB_PATH = os.environ.get("B200RS_LIB") or os.path.join(_HERE, "libb200rs.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "b200rs.h")

c_dev = ctypes.c_void_p
_vp, _sz, _u64, _int = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_int
_P = ctypes.POINTER


class ProfileEntry(ctypes.Structure):
    _fields_ = [("kernel", ctypes.c_char * 48), ("ms", ctypes.c_float), ("elements", ctypes.c_uint64), ("bytes", ctypes.c_uint64)]


# name -> (restype, argtypes); must list every symbol include/b200rs.h declares (tests check that)
SIGNATURES = {
    "b200rs_version": (_int, []),
    "b200rs_error_string": (ctypes.c_char_p, [_int]),
    "b200rs_device_count": (_int, [_P(_int)]),
    "b200rs_device_create": (_int, [_int, _P(c_dev)]),
    "b200rs_device_create_on_stream": (_int, [_int, _vp, _P(c_dev)]),
    "b200rs_device_destroy": (_int, [c_dev]),
    "b200rs_device_sync": (_int, [c_dev]),
    "b200rs_device_num_sms": (_int, [c_dev, _P(_int)]),
    "b200rs_device_name": (_int, [c_dev, ctypes.c_char_p]),
    "b200rs_device_index": (_int, [c_dev, _P(_int)]),
    "b200rs_device_mem_info": (_int, [c_dev, _P(_sz), _P(_sz)]),
    "b200rs_device_stream": (_vp, [c_dev]),
    "b200rs_malloc": (_int, [c_dev, _sz, _P(_vp)]),
    "b200rs_free": (_int, [c_dev, _vp]),
    "b200rs_host_alloc": (_int, [c_dev, _sz, _P(_vp)]),
    "b200rs_host_free": (_int, [c_dev, _vp]),
    "b200rs_memcpy_h2d": (_int, [c_dev, _vp, _vp, _sz]),
    "b200rs_memcpy_d2h": (_int, [c_dev, _vp, _vp, _sz]),
    "b200rs_memcpy_d2d": (_int, [c_dev, _vp, _vp, _sz]),
    "b200rs_memset": (_int, [c_dev, _vp, _int, _sz]),
    "b200rs_sort_keys_u32": (_int, [c_dev, _vp, _u64, _int, _vp, _P(_sz)]),
    "b200rs_sort_keys_u32_msd": (_int, [c_dev, _vp, _u64, _vp, _P(_sz), _P(_int)]),
    "b200rs_sort_pairs_u32": (_int, [c_dev, _vp, _u64, _int, _vp, _P(_sz)]),
    "b200rs_exclusive_scan_u32": (_int, [c_dev, _vp, _vp, _u64, _vp, _vp, _P(_sz)]),
    "b200rs_copy_u32": (_int, [c_dev, _vp, _vp, _u64]),
    "b200rs_copy_u128": (_int, [c_dev, _vp, _vp, _u64]),
    "b200rs_fill_u32": (_int, [c_dev, _vp, ctypes.c_uint32, _u64]),
    "b200rs_fill_u128": (_int, [c_dev, _vp, _P(ctypes.c_uint32), _u64]),
    "b200rs_digit_histogram_pairs": (_int, [c_dev, _vp, _u64, _int, _int, _vp]),
    "b200rs_partition_pairs": (_int, [c_dev, _vp, _vp, _u64, _int, _int, _vp, _vp, _vp, _P(_sz)]),
    "b200rs_scatter_pairs_to_parts": (_int, [c_dev, _vp, _u64, _int, _int, _vp, _vp, _vp, _vp, _P(_sz)]),
    "b200rs_exchange_pairs": (_int, [c_dev, _vp, _u64, _int, _int, _vp, _vp, _int, _vp, _vp, _P(_sz)]),
    "b200rs_dist_sort_pairs_u32": (_int, [c_dev, _vp, _vp, _u64, _vp, _u64, _vp, _vp, _vp, _P(_sz)]),
    "b200rs_exchange_pairs_by_splitters": (_int, [c_dev, _vp, _u64, _vp, _vp, _int, _vp, _P(_sz)]),
    "b200rs_filtered_histograms_pairs": (_int, [c_dev, _vp, _u64, _int, _vp, _int, _vp]),
    "b200rs_dist_plan": (_int, [c_dev, _vp, _int, _int, _vp, _u64, _u64, _vp, _vp, _vp, _vp]),
    "b200rs_dist_plan_halves": (_int, [c_dev, _vp, _int, _int, _vp, _u64, _u64, _u64, _int, _vp, _vp, _vp, _vp, _vp]),
    "b200rs_sort_pairs_u32_devn": (_int, [c_dev, _vp, _u64, _vp, _int, _vp, _P(_sz)]),
    "b200rs_enable_peer_access": (_int, [c_dev, _int]),
    "b200rs_ipc_export": (_int, [c_dev, _vp, ctypes.c_char_p]),
    "b200rs_ipc_import": (_int, [c_dev, ctypes.c_char_p, _P(_vp)]),
    "b200rs_ipc_release": (_int, [c_dev, _vp]),
    "b200rs_sort_keys_u32_host": (_int, [c_dev, _vp, _u64, _int]),
    "b200rs_sort_pairs_u32_host": (_int, [c_dev, _vp, _u64, _int]),
    "b200rs_sort_keys_u32_host_batch": (_int, [c_dev, _vp, _int, _u64, _int]),
    "b200rs_sort_pairs_u32_host_batch": (_int, [c_dev, _vp, _int, _u64, _int]),
    "b200rs_exclusive_scan_u32_host": (_int, [c_dev, _vp, _vp, _u64, _P(ctypes.c_uint32)]),
    "b200rs_device_release_scratch": (_int, [c_dev]),
    "b200rs_profile_enable": (_int, [c_dev, _int]),
    "b200rs_profile_read": (_int, [c_dev, _P(ProfileEntry), _int, _P(_int)]),
    "b200rs_device_launch_count": (_int, [c_dev, _P(_u64)]),
    "b200rs_event_create": (_int, [c_dev, _P(_vp)]),
    "b200rs_event_record": (_int, [c_dev, _vp]),
    "b200rs_event_elapsed_ms": (_int, [c_dev, _vp, _vp, _P(ctypes.c_float)]),
    "b200rs_event_destroy": (_int, [c_dev, _vp]),
    "b200rs_event_query": (_int, [c_dev, _vp, _P(_int)]),
    "b200rs_event_synchronize": (_int, [c_dev, _vp]),
}

_lib = None


class B200RSError(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        msg = lib().b200rs_error_string(code)
        super().__init__(f"{where} failed: {code} ({msg.decode() if msg else '?'})")


def lib() -> ctypes.CDLL:
    """Load libb200rs.so; raise if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: build it with `make` or `python -c 'import __graft_entry__ as g; g.build()'`")
        L = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = L
    return _lib


def check(code: int, where: str) -> None:
    if code != 0:
        raise B200RSError(code, where)


def declared_symbols() -> list[str]:
    """Function names declared in include/b200rs.h (used by the CPU-side ABI test)."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200rs_[a-z0-9_]+)\s*\(", text)))
