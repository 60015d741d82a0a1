"""Host-side mirror of Tahoe::Pprims (Tahoe/ParallelPrimitives/Pprims.h:11-48) over the C ABI."""
from __future__ import annotations

import ctypes

import numpy as np

from ._lib import check, lib
from .adl import PAIR_DTYPE, Buffer, Device


class Pprims:
    """Same calls as the reference: radixSort(device, buffer, n, sortBits=32), scan(device, dst, src, n).

    Owns grow-only device scratch (the reference's m_u32WorkBuffer / m_u2WorkBuffer uArrays,
    Pprims.h:44-45); release it (or drop the object) before DeviceUtils.deallocate.
    """

    # tuning enum of Pprims.h:22-33, values as realised by the sm_100a kernels
    SCAN_BLOCK_SIZE = 256
    RSORT_BITS_PER_PASS = 8
    RSORT_NUM_TABLES = 256
    R32SORT_DATA_ALIGNMENT = 1  # any n is accepted
    R32SORT_BITS_PER_PASS = 8

    def __init__(self):
        self._temp: Buffer | None = None
        self._word: Buffer | None = None
        self.m_cacheKernel = True

    def cacheKernel(self, cache: bool) -> None:  # Pprims.h:20 -- kernels are compiled ahead of time
        self.m_cacheKernel = cache

    def release(self) -> None:
        for b in (self._temp, self._word):
            if b is not None:
                b.release()
        self._temp = self._word = None

    def _scratch(self, device: Device, nbytes: int) -> Buffer:
        if self._temp is None or self._temp.m_device is not device:
            if self._temp is not None:
                self._temp.release()
            self._temp = Buffer(device, 0, np.uint8)
        self._temp.setSize(max(nbytes, 256))
        return self._temp

    def radixSort(self, device: Device, inout: Buffer, n: int, sortBits: int = 32) -> None:
        """Pprims::radixSort, Pprims.cpp:200-302 (pairs) / :304-406 (keys).  Asynchronous."""
        if device is None:
            raise ValueError("device == 0: there is no Host fallback (reference: Pprims.cpp:202-212,306-316)")
        assert n <= inout.getSize()
        if inout.dtype == PAIR_DTYPE:
            fn, name = lib().b200rs_sort_pairs_u32, "b200rs_sort_pairs_u32"
        elif inout.dtype.itemsize == 4:
            fn, name = lib().b200rs_sort_keys_u32, "b200rs_sort_keys_u32"
        else:
            raise TypeError(f"unsupported buffer dtype {inout.dtype}")
        need = ctypes.c_size_t(0)
        check(fn(device.handle, None, n, sortBits, None, ctypes.byref(need)), name + " (size query)")
        temp = self._scratch(device, need.value)
        have = ctypes.c_size_t(temp.getSize())
        check(fn(device.handle, ctypes.c_void_p(inout.m_ptr), n, sortBits, ctypes.c_void_p(temp.m_ptr), ctypes.byref(have)), name)
        inout.markDeviceWritten()

    def scan(self, device: Device, dst: Buffer, src: Buffer, n: int, sumOut: bool = False):
        """Pprims::scan, Pprims.cpp:122-179.  With sumOut=True returns the total of all n inputs
        (the reference writes it through a u32* after a non-blocking read; here the call syncs)."""
        if device is None:
            raise ValueError("device == 0 (reference asserts: Pprims.cpp:124-127)")
        assert n <= dst.getSize() and n <= src.getSize()
        need = ctypes.c_size_t(0)
        fn = lib().b200rs_exclusive_scan_u32
        check(fn(device.handle, None, None, n, None, None, ctypes.byref(need)), "b200rs_exclusive_scan_u32 (size query)")
        temp = self._scratch(device, need.value)
        have = ctypes.c_size_t(temp.getSize())
        total_ptr = None
        if sumOut:
            if self._word is None or self._word.m_device is not device:
                self._word = Buffer(device, 64, np.uint32)
            total_ptr = ctypes.c_void_p(self._word.m_ptr)
        check(fn(device.handle, ctypes.c_void_p(dst.m_ptr), ctypes.c_void_p(src.m_ptr), n, total_ptr, ctypes.c_void_p(temp.m_ptr),
                 ctypes.byref(have)), "b200rs_exclusive_scan_u32")
        dst.markDeviceWritten()
        if sumOut:
            return int(self._word.read(1)[0])
        return None

    def copy(self, device: Device, dst: Buffer, src: Buffer, n: int) -> None:
        """Pprims::copy, Pprims.cpp:32-65 (CopyIntKernel / CopyF4Kernel, PprimsKernels.cl:9-24): dst[i] = src[i], i < n.
        4-byte elements (int / u32) or 16-byte elements (float4).  Asynchronous."""
        if device is None:
            raise ValueError("device == 0: there is no Host path (reference: the CPU loop of Pprims.cpp:34-38)")
        assert n <= dst.getSize() and n <= src.getSize() and dst.dtype.itemsize == src.dtype.itemsize
        dst.markDeviceWritten()
        if dst.dtype.itemsize == 4:
            check(lib().b200rs_copy_u32(device.handle, ctypes.c_void_p(dst.m_ptr), ctypes.c_void_p(src.m_ptr), n), "b200rs_copy_u32")
        elif dst.dtype.itemsize == 16:
            check(lib().b200rs_copy_u128(device.handle, ctypes.c_void_p(dst.m_ptr), ctypes.c_void_p(src.m_ptr), n), "b200rs_copy_u128")
        else:
            raise TypeError(f"unsupported element size {dst.dtype.itemsize}")

    def fill(self, device: Device, dst: Buffer, value, n: int) -> None:
        """Pprims::fill, Pprims.cpp:67-120 (FillIntKernel / FillU32Kernel / FillF4Kernel, PprimsKernels.cl:28-48):
        dst[i] = value, i < n.  `value`: an int for 4-byte elements, 4 floats (or a 16-byte numpy scalar) for float4."""
        if device is None:
            raise ValueError("device == 0: there is no Host path (reference: the CPU loop of Pprims.cpp:69-73)")
        assert n <= dst.getSize()
        dst.markDeviceWritten()
        if dst.dtype.itemsize == 4:
            word = int(np.asarray(value, dtype=dst.dtype).view(np.uint32)) if dst.dtype.kind != "u" else int(value) & 0xFFFFFFFF
            check(lib().b200rs_fill_u32(device.handle, ctypes.c_void_p(dst.m_ptr), word, n), "b200rs_fill_u32")
        elif dst.dtype.itemsize == 16:
            raw = np.asarray(value, dtype=np.float32).reshape(4) if not isinstance(value, np.void) else np.frombuffer(value.tobytes(), np.float32)
            words = (ctypes.c_uint32 * 4)(*[int(w) for w in raw.view(np.uint32)])
            check(lib().b200rs_fill_u128(device.handle, ctypes.c_void_p(dst.m_ptr), words, n), "b200rs_fill_u128")
        else:
            raise TypeError(f"unsupported element size {dst.dtype.itemsize}")
