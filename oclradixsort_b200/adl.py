"""Host-side mirror of the reference's `adl` device/buffer wrapper (Adl/Adl.h:39-222) over the C ABI.

One device type: a CUDA device (the reference's TYPE_CL enumerator is kept as its name so caller
code reads the same).  TYPE_HOST exists as an enumerator only: asking for it raises, there is no
CPU backend here.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import check, lib

TYPE_CL = 0    # Adl.h:41 -- routed to the CUDA device
TYPE_DX11 = 1  # Adl.h:42 -- not available
TYPE_HOST = 2  # Adl.h:43 -- not available (no CPU fallback)

PAIR_DTYPE = np.dtype([("key", "<u4"), ("value", "<u4")])  # SortData / uint2, RadixSort.h:10-21


class Device:
    """adl::Device for one CUDA GPU (Adl.h:123-155) -- owns the C-ABI handle and its stream."""

    def __init__(self, device_idx: int = 0, cuda_stream: int | None = None):
        self._h = _lib.c_dev()
        if cuda_stream is None:
            check(lib().b200rs_device_create(device_idx, ctypes.byref(self._h)), "b200rs_device_create")
        else:
            check(lib().b200rs_device_create_on_stream(device_idx, ctypes.c_void_p(cuda_stream), ctypes.byref(self._h)),
                  "b200rs_device_create_on_stream")
        self.m_type = TYPE_CL
        self.m_memoryUsage = 0  # bytes held by Buffers (Adl.h:150; checked in DeviceUtils.deallocate)
        self.device_idx = device_idx
        # pinned staging ring of Buffer.getHostPtr / returnHostPtr, as in include/Adl/Adl.h (Device::acquireStage)
        self._stages = [{"ptr": 0, "bytes": 0, "event": None, "mapped": False, "in_flight": False} for _ in range(2)]

    def _acquire_stage(self, nbytes: int):
        """(slot, pinned address): a block no live mapping holds and no copy still reads."""
        for wait in (False, True):
            for i, st in enumerate(self._stages):
                if st["mapped"]:
                    continue
                if st["in_flight"]:
                    if wait:
                        check(lib().b200rs_event_synchronize(self.handle, st["event"]), "b200rs_event_synchronize")
                    else:
                        done = ctypes.c_int(0)
                        check(lib().b200rs_event_query(self.handle, st["event"], ctypes.byref(done)), "b200rs_event_query")
                        if not done.value:
                            continue
                    st["in_flight"] = False
                if nbytes > st["bytes"]:
                    if st["ptr"]:
                        check(lib().b200rs_host_free(self.handle, ctypes.c_void_p(st["ptr"])), "b200rs_host_free")
                    p = ctypes.c_void_p()
                    check(lib().b200rs_host_alloc(self.handle, nbytes, ctypes.byref(p)), "b200rs_host_alloc")
                    st["ptr"], st["bytes"] = int(p.value), nbytes
                st["mapped"] = True
                return i, st["ptr"]
        raise RuntimeError("more than two buffers mapped at once")

    def _release_stage(self, slot: int, copy_enqueued: bool) -> None:
        st = self._stages[slot]
        st["mapped"] = False
        if copy_enqueued:
            if st["event"] is None:
                e = ctypes.c_void_p()
                check(lib().b200rs_event_create(self.handle, ctypes.byref(e)), "b200rs_event_create")
                st["event"] = e
            check(lib().b200rs_event_record(self.handle, st["event"]), "b200rs_event_record")
            st["in_flight"] = True

    @property
    def handle(self):
        if not self._h:
            raise RuntimeError("device already released")
        return self._h

    def getType(self) -> int:
        return self.m_type

    def getUsedMemory(self) -> int:
        return self.m_memoryUsage

    def getDeviceName(self) -> str:
        buf = ctypes.create_string_buffer(128)
        check(lib().b200rs_device_name(self.handle, buf), "b200rs_device_name")
        return buf.value.decode()

    def waitForCompletion(self) -> None:
        check(lib().b200rs_device_sync(self.handle), "b200rs_device_sync")

    def stream(self) -> int:
        return int(lib().b200rs_device_stream(self.handle) or 0)

    def launch_count(self) -> int:
        n = ctypes.c_uint64(0)
        check(lib().b200rs_device_launch_count(self.handle, ctypes.byref(n)), "b200rs_device_launch_count")
        return int(n.value)

    def toggleProfiling(self, enable: bool) -> None:  # Adl.h:142
        check(lib().b200rs_profile_enable(self.handle, 1 if enable else 0), "b200rs_profile_enable")

    def readProfile(self, capacity: int = 256) -> list[dict]:
        arr = (_lib.ProfileEntry * capacity)()
        cnt = ctypes.c_int(0)
        check(lib().b200rs_profile_read(self.handle, arr, capacity, ctypes.byref(cnt)), "b200rs_profile_read")
        return [{"kernel": arr[i].kernel.decode(), "ms": float(arr[i].ms), "elements": int(arr[i].elements), "bytes": int(arr[i].bytes)}
                for i in range(cnt.value)]

    def release(self) -> None:
        if self._h:
            lib().b200rs_device_sync(self._h)
            for st in self._stages:
                if st["ptr"]:
                    lib().b200rs_host_free(self._h, ctypes.c_void_p(st["ptr"]))
                if st["event"] is not None:
                    lib().b200rs_event_destroy(self._h, st["event"])
                st.update(ptr=0, bytes=0, event=None, mapped=False, in_flight=False)
            check(lib().b200rs_device_destroy(self._h), "b200rs_device_destroy")
            self._h = _lib.c_dev()

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class DeviceUtils:
    """adl::DeviceUtils (Adl.h:71-116)."""

    @staticmethod
    def getNDevices(type_: int = TYPE_CL) -> int:
        n = ctypes.c_int(0)
        rc = lib().b200rs_device_count(ctypes.byref(n))
        return int(n.value) if rc == 0 else 0

    @staticmethod
    def allocate(type_: int = TYPE_CL, device_idx: int = 0, cuda_stream: int | None = None) -> Device:
        if type_ != TYPE_CL:
            raise NotImplementedError("only the CUDA device exists (TYPE_CL); there is no Host/CPU or DX11 backend")
        return Device(device_idx, cuda_stream)

    @staticmethod
    def deallocate(device: Device) -> None:
        assert device.getUsedMemory() == 0, "buffers still allocated at device teardown (Adl.inl:102)"
        device.release()

    @staticmethod
    def waitForCompletion(device: Device) -> None:
        device.waitForCompletion()

    @staticmethod
    def getNCUs(device: Device) -> int:
        n = ctypes.c_int(0)
        check(lib().b200rs_device_num_sms(device.handle, ctypes.byref(n)), "b200rs_device_num_sms")
        return int(n.value)


class Buffer:
    """adl::Buffer<T> (Adl.h:164-222): a typed device allocation.

    dtype is np.uint32, np.int32 or PAIR_DTYPE.  `ptr` may wrap foreign device memory
    (setRawPtr, Adl.inl:238-253), e.g. a torch tensor's data_ptr(); such buffers are not owned.
    """

    def __init__(self, device: Device, nElems: int = 0, dtype=np.uint32, ptr: int | None = None):
        self.m_device = device
        self.dtype = np.dtype(dtype)
        self.m_size = 0
        self.m_ptr = 0
        self.m_allocated = False
        self._written = False  # contents defined on the device (see include/Adl/Adl.h: markDeviceWritten)
        self._map = None
        if ptr is not None:
            self.m_ptr, self.m_size = int(ptr), int(nElems)
            self._written = True
        elif nElems:
            self.allocate(nElems)

    def allocate(self, nElems: int) -> None:
        assert not self.m_allocated
        p = ctypes.c_void_p()
        nbytes = int(nElems) * self.dtype.itemsize
        check(lib().b200rs_malloc(self.m_device.handle, nbytes, ctypes.byref(p)), "b200rs_malloc")
        self.m_ptr, self.m_size, self.m_allocated = int(p.value or 0), int(nElems), True
        self.m_device.m_memoryUsage += nbytes

    def release(self) -> None:
        if self.m_allocated and self.m_ptr:
            check(lib().b200rs_free(self.m_device.handle, ctypes.c_void_p(self.m_ptr)), "b200rs_free")
            self.m_device.m_memoryUsage -= self.m_size * self.dtype.itemsize
        self.m_ptr, self.m_size, self.m_allocated = 0, 0, False

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def getSize(self) -> int:
        return self.m_size

    def setSize(self, size: int) -> None:  # grow-only, contents not preserved (Adl.inl:327-356)
        if not self.m_allocated:
            self.allocate(size)
        elif self.m_size < size:
            self.release()
            self.allocate(size)

    def write(self, host: np.ndarray, nElems: int | None = None, dstOffsetNElems: int = 0) -> None:
        host = np.ascontiguousarray(host)
        n = host.shape[0] if nElems is None else nElems
        assert n + dstOffsetNElems <= self.m_size
        self._written = True
        isz = self.dtype.itemsize
        check(lib().b200rs_memcpy_h2d(self.m_device.handle, ctypes.c_void_p(self.m_ptr + dstOffsetNElems * isz),
                                      ctypes.c_void_p(host.ctypes.data), n * isz), "b200rs_memcpy_h2d")
        self.m_device.waitForCompletion()  # the numpy source may be a temporary

    def read(self, nElems: int | None = None, srcOffsetNElems: int = 0) -> np.ndarray:
        n = self.m_size - srcOffsetNElems if nElems is None else nElems
        out = np.empty(n, dtype=self.dtype)
        isz = self.dtype.itemsize
        check(lib().b200rs_memcpy_d2h(self.m_device.handle, ctypes.c_void_p(out.ctypes.data),
                                      ctypes.c_void_p(self.m_ptr + srcOffsetNElems * isz), n * isz), "b200rs_memcpy_d2h")
        self.m_device.waitForCompletion()
        return out

    def markDeviceWritten(self) -> None:
        self._written = True

    # map/unmap semantics of the CL backend (AdlCL.inl:544-565): a read+write host view in pinned memory.  The copy out is
    # stream-ordered (wait before reading, as with the reference's non-blocking map; skipped for a buffer nothing was ever
    # written to), returnHostPtr enqueues the copy back and returns (the pinned block is recycled once it has passed).
    def getHostPtr(self, size: int | None = None) -> np.ndarray:
        n = self.m_size if size is None or size < 0 or size > self.m_size else size
        assert self._map is None, "one mapping at a time per buffer"
        nbytes = n * self.dtype.itemsize
        slot, addr = self.m_device._acquire_stage(max(nbytes, 1))
        view = np.ctypeslib.as_array(ctypes.cast(ctypes.c_void_p(addr), ctypes.POINTER(ctypes.c_uint8)), shape=(nbytes,)).view(self.dtype)
        if self._written and n:
            check(lib().b200rs_memcpy_d2h(self.m_device.handle, ctypes.c_void_p(addr), ctypes.c_void_p(self.m_ptr), nbytes), "b200rs_memcpy_d2h")
        self._map = (slot, addr, n)
        return view

    def returnHostPtr(self, host: np.ndarray) -> None:
        assert self._map is not None and host.ctypes.data == self._map[1], "not the view getHostPtr returned"
        slot, addr, n = self._map
        if n:
            check(lib().b200rs_memcpy_h2d(self.m_device.handle, ctypes.c_void_p(self.m_ptr), ctypes.c_void_p(addr), n * self.dtype.itemsize), "b200rs_memcpy_h2d")
        self._written = True
        self.m_device._release_stage(slot, bool(n))
        self._map = None
