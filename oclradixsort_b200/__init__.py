"""oclradixsort_b200 -- B200 (sm_100a) implementation of OCLRadixSort's hot path.

Host-side Python mirror of the reference interface for this path:

    reference (C++)                                       here
    adl::DeviceUtils::allocate / adl::Device              adl.Device, adl.DeviceUtils
    adl::Buffer<T>                                        adl.Buffer
    Tahoe::Pprims::radixSort(device, Buffer<u32>, n, bits) Pprims.radixSort
    Tahoe::Pprims::radixSort(device, Buffer<uint2>, ...)   Pprims.radixSort (pair buffer)
    Tahoe::Pprims::scan(device, dst, src, n, sumOut)       Pprims.scan

All compute goes through the C ABI of libb200rs.so (include/b200rs.h); nothing here computes on
the CPU.  The C++ drop-in headers (include/Adl, include/Tahoe) are the primary boundary; this
package exists for tests and bench.py.
"""
from . import _lib  # noqa: F401
from .adl import Buffer, Device, DeviceUtils, TYPE_CL, TYPE_HOST, PAIR_DTYPE  # noqa: F401
from .pprims import Pprims  # noqa: F401

__all__ = ["Buffer", "Device", "DeviceUtils", "Pprims", "TYPE_CL", "TYPE_HOST", "PAIR_DTYPE"]
