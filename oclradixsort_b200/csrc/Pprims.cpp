// Pprims.cpp -- Tahoe::Pprims (include/Tahoe/ParallelPrimitives/Pprims.h) over the C ABI of libb200rs.so.
// Replaces the reference's host orchestration of OpenCL launches (Tahoe/ParallelPrimitives/Pprims.cpp:122-406).
// Plain C++ (no CUDA headers): link against libb200rs.so.
#include <Tahoe/ParallelPrimitives/Pprims.h>

#include <string.h>

namespace Tahoe {

namespace {
// GPU path iff the device is the GPU type -- the reference's enableSortOnDevice (Pprims.cpp:189-198);
// the "else" branch there is the Host fallback, which does not exist here.
bool isGpuDevice(const adl::Device* device) {
    return device && device->getType() == adl::TYPE_CL && device->getProcType() == adl::Device::Config::DEVICE_GPU && device->getHandle();
}
}  // namespace

Pprims::Pprims() : m_device(0), m_temp(0), m_tempBytes(0), m_cacheKernel(true) {}

Pprims::~Pprims() { releaseTemp(); }

void Pprims::releaseTemp() {
    if (m_temp && m_device) {
        adl::adlCheck(b200rs_free(m_device->getHandle(), m_temp), "b200rs_free");  // waits for in-flight work
        m_device->m_memoryUsage -= m_tempBytes;
    }
    m_temp = 0;
    m_tempBytes = 0;
}

void* Pprims::reserveTemp(const adl::Device* device, size_t bytes) {
    if (m_temp && (m_device != device || m_tempBytes < bytes)) releaseTemp();
    if (!m_temp) {
        void* p = 0;
        if (!adl::adlCheck(b200rs_malloc(device->getHandle(), bytes, &p), "b200rs_malloc")) return 0;
        m_temp = p;
        m_tempBytes = bytes;
        m_device = device;
        device->m_memoryUsage += bytes;  // counted like any buffer so the teardown check stays meaningful
    }
    return m_temp;
}

// ---- copy / fill (reference: Pprims.cpp:31-121, PprimsKernels.cl:9-48) ----
namespace {
bool primArgsOk(const adl::Device* device, int n, u64 dstSize, u64 srcSize) {
    if (!isGpuDevice(device) || n < 0) {
        ADLASSERT(0);  // the reference's device == 0 branch is a CPU loop; there is no Host path here
        return false;
    }
    ADLASSERT((u64)n <= dstSize && (u64)n <= srcSize);
    return (u64)n <= dstSize && (u64)n <= srcSize;
}
}  // namespace

void Pprims::copy(const adl::Device* device, adl::Buffer<int>& dst, const adl::Buffer<int>& src, int n) {
    if (!primArgsOk(device, n, dst.getSize(), src.getSize())) return;
    dst.markDeviceWritten();
    adl::adlCheck(b200rs_copy_u32(device->getHandle(), (uint32_t*)dst.m_ptr, (const uint32_t*)src.m_ptr, (uint64_t)n), "b200rs_copy_u32");
}
void Pprims::copy(const adl::Device* device, adl::Buffer<u32>& dst, const adl::Buffer<u32>& src, int n) {
    if (!primArgsOk(device, n, dst.getSize(), src.getSize())) return;
    dst.markDeviceWritten();
    adl::adlCheck(b200rs_copy_u32(device->getHandle(), (uint32_t*)dst.m_ptr, (const uint32_t*)src.m_ptr, (uint64_t)n), "b200rs_copy_u32");
}
void Pprims::copy(const adl::Device* device, adl::Buffer<float4>& dst, const adl::Buffer<float4>& src, int n) {
    if (!primArgsOk(device, n, dst.getSize(), src.getSize())) return;
    dst.markDeviceWritten();
    adl::adlCheck(b200rs_copy_u128(device->getHandle(), dst.m_ptr, src.m_ptr, (uint64_t)n), "b200rs_copy_u128");
}
void Pprims::fill(const adl::Device* device, adl::Buffer<int>& dst, int src, int n) {
    if (!primArgsOk(device, n, dst.getSize(), (u64)n)) return;
    dst.markDeviceWritten();
    adl::adlCheck(b200rs_fill_u32(device->getHandle(), (uint32_t*)dst.m_ptr, (uint32_t)src, (uint64_t)n), "b200rs_fill_u32");
}
void Pprims::fill(const adl::Device* device, adl::Buffer<u32>& dst, u32 src, int n) {
    if (!primArgsOk(device, n, dst.getSize(), (u64)n)) return;
    dst.markDeviceWritten();
    adl::adlCheck(b200rs_fill_u32(device->getHandle(), (uint32_t*)dst.m_ptr, src, (uint64_t)n), "b200rs_fill_u32");
}
void Pprims::fill(const adl::Device* device, adl::Buffer<float4>& dst, const float4& src, int n) {
    if (!primArgsOk(device, n, dst.getSize(), (u64)n)) return;
    dst.markDeviceWritten();
    uint32_t words[4];
    memcpy(words, &src, sizeof(words));
    adl::adlCheck(b200rs_fill_u128(device->getHandle(), dst.m_ptr, words, (uint64_t)n), "b200rs_fill_u128");
}
// uArray forms: the device copy is brought up to date and the CPU copy marked stale (uArray::getGpuBuffer), exactly
// what setToLauncher did for the reference's launches (uArray.h:214-228).
void Pprims::copy(const adl::Device* device, uArray<int>& dst, const uArray<int>& src, int n) {
    copy(device, *const_cast<adl::Buffer<int>*>(dst.getGpuBuffer(device)), *src.getGpuBuffer(device), n);
}
void Pprims::copy(const adl::Device* device, uArray<float4>& dst, const uArray<float4>& src, int n) {
    copy(device, *const_cast<adl::Buffer<float4>*>(dst.getGpuBuffer(device)), *src.getGpuBuffer(device), n);
}
void Pprims::fill(const adl::Device* device, uArray<int>& dst, int src, int n) {
    fill(device, *const_cast<adl::Buffer<int>*>(dst.getGpuBuffer(device)), src, n);
}
void Pprims::fill(const adl::Device* device, uArray<u32>& dst, u32 src, int n) {
    fill(device, *const_cast<adl::Buffer<u32>*>(dst.getGpuBuffer(device)), src, n);
}
void Pprims::fill(const adl::Device* device, uArray<float4>& dst, const float4& src, int n) {
    fill(device, *const_cast<adl::Buffer<float4>*>(dst.getGpuBuffer(device)), src, n);
}

void Pprims::scan(const adl::Device* device, adl::Buffer<int>& dst, const adl::Buffer<int>& src, int n, u32* sumOut) {
    if (!isGpuDevice(device) || n < 0) {
        ADLASSERT(0);  // reference: Pprims.cpp:124-127
        return;
    }
    ADLASSERT((u64)n <= dst.getSize() && (u64)n <= src.getSize());
    dst.markDeviceWritten();
    size_t need = 0;
    if (!adl::adlCheck(b200rs_exclusive_scan_u32(device->getHandle(), 0, 0, (uint64_t)n, 0, 0, &need), "b200rs_exclusive_scan_u32(size)")) return;
    const size_t totalSlot = (need + 255) / 256 * 256;  // one extra word behind the scan's own scratch holds the total
    char* temp = (char*)reserveTemp(device, totalSlot + 256);
    if (!temp) return;
    uint32_t* total = sumOut ? (uint32_t*)(temp + totalSlot) : 0;
    size_t have = totalSlot;
    if (!adl::adlCheck(b200rs_exclusive_scan_u32(device->getHandle(), (uint32_t*)dst.m_ptr, (const uint32_t*)src.m_ptr, (uint64_t)n, total, temp, &have),
                       "b200rs_exclusive_scan_u32"))
        return;
    if (sumOut) {
        // the reference enqueues a non-blocking read here (Pprims.cpp:164-167); this one completes before returning
        adl::adlCheck(b200rs_memcpy_d2h(device->getHandle(), sumOut, total, sizeof(u32)), "b200rs_memcpy_d2h");
        device->waitForCompletion();
    }
}

void Pprims::radixSort(const adl::Device* device, const adl::Buffer<uint2>& inout, int n, int sortBits) {
    if (!isGpuDevice(device) || n < 0) {
        ADLASSERT(0);  // no Host fallback (the reference would run RadixSort::sort here, Pprims.cpp:202-212)
        return;
    }
    ADLASSERT(sortBits >= 0 && sortBits <= 32);
    ADLASSERT((u64)n <= inout.getSize());
    size_t need = 0;
    if (!adl::adlCheck(b200rs_sort_pairs_u32(device->getHandle(), 0, (uint64_t)n, sortBits, 0, &need), "b200rs_sort_pairs_u32(size)")) return;
    void* temp = reserveTemp(device, need);
    if (!temp) return;
    size_t have = m_tempBytes;
    adl::adlCheck(b200rs_sort_pairs_u32(device->getHandle(), (b200rs_pair*)inout.m_ptr, (uint64_t)n, sortBits, temp, &have), "b200rs_sort_pairs_u32");
}

long long Pprims::radixSortDistributed(const adl::Device* device, const b200rs_dist_comm& comm, const u64* recvBases, u64 recvCapacity,
                                       const adl::Buffer<uint2>& in, int n) {
    if (!isGpuDevice(device) || n < 0 || !recvBases) {
        ADLASSERT(0);
        return -1;
    }
    ADLASSERT((u64)n <= in.getSize());
    size_t need = 0;
    if (!adl::adlCheck(b200rs_dist_sort_pairs_u32(device->getHandle(), &comm, 0, recvCapacity, 0, (uint64_t)n, 0, 0, 0, &need), "b200rs_dist_sort_pairs_u32(size)"))
        return -1;
    char* temp = (char*)reserveTemp(device, (need + 255) / 256 * 256 + 256);  // + [counts 2 x u64 | status u32]
    if (!temp) return -1;
    uint64_t* counts = (uint64_t*)(temp + (need + 255) / 256 * 256);
    uint32_t* status = (uint32_t*)(counts + 2);
    size_t have = need;
    uint64_t recv[32];
    for (int r = 0; r < comm.world && r < 32; ++r) recv[r] = recvBases[r];
    if (!adl::adlCheck(b200rs_dist_sort_pairs_u32(device->getHandle(), &comm, recv, recvCapacity, (const b200rs_pair*)in.m_ptr, (uint64_t)n, counts, status,
                                                  temp, &have), "b200rs_dist_sort_pairs_u32"))
        return -1;
    uint64_t host[3] = {0, 0, 0};
    adl::adlCheck(b200rs_memcpy_d2h(device->getHandle(), host, counts, sizeof(host)), "b200rs_memcpy_d2h");
    device->waitForCompletion();
    return (uint32_t)host[2] != 0 ? -1 : (long long)host[1];
}

void Pprims::radixSort(const adl::Device* device, const adl::Buffer<u32>& inout, int n, int sortBits) {
    if (!isGpuDevice(device) || n < 0) {
        ADLASSERT(0);  // no Host fallback (Pprims.cpp:306-316 in the reference)
        return;
    }
    ADLASSERT(sortBits >= 0 && sortBits <= 32);
    ADLASSERT((u64)n <= inout.getSize());
    size_t need = 0;
    if (!adl::adlCheck(b200rs_sort_keys_u32(device->getHandle(), 0, (uint64_t)n, sortBits, 0, &need), "b200rs_sort_keys_u32(size)")) return;
    void* temp = reserveTemp(device, need);
    if (!temp) return;
    size_t have = m_tempBytes;
    adl::adlCheck(b200rs_sort_keys_u32(device->getHandle(), (uint32_t*)inout.m_ptr, (uint64_t)n, sortBits, temp, &have), "b200rs_sort_keys_u32");
}

}  // namespace Tahoe
