// Tahoe/ParallelPrimitives/Pprims.h -- the reference's parallel-primitive entry points
// (reference: Tahoe/ParallelPrimitives/Pprims.h:11-48) over libb200rs.so:
//     scan      -> b200rs_exclusive_scan_u32   single-pass decoupled look-back, any n
//     radixSort -> b200rs_sort_keys_u32 / b200rs_sort_pairs_u32   one histogram + one scatter pass per 8-bit digit
// Same signatures and defaults.  All calls are asynchronous on the device's stream (results are visible
// after DeviceUtils::waitForCompletion), except that *sumOut is already valid when scan() returns.
// There is no Host fallback: a null device, or a device that is not the GPU, asserts and does nothing.
#pragma once

#include <Tahoe/ParallelPrimitives/uArray.h>

namespace Tahoe {

class Pprims {
public:
    TH_DECLARE_ALLOCATOR(Pprims);

    Pprims();
    ~Pprims();  // frees the scratch; destroy before DeviceUtils::deallocate (it asserts used memory == 0)

    void cacheKernel(bool cache) { m_cacheKernel = cache; }  // kept for source compatibility; kernels are built ahead of time

    // Names of the reference's tuning enum (Pprims.h:22-33) with the values the sm_100a kernels realise.
    enum {
        SCAN_BLOCK_SIZE = 256,            // threads per scan tile (4096 elements)
        RSORT_BITS_PER_PASS = 8,
        RSORT_NUM_TABLES = (1 << RSORT_BITS_PER_PASS),
        R32SORT_DATA_ALIGNMENT = 1,       // any n (the reference's key-only kernels needed multiples of 256)
        R32SORT_WG_SIZE = 256,
        R32SORT_ELEMENTS_PER_WORK_ITEM = 35,
        R32SORT_BITS_PER_PASS = 8,        // the reference's device path used 4
    };

    // dst[i] = src[0] + ... + src[i-1] in u32 arithmetic; *sumOut (optional) = sum of all n inputs
    void scan(const adl::Device* device, adl::Buffer<int>& dst, const adl::Buffer<int>& src, int n, u32* sumOut = 0);

    // inout.x: key, inout.y: value; stable
    void radixSort(const adl::Device* device, const adl::Buffer<uint2>& inout, int n, int sortBits = 32);

    // any n >= 0
    void radixSort(const adl::Device* device, const adl::Buffer<u32>& inout, int n, int sortBits = 32);

    // New capability (the reference is single-device, Adl/CL/AdlCL.inl:284-303): this rank's part of the key-value sort of an
    // input partitioned over `comm.world` GPUs, one Pprims per rank -- b200rs_dist_sort_pairs_u32 (include/b200rs.h): top-digit
    // histogram, all-gather, on-device plan, fused partition + peer stores, barrier, local stable sort.  `recvBases[r]` is a
    // device-visible address of rank r's receive buffer (own included; capacity `recvCapacity` pairs on every rank), the two
    // collectives are the caller's (b200rs_dist_comm).  Returns the number of pairs now sorted at recvBases[comm.rank] (the
    // ranks' results in rank order are the stable sort of the inputs in rank order), or -1 when a rank's share would exceed
    // recvCapacity (nothing is exchanged).  Waits for the device once, to read that count.
    long long radixSortDistributed(const adl::Device* device, const b200rs_dist_comm& comm, const u64* recvBases, u64 recvCapacity,
                                   const adl::Buffer<uint2>& in, int n);

    // Element-wise primitives: dst[i] = src[i] / dst[i] = src for i < n.  The reference declares them on uArray
    // (Pprims.cpp:31-121; commented out there, the OpenCL kernels CopyIntKernel ... FillF4Kernel are still
    // shipped, PprimsKernels.cl:9-48); the Buffer overloads are what those forward to.  Asynchronous.
    void copy(const adl::Device* device, uArray<int>& dst, const uArray<int>& src, int n);
    void copy(const adl::Device* device, uArray<float4>& dst, const uArray<float4>& src, int n);
    void fill(const adl::Device* device, uArray<int>& dst, int src, int n);
    void fill(const adl::Device* device, uArray<u32>& dst, u32 src, int n);
    void fill(const adl::Device* device, uArray<float4>& dst, const float4& src, int n);
    void copy(const adl::Device* device, adl::Buffer<int>& dst, const adl::Buffer<int>& src, int n);
    void copy(const adl::Device* device, adl::Buffer<u32>& dst, const adl::Buffer<u32>& src, int n);
    void copy(const adl::Device* device, adl::Buffer<float4>& dst, const adl::Buffer<float4>& src, int n);
    void fill(const adl::Device* device, adl::Buffer<int>& dst, int src, int n);
    void fill(const adl::Device* device, adl::Buffer<u32>& dst, u32 src, int n);
    void fill(const adl::Device* device, adl::Buffer<float4>& dst, const float4& src, int n);

private:
    void* reserveTemp(const adl::Device* device, size_t bytes);
    void releaseTemp();

    const adl::Device* m_device;  // device the scratch lives on
    void* m_temp;                 // grow-only device scratch (alternate buffer + histograms + look-back table)
    size_t m_tempBytes;
    bool m_cacheKernel;
};

}  // namespace Tahoe
