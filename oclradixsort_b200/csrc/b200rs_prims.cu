// b200rs_prims.cu -- the element-wise primitives of Tahoe/ParallelPrimitives: copy and fill.
//
// Replaces CopyIntKernel / CopyF4Kernel / FillIntKernel / FillU32Kernel / FillF4Kernel
// (Tahoe/ClKernels/PprimsKernels.cl:9-48: one work-item per element) and their host side
// Pprims::copy / Pprims::fill (Tahoe/ParallelPrimitives/Pprims.cpp:31-121, dormant in the reference:
// the definitions are commented out, the kernels are still shipped).  SURVEY.md section 8f row 3.
//
// Both are pure HBM streams (copy: 4 B read + 4 B write per u32, fill: 4 B write), so the kernels
// are 128-bit grid-stride loops with COPY_UNROLL / FILL_UNROLL independent accesses per thread in flight and a
// grid of a few CTAs per SM; ragged ends (pointers that are only 4-byte aligned, n not a multiple
// of 4) are handled element-wise by the first CTA.
#include "b200rs_internal.h"

namespace {

constexpr int PRIM_THREADS = 256;
constexpr int COPY_UNROLL = 8;   // uint4 loads in flight per thread (copy)
constexpr int FILL_UNROLL = 4;   // uint4 stores per thread and iteration (fill)
constexpr int CTAS_PER_SM = 8;

// dst[0..n) = src[0..n) for 4-byte elements.  `head` elements (0..3) precede the first 16-byte boundary of dst;
// the vector body is used only when src has the same misalignment (otherwise everything is `head`-style).
template <int UNROLL, bool STREAM>
__global__ void __launch_bounds__(PRIM_THREADS)
copy_u32_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, uint64_t n, uint32_t head, uint64_t nvec) {
    const uint4* __restrict__ s4 = reinterpret_cast<const uint4*>(src + head);
    uint4* __restrict__ d4 = reinterpret_cast<uint4*>(dst + head);
    const uint64_t chunk = (uint64_t)PRIM_THREADS * UNROLL;
    for (uint64_t base = (uint64_t)blockIdx.x * chunk; base < nvec; base += (uint64_t)gridDim.x * chunk) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint64_t i = base + (uint64_t)u * PRIM_THREADS + threadIdx.x;
            if (i < nvec) v[u] = STREAM ? __ldcs(s4 + i) : __ldg(s4 + i);  // streamed: neither side is reused by this kernel
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint64_t i = base + (uint64_t)u * PRIM_THREADS + threadIdx.x;
            if (i < nvec) {
                if (STREAM) __stcs(d4 + i, v[u]);
                else d4[i] = v[u];
            }
        }
    }
    if (blockIdx.x == 0) {
        if (threadIdx.x < head) dst[threadIdx.x] = src[threadIdx.x];
        const uint64_t t = head + nvec * 4 + threadIdx.x;  // at most 3 trailing elements
        if (t < n) dst[t] = src[t];
    }
}

// pointers with different misalignment: element-wise (still coalesced, 4 B per lane)
__global__ void __launch_bounds__(PRIM_THREADS)
copy_u32_elementwise_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * PRIM_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * PRIM_THREADS + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

// dst[0..n) = value pattern.  pattern = the 16 bytes to replicate, already rotated so that it lines up with the
// first 16-byte boundary of dst (for u32 fills all four words are the same value).
__global__ void __launch_bounds__(PRIM_THREADS)
fill_kernel(uint32_t* __restrict__ dst, uint4 pattern, uint64_t n, uint32_t head, uint64_t nvec, uint32_t value_if_u32) {
    uint4* __restrict__ d4 = reinterpret_cast<uint4*>(dst + head);
    const uint64_t chunk = (uint64_t)PRIM_THREADS * FILL_UNROLL;
    for (uint64_t base = (uint64_t)blockIdx.x * chunk; base < nvec; base += (uint64_t)gridDim.x * chunk) {
#pragma unroll
        for (int u = 0; u < FILL_UNROLL; ++u) {
            const uint64_t i = base + (uint64_t)u * PRIM_THREADS + threadIdx.x;
            if (i < nvec) __stcs(d4 + i, pattern);
        }
    }
    if (blockIdx.x == 0) {  // only the u32 form has ragged ends (16-byte elements are 16-byte aligned)
        if (threadIdx.x < head) dst[threadIdx.x] = value_if_u32;
        const uint64_t t = head + nvec * 4 + threadIdx.x;
        if (t < n) dst[t] = value_if_u32;
    }
}

int grid_for(const b200rs_device* dev, uint64_t work_items, uint64_t items_per_cta) {
    const uint64_t want = (work_items + items_per_cta - 1) / items_per_cta;
    const uint64_t cap = (uint64_t)dev->num_sms * CTAS_PER_SM;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

int launch_copy_u32(b200rs_device* dev, uint32_t* dst, const uint32_t* src, uint64_t n, const char* name) {
    if (n == 0 || dst == src) return B200RS_OK;
    if (((uintptr_t)dst | (uintptr_t)src) & 3u) return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    b200rs_launch_scope scope(dev, name, n, n * 8);
    if ((((uintptr_t)dst ^ (uintptr_t)src) & 15u) != 0) {
        copy_u32_elementwise_kernel<<<grid_for(dev, n, PRIM_THREADS * 16), PRIM_THREADS, 0, dev->stream>>>(dst, src, n);
    } else {
        uint64_t head = ((16u - ((uintptr_t)dst & 15u)) & 15u) / 4;
        if (head > n) head = n;
        const uint64_t nvec = (n - head) / 4;
        // measured at 2^28 u32: 8 loads in flight + streaming hints 6.31 TB/s, 4 + hints 5.89, 8 plain 6.16, 4 plain 5.93
        copy_u32_kernel<COPY_UNROLL, true><<<grid_for(dev, nvec, PRIM_THREADS * COPY_UNROLL), PRIM_THREADS, 0, dev->stream>>>(dst, src, n, (uint32_t)head, nvec);
    }
    B200RS_CUDA(cudaGetLastError());
    return B200RS_OK;
}

}  // namespace

extern "C" {

int b200rs_copy_u32(b200rs_device* dev, uint32_t* dst, const uint32_t* src, uint64_t n) {
    if (!dev || (n && (!dst || !src))) return B200RS_ERR_INVALID_ARGUMENT;
    return launch_copy_u32(dev, dst, src, n, "copy_u32");
}

int b200rs_copy_u128(b200rs_device* dev, void* dst, const void* src, uint64_t n) {
    if (!dev || (n && (!dst || !src))) return B200RS_ERR_INVALID_ARGUMENT;
    if (((uintptr_t)dst | (uintptr_t)src) & 15u) return B200RS_ERR_INVALID_ARGUMENT;  // float4 elements are 16-byte aligned
    if (n >> 60) return B200RS_ERR_TOO_LARGE;
    return launch_copy_u32(dev, (uint32_t*)dst, (const uint32_t*)src, n * 4, "copy_u128");
}

int b200rs_fill_u32(b200rs_device* dev, uint32_t* dst, uint32_t value, uint64_t n) {
    if (!dev || (n && !dst) || ((uintptr_t)dst & 3u)) return B200RS_ERR_INVALID_ARGUMENT;
    if (n == 0) return B200RS_OK;
    b200rs_device_guard guard(dev);
    b200rs_launch_scope scope(dev, "fill_u32", n, n * 4);
    uint64_t head = ((16u - ((uintptr_t)dst & 15u)) & 15u) / 4;
    if (head > n) head = n;
    const uint64_t nvec = (n - head) / 4;
    fill_kernel<<<grid_for(dev, nvec, PRIM_THREADS * FILL_UNROLL), PRIM_THREADS, 0, dev->stream>>>(dst, make_uint4(value, value, value, value), n,
                                                                                                 (uint32_t)head, nvec, value);
    B200RS_CUDA(cudaGetLastError());
    return B200RS_OK;
}

int b200rs_fill_u128(b200rs_device* dev, void* dst, const uint32_t value[4], uint64_t n) {
    if (!dev || !value || (n && !dst) || ((uintptr_t)dst & 15u)) return B200RS_ERR_INVALID_ARGUMENT;
    if (n == 0) return B200RS_OK;
    if (n >> 60) return B200RS_ERR_TOO_LARGE;
    b200rs_device_guard guard(dev);
    b200rs_launch_scope scope(dev, "fill_u128", n, n * 16);
    fill_kernel<<<grid_for(dev, n, PRIM_THREADS * FILL_UNROLL), PRIM_THREADS, 0, dev->stream>>>((uint32_t*)dst, make_uint4(value[0], value[1], value[2], value[3]),
                                                                                              n * 4, 0u, n, 0u);
    B200RS_CUDA(cudaGetLastError());
    return B200RS_OK;
}

}  // extern "C"
