// b200rs_scan.cu -- exclusive u32 prefix sum, single pass, decoupled look-back (sm_100a).
//
// Replaces the reference's three-kernel chain LocalScanKernel / TopLevelScanKernel / AddOffsetKernel
// (Tahoe/ClKernels/PrefixScanKernels.cl:72-143, driven by Pprims::scan, Pprims.cpp:122-179), which
// moves 16 B/element and refuses n >= 1048576 (Pprims.cpp:132-138).  This kernel reads every element
// once and writes it once (8 B/element, the roofline figure of SURVEY.md section 8d) at any n.
//
// Layout: a tile is 4096 consecutive elements; warp w of the CTA owns the 512-element slice
// [w*512, (w+1)*512) and reads it as four fully coalesced 512-byte rows (one uint4 per lane per row),
// so element order inside a warp is (row, lane, component).  Tile ids come from an atomic ticket so a
// tile only ever waits on tiles that are already running.  Per tile one 64-bit descriptor
// {status:32 | value:32} is published with a single relaxed store: status 1 = tile aggregate,
// status 2 = inclusive prefix; warp 0 looks back 32 descriptors at a time.
#include "b200rs_internal.h"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
constexpr int SCAN_ROWS = 4;                            // uint4 rows per warp slice
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ROWS * 4;  // 4096 elements = 16 KiB
constexpr int SCAN_WARP_SLICE = 32 * SCAN_ROWS * 4;      // 512 elements

constexpr uint64_t DESC_AGGREGATE = 1ull << 32;
constexpr uint64_t DESC_INCLUSIVE = 2ull << 32;

__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ uint32_t warp_inclusive_sum(uint32_t x, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    return x;
}

template <bool VEC16>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_lookback_kernel(const uint32_t* src, uint32_t* dst, uint64_t n, uint64_t* desc, uint32_t* ticket, uint32_t* total_out,
                     uint32_t num_tiles) {
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp_total[SCAN_WARPS];
    __shared__ uint32_t s_warp_excl[SCAN_WARPS];
    __shared__ uint32_t s_tile_prefix;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t tile_base = (uint64_t)tile * SCAN_TILE;
    const uint64_t slice_base = tile_base + (uint64_t)warp * SCAN_WARP_SLICE;
    const bool full = tile_base + SCAN_TILE <= n;

    // ---- load (whole tile lands in registers before anything is stored: dst may alias src) ----
    uint32_t v[SCAN_ROWS][4];
    if (VEC16 && full) {
#pragma unroll
        for (int r = 0; r < SCAN_ROWS; ++r) {
            const uint4 q = *reinterpret_cast<const uint4*>(src + slice_base + (uint64_t)(r * 32 + lane) * 4);
            v[r][0] = q.x; v[r][1] = q.y; v[r][2] = q.z; v[r][3] = q.w;
        }
    } else {
#pragma unroll
        for (int r = 0; r < SCAN_ROWS; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint64_t i = slice_base + (uint64_t)(r * 32 + lane) * 4 + c;
                v[r][c] = i < n ? src[i] : 0u;
            }
    }

    // ---- warp-level exclusive scan in (row, lane, component) order ----
    uint32_t excl[SCAN_ROWS];
    uint32_t carry = 0;
#pragma unroll
    for (int r = 0; r < SCAN_ROWS; ++r) {
        const uint32_t s = v[r][0] + v[r][1] + v[r][2] + v[r][3];
        const uint32_t inc = warp_inclusive_sum(s, lane);
        excl[r] = carry + inc - s;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 31) s_warp_total[warp] = carry;
    __syncthreads();

    // ---- warp 0: scan of the warp totals, publish, look back ----
    if (warp == 0) {
        const uint32_t wt = lane < SCAN_WARPS ? s_warp_total[lane] : 0u;
        const uint32_t winc = warp_inclusive_sum(wt, lane);
        if (lane < SCAN_WARPS) s_warp_excl[lane] = winc - wt;
        const uint32_t aggregate = __shfl_sync(0xffffffffu, winc, 31);

        uint32_t prefix = 0;
        if (tile == 0) {
            if (lane == 0) st_relaxed_u64(&desc[0], DESC_INCLUSIVE | aggregate);
        } else {
            if (lane == 0) st_relaxed_u64(&desc[tile], DESC_AGGREGATE | aggregate);
            int64_t look = (int64_t)tile - 1;  // lane L inspects tile look-L; tiles < 0 act as {inclusive, 0}
            while (true) {
                const int64_t idx = look - lane;
                uint64_t w = idx >= 0 ? ld_relaxed_u64(&desc[idx]) : DESC_INCLUSIVE;
                while (__any_sync(0xffffffffu, (w >> 32) == 0)) {
                    if ((w >> 32) == 0) w = ld_relaxed_u64(&desc[idx]);
                }
                const uint32_t inc_lanes = __ballot_sync(0xffffffffu, (w >> 32) == 2);
                const int stop = inc_lanes ? (__ffs(inc_lanes) - 1) : 31;  // nearest tile with a full prefix
                uint32_t contrib = lane <= stop ? (uint32_t)w : 0u;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
                prefix += contrib;
                if (inc_lanes) break;
                look -= 32;
            }
            if (lane == 0) st_relaxed_u64(&desc[tile], DESC_INCLUSIVE | (uint32_t)(prefix + aggregate));
        }
        if (lane == 0) {
            s_tile_prefix = prefix;
            if (total_out && tile == num_tiles - 1) *total_out = prefix + aggregate;
        }
    }
    __syncthreads();

    // ---- store ----
    const uint32_t base = s_tile_prefix + s_warp_excl[warp];
    if (VEC16 && full) {
#pragma unroll
        for (int r = 0; r < SCAN_ROWS; ++r) {
            uint4 q;
            uint32_t run = base + excl[r];
            q.x = run; run += v[r][0];
            q.y = run; run += v[r][1];
            q.z = run; run += v[r][2];
            q.w = run;
            *reinterpret_cast<uint4*>(dst + slice_base + (uint64_t)(r * 32 + lane) * 4) = q;
        }
    } else {
#pragma unroll
        for (int r = 0; r < SCAN_ROWS; ++r) {
            uint32_t run = base + excl[r];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint64_t i = slice_base + (uint64_t)(r * 32 + lane) * 4 + c;
                if (i < n) dst[i] = run;
                run += v[r][c];
            }
        }
    }
}

constexpr size_t SCAN_HEADER_BYTES = 256;  // ticket word, padded so descriptors stay 256 B aligned

}  // namespace

extern "C" int b200rs_exclusive_scan_u32(b200rs_device* dev, uint32_t* dst, const uint32_t* src, uint64_t n, uint32_t* total_out,
                                         void* temp, size_t* temp_bytes) {
    if (!dev || !temp_bytes) return B200RS_ERR_INVALID_ARGUMENT;
    const uint64_t num_tiles64 = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (num_tiles64 > 0x7fffffffull) return B200RS_ERR_TOO_LARGE;
    const uint32_t num_tiles = (uint32_t)num_tiles64;
    const size_t need = SCAN_HEADER_BYTES + b200rs_align_up((size_t)num_tiles * sizeof(uint64_t), 256);
    if (!temp) {
        *temp_bytes = need;
        return B200RS_OK;
    }
    if (*temp_bytes < need) return B200RS_ERR_TEMP_TOO_SMALL;
    if (n && (!dst || !src)) return B200RS_ERR_INVALID_ARGUMENT;
    if ((((uintptr_t)dst | (uintptr_t)src | (uintptr_t)total_out) & 3u) || ((uintptr_t)temp & 7u)) return B200RS_ERR_INVALID_ARGUMENT;

    b200rs_device_guard guard(dev);
    if (n == 0) {
        if (total_out) B200RS_CUDA(cudaMemsetAsync(total_out, 0, sizeof(uint32_t), dev->stream));
        return B200RS_OK;
    }
    B200RS_CUDA(cudaMemsetAsync(temp, 0, need, dev->stream));
    uint32_t* ticket = reinterpret_cast<uint32_t*>(temp);
    uint64_t* desc = reinterpret_cast<uint64_t*>(static_cast<char*>(temp) + SCAN_HEADER_BYTES);
    const bool vec16 = (((uintptr_t)dst | (uintptr_t)src) & 15u) == 0;
    {
        b200rs_launch_scope scope(dev, "scan_lookback", n, n * 8ull);
        if (vec16)
            scan_lookback_kernel<true><<<num_tiles, SCAN_THREADS, 0, dev->stream>>>(src, dst, n, desc, ticket, total_out, num_tiles);
        else
            scan_lookback_kernel<false><<<num_tiles, SCAN_THREADS, 0, dev->stream>>>(src, dst, n, desc, ticket, total_out, num_tiles);
    }
    B200RS_CUDA(cudaGetLastError());
    return B200RS_OK;
}
