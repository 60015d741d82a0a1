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

// Tile shape: THREADS x ROWS uint4 per thread.  The default (512 x 8 = 16384 elements = 64 KiB per tile) keeps
// 8 independent 16-byte loads in flight per thread and makes one look-back serve 64 KiB of data; the smaller
// shapes are kept for measurement (B200RS_SCAN_VARIANT).
constexpr int SCAN_MIN_TILE = 4096;  // temp storage is sized for the smallest tile among the variants

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

template <int SCAN_THREADS, int SCAN_ROWS, bool VEC16>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_lookback_kernel(const uint32_t* src, uint32_t* dst, uint64_t n, uint64_t* desc, uint32_t* ticket, uint32_t* total_out,
                     uint32_t num_tiles) {
    constexpr int SCAN_WARPS = SCAN_THREADS / 32;
    constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ROWS * 4;
    constexpr int SCAN_WARP_SLICE = 32 * SCAN_ROWS * 4;
    static_assert(SCAN_WARPS <= 32, "warp 0 scans the warp totals in one step");
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

// -------------------------------------------------------------------------------------------------
// Persistent, software-pipelined variant (the default).  The grid is exactly the number of CTAs that are
// resident at once; every CTA loops over tiles taken from the atomic ticket and keeps the NEXT tile's loads in
// flight (second register set) while warp 0 looks back and the CTA stores the current tile, so an SM always has
// tile-sized reads outstanding.  The ticket for tile k+1 is fetched while tile k is being scanned.
// Forward progress: all CTAs are co-resident and a tile only waits on lower tickets.
// dst == src stays legal: the prefetched tile k+1 is only ever written by this same CTA, later.
// -------------------------------------------------------------------------------------------------
template <int SCAN_THREADS, int SCAN_ROWS, bool VEC16>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_persistent_kernel(const uint32_t* src, uint32_t* dst, uint64_t n, uint64_t* desc, uint32_t* ticket, uint32_t* total_out,
                       uint32_t num_tiles) {
    constexpr int SCAN_WARPS = SCAN_THREADS / 32;
    constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ROWS * 4;
    constexpr int SCAN_WARP_SLICE = 32 * SCAN_ROWS * 4;
    static_assert(SCAN_WARPS <= 32, "warp 0 scans the warp totals in one step");
    __shared__ uint32_t s_next_tile[2];
    __shared__ uint32_t s_warp_total[SCAN_WARPS];
    __shared__ uint32_t s_warp_excl[SCAN_WARPS];
    __shared__ uint32_t s_tile_prefix;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t in_slice = (uint32_t)warp * SCAN_WARP_SLICE + (uint32_t)lane * 4;  // element offset of row 0 inside a tile

    auto load_tile = [&](uint32_t tile, uint4 (&v)[SCAN_ROWS]) {
        if (tile >= num_tiles) return;
        const uint64_t base = (uint64_t)tile * SCAN_TILE + in_slice;
        if (VEC16 && (uint64_t)tile * SCAN_TILE + SCAN_TILE <= n) {
#pragma unroll
            for (int r = 0; r < SCAN_ROWS; ++r) v[r] = *reinterpret_cast<const uint4*>(src + base + r * 128);
        } else {
#pragma unroll
            for (int r = 0; r < SCAN_ROWS; ++r) {
                const uint64_t i = base + r * 128;
                v[r].x = i + 0 < n ? src[i + 0] : 0u;
                v[r].y = i + 1 < n ? src[i + 1] : 0u;
                v[r].z = i + 2 < n ? src[i + 2] : 0u;
                v[r].w = i + 3 < n ? src[i + 3] : 0u;
            }
        }
    };

    // One pipeline step: `cur` holds tile `tile` (already loaded or in flight); `nxt` receives the following tile.
    // Returns that following tile's id.
    auto step = [&](uint32_t tile, uint4 (&cur)[SCAN_ROWS], uint4 (&nxt)[SCAN_ROWS], int parity) -> uint32_t {
        if (tid == 0) s_next_tile[parity] = atomicAdd(ticket, 1u);  // latency hidden behind the row scans below
        // ---- warp-level exclusive scan in (row, lane, component) order ----
        uint32_t excl[SCAN_ROWS];
        uint32_t carry = 0;
#pragma unroll
        for (int r = 0; r < SCAN_ROWS; ++r) {
            const uint32_t sum = cur[r].x + cur[r].y + cur[r].z + cur[r].w;
            const uint32_t inc = warp_inclusive_sum(sum, lane);
            excl[r] = carry + inc - sum;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 31) s_warp_total[warp] = carry;
        __syncthreads();
        const uint32_t next_tile = s_next_tile[parity];
        load_tile(next_tile, nxt);  // in flight during the look-back and the stores

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
        const uint32_t base_sum = s_tile_prefix + s_warp_excl[warp];
        const uint64_t base = (uint64_t)tile * SCAN_TILE + in_slice;
        if (VEC16 && (uint64_t)tile * SCAN_TILE + SCAN_TILE <= n) {
#pragma unroll
            for (int r = 0; r < SCAN_ROWS; ++r) {
                uint4 q;
                uint32_t run = base_sum + excl[r];
                q.x = run; run += cur[r].x;
                q.y = run; run += cur[r].y;
                q.z = run; run += cur[r].z;
                q.w = run;
                *reinterpret_cast<uint4*>(dst + base + r * 128) = q;
            }
        } else {
#pragma unroll
            for (int r = 0; r < SCAN_ROWS; ++r) {
                const uint64_t i = base + r * 128;
                uint32_t run = base_sum + excl[r];
                if (i + 0 < n) dst[i + 0] = run; run += cur[r].x;
                if (i + 1 < n) dst[i + 1] = run; run += cur[r].y;
                if (i + 2 < n) dst[i + 2] = run; run += cur[r].z;
                if (i + 3 < n) dst[i + 3] = run;
            }
        }
        return next_tile;
    };

    if (tid == 0) s_next_tile[1] = atomicAdd(ticket, 1u);
    __syncthreads();
    uint32_t tile = s_next_tile[1];
    uint4 a[SCAN_ROWS], b[SCAN_ROWS];
    load_tile(tile, a);
    while (tile < num_tiles) {
        tile = step(tile, a, b, 0);
        if (tile >= num_tiles) break;
        tile = step(tile, b, a, 1);
    }
}

constexpr size_t SCAN_HEADER_BYTES = 256;  // ticket word, padded so descriptors stay 256 B aligned

struct ScanVariant {
    void (*aligned)(const uint32_t*, uint32_t*, uint64_t, uint64_t*, uint32_t*, uint32_t*, uint32_t);
    void (*unaligned)(const uint32_t*, uint32_t*, uint64_t, uint64_t*, uint32_t*, uint32_t*, uint32_t);
    int threads, tile;
    bool persistent;
};
#define B200RS_SCAN_VARIANT(T, R) ScanVariant{scan_lookback_kernel<T, R, true>, scan_lookback_kernel<T, R, false>, T, T * R * 4, false}
#define B200RS_SCAN_PERSISTENT(T, R) ScanVariant{scan_persistent_kernel<T, R, true>, scan_persistent_kernel<T, R, false>, T, T * R * 4, true}
const ScanVariant& pick_scan_variant() {
    static const ScanVariant v[] = {
        B200RS_SCAN_VARIANT(512, 8),  // default: 16384 elements per tile
        B200RS_SCAN_VARIANT(256, 8),
        B200RS_SCAN_VARIANT(512, 4),
        B200RS_SCAN_VARIANT(256, 4),  // the round-1 shape, 4096 elements
        B200RS_SCAN_VARIANT(1024, 4),
        B200RS_SCAN_VARIANT(256, 16),
        B200RS_SCAN_PERSISTENT(256, 8),   // 6
        B200RS_SCAN_PERSISTENT(512, 8),   // 7
        B200RS_SCAN_PERSISTENT(256, 4),   // 8
        B200RS_SCAN_PERSISTENT(512, 4),   // 9
        B200RS_SCAN_PERSISTENT(1024, 4),  // 10
        B200RS_SCAN_PERSISTENT(256, 16),  // 11
    };
    const char* e = getenv("B200RS_SCAN_VARIANT");
    int idx = e ? atoi(e) : 0;
    if (idx < 0 || idx >= (int)(sizeof(v) / sizeof(v[0]))) idx = 0;
    return v[idx];
}

}  // namespace

extern "C" int b200rs_exclusive_scan_u32(b200rs_device* dev, uint32_t* dst, const uint32_t* src, uint64_t n, uint32_t* total_out,
                                         void* temp, size_t* temp_bytes) {
    if (!dev || !temp_bytes) return B200RS_ERR_INVALID_ARGUMENT;
    const uint64_t max_tiles = (n + SCAN_MIN_TILE - 1) / SCAN_MIN_TILE;
    if (max_tiles > 0x7fffffffull) return B200RS_ERR_TOO_LARGE;
    const size_t need = SCAN_HEADER_BYTES + b200rs_align_up((size_t)max_tiles * sizeof(uint64_t), 256);
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
    const ScanVariant& var = pick_scan_variant();
    const uint32_t num_tiles = (uint32_t)((n + var.tile - 1) / var.tile);
    // ticket + the descriptors this tiling uses
    B200RS_CUDA(cudaMemsetAsync(temp, 0, SCAN_HEADER_BYTES + (size_t)num_tiles * sizeof(uint64_t), dev->stream));
    uint32_t* ticket = reinterpret_cast<uint32_t*>(temp);
    uint64_t* desc = reinterpret_cast<uint64_t*>(static_cast<char*>(temp) + SCAN_HEADER_BYTES);
    const bool vec16 = (((uintptr_t)dst | (uintptr_t)src) & 15u) == 0;
    {
        b200rs_launch_scope scope(dev, "scan_lookback", n, n * 8ull);
        auto kernel = vec16 ? var.aligned : var.unaligned;
        uint32_t grid = num_tiles;
        if (var.persistent) {
            int per_sm = 0;
            B200RS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, var.threads, 0));
            const uint64_t resident = (uint64_t)(per_sm > 0 ? per_sm : 1) * dev->num_sms;  // every CTA must be resident: tiles spin on lower tickets
            if (grid > resident) grid = (uint32_t)resident;
        }
        kernel<<<grid, var.threads, 0, dev->stream>>>(src, dst, n, desc, ticket, total_out, num_tiles);
    }
    B200RS_CUDA(cudaGetLastError());
    return B200RS_OK;
}
