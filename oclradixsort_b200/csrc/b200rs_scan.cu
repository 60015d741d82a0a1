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

// -------------------------------------------------------------------------------------------------
// TMA ring variant.  Persistent CTAs (the grid is exactly what is co-resident), tile k of CTA c is tile c + k*grid.
// One thread keeps STAGES-1 tiles in flight per CTA with cp.async.bulk (global -> shared, completion on an mbarrier
// per stage), so the HBM read stream never depends on what the CTA's warps are doing (scan, look-back, stores); the
// registers only ever hold the tile being scanned.  No ticket: a tile waits only on lower tiles, those belong to
// co-resident CTAs that walk their own tiles in increasing order, so the lowest unfinished tile can always finish.
// dst == src stays legal: a prefetched tile is only ever written by this same CTA, after it has been read.
// Used for 16-byte aligned src/dst; the ragged last tile is read with plain loads.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t scan_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void scan_mbar_init(uint32_t mbar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void scan_bulk_load(uint32_t sdst, const void* gsrc, uint32_t bytes, uint32_t mbar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sdst), "l"(gsrc), "r"(bytes), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void scan_mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "B200RS_SCAN_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra B200RS_SCAN_WAIT_%=;\n\t}" ::"r"(mbar), "r"(parity)
        : "memory");
}

template <int SCAN_THREADS, int SCAN_ROWS, int STAGES>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_ring_kernel(const uint32_t* src, uint32_t* dst, uint64_t n, uint64_t* desc, uint32_t* /*ticket, unused*/, uint32_t* total_out,
                 uint32_t num_tiles) {
    constexpr int SCAN_WARPS = SCAN_THREADS / 32;
    constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ROWS * 4;
    constexpr int SCAN_WARP_SLICE = 32 * SCAN_ROWS * 4;
    constexpr uint32_t TILE_BYTES = SCAN_TILE * 4u;
    static_assert(SCAN_WARPS <= 32 && STAGES >= 2, "warp-wide combine of the warp totals; the next tile must be resident while this one is scanned");
    extern __shared__ __align__(128) unsigned char ring_raw[];  // [STAGES][TILE_BYTES]
    __shared__ __align__(8) uint64_t s_full[STAGES];
    __shared__ uint32_t s_warp_total[2][SCAN_WARPS];  // double-buffered across iterations (tile 0 has no barrier in its look-back)
    __shared__ uint32_t s_ahead[2][SCAN_WARPS];       // per-warp sums of the NEXT tile
    __shared__ uint32_t s_lb_sum[SCAN_WARPS];
    __shared__ uint32_t s_lb_inc[SCAN_WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t in_slice = (uint32_t)warp * SCAN_WARP_SLICE + (uint32_t)lane * 4;  // element offset of row 0 inside a tile
    const uint32_t ring = scan_smem_addr(ring_raw);
    const uint32_t full0 = scan_smem_addr(&s_full[0]);
    const uint64_t full_tiles = n / SCAN_TILE;  // tiles below this index are complete: bulk-loadable

    auto tile_of = [&](uint32_t k) { return (uint64_t)blockIdx.x + (uint64_t)k * gridDim.x; };
    auto issue = [&](uint32_t k) {  // thread 0 only: start the load of this CTA's k-th tile
        const uint64_t t = tile_of(k);
        if (t < full_tiles) scan_bulk_load(ring + (k % STAGES) * TILE_BYTES, src + t * SCAN_TILE, TILE_BYTES, full0 + 8u * (k % STAGES));
    };
    // this thread's part of the CTA's k-th tile: from the ring (waiting for the bulk copy if `wait`) or, ragged last tile, from global
    auto fetch = [&](uint32_t k, bool wait, uint4 (&v)[SCAN_ROWS]) {
        const uint64_t t = tile_of(k);
        if (t < full_tiles) {
            if (wait) scan_mbar_wait(full0 + 8u * (k % STAGES), (k / STAGES) & 1u);
            const uint4* sp = reinterpret_cast<const uint4*>(ring_raw + (size_t)(k % STAGES) * TILE_BYTES) + (in_slice >> 2);
#pragma unroll
            for (int r = 0; r < SCAN_ROWS; ++r) v[r] = sp[r * 32];
        } else {
            const uint64_t base = t * SCAN_TILE + in_slice;
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
    // REDUCE-AHEAD: the aggregate of tile k+1 is published while tile k is being scanned, i.e. one whole iteration before
    // anybody needs it.  Without it every tile of a wave waits for the slowest CTA of the wave (the tiles of a wave are
    // consecutive), all resident CTAs stall in the same phase and nothing hides the latency (measured: 34 % of roofline).
    auto reduce_ahead = [&](uint32_t k) {
        if (tile_of(k) >= num_tiles) return;
        uint4 v[SCAN_ROWS];
        fetch(k, true, v);
        uint32_t sum = 0;
#pragma unroll
        for (int r = 0; r < SCAN_ROWS; ++r) sum += (v[r].x + v[r].y) + (v[r].z + v[r].w);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        if (lane == 0) s_ahead[k & 1][warp] = sum;
    };
    auto publish_ahead = [&](uint32_t k) {  // warp 1 (or warp 0 in a one-warp CTA), after the barrier that follows reduce_ahead(k)
        const uint64_t t = tile_of(k);
        if (t >= num_tiles || t == 0) return;  // tile 0 publishes its inclusive prefix directly
        uint32_t sum = lane < SCAN_WARPS ? s_ahead[k & 1][lane] : 0u;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        if (lane == 0) st_relaxed_u64(&desc[t], DESC_AGGREGATE | sum);
    };
    constexpr int PUBLISH_WARP = SCAN_WARPS > 1 ? 1 : 0;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < STAGES; ++i) scan_mbar_init(full0 + 8u * i);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
        for (int i = 0; i < STAGES; ++i) issue(i);
    }
    __syncthreads();
    reduce_ahead(0);
    __syncthreads();
    if (warp == PUBLISH_WARP) publish_ahead(0);

    uint32_t k = 0;
    for (uint64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++k) {
        const bool full = tile < full_tiles;
        // first look-back window, issued before the scan so that its L2 round trip overlaps it: every aggregate it can meet
        // was published an iteration ago (reduce-ahead), and this CTA's own previous tile is an inclusive prefix
        const int64_t look0 = (int64_t)tile - 1 - tid;
        uint64_t w = (tile != 0 && look0 >= 0) ? ld_relaxed_u64(&desc[look0]) : DESC_INCLUSIVE;  // tiles < 0 act as {inclusive, 0}
        uint4 v[SCAN_ROWS];
        fetch(k, false, v);  // landed: reduce_ahead(k) waited for it
        // ---- warp-level exclusive scan in (row, lane, component) order ----
        uint32_t excl[SCAN_ROWS];
        uint32_t carry = 0;
#pragma unroll
        for (int r = 0; r < SCAN_ROWS; ++r) {
            const uint32_t sum = v[r].x + v[r].y + v[r].z + v[r].w;
            const uint32_t inc = warp_inclusive_sum(sum, lane);
            excl[r] = carry + inc - sum;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 31) s_warp_total[k & 1][warp] = carry;
        reduce_ahead(k + 1);
        __syncthreads();  // every thread holds its part of stage k in registers: the stage can be refilled
        if (tid == 0) issue(k + STAGES);
        if (warp == PUBLISH_WARP) publish_ahead(k + 1);

        // ---- every warp: exclusive scan of the warp totals (redundantly; no extra barrier) ----
        const uint32_t wt = lane < SCAN_WARPS ? s_warp_total[k & 1][lane] : 0u;
        const uint32_t winc = warp_inclusive_sum(wt, lane);
        const uint32_t aggregate = __shfl_sync(0xffffffffu, winc, 31);
        const uint32_t warp_excl = __shfl_sync(0xffffffffu, winc - wt, warp);

        // ---- block-wide look-back: thread i inspects tile (look - i), SCAN_THREADS descriptors per step.  With static tile
        //      assignment the nearest inclusive prefix is up to one wave (the grid) back -- this CTA's own previous tile at
        //      the latest -- and one step of the whole CTA reaches it in a single L2 round trip ----
        uint32_t prefix = 0;
        bool done = tile == 0;
        int64_t look = (int64_t)tile - 1;
        bool first_window = true;
        while (!done) {
            const int64_t idx = look - tid;
            if (!first_window) w = idx >= 0 ? ld_relaxed_u64(&desc[idx]) : DESC_INCLUSIVE;
            first_window = false;
            while (__any_sync(0xffffffffu, (w >> 32) == 0)) {
                if ((w >> 32) == 0) w = ld_relaxed_u64(&desc[idx]);
            }
            const uint32_t inc_lanes = __ballot_sync(0xffffffffu, (w >> 32) == 2);
            const int stop = inc_lanes ? (__ffs(inc_lanes) - 1) : 31;  // nearest tile with a full prefix inside this warp's window
            uint32_t contrib = lane <= stop ? (uint32_t)w : 0u;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
            if (lane == 0) {
                s_lb_sum[warp] = contrib;
                s_lb_inc[warp] = inc_lanes != 0;
            }
            __syncthreads();
            // combine the warps' windows nearest-first, up to and including the first one that met an inclusive prefix
            const uint32_t ws = lane < SCAN_WARPS ? s_lb_sum[lane] : 0u;
            const uint32_t wi = __ballot_sync(0xffffffffu, lane < SCAN_WARPS && s_lb_inc[lane]);
            const int wstop = wi ? (__ffs(wi) - 1) : 31;
            uint32_t c2 = lane <= wstop ? ws : 0u;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) c2 += __shfl_xor_sync(0xffffffffu, c2, d);
            prefix += c2;
            done = wi != 0;
            look -= SCAN_THREADS;
            if (!done) __syncthreads();  // s_lb_* are rewritten by the next step
        }
        if (tid == 0) {
            st_relaxed_u64(&desc[tile], DESC_INCLUSIVE | (uint32_t)(prefix + aggregate));
            if (total_out && tile == num_tiles - 1) *total_out = prefix + aggregate;
        }

        // ---- store ----
        const uint32_t base_sum = prefix + warp_excl;
        const uint64_t base = tile * SCAN_TILE + in_slice;
        if (full) {
#pragma unroll
            for (int r = 0; r < SCAN_ROWS; ++r) {
                uint4 q;
                uint32_t run = base_sum + excl[r];
                q.x = run; run += v[r].x;
                q.y = run; run += v[r].y;
                q.z = run; run += v[r].z;
                q.w = run;
                *reinterpret_cast<uint4*>(dst + base + r * 128) = q;
            }
        } else {
#pragma unroll
            for (int r = 0; r < SCAN_ROWS; ++r) {
                const uint64_t i = base + r * 128;
                uint32_t run = base_sum + excl[r];
                if (i + 0 < n) dst[i + 0] = run; run += v[r].x;
                if (i + 1 < n) dst[i + 1] = run; run += v[r].y;
                if (i + 2 < n) dst[i + 2] = run; run += v[r].z;
                if (i + 3 < n) dst[i + 3] = run;
            }
        }
    }
}

constexpr size_t SCAN_HEADER_BYTES = 256;  // ticket word, padded so descriptors stay 256 B aligned

struct ScanVariant {
    void (*aligned)(const uint32_t*, uint32_t*, uint64_t, uint64_t*, uint32_t*, uint32_t*, uint32_t);
    void (*unaligned)(const uint32_t*, uint32_t*, uint64_t, uint64_t*, uint32_t*, uint32_t*, uint32_t);
    int threads, tile;
    bool persistent;
    size_t ring_bytes;  // dynamic shared memory of the aligned kernel (TMA ring variants)
};
#define B200RS_SCAN_VARIANT(T, R) ScanVariant{scan_lookback_kernel<T, R, true>, scan_lookback_kernel<T, R, false>, T, T * R * 4, false, 0}
#define B200RS_SCAN_PERSISTENT(T, R) ScanVariant{scan_persistent_kernel<T, R, true>, scan_persistent_kernel<T, R, false>, T, T * R * 4, true, 0}
#define B200RS_SCAN_RING(T, R, S) ScanVariant{scan_ring_kernel<T, R, S>, scan_lookback_kernel<T, R, false>, T, T * R * 4, true, (size_t)S * T * R * 16}
const ScanVariant& pick_scan_variant(uint64_t n, int num_sms) {
    // production: the TMA ring kernel with 96 KiB tiles (one persistent CTA per SM); inputs too small to give every SM a
    // couple of those use 32 KiB tiles (three CTAs per SM).  The other shapes are compiled only with -DB200RS_EXPERIMENTS.
    static const ScanVariant big = B200RS_SCAN_RING(512, 12, 2), small = B200RS_SCAN_RING(256, 8, 2);
    const ScanVariant& production = n >= (uint64_t)num_sms * 2 * 24576 ? big : small;
#ifdef B200RS_EXPERIMENTS
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
        B200RS_SCAN_RING(512, 4, 3),      // 12: 32 KiB tiles, 96 KiB ring
        B200RS_SCAN_RING(256, 4, 4),      // 13: 16 KiB tiles, 64 KiB ring
        B200RS_SCAN_RING(256, 8, 3),      // 14
        B200RS_SCAN_RING(512, 4, 2),      // 15
        B200RS_SCAN_RING(512, 2, 4),      // 16
        B200RS_SCAN_RING(1024, 4, 3),     // 17: 64 KiB tiles, 192 KiB ring
        B200RS_SCAN_RING(256, 4, 3),      // 18
        B200RS_SCAN_RING(512, 8, 3),      // 19: 64 KiB tiles, 192 KiB ring
        B200RS_SCAN_RING(256, 8, 2),      // 20
        B200RS_SCAN_RING(256, 16, 2),     // 21
        B200RS_SCAN_RING(512, 8, 2),      // 22
        B200RS_SCAN_RING(1024, 4, 2),     // 23
        B200RS_SCAN_RING(512, 12, 2),     // 24: 96 KiB tiles
        B200RS_SCAN_RING(1024, 6, 2),     // 25: 96 KiB tiles
        B200RS_SCAN_RING(768, 8, 2),      // 26: 96 KiB tiles
    };
    if (const char* e = getenv("B200RS_SCAN_VARIANT")) {
        const int idx = atoi(e);
        if (idx >= 0 && idx < (int)(sizeof(v) / sizeof(v[0]))) return v[idx];
    }
#endif
    return production;
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
    const ScanVariant& var = pick_scan_variant(n, dev->num_sms);
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
        const size_t dyn_smem = vec16 ? var.ring_bytes : 0;
        int per_sm = 0;
        B200RS_TRY(b200rs_kernel_setup(dev, (const void*)kernel, dyn_smem, var.threads, &per_sm));
        if (var.persistent && (vec16 || var.ring_bytes == 0)) {
            const uint64_t resident = (uint64_t)(per_sm > 0 ? per_sm : 1) * dev->num_sms;  // every CTA must be resident: tiles spin on lower ones
            if (grid > resident) grid = (uint32_t)resident;
        }
        kernel<<<grid, var.threads, dyn_smem, dev->stream>>>(src, dst, n, desc, ticket, total_out, num_tiles);
    }
    B200RS_CUDA(cudaGetLastError());
    return B200RS_OK;
}
