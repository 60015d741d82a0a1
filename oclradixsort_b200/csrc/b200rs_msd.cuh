// b200rs_msd.cuh -- 32-bit KEY-ONLY sort as a most-significant-digit-first pipeline (included by b200rs_sort.cu, inside its
// anonymous namespace; also by tools/msd_dev.cu, the stand-alone measurement harness).
//
// Why a second path.  Pprims::radixSort(Buffer<u32>) (Pprims.cpp:304-406) has no payload, so equal keys are
// indistinguishable and NOTHING has to be stable when all 32 bits are sorted -- the result is the same bits whatever order
// equal digits are moved in.  The LSD scatter pass (b200rs_onesweep2.cuh) pays for stability three times per key and pass:
// a counting RED, a ballot search + claiming ATOMS, and a scattering STS, every one of them a random shared-memory access
// that costs ~3.3 L1 wavefronts per warp instruction; it is bound by that pipe at 50 % of the HBM roofline,
// and four such passes cannot reach 70 %.  Here:
//
//   H   msd_hist16_kernel      one read: the JOINT histogram of the top 16 bits (65 536 bins, 16-bit counters in 128 KiB of
//                              shared memory, overflow detected by a checksum) -> every bucket boundary of both partition
//                              passes and of the final step.                                               4 B/key
//   PL  msd_plan_kernel        exclusive scan of the 65 536 bins, bucket cursors, tile table, eligibility.
//   P1  msd_partition_kernel   partition by bits 24..31 (256 buckets), UNSTABLE: one ATOMS with return per key gives its slot
//                              in a fixed-capacity bin of the tile (two random accesses per key instead of three, no ballots),
//                              a tile's run of every digit is placed by ONE global atomicAdd on the bucket cursor -- no
//                              look-back chain, no ticket, tiles finish in any order.                            8 B/key
//   P2  msd_partition_kernel   the same kernel, segmented: every tile lies inside one bucket of P1 and partitions it by bits
//                              16..23 into that bucket's 256 sub-buckets.                                        8 B/key
//   F   msd_bucket_kernel      every top-16 bucket (n / 65 536 keys, all in shared memory) is sorted on its low 16 bits by
//                              COUNTING: 4-bit counters for the 65 536 possible values, one ATOMS per key (returns the key's
//                              rank among equal keys), a scan of the counters, one lookup per key.  In place.     8 B/key
//
// 28 B/key of HBM traffic instead of 36, and 2 + 2 + 2 random shared-memory accesses per key instead of 12.
// Used when every top-16 bucket fits F's shared memory (uniform-like, presorted, reversed ... inputs); otherwise -- heavy
// duplicates in the top bits -- the LSD path runs (b200rs_sort.cu decides after PL; see sort_keys_msd there).
#pragma once

constexpr int MSD_BUCKET_BITS = 16;
constexpr int MSD_BUCKETS = 1 << MSD_BUCKET_BITS;
constexpr int MSD_LOW_BITS = 32 - MSD_BUCKET_BITS;

// control words (device, zeroed per sort)
enum MsdCtl {
    MSD_CTL_INELIGIBLE = 0,  // != 0: a counter of H overflowed, or a bucket exceeds F's capacity -> the caller runs the LSD path
    MSD_CTL_MAX_BUCKET = 1,  // largest top-16 bucket
    MSD_CTL_P2_TILES = 2,    // tiles of the second partition pass
    MSD_CTL_WORDS = 8,
};

// =================================================================================================
// H: joint histogram of the top 16 bits
// =================================================================================================
// One 1024-thread CTA per SM.  Table: 32 768 words, word (bin >> 1) holds bin's count in its half (bin & 1), so a pair of
// neighbouring bins is flushed with one 64-bit global reduction.  A half is 16 bits: it can wrap, and a wrap of the low half
// carries into the high half.  Both are caught by a checksum -- as an integer the word is exactly c_lo + 65536 c_hi, so the
// sum of all halves read back equals the number of keys the CTA counted unless some half wrapped (each wrap lowers it by
// 65 535 or 65 536: wraps cannot cancel) -- and a wrap means a bucket beyond anything F can take, so the input is simply
// marked ineligible.  A warp whose 128 keys (one LDG.128 each) share one bin -- presorted input -- adds them with one RED.
constexpr int MSD_HIST_THREADS = 1024;
constexpr int MSD_HIST_VECS = 4;  // LDG.128 in flight per thread
constexpr size_t MSD_HIST_SMEM = (size_t)(MSD_BUCKETS / 2) * sizeof(uint32_t);

__global__ void __launch_bounds__(MSD_HIST_THREADS, 1)
msd_hist16_kernel(const uint32_t* __restrict__ in, uint64_t n, unsigned long long* __restrict__ joint2 /*[32768] pairs of u32 counters*/,
                  uint32_t* __restrict__ ctl) {
    extern __shared__ __align__(16) uint32_t msd_tab[];  // [32768]
    __shared__ uint32_t s_sum[MSD_HIST_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < MSD_BUCKETS / 8; i += MSD_HIST_THREADS) reinterpret_cast<uint4*>(msd_tab)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const uint32_t table = smem_addr(msd_tab);
    uint32_t counted = 0;  // keys this thread added to the table
    auto count1 = [&](uint32_t key, uint32_t mult) {
        const uint32_t addr = table + ((key >> 15) & 0x1fffcu);           // word (key >> 17)
        const uint32_t inc = ((key >> 16) & 1u) * (0xffffu * mult) + mult;  // mult << (16 * (bin & 1))
        red_add_shared(addr, inc);
    };
    const uint64_t nvec = n / 4;
    const uint4* in4 = reinterpret_cast<const uint4*>(in);
    const uint64_t round_vecs = (uint64_t)MSD_HIST_THREADS * MSD_HIST_VECS;
    for (uint64_t base = (uint64_t)blockIdx.x * round_vecs; base < nvec; base += (uint64_t)gridDim.x * round_vecs) {
        uint4 q[MSD_HIST_VECS];
        bool have[MSD_HIST_VECS];
#pragma unroll
        for (int u = 0; u < MSD_HIST_VECS; ++u) {
            const uint64_t v = base + (uint64_t)u * MSD_HIST_THREADS + tid;
            have[u] = v < nvec;
            if (have[u]) q[u] = __ldg(in4 + v);
        }
#pragma unroll
        for (int u = 0; u < MSD_HIST_VECS; ++u) {
            // do all 128 keys of this warp instruction fall into one bin?  (have[u] is warp-uniform except in the last round)
            const uint32_t x0 = __shfl_sync(0xffffffffu, q[u].x, 0);
            const uint32_t diff = ((q[u].x ^ q[u].y) | (q[u].x ^ q[u].z)) | ((q[u].x ^ q[u].w) | (q[u].x ^ x0));
            const bool uniform = __all_sync(0xffffffffu, have[u] && (diff >> 16) == 0);
            if (uniform) {
                if (lane == 0) count1(x0, 128u);
                counted += 4;
            } else if (have[u]) {
                count1(q[u].x, 1u); count1(q[u].y, 1u); count1(q[u].z, 1u); count1(q[u].w, 1u);
                counted += 4;
            }
        }
    }
    if (blockIdx.x == 0) {  // ragged tail (n not a multiple of 4)
        const uint64_t i = nvec * 4 + tid;
        if (i < n) { count1(in[i], 1u); counted += 1; }
    }
    __syncthreads();
    // flush + checksum.  Thread t owns words t, t + 1024, ... (conflict-free reads, coalesced reductions)
    uint32_t seen = 0;
    for (int w = tid; w < MSD_BUCKETS / 2; w += MSD_HIST_THREADS) {
        const uint32_t x = msd_tab[w];
        const uint32_t lo = x & 0xffffu, hi = x >> 16;
        seen += lo + hi;
        if (x) atomicAdd(&joint2[w], (unsigned long long)lo | ((unsigned long long)hi << 32));
    }
    // CTA-wide: sum(seen) == sum(counted) ?  (both fit 32 bits: a CTA sees fewer than 2^32 keys)
    uint32_t d = seen - counted;  // per-thread differences are arbitrary; only their sum must vanish
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0) s_sum[tid >> 5] = d;
    __syncthreads();
    if (tid == 0) {
        uint32_t t = 0;
        for (int w = 0; w < MSD_HIST_THREADS / 32; ++w) t += s_sum[w];
        if (t != 0) atomicOr(&ctl[MSD_CTL_INELIGIBLE], 1u);
    }
}

// =================================================================================================
// PL: bucket offsets, cursors, tile table of the segmented pass, eligibility (one CTA)
// =================================================================================================
struct MsdTile {
    uint32_t start;  // first key of the tile in the intermediate buffer (a multiple of 4)
    uint32_t meta;   // (bucket of the first partition pass << 16) | keys in the tile
};
constexpr int MSD_PLAN_THREADS = 1024;
constexpr int MSD_PLAN_BINS = MSD_BUCKETS / MSD_PLAN_THREADS;  // 64 consecutive bins per thread

// Layout of the intermediate buffer (output of P1, input of P2): bucket b of the first pass starts at
// (final start of b rounded down to a multiple of 4) + 4 b, so every tile of P2 can be read with aligned 128-bit loads;
// buckets cannot overlap (the rounding loses at most 3, the 4 b term adds 4 per bucket); the buffer needs n + 1024 keys.
__device__ __forceinline__ uint32_t msd_mid_start(uint32_t final_start, uint32_t bucket) { return (final_start & ~3u) + 4u * bucket; }

__global__ void __launch_bounds__(MSD_PLAN_THREADS, 1)
msd_plan_kernel(const uint32_t* __restrict__ joint /*[65536]*/, uint32_t n, uint32_t tile_keys /* P2 tile */, uint32_t bucket_cap /* F capacity */,
                uint32_t* __restrict__ bucket_off /*[65537]: final start of every top-16 bucket*/,
                uint32_t* __restrict__ cursor2 /*[65536]: = bucket_off, consumed by P2*/,
                uint32_t* __restrict__ cursor1 /*[256]: start of every P1 bucket in the intermediate buffer, consumed by P1*/,
                MsdTile* __restrict__ tiles, uint32_t* __restrict__ ctl) {
    __shared__ uint32_t s_warp[MSD_PLAN_THREADS / 32];
    __shared__ uint32_t s_max[MSD_PLAN_THREADS / 32];
    __shared__ uint32_t s_start3[257];      // final start of every byte-3 bucket
    __shared__ uint32_t s_tile_first[257];  // first tile of every byte-3 bucket
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // 64 consecutive bins per thread, read as 16 x uint4; lanes are 256 B apart (the table is 256 KiB, read once, from L2)
    const uint4* j4 = reinterpret_cast<const uint4*>(joint) + (size_t)tid * (MSD_PLAN_BINS / 4);
    uint32_t sum = 0, mx = 0;
#pragma unroll 4
    for (int i = 0; i < MSD_PLAN_BINS / 4; ++i) {
        const uint4 v = __ldcg(j4 + i);
        sum += (v.x + v.y) + (v.z + v.w);
        mx = max(max(mx, max(v.x, v.y)), max(v.z, v.w));
    }
    uint32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 31) s_warp[warp] = inc;
    if (lane == 0) s_max[warp] = mx;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; ++w) base += s_warp[w];
    uint32_t run = base + inc - sum;  // exclusive prefix of this thread's first bin
    if ((tid & 3) == 0) s_start3[tid >> 2] = run;  // bins 256 b .. 256 b + 255 belong to byte-3 bucket b = tid / 4
    if (tid == 0) s_start3[256] = n;
    uint4* o4 = reinterpret_cast<uint4*>(bucket_off) + (size_t)tid * (MSD_PLAN_BINS / 4);
    uint4* c4 = reinterpret_cast<uint4*>(cursor2) + (size_t)tid * (MSD_PLAN_BINS / 4);
#pragma unroll 4
    for (int i = 0; i < MSD_PLAN_BINS / 4; ++i) {
        const uint4 v = __ldcg(j4 + i);
        uint4 o;
        o.x = run; run += v.x;
        o.y = run; run += v.y;
        o.z = run; run += v.z;
        o.w = run; run += v.w;
        o4[i] = o;
        c4[i] = o;
    }
    if (tid == 0) bucket_off[MSD_BUCKETS] = n;
    __syncthreads();
    // byte-3 buckets: cursors of the first pass, tiles of the second
    if (tid < 256) {
        const uint32_t start = s_start3[tid], size = s_start3[tid + 1] - start;
        cursor1[tid] = msd_mid_start(start, (uint32_t)tid);
        const uint32_t t = (size + tile_keys - 1) / tile_keys;
        uint32_t tinc = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, tinc, d);
            if (lane >= d) tinc += y;
        }
        if (lane == 31) s_warp[warp] = tinc;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        uint32_t tb = 0;
        for (int w = 0; w < warp; ++w) tb += s_warp[w];
        s_tile_first[tid] = tb + tinc - t;
        if (tid == 255) s_tile_first[256] = tb + tinc;
    }
    __syncthreads();
    const uint32_t total_tiles = s_tile_first[256];
    for (uint32_t t = tid; t < total_tiles; t += MSD_PLAN_THREADS) {
        // bucket of tile t: the last b with tile_first[b] <= t (binary search over 256 entries; empty buckets are skipped by construction)
        uint32_t lo = 0, hi = 256;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_tile_first[mid] <= t) lo = mid; else hi = mid;
        }
        const uint32_t b = lo, k = t - s_tile_first[b];
        const uint32_t start = s_start3[b], size = s_start3[b + 1] - start;
        MsdTile e;
        e.start = msd_mid_start(start, b) + k * tile_keys;
        e.meta = (b << 16) | min(tile_keys, size - k * tile_keys);
        tiles[t] = e;
    }
    if (tid == 0) {
        uint32_t m = 0;
        for (int w = 0; w < MSD_PLAN_THREADS / 32; ++w) m = max(m, s_max[w]);
        ctl[MSD_CTL_MAX_BUCKET] = m;
        ctl[MSD_CTL_P2_TILES] = total_tiles;
        if (m > bucket_cap) atomicOr(&ctl[MSD_CTL_INELIGIBLE], 2u);
    }
}

// =================================================================================================
// P1 / P2: unstable partition of a tile by one 8-bit digit
// =================================================================================================
// A tile is THREADS x VPT aligned 128-bit loads.  Every key claims a slot in its digit's bin of the tile with ONE shared
// atomic with return (rank = old count); bins have a fixed capacity CAP (about twice the mean), so the slot address needs no
// table: bins[digit * CAP + rank].  After the barrier one thread per digit reserves the digit's run in the output with one
// global atomicAdd on the bucket's cursor (any order of tiles is fine: nothing is stable), and each warp copies whole bins,
// lanes aligned to the 128-byte lines of the destination.  A tile in which some bin overflows -- skewed or presorted input --
// takes the dense route instead: counts are known by then, a second claim on counters pre-seeded with the bins' dense
// starts yields the final slot, and the tile is written like an LSD tile (flat loop, per-digit base).  Warps whose 32 keys
// of one instruction share the digit (presorted input) claim with one atomic.
template <int THREADS, int VPT, int CAP>
struct MsdPartitionConfig {
    static constexpr int TILE = THREADS * VPT * 4;
    static constexpr int WARPS = THREADS / 32;
    static constexpr int SLOTS = RADIX * CAP > TILE ? RADIX * CAP : TILE;
    struct Smem {
        alignas(16) uint32_t bins[SLOTS];
        uint32_t cnt[RADIX];     // keys of the tile per digit
        uint32_t cnt2[RADIX];    // dense route: running dense slot per digit
        uint2 info[RADIX];       // {count, global start of the run} (dense route: {dense start, global start - dense start})
        uint32_t scan_scratch[RADIX / 32];
        uint32_t overflow;
    };
};

template <int THREADS, int VPT, int CAP, int MIN_CTAS, bool SEGMENTED>
__global__ void __launch_bounds__(THREADS, MIN_CTAS)
msd_partition_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n, int shift,
                     uint32_t* __restrict__ cursor /* P1: [256]; P2: [65536], bucket b's 256 cursors at 256 b */,
                     const MsdTile* __restrict__ tiles, const uint32_t* __restrict__ ctl) {
    using Cfg = MsdPartitionConfig<THREADS, VPT, CAP>;
    static_assert(THREADS >= RADIX, "one thread per digit");
    extern __shared__ __align__(16) unsigned char msd_smem_raw[];
    typename Cfg::Smem& s = *reinterpret_cast<typename Cfg::Smem*>(msd_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    uint32_t start, count;
    if (SEGMENTED) {
        if (blockIdx.x >= ctl[MSD_CTL_P2_TILES]) return;  // the grid was sized for the upper bound n / TILE + 256
        const MsdTile t = tiles[blockIdx.x];
        start = t.start;
        count = t.meta & 0xffffu;
        cursor += (size_t)(t.meta >> 16) * RADIX;
    } else {
        start = blockIdx.x * (uint32_t)Cfg::TILE;
        count = min((uint32_t)Cfg::TILE, n - start);
    }
    for (int i = tid; i < RADIX; i += THREADS) s.cnt[i] = 0;
    if (tid == 0) s.overflow = 0;

    // ---- load: key (v, tid, c) of the tile is in[start + (v * THREADS + tid) * 4 + c] ----
    uint32_t key[VPT][4];
    const bool full = count == (uint32_t)Cfg::TILE;
    if (full) {
        const uint4* src = reinterpret_cast<const uint4*>(in + start) + tid;
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
            const uint4 q = __ldg(src + v * THREADS);
            key[v][0] = q.x; key[v][1] = q.y; key[v][2] = q.z; key[v][3] = q.w;
        }
    } else {
#pragma unroll
        for (int v = 0; v < VPT; ++v)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t j = (uint32_t)(v * THREADS + tid) * 4 + c;
                key[v][c] = j < count ? in[start + j] : 0u;
            }
    }
    __syncthreads();

    // ---- claim: rank inside the digit's bin; keys below the capacity go straight to their slot ----
    const uint32_t cnt_base = smem_addr(&s.cnt[0]), bins_base = smem_addr(&s.bins[0]);
    const uint32_t prmt_sel = 0x4440u | (uint32_t)(shift >> 3);
    auto claim_all = [&](uint32_t counters, bool sparse) {
#pragma unroll
        for (int v = 0; v < VPT; ++v)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const bool live = full || (uint32_t)(v * THREADS + tid) * 4 + c < count;
                const uint32_t d = __byte_perm(key[v][c], 0u, prmt_sel);
                const uint32_t d0 = __shfl_sync(0xffffffffu, d, 0);
                uint32_t r;
                if (__all_sync(0xffffffffu, live && d == d0)) {
                    // the warp's 32 keys share the digit: one atomic for all of them
                    uint32_t b = 0;
                    if (lane == 0) b = atom_add_shared(counters + 4u * d0, 32u);
                    r = __shfl_sync(0xffffffffu, b, 0) + (uint32_t)lane;
                } else {
                    r = live ? atom_add_shared(counters + 4u * d, 1u) : 0xffffffffu;
                }
                if (sparse) {
                    if (r < (uint32_t)CAP) st_shared(bins_base + 4u * (d * (uint32_t)CAP + r), key[v][c]);
                } else {
                    if (live) st_shared(bins_base + 4u * r, key[v][c]);
                }
            }
    };
    claim_all(cnt_base, true);
    __syncthreads();

    // ---- one thread per digit: the tile's run of that digit is reserved in the output ----
    uint32_t c = 0, g = 0;
    if (tid < RADIX) {
        c = s.cnt[tid];
        if (c) g = atomicAdd(&cursor[tid], c);
        if (c > (uint32_t)CAP) s.overflow = 1;
        s.info[tid] = make_uint2(c, g);
    }
    __syncthreads();
    if (!s.overflow) {
        // ---- sparse route: warp w copies bins w, w + WARPS, ... (two bins in flight); lane l of step k handles the key that
        //      lands on word l of the k-th 128-byte line touched by the bin's run, so every store fills one aligned line ----
        const uint32_t out_word = (uint32_t)((uintptr_t)out >> 2);
        for (int d = warp; d < RADIX; d += 2 * Cfg::WARPS) {
            const int d1 = d + Cfg::WARPS;
            const uint2 c0 = s.info[d];
            const uint2 c1 = d1 < RADIX ? s.info[d1] : make_uint2(0u, 0u);
            const uint32_t a0 = (out_word + c0.y) & 31u, a1 = (out_word + c1.y) & 31u;  // position of the run's first key inside its line
            const uint32_t j0 = (uint32_t)lane - a0, j1 = (uint32_t)lane - a1;          // (wraps to a huge value where the line starts before the run)
            const uint32_t* bin0 = &s.bins[d * CAP];
            const uint32_t* bin1 = &s.bins[(d1 < RADIX ? d1 : d) * CAP];
            uint32_t* dst0 = out + c0.y;
            uint32_t* dst1 = out + c1.y;
            // the first two lines of each run: all of it unless count + a > 64
            const bool p00 = j0 < c0.x, p01 = j0 + 32u < c0.x, p10 = j1 < c1.x, p11 = j1 + 32u < c1.x;
            uint32_t v00 = 0, v01 = 0, v10 = 0, v11 = 0;
            if (p00) v00 = bin0[j0];
            if (p01) v01 = bin0[j0 + 32u];
            if (p10) v10 = bin1[j1];
            if (p11) v11 = bin1[j1 + 32u];
            if (p00) dst0[j0] = v00;
            if (p01) dst0[j0 + 32u] = v01;
            if (p10) dst1[j1] = v10;
            if (p11) dst1[j1 + 32u] = v11;
            for (uint32_t k32 = 64; k32 < c0.x + a0; k32 += 32) {  // warp-uniform trip counts
                const uint32_t j = k32 + j0;
                if (j < c0.x) dst0[j] = bin0[j];
            }
            for (uint32_t k32 = 64; k32 < c1.x + a1; k32 += 32) {
                const uint32_t j = k32 + j1;
                if (j < c1.x) dst1[j] = bin1[j];
            }
        }
        return;
    }
    // ---- dense route ----
    if (tid < RADIX) {
        const uint32_t dense = block_exclusive_scan_256<uint32_t>(c, s.scan_scratch, tid);
        s.cnt2[tid] = dense;
        s.info[tid] = make_uint2(dense, g - dense);
    }
    __syncthreads();
    claim_all(smem_addr(&s.cnt2[0]), false);
    __syncthreads();
    for (uint32_t j = tid; j < count; j += THREADS) {
        const uint32_t k = s.bins[j];
        out[s.info[__byte_perm(k, 0u, prmt_sel)].y + j] = k;
    }
}

// =================================================================================================
// F: every top-16 bucket sorted on its low 16 bits by counting, in shared memory, in place
// =================================================================================================
// Counters: 65 536 bins x 4 bits = 8 192 words (bin b: nibble b & 7 of word b >> 3).  Per key one ATOMS with return: the old
// nibble is the key's rank among equal keys.  Scan: thread t adds up the nibbles of its 32 consecutive words, a block scan
// makes that the exclusive prefix wp[] of every word.  Lookup: slot = wp[word] + (sum of the word's nibbles below mine) +
// rank; the nibble sum is one multiply ((x * 0x11111111) >> 28 adds the eight nibbles when the sum is below 16).
// Anything that would break the nibble arithmetic -- 16 or more equal keys, a word whose eight bins hold more than 15 keys
// -- is caught exactly (old nibble == 15; word total > 15; and, for nibble wraps, the checksum of all words against the
// bucket size) and sends the bucket to the robust route: a stable two-pass LSD sort in the same shared memory.
// The words are stored XOR-swizzled inside each thread's 128-byte scan range so the scan's LDS.128 are conflict-free.
template <int THREADS, int IPT>
struct MsdBucketConfig {
    static constexpr int CAP = THREADS * IPT;
    static constexpr int WORDS = (1 << MSD_LOW_BITS) / 8;
    static constexpr int WORDS_PER_THREAD = WORDS / THREADS;
    struct Smem {
        alignas(16) uint32_t cnt[WORDS];
        alignas(16) uint16_t wp[WORDS];
        alignas(16) uint32_t staged[CAP];
        uint32_t warp_total[THREADS / 32];
        uint32_t flags;
    };
    // the robust route needs two key buffers, the per-warp counters of the stable ranking and its scratch
    static constexpr size_t ROBUST_BYTES = 2 * (size_t)CAP * 4 + (size_t)(THREADS / 32) * RADIX * 4 + 64 * 4;
    static constexpr size_t SMEM_BYTES = sizeof(Smem) > ROBUST_BYTES ? sizeof(Smem) : ROBUST_BYTES;
};

// byte offset of counter word (key's low 16 bits >> 3), swizzled: chunk (16 B) index inside the 128-byte group ^= group & 7
__device__ __forceinline__ uint32_t msd_word_offset(uint32_t key) {
    const uint32_t plain = (key & 0xfff8u) >> 1;  // 4 * (bin >> 3)
    return plain ^ ((key >> 4) & 0x70u);          // group = bin >> 8
}

template <int THREADS>
__device__ void msd_bucket_robust(unsigned char* smem, uint32_t* __restrict__ data, uint32_t s, uint32_t minus_one);

template <int THREADS, int IPT, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS)
msd_bucket_kernel(uint32_t* __restrict__ data, const uint32_t* __restrict__ bucket_off, uint32_t minus_one) {
    using Cfg = MsdBucketConfig<THREADS, IPT>;
    static_assert(Cfg::WORDS_PER_THREAD == 32, "the scan gives every thread one 128-byte group of counter words");
    extern __shared__ __align__(16) unsigned char msd_smem_raw[];
    typename Cfg::Smem& s = *reinterpret_cast<typename Cfg::Smem*>(msd_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t start = bucket_off[blockIdx.x];
    const uint32_t size = bucket_off[blockIdx.x + 1] - start;
    if (size <= 1) return;
    uint32_t* __restrict__ keys = data + start;

    {   // zero the counters
        uint4* z = reinterpret_cast<uint4*>(s.cnt);
#pragma unroll
        for (int i = 0; i < Cfg::WORDS / 4 / THREADS; ++i) z[i * THREADS + tid] = make_uint4(0, 0, 0, 0);
        if (tid == 0) s.flags = 0;
    }
    uint32_t key[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i)
        if ((uint32_t)(i * THREADS + tid) < size) key[i] = keys[i * THREADS + tid];
    __syncthreads();

    // ---- count; the returned nibble is the rank among equal keys ----
    const uint32_t cnt_base = smem_addr(&s.cnt[0]);
    uint32_t ranks[(IPT + 7) / 8];
#pragma unroll
    for (int i = 0; i < (IPT + 7) / 8; ++i) ranks[i] = 0;
    uint32_t bad = 0;
#pragma unroll
    for (int i = 0; i < IPT; ++i)
        if ((uint32_t)(i * THREADS + tid) < size) {
            const uint32_t sh = (key[i] & 7u) << 2;
            const uint32_t old = atom_add_shared(cnt_base + msd_word_offset(key[i]), 1u << sh);
            const uint32_t r = (old >> sh) & 15u;
            bad |= (r == 15u) ? 1u : 0u;  // the 16th equal key: the nibble wrapped
            ranks[i >> 3] |= r << (4 * (i & 7));
        }
    __syncthreads();

    // ---- scan of the counter words: thread t owns words 32 t .. 32 t + 31 (chunk j of them sits at chunk j ^ (t & 7)) ----
    uint32_t total = 0;
    {
        const uint4* mine = reinterpret_cast<const uint4*>(s.cnt) + tid * 8;
        uint4* wp4 = reinterpret_cast<uint4*>(s.wp) + tid * 4;  // wp[] of my 32 words, two per register (not swizzled)
        uint32_t run = 0, wide = 0;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {  // eight words at a time: their prefixes (relative to my first word) fill one 16-byte store
            uint32_t pre[8];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint4 q = mine[(2 * jj + h) ^ (tid & 7)];
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    pre[h * 4 + k] = run;
                    const uint32_t before = run;
                    run = __dp4a(w[k] & 0x0f0f0f0fu, 0x01010101u, run);
                    run = __dp4a((w[k] >> 4) & 0x0f0f0f0fu, 0x01010101u, run);
                    wide |= (run - before) >> 4;  // a word whose eight bins hold 16 or more keys
                }
            }
            wp4[jj] = make_uint4(pre[0] | (pre[1] << 16), pre[2] | (pre[3] << 16), pre[4] | (pre[5] << 16), pre[6] | (pre[7] << 16));
        }
        total = run;
        bad |= wide ? 1u : 0u;
        // block-wide exclusive scan of the threads' totals
        uint32_t inc = total;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += y;
        }
        if (lane == 31) s.warp_total[warp] = inc;
        __syncthreads();
        uint32_t base = 0, grand = 0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) {
            const uint32_t x = s.warp_total[w];
            grand += x;
            if (w < warp) base += x;
        }
        bad |= grand != size ? 1u : 0u;  // a nibble wrapped somewhere (each wrap lowers the sum by 15 or 16)
        base += inc - total;
        const uint32_t base2 = base | (base << 16);  // slots are below 2^16: the halves cannot carry into each other
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint4 o = wp4[j];
            o.x += base2; o.y += base2; o.z += base2; o.w += base2;
            wp4[j] = o;
        }
    }
    if (__syncthreads_or((int)bad)) {
        msd_bucket_robust<THREADS>(msd_smem_raw, keys, size, minus_one);  // (reads the bucket again: the keys in registers are not needed)
        return;
    }

    // ---- lookup: slot = wp[word] + nibbles of the word below mine + rank among equal keys ----
    const uint32_t wp_base = smem_addr(&s.wp[0]), staged_base = smem_addr(&s.staged[0]);
#pragma unroll
    for (int i = 0; i < IPT; ++i)
        if ((uint32_t)(i * THREADS + tid) < size) {
            const uint32_t sh = (key[i] & 7u) << 2;
            uint32_t word, wpv;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(cnt_base + msd_word_offset(key[i])));
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(wpv) : "r"(wp_base + ((key[i] & 0xfff8u) >> 2)));
            const uint32_t below = word & ~(0xffffffffu << sh);
            const uint32_t slot = wpv + ((below * 0x11111111u) >> 28) + ((ranks[i >> 3] >> (4 * (i & 7))) & 15u);
            st_shared(staged_base + 4u * slot, key[i]);
        }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < IPT; ++i)
        if ((uint32_t)(i * THREADS + tid) < size) keys[i * THREADS + tid] = s.staged[i * THREADS + tid];
}

// Robust route of F: stable LSD sort of the bucket on its low 16 bits, two 8-bit passes between two shared-memory buffers,
// ranking exactly like small_sort_kernel (warp-striped order, per-warp counters, ballot multisplit).
template <int THREADS>
__device__ __noinline__ void msd_bucket_robust(unsigned char* smem, uint32_t* __restrict__ data, uint32_t n, uint32_t minus_one) {
    constexpr int WARPS = THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t cap = (n + 3u) & ~3u;
    uint32_t* buf0 = reinterpret_cast<uint32_t*>(smem);
    uint32_t* buf1 = buf0 + cap;
    uint32_t* warp_offset = buf1 + cap;  // [WARPS][RADIX]
    uint32_t* scratch = warp_offset + WARPS * RADIX;  // [8] scan scratch, [32] dummy
    __syncthreads();  // everybody is done with the counting layout of this memory
    for (uint32_t j = tid; j < n; j += THREADS) buf0[j] = data[j];
    const uint32_t ipt = (n + THREADS - 1) / THREADS;
    const uint32_t slice_base = (uint32_t)warp * ipt * 32u + (uint32_t)lane;
    const uint32_t my_offset = smem_addr(warp_offset + warp * RADIX);
    const uint32_t le = lanemask_le(), gt = lanemask_gt();
    const uint32_t dummy = smem_addr(scratch + 8 + lane);
    uint32_t* src = buf0;
    uint32_t* dstp = buf1;
    for (int shift = 0; shift < MSD_LOW_BITS; shift += RADIX_BITS) {
        for (int i = tid; i < WARPS * RADIX; i += THREADS) warp_offset[i] = 0;
        __syncthreads();
        const uint32_t dst = smem_addr(dstp);
        for (uint32_t i = 0; i < ipt; ++i) {
            const uint32_t j = slice_base + i * 32u;
            if (j < n) red_add_shared(my_offset + 4u * ((src[j] >> shift) & 255u), 1u);
        }
        __syncthreads();
        if (tid < RADIX) {
            uint32_t total = 0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) total += warp_offset[w * RADIX + tid];
            const uint32_t sbase = block_exclusive_scan_256<uint32_t>(total, scratch, tid);
            uint32_t run = dst + 4u * sbase - 4u;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                const uint32_t c = warp_offset[w * RADIX + tid];
                warp_offset[w * RADIX + tid] = run;
                run += 4u * c;
            }
        }
        __syncthreads();
        for (uint32_t i = 0; i < ipt; ++i) {
            const uint32_t j = slice_base + i * 32u;
            const bool live = j < n;
            const uint32_t e = live ? src[j] : 0u;
            const uint32_t digit = live ? ((e >> shift) & 255u) : 255u;
            uint32_t peers = same_digit_lanes<RANK_BALLOT>(digit, minus_one);
            const uint32_t live_lanes = __ballot_sync(0xffffffffu, live);
            peers = live ? (peers & live_lanes) : (1u << lane);
            const uint32_t upto = 4u * (uint32_t)__popc(peers & le);
            const bool leader = (peers & gt) == 0 && live;
            uint32_t base = atom_add_shared(leader ? my_offset + 4u * digit : dummy, upto);
            base = __shfl_sync(0xffffffffu, base, 31 - __clz(peers));
            if (live) st_shared(base + upto, e);
        }
        uint32_t* t = src; src = dstp; dstp = t;
        __syncthreads();
    }
    for (uint32_t j = tid; j < n; j += THREADS) data[j] = src[j];
}
