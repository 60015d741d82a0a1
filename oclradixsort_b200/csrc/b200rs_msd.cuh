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
// Programmatic dependent launch (the chain H -> PL -> PL -> P1 -> P2 -> F is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization): a kernel of the chain does what needs no global data (zeroing its
// shared memory), then WAITS for the previous grid to complete (griddepcontrol.wait: all of its memory operations are
// visible), then lets the next grid's CTAs be scheduled as SMs drain.  Wait before release, in every CTA, so "the previous
// grid completed" is transitive down the chain; every global access of a chain kernel sits behind its wait.  Both
// instructions are no-ops in a kernel that was launched the ordinary way (profiling on; ncu).  About 5 us per kernel
// boundary otherwise (profiles/r2a_launch_cost.txt).
__device__ __forceinline__ void chain_wait_then_release() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

enum MsdIneligible {
    MSD_WHY_CHECKSUM = 1,  // a 16-bit counter of H wrapped
    MSD_WHY_BUCKET = 2,    // PL: a bucket is larger than F's capacity
    MSD_WHY_EARLY = 4,     // H: a CTA's own count of one bucket already exceeds F's capacity -- H stops early
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
constexpr int MSD_HIST_THREADS_DEFAULT = 1024;
constexpr int MSD_HIST_VECS_DEFAULT = 8;  // LDG.128 in flight per thread (4: 0.190 ms at 2^28 keys, 6: 0.184, 8: 0.182; 512 threads x 8: 0.219)
constexpr size_t MSD_HIST_SMEM = (size_t)(MSD_BUCKETS / 2) * sizeof(uint32_t);

// Every MSD_HIST_CHECK_ROUNDS rounds the CTA looks for a counter above `bucket_cap`: its own share of one bucket already
// exceeds what F can take, so the input is not eligible and the CTA stops (skewed inputs make this kernel slow -- lanes
// that hit the same counter are serialised -- and would be sent to the LSD path anyway).  The decision is local: the
// CTAs see statistically the same data (grid-stride rounds), so they all stop within a check or two of each other,
// and nobody waits for a global flag.
constexpr int MSD_HIST_FIRST_CHECK = 4, MSD_HIST_CHECK_ROUNDS = 32;  // rounds (32 Ki keys per CTA each): one early look, then one per 1 Mi keys
template <int MSD_HIST_THREADS, int MSD_HIST_VECS>
__global__ void __launch_bounds__(MSD_HIST_THREADS, 1)
msd_hist16_kernel(const uint32_t* __restrict__ in, uint64_t n, unsigned long long* __restrict__ joint2 /*[32768] pairs of u32 counters*/,
                  uint32_t* __restrict__ ctl, uint32_t bucket_cap) {
    extern __shared__ __align__(16) uint32_t msd_tab[];  // [32768]
    __shared__ uint32_t s_sum[MSD_HIST_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < MSD_BUCKETS / 8; i += MSD_HIST_THREADS) reinterpret_cast<uint4*>(msd_tab)[i] = make_uint4(0, 0, 0, 0);
    chain_wait_then_release();
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
    int rounds = 0;
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
        ++rounds;
        if (rounds == MSD_HIST_FIRST_CHECK || rounds % MSD_HIST_CHECK_ROUNDS == 0) {  // (every thread of the CTA runs the same number of rounds)
            // a heuristic look at the table (no barrier in front: reductions still in flight only make it a little stale);
            // counters that wrapped in between are caught by the checksum at the end
            uint32_t over = 0;
            for (int w = tid; w < MSD_BUCKETS / 2; w += MSD_HIST_THREADS) {
                const uint32_t x = msd_tab[w];
                over |= ((x & 0xffffu) > bucket_cap || (x >> 16) > bucket_cap) ? 1u : 0u;
            }
            if (__syncthreads_or((int)over)) {  // the histogram is abandoned: nobody will use it
                if (tid == 0) atomicOr(&ctl[MSD_CTL_INELIGIBLE], (uint32_t)MSD_WHY_EARLY);
                return;
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
        if (t != 0) atomicOr(&ctl[MSD_CTL_INELIGIBLE], (uint32_t)MSD_WHY_CHECKSUM);
    }
}

// =================================================================================================
// PL: bucket offsets, cursors, tile table of the segmented pass, eligibility (one CTA)
// =================================================================================================
struct MsdTile {
    uint32_t start;  // first key of the tile in the intermediate buffer (a multiple of 4)
    uint32_t meta;   // (bucket of the first partition pass << 16) | keys in the tile
};
// Layout of the intermediate buffer (output of P1, input of P2): bucket b of the first pass starts at
// (final start of b rounded down to a multiple of 4) + 4 b, so every tile of P2 can be read with aligned 128-bit loads;
// buckets cannot overlap (the rounding loses at most 3, the 4 b term adds 4 per bucket); the buffer needs n + 1024 keys.
__device__ __forceinline__ uint32_t msd_mid_start(uint32_t final_start, uint32_t bucket) { return (final_start & ~3u) + 4u * bucket; }

// PL-a: 256 CTAs x 256 threads, CTA b sums the 256 joint bins of first-pass bucket b (-> hist3[b]) and folds the largest
// bin into the control words.
__global__ void __launch_bounds__(RADIX)
msd_plan_sums_kernel(const uint32_t* __restrict__ joint, uint32_t* __restrict__ hist3 /*[256]*/, uint32_t* __restrict__ ctl) {
    __shared__ uint32_t s_sum[RADIX / 32], s_max[RADIX / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    chain_wait_then_release();
    const uint32_t x = __ldcg(joint + blockIdx.x * RADIX + tid);
    uint32_t sum = x, mx = x;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { s_sum[warp] = sum; s_max[warp] = mx; }
    __syncthreads();
    if (tid == 0) {
        sum = 0; mx = 0;
        for (int w = 0; w < RADIX / 32; ++w) { sum += s_sum[w]; mx = max(mx, s_max[w]); }
        hist3[blockIdx.x] = sum;
        atomicMax(&ctl[MSD_CTL_MAX_BUCKET], mx);
    }
}

// PL-b: 256 CTAs x 256 threads.  Every CTA scans hist3 (256 values) for the bucket starts and the first tile of every
// bucket; CTA b then scans its own 256 joint bins into bucket_off / cursor2 and writes the tile entries of bucket b.
__global__ void __launch_bounds__(RADIX)
msd_plan_kernel(const uint32_t* __restrict__ joint /*[65536]*/, const uint32_t* __restrict__ hist3, uint32_t n, uint32_t tile_keys /* P2 tile */,
                uint32_t bucket_cap /* F capacity */,
                uint32_t* __restrict__ bucket_off /*[65537]: final start of every top-16 bucket*/,
                uint32_t* __restrict__ cursor2 /*[65536]: = bucket_off, consumed by P2*/,
                uint32_t* __restrict__ cursor1 /*[256]: start of every P1 bucket in the intermediate buffer, consumed by P1*/,
                MsdTile* __restrict__ tiles, uint32_t* __restrict__ ctl,
                volatile uint32_t* __restrict__ host_verdict /* pinned host memory: [1] = largest bucket, then [0] = why not eligible (0: eligible) */) {
    __shared__ uint32_t scratch[RADIX / 32];
    __shared__ uint32_t s_bcast[4];
    const int tid = threadIdx.x;
    const uint32_t b = blockIdx.x;
    chain_wait_then_release();
    const uint32_t size3 = __ldcg(hist3 + tid);
    const uint32_t start3 = block_exclusive_scan_256<uint32_t>(size3, scratch, tid);          // final start of first-pass bucket tid
    const uint32_t t3 = (size3 + tile_keys - 1) / tile_keys;
    const uint32_t tile_first = block_exclusive_scan_256<uint32_t>(t3, scratch, tid);          // first P2 tile of bucket tid
    if ((uint32_t)tid == b) { s_bcast[0] = start3; s_bcast[1] = size3; s_bcast[2] = tile_first; s_bcast[3] = t3; }
    if (b == 0) cursor1[tid] = msd_mid_start(start3, (uint32_t)tid);
    if (b == 0 && tid == RADIX - 1) {
        ctl[MSD_CTL_P2_TILES] = tile_first + t3;
        bucket_off[MSD_BUCKETS] = n;
        const uint32_t mx = __ldcg(ctl + MSD_CTL_MAX_BUCKET);
        uint32_t why = __ldcg(ctl + MSD_CTL_INELIGIBLE);  // (H's verdicts: complete, H is two grids back)
        if (mx > bucket_cap) { why |= (uint32_t)MSD_WHY_BUCKET; ctl[MSD_CTL_INELIGIBLE] = why; }
        // the host is polling these two words (it picks the shape of F and decides between this pipeline and the LSD path)
        host_verdict[1] = mx;
        __threadfence_system();
        host_verdict[0] = why;
    }
    __syncthreads();
    const uint32_t my_start = s_bcast[0], my_size = s_bcast[1], my_tile0 = s_bcast[2], my_tiles = s_bcast[3];
    const uint32_t x = __ldcg(joint + b * RADIX + tid);
    const uint32_t off = my_start + block_exclusive_scan_256<uint32_t>(x, scratch, tid);
    bucket_off[b * RADIX + tid] = off;
    cursor2[b * RADIX + tid] = off;
    for (uint32_t k = tid; k < my_tiles; k += RADIX) {
        MsdTile e;
        e.start = msd_mid_start(my_start, b) + k * tile_keys;
        e.meta = (b << 16) | min(tile_keys, my_size - k * tile_keys);
        tiles[my_tile0 + k] = e;
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
                     const MsdTile* __restrict__ tiles, const uint32_t* __restrict__ ctl, uint32_t pf_tiles /* L2 prefetch distance, 0 = off */) {
    using Cfg = MsdPartitionConfig<THREADS, VPT, CAP>;
    static_assert(THREADS >= RADIX, "one thread per digit");
    extern __shared__ __align__(16) unsigned char msd_smem_raw[];
    typename Cfg::Smem& s = *reinterpret_cast<typename Cfg::Smem*>(msd_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < RADIX; i += THREADS) s.cnt[i] = 0;
    if (tid == 0) s.overflow = 0;
    chain_wait_then_release();
    uint32_t start, count;
    if (SEGMENTED) {
        const uint32_t num_tiles = ctl[MSD_CTL_P2_TILES];
        if (blockIdx.x >= num_tiles) return;  // the grid was sized for the upper bound n / TILE + 256
        const MsdTile t = tiles[blockIdx.x];
        start = t.start;
        count = t.meta & 0xffffu;
        cursor += (size_t)(t.meta >> 16) * RADIX;
        // the tile a CTA will load about one wave of CTAs from now is pulled into L2 (one bulk prefetch, nobody waits for it)
        if (tid == 32 && pf_tiles && blockIdx.x + pf_tiles < num_tiles) {
            const MsdTile p = tiles[blockIdx.x + pf_tiles];
            const uint32_t bytes = (((p.meta & 0xffffu) + 3u) & ~3u) * 4u;
            bulk_prefetch_l2(in + p.start, bytes);
        }
    } else {
        // P1 is launched before the host has read the verdict of H + PL (it runs while the host waits): it checks for itself
        if (ctl[MSD_CTL_INELIGIBLE]) return;
        start = blockIdx.x * (uint32_t)Cfg::TILE;
        count = min((uint32_t)Cfg::TILE, n - start);
        if (tid == 32 && pf_tiles) {
            const uint64_t p = ((uint64_t)blockIdx.x + pf_tiles) * Cfg::TILE;
            if (p + Cfg::TILE <= n) bulk_prefetch_l2(in + p, (uint32_t)Cfg::TILE * 4u);
        }
    }

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

    // ---- claim: rank inside the digit's bin; keys below the capacity go straight to their slot ----
    // The 128 keys of one vector (4 per lane) are tested together: if they all share the digit -- presorted input -- lane 0
    // claims for the whole warp; otherwise the four atomics of a lane are issued back to back.
    const uint32_t cnt_base = smem_addr(&s.cnt[0]), bins_base = smem_addr(&s.bins[0]);
    const uint32_t prmt_sel = 0x4440u | (uint32_t)(shift >> 3);
    auto claim_all = [&](uint32_t counters, bool sparse) {
        bool try_same = true;  // warp-uniform: the test is dropped once a vector fails it (random keys never pass; presorted ones always do)
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
            uint32_t d[4], r[4];
            bool live[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                live[c] = full || (uint32_t)(v * THREADS + tid) * 4 + c < count;
                d[c] = __byte_perm(key[v][c], 0u, prmt_sel);
            }
            uint32_t d0 = 0;
            if (try_same) {
                d0 = __shfl_sync(0xffffffffu, d[0], 0);
                const bool same = live[3] && ((d[0] ^ d0) | (d[1] ^ d0) | (d[2] ^ d0) | (d[3] ^ d0)) == 0;  // (live[3] implies live[0..2])
                try_same = __all_sync(0xffffffffu, same);
            }
            if (try_same) {
                uint32_t b = 0;
                if (lane == 0) b = atom_add_shared(counters + 4u * d0, 128u);
                b = __shfl_sync(0xffffffffu, b, 0) + 4u * (uint32_t)lane;
#pragma unroll
                for (int c = 0; c < 4; ++c) r[c] = b + c;
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) r[c] = live[c] ? atom_add_shared(counters + 4u * d[c], 1u) : 0xffffffffu;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (sparse) {
                    if (r[c] < (uint32_t)CAP) st_shared(bins_base + 4u * (d[c] * (uint32_t)CAP + r[c]), key[v][c]);
                } else {
                    if (live[c]) st_shared(bins_base + 4u * r[c], key[v][c]);
                }
            }
        }
    };

    // ---- a tile whose keys all share the digit (presorted / reversed input: every tile but the ones on bucket edges) is
    //      copied straight from the registers, in input order, behind ONE cursor atomic ----
    {
        // cheap first: only if every thread's FIRST vector already agrees with thread 0's digit are all the keys looked at
        const uint32_t mine = __byte_perm(key[0][0], 0u, prmt_sel);
        if (tid == 0) s.overflow = mine;
        __syncthreads();
        const uint32_t dtile = s.overflow;
        const bool first_ok = full && __byte_perm((key[0][0] ^ key[0][1]) | (key[0][0] ^ key[0][2]) | (key[0][0] ^ key[0][3]), 0u, prmt_sel) == 0 && mine == dtile;
        bool uniform = false;
        const int all_first = __syncthreads_and(first_ok);  // CTA-uniform
        if (all_first) {
            uint32_t k_or = key[0][0], k_and = key[0][0];
#pragma unroll
            for (int v = 0; v < VPT; ++v)
#pragma unroll
                for (int c = 0; c < 4; ++c) { k_or |= key[v][c]; k_and &= key[v][c]; }
            uniform = __byte_perm(k_or ^ k_and, 0u, prmt_sel) == 0;  // this thread's keys share the digit (and it is thread 0's)
        }
        const int one_digit = all_first ? __syncthreads_and(uniform) : 0;
        if (one_digit) {
            if (tid == 0) s.cnt2[0] = atomicAdd(&cursor[dtile], (uint32_t)Cfg::TILE);
            __syncthreads();
            uint32_t* dst = out + s.cnt2[0];
            if ((((uintptr_t)dst) & 15u) == 0) {
#pragma unroll
                for (int v = 0; v < VPT; ++v) reinterpret_cast<uint4*>(dst)[v * THREADS + tid] = make_uint4(key[v][0], key[v][1], key[v][2], key[v][3]);
            } else {
#pragma unroll
                for (int v = 0; v < VPT; ++v)
#pragma unroll
                    for (int c = 0; c < 4; ++c) dst[(v * THREADS + tid) * 4 + c] = key[v][c];
            }
            return;
        }
        if (tid == 0) s.overflow = 0;
        // (the barrier after claim_all orders this store before the digit threads' stores to s.overflow)
    }
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
        // ---- sparse route: warp w copies bins w, w + WARPS, ...; lane l of step k handles the key that lands on word l of
        //      the k-th 128-byte line touched by the bin's run, so every store fills one aligned line.  Three steps cover
        //      count + misalignment <= 96 (all but a few bins); a loop takes the rest ----
        const uint32_t out_word = (uint32_t)((uintptr_t)out >> 2);
        auto lds = [](uint32_t addr) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; };
        uint2 next = s.info[warp];
        for (int d = warp; d < RADIX; d += Cfg::WARPS) {
            const uint2 ci = next;
            if (d + Cfg::WARPS < RADIX) next = s.info[d + Cfg::WARPS];  // the next bin's descriptor is in flight during this bin's copy
            const uint32_t a = (out_word + ci.y) & 31u;         // position of the run's first key inside its line
            const int32_t j0 = (int32_t)lane - (int32_t)a;        // negative where the line starts before the run
            const uint32_t sbin = bins_base + 4u * (uint32_t)(d * CAP + j0);
            uint32_t* dst = out + (uint32_t)((int32_t)ci.y + j0);  // (not dereferenced where j0 < 0)
            const int32_t cnt_i = (int32_t)ci.x;
            const bool p0 = j0 >= 0 && j0 < cnt_i, p1 = j0 + 32 < cnt_i, p2 = j0 + 64 < cnt_i;
            uint32_t v0, v1, v2;
            if (p0) v0 = lds(sbin);
            if (p1) v1 = lds(sbin + 128u);
            if (p2) v2 = lds(sbin + 256u);
            if (p0) dst[0] = v0;
            if (p1) dst[32] = v1;
            if (p2) dst[64] = v2;
            if (cnt_i + (int32_t)a > 96) {  // warp-uniform, rare
                for (int32_t k32 = 96; k32 < cnt_i + (int32_t)a; k32 += 32)
                    if (j0 + k32 < cnt_i) dst[k32] = lds(sbin + 4u * (uint32_t)k32);
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
// Counters: 65 536 values x 4 bits = 8 192 words (value v: nibble v & 7 of word v >> 3).  Per key one ATOMS with return: the
// old nibble is the key's rank among equal keys.  Scan: thread t adds up the nibbles of its 32 consecutive words, a block
// scan makes that the exclusive prefix wp[] of every word.  Lookup: slot = wp[word] + (sum of the word's nibbles below
// mine) + rank; the nibble sum is one multiply ((x * 0x11111111) >> 28 adds the eight nibbles when the sum is below 16).
// The low halves go to their slots in a 16-bit staging array (the high half is the bucket's number) and are written back
// with whole-line stores.  Anything that would break the nibble arithmetic -- 16 or more equal keys, a word whose eight
// values hold more than 15 keys -- is caught exactly (old nibble == 15; word total > 15; and, for nibble wraps, the checksum
// of all words against the bucket size) and sends the bucket to the robust route: a stable two-pass LSD sort in the same
// shared memory.  Counter words and wp[] are stored XOR-swizzled inside each thread's scan range so the scan's 128-bit
// accesses are conflict-free.
//
// Measured and rejected (profiles/r2f_msd_perf_count_emit_rejected.txt): writing the sorted bucket back FROM the counters (no ranks, no lookup,
// no staging: lane l walks word 32 r + l of row r and peels its non-empty nibbles off with ffs) -- only 39 % of the lanes
// have a non-empty word and a row of 32 words holds 16 keys, so the walk costs 300 warp instructions per 32 keys against
// 27 for the lookup: 2.94 ms instead of 0.88 at 2^28.
template <int THREADS, int IPT>
struct MsdBucketConfig {
    static constexpr int CAP = THREADS * IPT;
    static constexpr int WORDS = (1 << MSD_LOW_BITS) / 8;
    static constexpr int WORDS_PER_THREAD = WORDS / THREADS;  // 32 (256 threads) or 16 (512 threads)
    static constexpr int CHUNKS = WORDS_PER_THREAD / 4;       // 16-byte chunks of counter words per thread
    struct Smem {
        alignas(16) uint32_t cnt[WORDS];
        alignas(16) uint16_t wp[WORDS];
        alignas(16) uint16_t staged[CAP + 64];  // low 16 bits of the keys in sorted order (the high 16 are the bucket's number) + dummy slots
        uint32_t warp_total[THREADS / 32];
        uint32_t dummy[32];                // keys outside the bucket aim here (bank = lane)
    };
    // the robust route needs two key buffers, the per-warp counters of the stable ranking and its scratch
    static constexpr size_t ROBUST_BYTES = 2 * (size_t)CAP * 4 + (size_t)(THREADS / 32) * RADIX * 4 + 64 * 4;
    static constexpr size_t SMEM_BYTES = sizeof(Smem) > ROBUST_BYTES ? sizeof(Smem) : ROBUST_BYTES;
};

// Byte offset of counter word (key's low 16 bits >> 3), swizzled so that the scan's LDS.128 are conflict-free: thread t of
// the scan owns CHUNKS consecutive 16-byte chunks, the eight threads of a quarter-warp would all hit the same bank group,
// so chunk j of thread t is stored at chunk j ^ f(t) of the thread's own range (f(t) = t & 7 for 8 chunks per thread,
// (t >> 1) & 3 for 4); both are bits 8.. of the key's low half.
template <int CHUNKS>
__device__ __forceinline__ uint32_t msd_word_offset(uint32_t key) {
    static_assert(CHUNKS == 8 || CHUNKS == 4, "256 or 512 threads");
    const uint32_t plain = (key & 0xfff8u) >> 1;  // 4 * (value >> 3)
    return plain ^ ((key >> 4) & (CHUNKS == 8 ? 0x70u : 0x30u));
}
// Byte offset of wp[word] (u16): thread t of the scan owns CHUNKS / 2 consecutive 16-byte chunks of wp (8 words each);
// chunk j of them is stored at j ^ g(t), g(t) = (t >> 1) & 3 for 4 chunks per thread (quarter-warps then cover all eight bank
// groups), (t >> 2) & 1 for 2.
template <int CHUNKS>
__device__ __forceinline__ uint32_t msd_wp_offset(uint32_t key) {
    const uint32_t plain = (key & 0xfff8u) >> 2;  // 2 * (value >> 3)
    return plain ^ (CHUNKS == 8 ? ((key >> 5) & 0x30u) : ((key >> 5) & 0x10u));
}

template <int THREADS>
__device__ void msd_bucket_robust(unsigned char* smem, uint32_t* __restrict__ data, uint32_t s, uint32_t minus_one);

template <int THREADS, int IPT, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS)
msd_bucket_kernel(uint32_t* __restrict__ data, const uint32_t* __restrict__ bucket_off, uint32_t minus_one, uint32_t pf_buckets) {
    using Cfg = MsdBucketConfig<THREADS, IPT>;
    constexpr int CHUNKS = Cfg::CHUNKS;
    extern __shared__ __align__(16) unsigned char msd_smem_raw[];
    typename Cfg::Smem& s = *reinterpret_cast<typename Cfg::Smem*>(msd_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {   // zero the counters
        uint4* z = reinterpret_cast<uint4*>(s.cnt);
#pragma unroll
        for (int i = 0; i < Cfg::WORDS / 4 / THREADS; ++i) z[i * THREADS + tid] = make_uint4(0, 0, 0, 0);
    }
    chain_wait_then_release();
    const uint32_t start = bucket_off[blockIdx.x];
    const uint32_t size = bucket_off[blockIdx.x + 1] - start;
    // the CTA that will take over this CTA's slot handles a bucket about pf_buckets further on: pull it into L2 now
    if (tid == 32 && pf_buckets && blockIdx.x + pf_buckets < (uint32_t)MSD_BUCKETS) {
        const uint32_t p0 = bucket_off[blockIdx.x + pf_buckets] & ~3u, p1 = (bucket_off[blockIdx.x + pf_buckets + 1] + 3u) & ~3u;
        if (p1 > p0 && (((uintptr_t)data & 15u) == 0)) bulk_prefetch_l2(data + p0, (p1 - p0) * 4u);
    }
    if (size <= 1) return;
    // The bucket is addressed through the 128-byte aligned window that starts `lead` keys before it: slot e of the window
    // (thread e % THREADS, item e / THREADS) holds key e - lead, so every warp-wide load and store covers whole lines.
    const uint32_t lead = (uint32_t)(((uintptr_t)(data + start) >> 2) & 31u);
    uint32_t* __restrict__ keys = data + start;           // (robust route)
    uint32_t* __restrict__ window = data + start - lead;  // (never dereferenced below data + start)

    // this thread holds window slots tid, tid + THREADS, ...: the first `mine_n` of its IPT items, minus item 0 when
    // tid < lead
    const uint32_t wend = lead + size;
    const int mine_n = wend > (uint32_t)tid ? (int)((wend - (uint32_t)tid + THREADS - 1) / THREADS) : 0;
    const bool skip0 = (uint32_t)tid < lead;
    auto is_valid = [&](int i) { return i < mine_n && !(i == 0 && skip0); };
    uint32_t key[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i) key[i] = is_valid(i) ? window[i * THREADS + tid] : 0u;
    __syncthreads();

    // ---- count; the returned nibble is the rank among equal keys ----
    const uint32_t cnt_base = smem_addr(&s.cnt[0]);
    uint32_t ranks[(IPT + 7) / 8];
#pragma unroll
    for (int i = 0; i < (IPT + 7) / 8; ++i) ranks[i] = 0;
    // Keys outside the bucket (the first or the last one or two of a thread): the atomic is predicated off, no branch per key.
    uint32_t worst = 0;  // largest old nibble seen: 15 means the nibble wrapped (the 16th equal key)
    constexpr int BATCH = 6;  // atomics in flight per thread before their results are used
#pragma unroll
    for (int i0 = 0; i0 < IPT; i0 += BATCH) {
        uint32_t old[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
            const int i = i0 + j;
            if (i < IPT) old[j] = atom_add_shared_if(is_valid(i), cnt_base + msd_word_offset<CHUNKS>(key[i]), 1u << ((key[i] & 7u) << 2));  // 0 where not valid
        }
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
            const int i = i0 + j;
            if (i < IPT) {
                const uint32_t r = (old[j] >> ((key[i] & 7u) << 2)) & 15u;  // (not valid: 0)
                worst = max(worst, r);
                ranks[i >> 3] |= r << (4 * (i & 7));
            }
        }
    }
    uint32_t bad = worst == 15u ? 1u : 0u;
    __syncthreads();

    // ---- scan of the counter words: thread t owns words WPT t .. WPT t + WPT - 1 ----
    uint32_t total = 0;
    {
        const uint4* mine = reinterpret_cast<const uint4*>(s.cnt) + tid * CHUNKS;
        // wp[] of my words, two per register; my CHUNKS / 2 16-byte chunks are swizzled among themselves like the counters
        // (msd_wp_offset), so that neither these stores nor the base update below conflict
        uint4* wp4 = reinterpret_cast<uint4*>(s.wp) + tid * (CHUNKS / 2);
        const int swz = CHUNKS == 8 ? (tid & 7) : ((tid >> 1) & 3);
        const int wswz = CHUNKS == 8 ? ((tid >> 1) & 3) : ((tid >> 2) & 1);
        uint32_t run = 0, wide = 0;
#pragma unroll
        for (int jj = 0; jj < CHUNKS / 2; ++jj) {  // eight words at a time: their prefixes (relative to my first word) fill one 16-byte store
            uint32_t pre[8];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint4 q = mine[(2 * jj + h) ^ swz];
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    pre[h * 4 + k] = run;
                    const uint32_t t = __dp4a(w[k] & 0x0f0f0f0fu, 0x01010101u, __dp4a((w[k] >> 4) & 0x0f0f0f0fu, 0x01010101u, 0u));
                    run += t;
                    wide |= t;  // bits 4.. set: a word whose eight values hold 16 or more keys
                }
            }
            wp4[jj ^ wswz] = make_uint4(pre[0] | (pre[1] << 16), pre[2] | (pre[3] << 16), pre[4] | (pre[5] << 16), pre[6] | (pre[7] << 16));
        }
        total = run;
        bad |= (wide >> 4) ? 1u : 0u;
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
        for (int j = 0; j < CHUNKS / 2; ++j) {
            uint4 o = wp4[j ^ wswz];
            o.x += base2; o.y += base2; o.z += base2; o.w += base2;
            wp4[j ^ wswz] = o;
        }
    }
    if (__syncthreads_or((int)bad)) {
        msd_bucket_robust<THREADS>(msd_smem_raw, keys, size, minus_one);  // (reads the bucket again: the keys in registers are not needed)
        return;
    }

    // ---- lookup: slot = wp[word] + nibbles of the word below mine + rank among equal keys ----
    const unsigned char* cnt_bytes = reinterpret_cast<const unsigned char*>(s.cnt);
    const unsigned char* wp_bytes = reinterpret_cast<const unsigned char*>(s.wp);
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const bool valid = is_valid(i);  // (a key outside the bucket reads the words of key 0 and stores to a dummy slot)
        const uint32_t sh = (key[i] & 7u) << 2;
        const uint32_t word = *reinterpret_cast<const uint32_t*>(cnt_bytes + msd_word_offset<CHUNKS>(key[i]));
        const uint32_t wpv = *reinterpret_cast<const uint16_t*>(wp_bytes + msd_wp_offset<CHUNKS>(key[i]));
        const uint32_t below = word & ~(0xffffffffu << sh);
        const uint32_t slot = wpv + ((below * 0x11111111u) >> 28) + ((ranks[i >> 3] >> (4 * (i & 7))) & 15u);
        s.staged[valid ? slot : (uint32_t)(Cfg::CAP + 2 * lane)] = (uint16_t)key[i];
    }
    __syncthreads();
    const uint32_t high = blockIdx.x << MSD_LOW_BITS;
#pragma unroll
    for (int i = 0; i < IPT; ++i)
        if (is_valid(i)) window[i * THREADS + tid] = high | s.staged[i * THREADS + tid - lead];
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
