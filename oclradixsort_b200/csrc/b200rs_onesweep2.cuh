// b200rs_onesweep2.cuh -- second-generation scatter pass (included by b200rs_sort.cu, inside its anonymous namespace).
//
// Same contract as onesweep_kernel (one stable scatter pass on an 8-bit digit; replaces SortAndScatterKernel /
// SortAndScatterKeyValueKernel, RadixSort32Kernels.cl:495-631, RadixSortKeyValueKernels.cl:513-663).  What the profile of
// onesweep_kernel showed (profiles/r1_*, gpurun session 3): per 32 pairs 82 warp instructions, of which 16.7 were the
// per-digit tagged look-back walk (5 iterations x 73 instructions per digit thread per tile: at B200 speed a tile
// completes every 50-100 ns, so the nearest tile with an inclusive prefix is 20-80 tiles back) and 10.7 the element-wise
// write-out; L1/TEX data pipe 77 % busy.  Here:
//
//   * TWO-LEVEL decoupled look-back.  Tiles form groups of LB_GROUP.  Every tile publishes its digit counts (u32 words
//     {tag:4 | count:28}); the last tile of a group adds up its group's counts and publishes the GROUP aggregate, later
//     the group's inclusive prefix, in one tagged u64 word per digit -- the classic single-word protocol, but at group
//     granularity, so the chain is LB_GROUP times shorter.  A tile's prefix = digit start (pre-scanned histogram) +
//     group chain (window of 4 independent loads per step) + counts of the tiles before it in its own group (at most
//     LB_GROUP-1 independent loads).  No fences anywhere: every word validates itself through its tag.
//   * ORDER_EARLY: the look-back runs BEFORE the ranking phase, so the global position of every digit's run is known
//     when the tile-local layout is decided: each run is placed in shared memory at the same offset modulo 16 bytes as
//     its destination in global memory, and the write-out is one cp.async.bulk (TMA, shared -> global) per digit for the
//     16-byte-aligned body of its run plus at most (16/sizeof(elem) - 1) element stores at each ragged end.
//   * ORDER_LATE: look-back after the ranking phase (its latency hides behind the ranking of the other resident CTAs),
//     element-wise write-out with per-digit 64-bit base pointers.
#pragma once

enum WriteOut { WO_ELEM = 0, WO_BULK = 1 };
enum LookbackOrder { ORDER_LATE = 0, ORDER_EARLY = 1 };
enum TileLoad { LOAD_LDG = 0, LOAD_BULK = 1 };
constexpr uint32_t PASS_IDENTITY = 1u;  // pass control word, see digit_start_kernel
constexpr uint32_t PASS_REGULAR = 2u;   // every non-empty digit bin of the pass holds the same count (+-): see SWZ / DSWZ below  // LOAD_BULK: one cp.async.bulk (TMA) brings the whole tile into shared memory

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <typename ElemT, int THREADS, int IPT, int WO>
struct Onesweep2Config {
    static constexpr int WARPS = THREADS / 32;
    static constexpr int TILE = THREADS * IPT;
    static constexpr int WARP_SLICE = 32 * IPT;
    static constexpr int M = WO == WO_BULK ? 16 / (int)sizeof(ElemT) : 1;         // elements per 16-byte chunk
    static constexpr int STAGE_SLOTS = TILE + (M > 1 ? 2 * (M - 1) * RADIX : 0);  // every run may be padded at both ends
    struct Smem {
        alignas(128) ElemT staged[STAGE_SLOTS];  // tile in sorted order; run d starts at a slot congruent to its global index mod M
        uint32_t warp_offset[WARPS][RADIX];     // per-warp digit counts -> running staged slot of (warp, digit)
        uint64_t run_ptr[RADIX];                // WO_ELEM: byte address in `out` of staged slot 0, as seen by digit d's run
        uint64_t mbar;                          // LOAD_BULK: completion of the tile's bulk copy
        uint64_t scan_scratch[RADIX / 32];
        uint32_t dummy[32];                     // lanes that are not their group's leader aim their atomic here (bank = lane)
        uint32_t tile;
        uint32_t ctl;
        uint64_t n_eff;
    };
};

// Look-back table (both arrays zeroed by the host once per sort; tags differ per pass):
//   partial[tile][256]  u32  {tag:4 = 2*pass + 1 | elements of the tile with this digit : 28}
//   group[tile / LB_GROUP][256] u64 {tag:8 | value:56}
//        tag 2*pass + 1: value = elements of the group's tiles with this digit
//        tag 2*pass + 2: value = elements of tiles 0 .. last tile of the group with this digit
constexpr int LB_GROUP = 8;
constexpr int LB2_PARTIAL_SHIFT = 28;
struct Lookback3 {
    uint32_t* partial;
    uint64_t* group;
};

// Sum of `count` (< LB_GROUP, CTA-uniform) consecutive tiles' counts of one digit; spins until every word carries `tag`.
__device__ __forceinline__ uint32_t sum_tile_partials(const uint32_t* p, uint32_t count, uint32_t tag) {
    uint32_t sum;
    while (true) {
        sum = 0;
        uint32_t bad = 0;
#pragma unroll
        for (int j = 0; j < LB_GROUP - 1; ++j)
            if ((uint32_t)j < count) {
                const uint32_t x = ld_relaxed_u32(p + j * RADIX) ^ tag;  // the count if the tag matches, >= 2^28 otherwise
                sum += x;
                bad |= x;
            }
        if ((bad >> LB2_PARTIAL_SHIFT) == 0) break;
    }
    return sum;
}

// Early half, digit threads, right after the tile's digit totals are known: publish them; the last tile of a group
// also publishes the group aggregate (it needs its group-mates' counts for that, and keeps their sum).
__device__ __forceinline__ uint32_t lookback3_publish(const Lookback3& lb, uint32_t tile, uint32_t pass, uint32_t total, int tid) {
    const uint32_t tag_partial = (2u * pass + 1u) << LB2_PARTIAL_SHIFT;
    st_relaxed_u32(&lb.partial[(uint64_t)tile * RADIX + tid], tag_partial | total);
    const uint32_t q = tile % LB_GROUP;
    uint32_t mates = 0;
    if (q == LB_GROUP - 1) {
        mates = sum_tile_partials(lb.partial + (uint64_t)(tile - q) * RADIX + tid, q, tag_partial);
        st_relaxed_u64(&lb.group[(uint64_t)(tile / LB_GROUP) * RADIX + tid], ((uint64_t)(2u * pass + 1u) << LB_TAG_SHIFT) | (uint64_t)(mates + total));
    }
    return mates;
}

// Late half: returns the number of elements with this digit in tiles 0 .. tile-1.
__device__ __forceinline__ uint64_t lookback3_resolve(const Lookback3& lb, uint32_t tile, uint32_t pass, uint32_t total, uint32_t mates, int tid) {
    const uint32_t q = tile % LB_GROUP, g = tile / LB_GROUP;
    if (q != LB_GROUP - 1) mates = sum_tile_partials(lb.partial + (uint64_t)(tile - q) * RADIX + tid, q, (2u * pass + 1u) << LB2_PARTIAL_SHIFT);
    uint64_t acc = 0;
    const uint32_t tag_aggregate_hi = (2u * pass + 1u) << (LB_TAG_SHIFT - 32);
    const uint64_t end_of_chain = (uint64_t)(2u * pass + 2u) << LB_TAG_SHIFT;  // "inclusive prefix 0": what lies below group 0
    int32_t h = (int32_t)g - 1;                                                  // nearest group not yet accounted for
    const uint64_t* p = lb.group + (int64_t)h * RADIX + tid;                     // not dereferenced when h < 0
    bool done = h < 0;
    while (!done) {
        uint64_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = j <= h ? ld_relaxed_u64(p - j * RADIX) : end_of_chain;
        bool open = true;  // entries are consumed in order, up to the first one that is not published or is inclusive
        int32_t consumed = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t t = (uint32_t)(w[j] >> 32) - tag_aggregate_hi;  // aggregate: [0, 2^24), inclusive: [2^24, 2^25)
            const bool take = open && t < (2u << (LB_TAG_SHIFT - 32));
            const bool inclusive = t >= (1u << (LB_TAG_SHIFT - 32));
            if (take) {
                acc += w[j] & LB_VALUE_MASK;
                ++consumed;
            }
            done = done || (take && inclusive);
            open = take && !inclusive;
        }
        p -= consumed * RADIX;
        h -= consumed;
    }
    if (q == LB_GROUP - 1) st_relaxed_u64(&lb.group[(uint64_t)g * RADIX + tid], ((uint64_t)(2u * pass + 2u) << LB_TAG_SHIFT) | (acc + mates + total));
    return acc + mates;
}

__device__ __forceinline__ void bulk_copy_s2g(void* gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t sdst, const void* gsrc, uint32_t bytes, uint32_t mbar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sdst), "l"(gsrc), "r"(bytes), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "B200RS_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra B200RS_WAIT_%=;\n\t}"
        ::"r"(mbar), "r"(parity)
        : "memory");
}
// One instruction pulls `bytes` (multiple of 16, 16-byte aligned source) into L2; nothing waits for it.
__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t lanemask_le() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
    return m;
}

// SWZ: the staged tile is stored with the bank bits of every element's shared-memory address XORed with a hash of its
// 128-byte row number (a permutation inside each row, so still a bijection).  Without it the slot of an element is
// run start + fill count, and inputs whose digits are evenly spread -- presorted or reversed keys k*i, any arithmetic
// progression -- make every run of a tile the same length C: lanes holding different digits then hit addresses that differ
// by multiples of C, and with C = 32 words (256 threads x 32 keys / 256 digits) all 32 lanes of a scatter store land in ONE
// bank (measured: 1.30 ms per pass against 0.66 for uniform keys).  The hash decorrelates bank and run start for any C.
template <typename ElemT, int THREADS, int IPT, int WO, int ORDER, int LOAD, bool FULL, bool BYTE_DIGIT, int SWZ>
__device__ __forceinline__ void onesweep2_tile(typename Onesweep2Config<ElemT, THREADS, IPT, WO>::Smem& s, const ElemT* __restrict__ in,
                                               ElemT* __restrict__ out, uint64_t tile_base, uint32_t valid, int shift,
                                               uint32_t digit_mask, uint32_t prmt_sel, uint32_t tile, uint32_t pass,
                                               const unsigned long long* __restrict__ digit_start, const Lookback3& lb, uint32_t minus_one,
                                               bool bulk_loaded) {
    using Cfg = Onesweep2Config<ElemT, THREADS, IPT, WO>;
    constexpr int M = Cfg::M;
    constexpr uint32_t E = (uint32_t)sizeof(ElemT);
    static_assert(WO == WO_ELEM || ORDER == ORDER_EARLY, "the bulk write-out needs the global positions before the tile is laid out");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t slice = warp * Cfg::WARP_SLICE + lane;  // tile-local index of item 0
    const uint32_t my_offset = smem_addr(&s.warp_offset[warp][0]);
    const uint32_t staged = smem_addr(&s.staged[0]);
    constexpr bool SWIZZLE = SWZ != 0 && WO == WO_ELEM;
    constexpr uint32_t SWZ_MASK = E == 4 ? 0x7cu : 0x78u;  // address bits that select the element inside its 128-byte row
    auto swz = [&](uint32_t a) { return SWIZZLE ? a ^ ((((a >> 7) * 0x9E3779B1u) >> 25) & SWZ_MASK) : a; };
    // ... and the counter rows are indexed by d ^ (d >> 5) (a bijection on 0..255): digits that are multiples of 16 or 32 --
    // all a pass over strided keys may contain -- would otherwise share two banks or one
    auto cswz = [&](uint32_t d) { return SWIZZLE ? d ^ (d >> 5) : d; };
    const uint32_t ctid = cswz((uint32_t)tid);  // digit threads: where digit tid's counter lives in a row

    // ---- 1. warp-striped load + per-warp digit counts ----
    ElemT elem[IPT];
    if (LOAD == LOAD_BULK && FULL && bulk_loaded) {
        // the tile was brought into the (not yet used) staging buffer by one bulk copy; every thread picks up its elements
        mbar_wait(smem_addr(&s.mbar), 0);
#pragma unroll
        for (int i = 0; i < IPT; ++i) elem[i] = s.staged[slice + i * 32];
    } else {
        const ElemT* __restrict__ src = in + tile_base + slice;
#pragma unroll
        for (int i = 0; i < IPT; ++i)
            if (FULL || slice + i * 32 < valid) elem[i] = __ldg(src + i * 32);  // read-only path: the input buffer is not written by this pass
    }
#pragma unroll
    for (int i = 0; i < IPT; ++i)
        if (FULL || slice + i * 32 < valid) {
            const uint32_t d = digit_of_opaque<BYTE_DIGIT>(Elem<ElemT>::key(elem[i]), shift, digit_mask, prmt_sel);
            red_add_shared(my_offset + 4u * cswz(d), 1u);
        }
    __syncthreads();

    // ---- 2. one thread per digit: tile totals, publication, (ORDER_EARLY: look-back,) layout of the staged tile ----
    uint32_t total = 0, sbase = 0, mates = 0;  // digit threads: elements of digit tid in this tile, staged slot of the first one
    uint64_t exclusive = 0;                    // digit threads: global index of the first one
    if (tid < RADIX) {
#pragma unroll
        for (int w = 0; w < Cfg::WARPS; ++w) total += s.warp_offset[w][ctid];
        mates = lookback3_publish(lb, tile, pass, total, tid);
        uint32_t a = 0, region = total;
        if (ORDER == ORDER_EARLY) {
            exclusive = (uint64_t)digit_start[tid] + lookback3_resolve(lb, tile, pass, total, mates, tid);
            if (M > 1) {
                a = (uint32_t)(((uintptr_t)out / sizeof(ElemT) + exclusive) & (uint64_t)(M - 1));  // position inside its 16-byte chunk
                region = total ? ((a + total + (uint32_t)(M - 1)) & ~(uint32_t)(M - 1)) : 0u;
            }
        }
        sbase = block_exclusive_scan_256<uint32_t>(region, reinterpret_cast<uint32_t*>(s.scan_scratch), tid) + a;
        // the running slot of (warp, digit) is kept as a shared-memory BYTE address biased by one element, so that the
        // ranking loop adds E * (lanes of the group up to and including me) and stores without any further arithmetic
        uint32_t run = staged + E * sbase - E;
#pragma unroll
        for (int w = 0; w < Cfg::WARPS; ++w) {
            const uint32_t c = s.warp_offset[w][ctid];
            s.warp_offset[w][ctid] = run;
            run += E * c;
        }
    }
    __syncthreads();

    // ---- 3. warp multisplit ranking; each element goes straight to its staged slot ----
    const uint32_t le = lanemask_le(), gt = lanemask_gt();
    const uint32_t dummy = smem_addr(&s.dummy[lane]);
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const bool live = FULL || (slice + i * 32 < valid);
        const uint32_t digit = live ? digit_of<BYTE_DIGIT>(Elem<ElemT>::key(elem[i]), shift, digit_mask, prmt_sel) : (uint32_t)(RADIX - 1);
        uint32_t peers = same_digit_lanes<RANK_BALLOT>(digit, minus_one);
        if (!FULL) {
            const uint32_t live_lanes = __ballot_sync(0xffffffffu, live);
            peers = live ? (peers & live_lanes) : (1u << lane);  // padding lanes are nobody's peers
        }
        // upto = E * (lanes of my group up to and including me).  The highest lane of each group claims the group's slots
        // (its upto is E * group size); every lane issues the atomic -- the others on a private dummy word -- so there is
        // no branch.  My slot's address = (claimed base) + upto, thanks to the -E bias of the counters.
        const uint32_t upto = E * (uint32_t)__popc(peers & le);
        const bool leader = (peers & gt) == 0 && live;
        uint32_t base = atom_add_shared(leader ? my_offset + 4u * cswz(digit) : dummy, upto);
        base = __shfl_sync(0xffffffffu, base, 31 - __clz(peers));
        if (live) st_shared(swz(base + upto), elem[i]);
    }

    // ---- 4. write-out ----
    if (WO == WO_BULK) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the staged tile is read by the async proxy below
        __syncthreads();
        if (tid < RADIX && total) {
            const uint32_t a = sbase & (uint32_t)(M - 1);
            uint32_t head = (uint32_t)(M - a) & (uint32_t)(M - 1);  // elements before the first 16-byte boundary
            if (head > total) head = total;
            const uint32_t body = (total - head) & ~(uint32_t)(M - 1);
            const uint32_t tail = total - head - body;
            ElemT* g = out + exclusive;
            const ElemT* sp = &s.staged[sbase];
            if (body) bulk_copy_s2g(g + head, smem_addr(sp + head), body * E);
#pragma unroll
            for (int j = 0; j < M - 1; ++j)
                if ((uint32_t)j < head) g[j] = sp[j];
#pragma unroll
            for (int j = 0; j < M - 1; ++j)
                if ((uint32_t)j < tail) g[head + body + j] = sp[head + body + j];
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must outlive the copy's reads
        }
    } else {
        if (tid < RADIX) {
            if (ORDER == ORDER_LATE) exclusive = (uint64_t)digit_start[tid] + lookback3_resolve(lb, tile, pass, total, mates, tid);
            // (a 32-bit element-index table + IMAD.WIDE per element was measured 5-6 % slower than this 64-bit pointer table)
            // SWIZZLE: the write-out adds the element's (un-swizzled) absolute shared address, so `staged` is taken off here
            s.run_ptr[tid] = (uint64_t)(uintptr_t)(out + exclusive) - (uint64_t)sbase * E - (SWIZZLE ? (uint64_t)staged : 0ull);
        }
        __syncthreads();
        {
            const uint64_t my_bytes = (uint64_t)tid * E;
            const uint32_t my_addr = staged + (uint32_t)tid * E;  // physical position tid of the staged tile
#pragma unroll
            for (int k = 0; k < IPT; ++k) {
                const uint32_t j = (uint32_t)tid + (uint32_t)k * THREADS;  // physical position read by this thread
                // its logical slot (position in sorted order), as an absolute shared address: the swizzle is its own inverse
                const uint32_t logical = swz(my_addr + (uint32_t)k * THREADS * E);
                if (FULL || (SWIZZLE ? (logical - staged) / E : j) < valid) {
                    const ElemT e = s.staged[j];
                    const uint32_t d = digit_of<BYTE_DIGIT>(Elem<ElemT>::key(e), shift, digit_mask, prmt_sel);
                    // A plain C++ store through a pointer rebuilt from an integer compiles to a GENERIC store (ST.E), and
                    // that is deliberate: telling the compiler the address is global (__isGlobal / st.global) lets it hoist
                    // all of the loop's shared-memory reads above the stores, which measured 4 % slower (keys and pairs).
                    if (SWIZZLE) *reinterpret_cast<ElemT*>(s.run_ptr[d] + (uint64_t)logical) = e;
                    else         *reinterpret_cast<ElemT*>(s.run_ptr[d] + my_bytes + (uint64_t)k * THREADS * E) = e;
                }
            }
        }
    }
}

// (DSWZ == 2) the swizzled tile body as a separate function: its register allocation and spills stay out of the plain bodies
template <typename ElemT, int THREADS, int IPT, int WO, int ORDER, int LOAD>
__device__ __noinline__ void onesweep2_tile_regular(typename Onesweep2Config<ElemT, THREADS, IPT, WO>::Smem& s, const ElemT* __restrict__ in, ElemT* __restrict__ out,
                                                    uint64_t tile_base, uint32_t valid, int shift, uint32_t digit_mask, uint32_t prmt_sel, uint32_t tile, uint32_t pass,
                                                    const unsigned long long* __restrict__ digit_start, Lookback3 lb, uint32_t minus_one, bool in_aligned) {
    onesweep2_tile<ElemT, THREADS, IPT, WO, ORDER, LOAD, true, true, 1>(s, in, out, tile_base, valid, shift, digit_mask, prmt_sel, tile, pass, digit_start, lb, minus_one, in_aligned);
}

// DSWZ: the swizzled body is compiled in next to the plain one and taken (CTA-uniform branch, like the identity-pass test)
// only in passes that digit_start flagged PASS_REGULAR -- presorted / reversed / strided keys, or a shuffled permutation of
// a full range -- so inputs with ordinary histograms never pay for the swizzle's 7 + 4 extra instructions per element.
template <typename ElemT, int THREADS, int IPT, int MIN_CTAS, int WO, int ORDER, int LOAD, int SWZ = 0, int DSWZ = 0>
__global__ void __launch_bounds__(THREADS, MIN_CTAS)
onesweep2_kernel(const ElemT* __restrict__ in, ElemT* __restrict__ out, uint64_t n, int shift, uint32_t digit_mask,
                 const unsigned long long* __restrict__ digit_start /*[RADIX]: exclusive scan of the pass's histogram*/, Lookback3 lb,
                 uint32_t* ticket, uint32_t pass, uint32_t minus_one /* 0xffffffff, opaque to ptxas: see same_digit_lanes */,
                 const unsigned long long* __restrict__ n_dev, const uint32_t* __restrict__ pass_ctl,
                 uint32_t pf_tiles /* L2 prefetch distance in tiles (about the number of co-resident CTAs); 0 = off */) {
    using Cfg = Onesweep2Config<ElemT, THREADS, IPT, WO>;
    static_assert(THREADS >= RADIX && THREADS % 32 == 0, "one thread per digit is needed");
    // (own symbol: the other kernel families of this translation unit declare their dynamic shared memory with 16-byte
    // alignment, and the swizzled body needs the staged tile on a 128-byte boundary)
    extern __shared__ __align__(128) unsigned char os2_smem_raw[];
    typename Cfg::Smem& s = *reinterpret_cast<typename Cfg::Smem*>(os2_smem_raw);
    // Thread 0 fetches, concurrently, the tile ticket, the pass control word (digit_start_kernel: PASS_IDENTITY = every
    // element has the same digit in this pass, so the pass moves nothing) and the device-side element count (multi-GPU
    // sort): one L2 round trip for all three.
    const int tid = threadIdx.x;
    const bool in_aligned = ((uintptr_t)in & 15u) == 0;  // the bulk copy needs a 16-byte aligned source
    if (tid == 0) {
        const uint32_t ctl = pass_ctl[pass];
        const uint64_t n_eff = n_dev ? min(n, (uint64_t)*n_dev) : n;  // the grid was sized for the upper bound n
        const uint32_t t = atomicAdd(ticket, 1u);
        s.tile = t;
        s.ctl = ctl;
        s.n_eff = n_eff;
        // The CTA that will take over this CTA's slot gets a ticket about pf_tiles higher: pull that tile into L2 now
        // (one bulk prefetch, nobody waits for it), so its counting phase sees L2 latency instead of HBM latency.
        if (pf_tiles) {  // (an input that is not 16-byte aligned -- the second half of a partitioned sort -- is prefetched from the boundary below)
            const uint64_t pf_base = ((uint64_t)t + pf_tiles) * Cfg::TILE;
            if (pf_base + Cfg::TILE <= n_eff)
                bulk_prefetch_l2(reinterpret_cast<const void*>(reinterpret_cast<uintptr_t>(in + pf_base) & ~(uintptr_t)15), Cfg::TILE * (uint32_t)sizeof(ElemT));
        }
        if (LOAD == LOAD_BULK) {
            const uint32_t mbar = smem_addr(&s.mbar);
            mbar_init(mbar, 1);
            const uint64_t base = (uint64_t)t * Cfg::TILE;
            if (!(ctl & PASS_IDENTITY) && in_aligned && base + Cfg::TILE <= n_eff)
                bulk_copy_g2s(smem_addr(&s.staged[0]), in + base, Cfg::TILE * (uint32_t)sizeof(ElemT), mbar);
        }
    }
#pragma unroll
    for (int i = tid; i < Cfg::WARPS * RADIX; i += THREADS) (&s.warp_offset[0][0])[i] = 0;
    __syncthreads();
    n = s.n_eff;
    const uint32_t tile = s.tile;
    const uint64_t tile_base = (uint64_t)tile * Cfg::TILE;
    if (tile_base >= n) return;  // only when n came from n_dev: surplus CTAs (nobody looks back at them)
    const uint32_t valid = (uint32_t)min((uint64_t)Cfg::TILE, n - tile_base);
    if (s.ctl & PASS_IDENTITY) {
        // identity pass: the tile is copied straight across (the buffers keep ping-ponging on the host's schedule);
        // no counting, ranking or look-back -- about half the time of a real pass for keys
        const ElemT* __restrict__ src = in + tile_base;
        ElemT* __restrict__ dst = out + tile_base;
        if (valid == Cfg::TILE && ((((uintptr_t)in | (uintptr_t)out) & 15u) == 0)) {
            constexpr int VECS = Cfg::TILE * (int)sizeof(ElemT) / 16;
            const uint4* s4 = reinterpret_cast<const uint4*>(src);
            uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll 4
            for (int v = tid; v < VECS; v += THREADS) d4[v] = __ldg(s4 + v);
        } else {
            for (uint32_t j = tid; j < valid; j += THREADS) dst[j] = src[j];
        }
        return;
    }
    const bool byte_digit = digit_mask == (uint32_t)(RADIX - 1);
    const uint32_t prmt_sel = 0x4440u | (uint32_t)(shift >> 3);

    if (DSWZ == 2 && valid == Cfg::TILE && byte_digit && (s.ctl & PASS_REGULAR)) {
        onesweep2_tile_regular<ElemT, THREADS, IPT, WO, ORDER, LOAD>(s, in, out, tile_base, valid, shift, digit_mask, prmt_sel, tile, pass, digit_start, lb, minus_one, in_aligned);
    } else if (DSWZ == 1 && valid == Cfg::TILE && byte_digit && (s.ctl & PASS_REGULAR)) {
        onesweep2_tile<ElemT, THREADS, IPT, WO, ORDER, LOAD, true, true, 1>(s, in, out, tile_base, valid, shift, digit_mask, prmt_sel, tile, pass, digit_start, lb, minus_one, in_aligned);
    } else if (valid == Cfg::TILE) {
        if (byte_digit) onesweep2_tile<ElemT, THREADS, IPT, WO, ORDER, LOAD, true, true, SWZ>(s, in, out, tile_base, valid, shift, digit_mask, prmt_sel, tile, pass, digit_start, lb, minus_one, in_aligned);
        else            onesweep2_tile<ElemT, THREADS, IPT, WO, ORDER, LOAD, true, false, SWZ>(s, in, out, tile_base, valid, shift, digit_mask, prmt_sel, tile, pass, digit_start, lb, minus_one, in_aligned);
    } else {
        onesweep2_tile<ElemT, THREADS, IPT, WO, ORDER, LOAD, false, false, SWZ>(s, in, out, tile_base, valid, shift, digit_mask, prmt_sel, tile, pass, digit_start, lb, minus_one, false);
    }
}

// PASS_REGULAR test for one 256-bin histogram (x = this thread's bin), 256 threads, block-wide barriers: between 2 and 255
// non-empty bins whose counts differ by at most 1.  Random keys practically never pass it (two bins of 2^27 random keys
// differ by ~16 000; a looser bound of max >> 12 let "two distinct values" through and cost it 17 %); a pass over strided
// keys that uses only every 16th or 32nd digit does.  With all 256 bins equally full the odd tile shapes are already conflict-free and the plain body is
// faster (presorted keys 0.576 against 0.674 ms), so that case is left alone.
__device__ __forceinline__ uint32_t bins_are_regular(uint64_t x, int tid) {
    __shared__ uint64_t s_min[RADIX / 32], s_max[RADIX / 32];
    __shared__ uint32_t s_cnt[RADIX / 32];
    uint64_t mn = x ? x : ~0ull, mx = x;
    uint32_t cnt = x ? 1u : 0u;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    }
    if ((tid & 31) == 0) { s_min[tid >> 5] = mn; s_max[tid >> 5] = mx; s_cnt[tid >> 5] = cnt; }
    __syncthreads();
    mn = ~0ull; mx = 0; cnt = 0;
#pragma unroll
    for (int w = 0; w < RADIX / 32; ++w) { mn = min(mn, s_min[w]); mx = max(mx, s_max[w]); cnt += s_cnt[w]; }
    __syncthreads();
    return cnt >= 2u && cnt < (uint32_t)RADIX && mx - mn <= 1ull ? 1u : 0u;
}

// One CTA, after the histogram kernel.  For each pass: in-place exclusive scan of its 256-bin histogram
// (digit_start[p][d] = number of elements with a smaller digit) and the pass control word: a pass in which one digit
// holds every element moves nothing, so its kernel only copies (PASS_IDENTITY).
__global__ void __launch_bounds__(RADIX) digit_start_kernel(unsigned long long* __restrict__ ghist /*[passes][RADIX]*/, int passes, uint64_t n,
                                                            const unsigned long long* __restrict__ n_dev, uint32_t* __restrict__ ctl) {
    __shared__ uint64_t scratch[RADIX / 32];
    if (n_dev) n = min(n, (uint64_t)*n_dev);
    for (int p = 0; p < passes; ++p) {
        unsigned long long* h = ghist + (size_t)p * RADIX;
        const uint64_t x = h[threadIdx.x];
        const int degenerate = __syncthreads_or(x == n);
        const uint32_t regular = bins_are_regular(x, threadIdx.x);  // (block-wide, uses __syncthreads)
        h[threadIdx.x] = block_exclusive_scan_256<uint64_t>(x, scratch, threadIdx.x);
        if (threadIdx.x == 0) ctl[p] = (degenerate ? PASS_IDENTITY : 0u) | (regular ? PASS_REGULAR : 0u);
    }
}
