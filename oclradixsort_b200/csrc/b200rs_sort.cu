// b200rs_sort.cu -- LSD radix sort of u32 keys / u32-u32 pairs, one scatter pass per 8-bit digit
// (Onesweep-style) for sm_100a.
//
// Replaces the reference's per-4-bit-pass chain StreamCount -> PrefixScan{16,32}PerWi -> SortAndScatter
// (Tahoe/ClKernels/RadixSort32Kernels.cl:178-631, RadixSortKeyValueKernels.cl:184-663, driven by
// Pprims::radixSort, Pprims.cpp:200-406): 8 passes x 3 launches and 96 B/key there; here
//   1 launch  digit_histogram_kernel : reads the input once, builds the histograms of ALL digits
//   P launches onesweep_kernel       : P = ceil(sort_bits/8); reads each element once, writes it once
// => 4 + 8P bytes per key (36 B at 32 bits), 8 + 16P per pair.
//
// onesweep_kernel, per tile of TILE consecutive elements (tile ids from an atomic ticket):
//   1. warp-striped load: warp w owns a contiguous slice, element (i, lane) sits at slice + i*32 + lane,
//      so (i, lane) order == input order -- this is what makes the ranking stable;
//   2. warp multisplit ranking: __match_any_sync groups the lanes holding the same digit; the lowest
//      lane of each group bumps that warp's private counter in shared memory; rank = old counter +
//      number of lower lanes in the group (no atomics, 32-lane warps);
//   3. per-digit totals across warps -> published to the look-back table as PARTIAL; tile-local bin
//      starts by a 256-wide block scan; elements scattered to their tile-local sorted slot in shared memory;
//   4. decoupled look-back, one thread per digit: sums predecessors' PARTIALs until an INCLUSIVE is met,
//      publishes its own INCLUSIVE; tile 0 seeds the chain with the exclusive scan of the global histogram,
//      so there is no separate scan launch;
//   5. the tile is written out from shared memory in sorted order: consecutive threads write consecutive
//      addresses within each digit's run.
// Look-back words are 64-bit {tag:8 | count:56}; tag = 2*pass + {1 PARTIAL, 2 INCLUSIVE}, so one table,
// zeroed once per sort, serves every pass and counts never overflow (n up to 2^56).
#include "b200rs_internal.h"

#include <cooperative_groups.h>
#include <atomic>

namespace {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int MAX_PASSES = 4;

// ---- element traits: u32 key, or the 8-byte AoS pair {key, value} (Tahoe::uint2, Math.h:175-188) ----
template <typename ElemT> struct Elem;
template <> struct Elem<uint32_t> {
    static __device__ __forceinline__ uint32_t key(uint32_t e) { return e; }
};
template <> struct Elem<uint2> {
    static __device__ __forceinline__ uint32_t key(uint2 e) { return e.x; }
};

__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// shared-memory accesses through 32-bit shared-window addresses (keeps address arithmetic to one LEA)
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void red_add_shared(uint32_t addr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// =================================================================================================
// Histogram of every digit in one read of the input.
// =================================================================================================
// Tail of the histogram kernels (defined below, next to digit_start_kernel's description): the LAST CTA to finish turns
// every pass's histogram into digit start offsets and writes the pass control words, which saves the separate one-CTA
// launch.  All threads of that CTA call it.
__device__ void digit_start_tail(unsigned long long* ghist, int passes, uint64_t n, uint32_t* ctl);
__device__ __forceinline__ void fused_digit_start(unsigned long long* ghist, int passes, uint64_t n, uint32_t* done_counter, uint32_t* pass_ctl) {
    __shared__ uint32_t s_last;
    __threadfence();  // this CTA's histogram atomics are performed before its arrival is counted
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(done_counter, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (s_last) {
        __threadfence();
        digit_start_tail(ghist, passes, n, pass_ctl);  // every 256 threads take one pass at a time (blockDim.x is a multiple of 256)
    }
}

constexpr int HIST_THREADS = 512;
constexpr int HIST_VEC_PER_THREAD = 4;  // uint4 loads in flight per thread
constexpr int HIST_ROWS = RADIX / 2;    // two 16-bit counters per word: digits r (low half) and r + 128 (high half)
constexpr int HIST_FLUSH_ROUNDS = 255;  // a counter half gains at most 16 warps x 16 keys = 256 per round: 255 rounds < 65536
constexpr size_t HIST_SMEM_BYTES = (size_t)MAX_PASSES * HIST_ROWS * 32 * sizeof(uint32_t);  // 64 KiB

// Shared-memory atomics cost one L1 wavefront per lane that collides on a bank, and a random 8-bit digit makes
// ~3.4 lanes of a warp collide: with one 256-bin table per digit place the round-1 kernel spent 27 SM-cycles per
// 32 keys on its four RED.ADDs (ncu: L1/TEX 94 % busy, DRAM 30 %).  Here every lane owns a private column:
// counter word (place p, row r, lane l) sits in bank l, so a warp's RED never conflicts.  A word packs two 16-bit
// counters (digit r in the low half, digit r + 128 in the high half); the CTA drains the table into the global
// 64-bit histogram before a half can overflow.
template <typename ElemT, int NUM_PASSES>
__global__ void __launch_bounds__(HIST_THREADS)
digit_histogram_kernel(const ElemT* __restrict__ in, uint64_t n, uint32_t key_mask, uint32_t mul_0x101 /* 0x101, kept in a register */,
                       unsigned long long* __restrict__ ghist /*[num_passes][RADIX]*/, const unsigned long long* __restrict__ n_dev,
                       uint32_t* done_counter /* non-null: fused digit starts */, uint32_t* pass_ctl) {
    extern __shared__ __align__(16) uint32_t s_cnt[];  // [MAX_PASSES][HIST_ROWS][32 lanes]
    if (n_dev) n = min(n, (uint64_t)*n_dev);  // element count decided on the device (multi-GPU sort): n is only the upper bound
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < MAX_PASSES * HIST_ROWS * 32; i += HIST_THREADS) s_cnt[i] = 0;
    __syncthreads();

    constexpr int KEYS_PER_VEC = 16 / sizeof(ElemT);  // 4 keys or 2 pairs per uint4
    const uint64_t nvec = n / KEYS_PER_VEC;
    const uint4* in4 = reinterpret_cast<const uint4*>(in);
    const uint32_t table = smem_addr(s_cnt);  // warp-uniform
    const uint32_t col = 4u * lane;           // this lane's column

    constexpr int num_passes = NUM_PASSES;
    auto count = [&](uint32_t key) {
        key &= key_mask;
#pragma unroll
        for (int p = 0; p < NUM_PASSES; ++p) {
                // row = low 7 bits of the digit, half = its top bit:  address = col + p*16K + row*128,
                // increment = 1 << (16 * top bit) = 1 + 0xffff * top bit  (PRMT replicates the byte's sign bit)
                uint32_t row_off;  // (shifted key & 0x3f80) | col in one LOP3
                asm("lop3.b32 %0, %1, 0x3f80, %2, 0xea;" : "=r"(row_off) : "r"(p == 0 ? key << 7 : key >> (8 * p - 7)), "r"(col));
                uint32_t top;  // 0xff if bit 7 of byte p is set, else 0 (prmt: selector bit 3 replicates the byte's sign)
                asm("prmt.b32 %0, %1, 0, %2;" : "=r"(top) : "r"(key), "r"(0x4440u | (8u + p)));
                red_add_shared(table + row_off + (uint32_t)p * (HIST_ROWS * 128u), top * mul_0x101 + 1u);
        }
    };
    // drain: thread t owns row t of the [num_passes*128] rows; lanes read the row rotated so banks stay distinct
    auto drain = [&]() {
        __syncthreads();
        if (tid < num_passes * HIST_ROWS) {
            uint32_t lo = 0, hi = 0;
#pragma unroll 8
            for (int j = 0; j < 32; ++j) {
                const int c = (j + tid) & 31;
                const uint32_t w = s_cnt[tid * 32 + c];
                s_cnt[tid * 32 + c] = 0;
                lo += w & 0xffffu;
                hi += w >> 16;
            }
            const int p = tid / HIST_ROWS, r = tid % HIST_ROWS;
            if (lo) atomicAdd(&ghist[p * RADIX + r], (unsigned long long)lo);
            if (hi) atomicAdd(&ghist[p * RADIX + r + HIST_ROWS], (unsigned long long)hi);
        }
        __syncthreads();
    };

    // every thread of the CTA runs the same number of rounds (the bounds test is inside), so the drain's barriers are uniform
    const uint64_t round_vecs = (uint64_t)HIST_THREADS * HIST_VEC_PER_THREAD;
    int rounds = 0;
    for (uint64_t base = (uint64_t)blockIdx.x * round_vecs; base < nvec; base += (uint64_t)gridDim.x * round_vecs) {
        uint4 q[HIST_VEC_PER_THREAD];
        bool have[HIST_VEC_PER_THREAD];
#pragma unroll
        for (int u = 0; u < HIST_VEC_PER_THREAD; ++u) {
            const uint64_t v = base + (uint64_t)u * HIST_THREADS + tid;
            have[u] = v < nvec;
            if (have[u]) q[u] = __ldg(in4 + v);
        }
#pragma unroll
        for (int u = 0; u < HIST_VEC_PER_THREAD; ++u)
            if (have[u]) {
                if (sizeof(ElemT) == 4) { count(q[u].x); count(q[u].y); count(q[u].z); count(q[u].w); }
                else                    { count(q[u].x); count(q[u].z); }
            }
        if (++rounds == HIST_FLUSH_ROUNDS) { drain(); rounds = 0; }
    }
    // ragged tail (n not a multiple of the vector width): block 0 only
    if (blockIdx.x == 0) {
        const uint64_t i = nvec * KEYS_PER_VEC + tid;
        if (i < n) count(Elem<ElemT>::key(in[i]));
    }
    drain();
    if (done_counter) fused_digit_start(ghist, NUM_PASSES, n, done_counter, pass_ctl);
}

// Variant with one 32-bit counter per (digit, lane): the 16-bit packing above costs two instructions per digit
// (PRMT + IMAD to build the increment) in a kernel that is bound by instruction issue (27 warp instructions per 32 keys,
// issue slots 73 % busy, HBM 61 %).  Here a digit costs SHF + LOP3 + ATOMS: 13 instructions per key instead of 21.  The
// table is [NUM_PASSES][256 digits][32 lanes] words = 32 KiB per digit place, so one 1024-thread CTA per SM; a counter
// cannot overflow (a lane sees fewer than 2^32 keys), so the table is drained once, at the end.
constexpr int HIST32_THREADS = 1024;
template <typename ElemT, int NUM_PASSES>
__global__ void __launch_bounds__(HIST32_THREADS, 1)
digit_histogram32_kernel(const ElemT* __restrict__ in, uint64_t n, uint32_t key_mask, unsigned long long* __restrict__ ghist /*[num_passes][RADIX]*/,
                         const unsigned long long* __restrict__ n_dev, uint32_t* done_counter /* non-null: fused digit starts */, uint32_t* pass_ctl) {
    extern __shared__ __align__(16) uint32_t s_cnt[];  // [NUM_PASSES][RADIX][32 lanes]
    if (n_dev) n = min(n, (uint64_t)*n_dev);
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < NUM_PASSES * RADIX * 32; i += HIST32_THREADS) s_cnt[i] = 0;
    __syncthreads();

    constexpr int KEYS_PER_VEC = 16 / sizeof(ElemT);
    const uint64_t nvec = n / KEYS_PER_VEC;
    const uint4* in4 = reinterpret_cast<const uint4*>(in);
    const uint32_t table = smem_addr(s_cnt);  // CTA-uniform
    const uint32_t col = 4u * lane;           // this lane's column: word (place p, digit d, lane l) sits in bank l
    auto count = [&](uint32_t key) {
        key &= key_mask;
#pragma unroll
        for (int p = 0; p < NUM_PASSES; ++p) {
            uint32_t off;  // digit * 128 | col in one LOP3: the shifted key has the digit at bits 7..14
            asm("lop3.b32 %0, %1, 0x7f80, %2, 0xea;" : "=r"(off) : "r"(p == 0 ? key << 7 : key >> (8 * p - 7)), "r"(col));
            red_add_shared(table + off + (uint32_t)p * (RADIX * 128u), 1u);
        }
    };
    const uint64_t round_vecs = (uint64_t)HIST32_THREADS * HIST_VEC_PER_THREAD;
    for (uint64_t base = (uint64_t)blockIdx.x * round_vecs; base < nvec; base += (uint64_t)gridDim.x * round_vecs) {
        uint4 q[HIST_VEC_PER_THREAD];
        bool have[HIST_VEC_PER_THREAD];
#pragma unroll
        for (int u = 0; u < HIST_VEC_PER_THREAD; ++u) {
            const uint64_t v = base + (uint64_t)u * HIST32_THREADS + tid;
            have[u] = v < nvec;
            if (have[u]) q[u] = __ldg(in4 + v);
        }
#pragma unroll
        for (int u = 0; u < HIST_VEC_PER_THREAD; ++u)
            if (have[u]) {
                if (sizeof(ElemT) == 4) { count(q[u].x); count(q[u].y); count(q[u].z); count(q[u].w); }
                else                    { count(q[u].x); count(q[u].z); }
            }
    }
    if (blockIdx.x == 0) {  // ragged tail (n not a multiple of the vector width)
        const uint64_t i = nvec * KEYS_PER_VEC + tid;
        if (i < n) count(Elem<ElemT>::key(in[i]));
    }
    __syncthreads();
    // drain: thread t owns row t of the [NUM_PASSES * 256] rows; lanes read the row rotated so banks stay distinct
    for (int row = tid; row < NUM_PASSES * RADIX; row += HIST32_THREADS) {
        uint32_t sum = 0;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) sum += s_cnt[row * 32 + ((j + tid) & 31)];
        if (sum) atomicAdd(&ghist[row], (unsigned long long)sum);
    }
    if (done_counter) fused_digit_start(ghist, NUM_PASSES, n, done_counter, pass_ctl);
}

// Same, for inputs whose base pointer is not 16-byte aligned (element-wise loads).
template <typename ElemT>
__global__ void __launch_bounds__(HIST_THREADS)
digit_histogram_unaligned_kernel(const ElemT* __restrict__ in, uint64_t n, int num_passes, uint32_t key_mask,
                                 unsigned long long* __restrict__ ghist, const unsigned long long* __restrict__ n_dev) {
    __shared__ uint32_t s_hist[MAX_PASSES][RADIX];
    if (n_dev) n = min(n, (uint64_t)*n_dev);
    for (int i = threadIdx.x; i < MAX_PASSES * RADIX; i += HIST_THREADS) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * HIST_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * HIST_THREADS + threadIdx.x; i < n; i += stride) {
        const uint32_t key = Elem<ElemT>::key(in[i]) & key_mask;
#pragma unroll
        for (int p = 0; p < MAX_PASSES; ++p)
            if (p < num_passes) atomicAdd(&s_hist[p][(key >> (p * RADIX_BITS)) & (RADIX - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < num_passes * RADIX; i += HIST_THREADS) {
        const uint32_t c = (&s_hist[0][0])[i];
        if (c) atomicAdd(&ghist[i], (unsigned long long)c);
    }
}

// =================================================================================================
// One scatter pass.
// =================================================================================================
constexpr int DEFAULT_LOOKBACK_WINDOW = 4;  // predecessors fetched per look-back step (independent loads in flight)
constexpr uint64_t LB_VALUE_MASK = (1ull << 56) - 1;
constexpr int LB_TAG_SHIFT = 56;

// How the lanes of a warp that hold the same digit find each other.
//   RANK_BALLOT : 8 x vote.ballot (one per digit bit), no shared memory.  Measured 13.8 SM-cycles per 32 keys.
//   RANK_MATCH  : __match_any_sync (MATCH.ANY).  Measured 62 SM-cycles per 32 keys on B200 (about 2 cycles per
//                 distinct value in the warp) -- kept only so the measurement can be reproduced.
//   RANK_ATOMIC_UNORDERED : EXPERIMENT, NOT SELECTABLE BY DEFAULT.  One shared-memory atomicAdd per key and no
//                 peer search: 3.5 SM-cycles per 32 keys, but the order of same-digit lanes inside one warp
//                 instruction is whatever the hardware does (undocumented), so stability is not guaranteed.
//                 Exists to measure what the guaranteed ranking costs.
enum RankMode { RANK_BALLOT = 0, RANK_MATCH = 1, RANK_ATOMIC_UNORDERED = 2 };

// `minus_one` is 0xffffffff passed in as a kernel argument, i.e. a value ptxas cannot see: ~b is then computed as
// b * minus_one + minus_one, an IMAD on the FMA pipe.  The obvious LOP3 form puts 16 logic operations per key on the
// ALU pipe (one warp instruction per 2 cycles per scheduler), which is what bounded the round-1 kernel
// (ncu: "math pipe throttle" + "not selected" = 37 % of the ranking samples).
template <int MODE>
__device__ __forceinline__ uint32_t same_digit_lanes(uint32_t digit, uint32_t minus_one) {
    if (MODE == RANK_MATCH) return __match_any_sync(0xffffffffu, digit);
    // same_bit(k) = lanes whose bit k equals mine: ballot of the bit, complemented where my bit is clear
    auto same_bit = [&](int k) {
        uint32_t b;
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
            "and.b32 t, %1, %2;\n\tsetp.ne.u32 p, t, 0;\n\t"
            "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t@!p mad.lo.u32 %0, %0, %3, %3;\n\t}"
            : "=r"(b)
            : "r"(digit), "r"(1u << k), "r"(minus_one));
        return b;
    };
    // three-input ANDs (4 LOP3 instead of 7), at most three ballots alive at a time
    uint32_t peers, x, y;
    peers = same_bit(0); x = same_bit(1); y = same_bit(2);
    asm volatile("lop3.b32 %0, %0, %1, %2, 0x80;" : "+r"(peers) : "r"(x), "r"(y));
    x = same_bit(3); y = same_bit(4);
    asm volatile("lop3.b32 %0, %0, %1, %2, 0x80;" : "+r"(peers) : "r"(x), "r"(y));
    x = same_bit(5); y = same_bit(6);
    asm volatile("lop3.b32 %0, %0, %1, %2, 0x80;" : "+r"(peers) : "r"(x), "r"(y));
    x = same_bit(7);
    return peers & x;
}

__device__ __forceinline__ uint32_t lanemask_gt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_gt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ uint32_t atom_add_shared(uint32_t addr, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
    return old;
}
// atomicAdd executed only where `pred` is set (predicated instruction, no divergent branch); returns 0 elsewhere
__device__ __forceinline__ uint32_t atom_add_shared_if(bool pred, uint32_t addr, uint32_t v) {
    uint32_t old = 0;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p atom.shared.add.u32 %0, [%1], %2;\n\t}" : "+r"(old) : "r"(addr), "r"(v), "r"((uint32_t)pred) : "memory");
    return old;
}
__device__ __forceinline__ void st_shared(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void st_shared(uint32_t addr, uint2 v) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}

template <typename ElemT, int THREADS, int IPT>
struct OnesweepConfig {
    static constexpr int WARPS = THREADS / 32;
    static constexpr int TILE = THREADS * IPT;
    static constexpr int WARP_SLICE = 32 * IPT;
    struct Smem {
        ElemT staged[TILE];                  // tile in tile-local sorted order
        uint32_t warp_offset[WARPS][RADIX];  // per-warp digit counts -> running tile-local slot of (warp, digit)
        uint64_t run_start[RADIX];           // (global start of this tile's run of digit d) - (tile-local start of d)
        uint32_t scan_warp_total[RADIX / 32];
        uint32_t tile;
        uint8_t lut[RADIX];                  // partition passes only: raw digit -> part (see onesweep_kernel<..., LUT>)
    };
};

// named barriers (0 is __syncthreads over the whole CTA, look-back warp included)
#define B200RS_BAR_DIGITS 1    /* the 256 digit threads, inside block_exclusive_scan_256 */


// 256-wide exclusive scan by threads 0..255 (8 warps); every one of those threads must call it.
template <typename T>
__device__ __forceinline__ T block_exclusive_scan_256(T x, T* warp_totals /*[8] shared*/, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    T inc = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T y = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += y;
    }
    if (lane == 31) warp_totals[warp] = inc;
    asm volatile("bar.sync 1, 256;" ::: "memory");  // named barrier: only the 256 digit threads
    T base = 0;
#pragma unroll
    for (int w = 0; w < RADIX / 32; ++w)
        if (w < warp) base += warp_totals[w];
    asm volatile("bar.sync 1, 256;" ::: "memory");  // warp_totals may be reused by the caller
    return base + inc - x;
}

// See fused_digit_start.  Same result as digit_start_kernel (b200rs_onesweep2.cuh): per pass, in-place exclusive scan of the
// 256-bin histogram and the pass control word (1 = one digit holds every element: the pass only copies).  Group g of 256
// threads handles passes g, g + groups, ... concurrently (each pass is an L2 round trip plus a scan: done one after the
// other by 256 threads it took as long as the separate launch it replaces); group g synchronises on named barrier 1 + g.
__device__ __noinline__ void digit_start_tail(unsigned long long* ghist, int passes, uint64_t n, uint32_t* ctl) {
    __shared__ uint64_t scratch[MAX_PASSES][RADIX / 32], s_min[MAX_PASSES][RADIX / 32], s_max[MAX_PASSES][RADIX / 32];
    __shared__ uint32_t s_cnt[MAX_PASSES][RADIX / 32];
    const int group = threadIdx.x >> 8, groups = blockDim.x >> 8, t = threadIdx.x & 255;
    const int lane = t & 31, warp = t >> 5;
    for (int p = group; p < passes; p += groups) {
        unsigned long long* h = ghist + (size_t)p * RADIX;
        const uint64_t x = __ldcg(h + t);  // written by other CTAs' atomics: read at L2
        uint64_t inc = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t y = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += y;
        }
        const uint32_t degenerate_here = __ballot_sync(0xffffffffu, x == n);
        // PASS_REGULAR (see bins_are_regular in b200rs_onesweep2.cuh): min / max / number of the non-empty bins
        uint64_t mn = x ? x : ~0ull, mx = x;
        uint32_t cnt = x ? 1u : 0u;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
            cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        }
        if (lane == 31) {
            scratch[group][warp] = inc | (degenerate_here ? 1ull << 63 : 0ull);  // counts are < 2^63: bit 63 carries the flag
            s_min[group][warp] = mn; s_max[group][warp] = mx; s_cnt[group][warp] = cnt;
        }
        asm volatile("bar.sync %0, 256;" ::"r"(1 + group) : "memory");
        uint64_t base = 0, any = 0;
        mn = ~0ull; mx = 0; cnt = 0;
#pragma unroll
        for (int w = 0; w < RADIX / 32; ++w) {
            const uint64_t v = scratch[group][w];
            any |= v;
            if (w < warp) base += v & ~(1ull << 63);
            mn = min(mn, s_min[group][w]); mx = max(mx, s_max[group][w]); cnt += s_cnt[group][w];
        }
        h[t] = base + inc - x;
        const uint32_t regular = cnt >= 2u && cnt < (uint32_t)RADIX && mx - mn <= 1ull ? 2u : 0u;  // PASS_REGULAR == 2
        if (t == 0) ctl[p] = (uint32_t)(any >> 63) | regular;  // PASS_IDENTITY == 1
        asm volatile("bar.sync %0, 256;" ::"r"(1 + group) : "memory");  // scratch[group] is reused by this group's next pass
    }
}

// digit of a key: byte `shift/8` (PRMT) when the pass is a whole byte, shift-and-mask otherwise
template <bool BYTE_DIGIT>
__device__ __forceinline__ uint32_t digit_of(uint32_t key, int shift, uint32_t digit_mask, uint32_t prmt_sel) {
    if (BYTE_DIGIT) return __byte_perm(key, 0u, prmt_sel);
    return (key >> shift) & digit_mask;
}
// Same value, but opaque to the optimiser: used in the counting phase so the digits are recomputed (one PRMT)
// in the ranking phase instead of being kept alive in IPT extra registers across two barriers.
template <bool BYTE_DIGIT>
__device__ __forceinline__ uint32_t digit_of_opaque(uint32_t key, int shift, uint32_t digit_mask, uint32_t prmt_sel) {
    uint32_t d;
    if (BYTE_DIGIT) {
        asm volatile("prmt.b32 %0, %1, 0, %2;" : "=r"(d) : "r"(key), "r"(prmt_sel));
    } else {
        asm volatile("shr.b32 %0, %1, %2;" : "=r"(d) : "r"(key), "r"(shift));
        d &= digit_mask;
    }
    return d;
}

template <typename ElemT, int THREADS, int IPT, int MODE, bool FULL, bool BYTE_DIGIT, bool LUT>
__device__ __forceinline__ void count_rank_scatter(typename OnesweepConfig<ElemT, THREADS, IPT>::Smem& s, const ElemT* __restrict__ in,
                                                   uint64_t tile_base, uint32_t valid, int shift, uint32_t digit_mask, uint32_t prmt_sel,
                                                   uint32_t tile, uint64_t* lookback, uint64_t tag_partial, uint32_t& total,
                                                   uint32_t& bin_start, uint32_t minus_one) {
    using Cfg = OnesweepConfig<ElemT, THREADS, IPT>;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t slice = warp * Cfg::WARP_SLICE + lane;  // tile-local index of item 0
    const uint32_t my_offset = smem_addr(&s.warp_offset[warp][0]);
    const uint32_t staged = smem_addr(&s.staged[0]);

    // ---- 1. warp-striped load + per-warp digit counts (shared-memory reductions, no return value) ----
    ElemT elem[IPT];
    const ElemT* __restrict__ src = in + tile_base + slice;  // one base register, immediate offsets
#pragma unroll
    for (int i = 0; i < IPT; ++i)
        if (FULL || slice + i * 32 < valid) elem[i] = src[i * 32];
#pragma unroll
    for (int i = 0; i < IPT; ++i)
        if (FULL || slice + i * 32 < valid)
        {
            uint32_t d = digit_of_opaque<BYTE_DIGIT>(Elem<ElemT>::key(elem[i]), shift, digit_mask, prmt_sel);
            if (LUT) d = s.lut[d];
            red_add_shared(my_offset + 4u * d, 1u);
        }
    __syncthreads();

    // ---- 2. one thread per digit: totals -> PARTIAL published before the (long) ranking phase, so successors
    //         never wait for it; tile-local slots for every (warp, digit) ----
    if (tid < RADIX) {
#pragma unroll
        for (int w = 0; w < Cfg::WARPS; ++w) total += s.warp_offset[w][tid];
        if (tile != 0) st_relaxed_u64(&lookback[(uint64_t)tile * RADIX + tid], tag_partial | total);
        bin_start = block_exclusive_scan_256<uint32_t>(total, s.scan_warp_total, tid);
        uint32_t run = bin_start;
#pragma unroll
        for (int w = 0; w < Cfg::WARPS; ++w) {
            const uint32_t c = s.warp_offset[w][tid];
            s.warp_offset[w][tid] = run;
            run += c;
        }
    }
    __syncthreads();

    // ---- 3. warp multisplit ranking; each element goes straight to its tile-local sorted slot ----
    const uint32_t lt = lanemask_lt(), gt = lanemask_gt();
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const bool live = FULL || (slice + i * 32 < valid);
        uint32_t digit = live ? digit_of<BYTE_DIGIT>(Elem<ElemT>::key(elem[i]), shift, digit_mask, prmt_sel) : (uint32_t)(RADIX - 1);
        if (LUT && live) digit = s.lut[digit];
        const uint32_t counter = my_offset + 4u * digit;
        if (MODE == RANK_ATOMIC_UNORDERED) {
            if (live) st_shared(staged + (uint32_t)sizeof(ElemT) * atom_add_shared(counter, 1u), elem[i]);
        } else {
            uint32_t peers = same_digit_lanes<MODE>(digit, minus_one);
            if (!FULL) {
                const uint32_t live_lanes = __ballot_sync(0xffffffffu, live);
                peers = live ? (peers & live_lanes) : (1u << lane);  // padding lanes are nobody's peers
            }
            // rank inside the group; the highest lane of each group claims slots for the whole group
            // (its rank + 1 is the group size, so one POPC serves both)
            const uint32_t rank = (uint32_t)__popc(peers & lt);
            uint32_t slot = atom_add_shared_if(live && (peers & gt) == 0, counter, rank + 1u);
            slot = __shfl_sync(0xffffffffu, slot, 31 - __clz(peers));
            if (live) st_shared(staged + (uint32_t)sizeof(ElemT) * rank + (uint32_t)sizeof(ElemT) * slot, elem[i]);
        }
    }
}

// LUT = true turns the pass into a stable PARTITION: the raw digit is mapped through digit_lut (256 bytes) to a part
// id, ghist_pass holds the element count of every part, and the output is the parts laid out one after another in
// part order.  Used by the multi-GPU sort to group elements by destination GPU (oclradixsort_b200/dist.py).
// ABS = true (with LUT): instead of one output array, every part p has its own absolute base address
// ghist_pass[p] (bytes, 8-byte aligned) -- possibly in ANOTHER GPU's memory, mapped through CUDA IPC: the
// partition and the exchange over NVLink are then one kernel (plain st.global to peer addresses).
template <typename ElemT, int THREADS, int IPT, int MODE, int MIN_CTAS, bool LUT = false, bool ABS = false, int LOOKBACK_WINDOW = DEFAULT_LOOKBACK_WINDOW>
__global__ void __launch_bounds__(THREADS, MIN_CTAS)
onesweep_kernel(const ElemT* __restrict__ in, ElemT* __restrict__ out, uint64_t n, int shift, uint32_t digit_mask,
                const unsigned long long* __restrict__ ghist_pass /*[RADIX]*/, uint64_t* lookback /*[tiles][RADIX]*/,
                uint32_t* ticket, uint32_t tag_base, const uint8_t* __restrict__ digit_lut, const unsigned long long* __restrict__ n_dev,
                uint32_t minus_one /* 0xffffffff, opaque to ptxas: see same_digit_lanes */) {
    using Cfg = OnesweepConfig<ElemT, THREADS, IPT>;
    static_assert(THREADS >= RADIX && THREADS % 32 == 0, "one thread per digit is needed");
    if (n_dev) n = min(n, (uint64_t)*n_dev);  // device-decided element count; the grid was sized for the upper bound
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typename Cfg::Smem& s = *reinterpret_cast<typename Cfg::Smem*>(smem_raw);

    const int tid = threadIdx.x;
    const uint64_t TAG_PARTIAL = (uint64_t)(tag_base + 1) << LB_TAG_SHIFT;
    const uint64_t TAG_INCLUSIVE = (uint64_t)(tag_base + 2) << LB_TAG_SHIFT;

    if (tid == 0) s.tile = atomicAdd(ticket, 1u);
#pragma unroll
    for (int i = tid; i < Cfg::WARPS * RADIX; i += THREADS) (&s.warp_offset[0][0])[i] = 0;
    if (LUT && tid < RADIX) s.lut[tid] = digit_lut[tid];
    __syncthreads();
    const uint32_t tile = s.tile;
    const uint64_t tile_base = (uint64_t)tile * Cfg::TILE;
    if (tile_base >= n) return;  // only when n came from n_dev: surplus CTAs of a grid sized for the upper bound
    const uint32_t valid = (uint32_t)min((uint64_t)Cfg::TILE, n - tile_base);  // elements of this tile that exist
    const bool byte_digit = !LUT && digit_mask == (uint32_t)(RADIX - 1);
    const uint32_t prmt_sel = 0x4440u | (uint32_t)(shift >> 3);

    uint32_t total = 0, bin_start = 0;  // meaningful in the 256 digit threads
    if (valid == Cfg::TILE) {
        if (byte_digit) count_rank_scatter<ElemT, THREADS, IPT, MODE, true, !LUT, LUT>(s, in, tile_base, valid, shift, digit_mask, prmt_sel, tile, lookback, TAG_PARTIAL, total, bin_start, minus_one);
        else            count_rank_scatter<ElemT, THREADS, IPT, MODE, true, false, LUT>(s, in, tile_base, valid, shift, digit_mask, prmt_sel, tile, lookback, TAG_PARTIAL, total, bin_start, minus_one);
    } else {
        count_rank_scatter<ElemT, THREADS, IPT, MODE, false, false, LUT>(s, in, tile_base, valid, shift, digit_mask, prmt_sel, tile, lookback, TAG_PARTIAL, total, bin_start, minus_one);
    }

    // ---- 4. decoupled look-back, one thread per digit, LOOKBACK_WINDOW predecessors per step ----
    // The chain depth is (latency of one step) / (interval between tile starts): tens of tiles at full speed,
    // so each step fetches a window of predecessors with independent loads instead of one.
    const bool wide = ABS || (n >> 32) != 0;  // element indices need 64 bits
    if (tid < RADIX) {
        uint64_t exclusive = 0;
        if (tile == 0) {
            if (!ABS) {
                // seed: global start of each digit = exclusive scan of the whole-input histogram
                uint64_t* scratch = s.run_start;  // 8 x u64 of it, not yet in use
                exclusive = block_exclusive_scan_256<uint64_t>((uint64_t)ghist_pass[tid], scratch, tid);
            }  // ABS: every part starts at its own base address, the chain carries only this source's counts
        } else {
            const uint32_t tag_partial_hi = (tag_base + 1) << (LB_TAG_SHIFT - 32), tag_inclusive_hi = (tag_base + 2) << (LB_TAG_SHIFT - 32);
            const uint64_t* p = lookback + (uint64_t)(tile - 1) * RADIX + tid;  // tile 0 always publishes INCLUSIVE: the walk ends there
            int32_t ahead = (int32_t)tile;                                     // predecessors that exist below and including *p
            bool done = false;
            while (!done) {
                uint64_t w[LOOKBACK_WINDOW];
#pragma unroll
                for (int j = 0; j < LOOKBACK_WINDOW; ++j) w[j] = j < ahead ? ld_relaxed_u64(p - j * RADIX) : 0ull;
                int consumed = 0;
#pragma unroll
                for (int j = 0; j < LOOKBACK_WINDOW; ++j) {
                    const uint32_t tag_hi = (uint32_t)(w[j] >> 32) & 0xff000000u;
                    const bool inclusive = tag_hi == tag_inclusive_hi;
                    if (!done && consumed == j && (inclusive || tag_hi == tag_partial_hi)) {
                        exclusive += w[j] & LB_VALUE_MASK;
                        consumed = j + 1;
                        done = inclusive;
                    }
                }
                p -= consumed * RADIX;  // entries not yet published for this pass are polled again
                ahead -= consumed;
            }
        }
        st_relaxed_u64(&lookback[(uint64_t)tile * RADIX + tid], TAG_INCLUSIVE | (exclusive + total));
        if (ABS) s.run_start[tid] = (uint64_t)ghist_pass[tid] + (exclusive - bin_start) * sizeof(ElemT);  // byte address of slot 0
        else     s.run_start[tid] = exclusive - bin_start;  // out index of the element at tile-local slot j of digit d: run_start[d] + j
    }
    __syncthreads();

    // ---- 5. write the tile out in sorted order ----
    if (!wide) {
        const uint32_t* run_start32 = reinterpret_cast<const uint32_t*>(s.run_start);  // low words (little endian)
#pragma unroll 4
        for (uint32_t j = tid; j < valid; j += THREADS) {
            const ElemT e = s.staged[j];
            uint32_t d = byte_digit ? __byte_perm(Elem<ElemT>::key(e), 0u, prmt_sel) : ((Elem<ElemT>::key(e) >> shift) & digit_mask);
            if (LUT) d = s.lut[d];
            out[run_start32[2 * d] + j] = e;  // 32-bit index arithmetic (wraps correctly: the true index is < 2^32)
        }
    } else {
#pragma unroll 4
        for (uint32_t j = tid; j < valid; j += THREADS) {
            const ElemT e = s.staged[j];
            uint32_t d = (Elem<ElemT>::key(e) >> shift) & digit_mask;
            if (LUT) d = s.lut[d];
            if (ABS) *reinterpret_cast<ElemT*>(s.run_start[d] + (uint64_t)j * sizeof(ElemT)) = e;
            else     out[s.run_start[d] + j] = e;
        }
    }
}

#include "b200rs_onesweep2.cuh"
#if defined(B200RS_EXPERIMENTS) && B200RS_EXPERIMENTS >= 2
#include "b200rs_onesweep3.cuh"
#endif
#include "b200rs_msd.cuh"

// =================================================================================================
// Small inputs: the whole sort in ONE launch of ONE CTA.
// =================================================================================================
// The multi-kernel chain has a floor of ~55 us (memset + histogram + digit starts + one launch and one look-back chain per
// digit), which the CPU reference beats below ~16K elements -- and the reference's own unit test sweeps 1K .. 1M
// (UnitTest/main.cpp:105).  Up to SMALL_CAP elements everything fits one CTA's shared memory: every pass counts, scans and
// ranks exactly like a scatter-pass tile (warp-striped order, per-warp counters, ballot multisplit: stable) but scatters
// into a second shared-memory buffer instead of global memory, and nothing needs a histogram, a ticket or a look-back.
constexpr int SMALL_THREADS = 256;
constexpr int SMALL_WARPS = SMALL_THREADS / 32;
constexpr int SMALL_CAP = 8192;  // elements: 2 x 64 KiB of staging for pairs
template <typename ElemT>
struct SmallSortSmem {
    alignas(16) ElemT buf[2][SMALL_CAP];
    uint32_t warp_offset[SMALL_WARPS][RADIX];
    uint32_t scan_scratch[RADIX / 32];
    uint32_t dummy[32];
};

template <typename ElemT>
__global__ void __launch_bounds__(SMALL_THREADS, 1)
small_sort_kernel(ElemT* __restrict__ inout, uint32_t n, int sort_bits, uint32_t minus_one) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmallSortSmem<ElemT>& s = *reinterpret_cast<SmallSortSmem<ElemT>*>(smem_raw);
    constexpr uint32_t E = (uint32_t)sizeof(ElemT);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t j = tid; j < n; j += SMALL_THREADS) s.buf[0][j] = inout[j];
    // warp w owns the contiguous slice [w * slice, (w + 1) * slice); element (i, lane) of it is slice_base + 32 i + lane,
    // so (warp, i, lane) order is input order: the stability argument of the scatter pass
    const uint32_t ipt = (n + SMALL_THREADS - 1) / SMALL_THREADS;  // <= 32
    const uint32_t slice_base = (uint32_t)warp * ipt * 32u + (uint32_t)lane;
    const uint32_t my_offset = smem_addr(&s.warp_offset[warp][0]);
    const uint32_t le = lanemask_le(), gt = lanemask_gt();
    const uint32_t dummy = smem_addr(&s.dummy[lane]);
    int cur = 0;
    for (int shift = 0; shift < sort_bits; shift += RADIX_BITS) {
        const int width = sort_bits - shift < RADIX_BITS ? sort_bits - shift : RADIX_BITS;
        const uint32_t digit_mask = (1u << width) - 1u;
        for (int i = tid; i < SMALL_WARPS * RADIX; i += SMALL_THREADS) (&s.warp_offset[0][0])[i] = 0;
        __syncthreads();  // also: the load is complete (first pass)
        const ElemT* src = s.buf[cur];
        const uint32_t dst = smem_addr(&s.buf[cur ^ 1][0]);
        for (uint32_t i = 0; i < ipt; ++i) {
            const uint32_t j = slice_base + i * 32u;
            if (j < n) red_add_shared(my_offset + 4u * ((Elem<ElemT>::key(src[j]) >> shift) & digit_mask), 1u);
        }
        __syncthreads();
        {   // one thread per digit: bin starts, then the running slot of every (warp, digit) as a biased byte address
            uint32_t total = 0;
#pragma unroll
            for (int w = 0; w < SMALL_WARPS; ++w) total += s.warp_offset[w][tid];
            const uint32_t sbase = block_exclusive_scan_256<uint32_t>(total, s.scan_scratch, tid);
            uint32_t run = dst + E * sbase - E;
#pragma unroll
            for (int w = 0; w < SMALL_WARPS; ++w) {
                const uint32_t c = s.warp_offset[w][tid];
                s.warp_offset[w][tid] = run;
                run += E * c;
            }
        }
        __syncthreads();
        for (uint32_t i = 0; i < ipt; ++i) {
            const uint32_t j = slice_base + i * 32u;
            const bool live = j < n;
            ElemT e = ElemT();
            if (live) e = src[j];
            const uint32_t digit = live ? ((Elem<ElemT>::key(e) >> shift) & digit_mask) : (uint32_t)(RADIX - 1);
            uint32_t peers = same_digit_lanes<RANK_BALLOT>(digit, minus_one);
            const uint32_t live_lanes = __ballot_sync(0xffffffffu, live);
            peers = live ? (peers & live_lanes) : (1u << lane);  // padding lanes are nobody's peers
            const uint32_t upto = E * (uint32_t)__popc(peers & le);
            const bool leader = (peers & gt) == 0 && live;
            uint32_t base = atom_add_shared(leader ? my_offset + 4u * digit : dummy, upto);
            base = __shfl_sync(0xffffffffu, base, 31 - __clz(peers));
            if (live) st_shared(base + upto, e);
        }
        cur ^= 1;
        __syncthreads();  // every warp is done with its counter row and the scatter is complete
    }
    __syncthreads();  // (covers sort_bits == 0: the load)
    for (uint32_t j = tid; j < n; j += SMALL_THREADS) inout[j] = s.buf[cur][j];
}

// =================================================================================================
// Mid-size inputs: the whole sort in ONE cooperative launch.
// =================================================================================================
// Between the single-CTA sort (<= 8192 elements) and inputs that fill the GPU, the multi-kernel chain is pure latency: memset,
// histogram and one launch per digit, each a few microseconds of work behind a few microseconds of launch (44-48 us for any
// n from 16K to 512K; the reference's own unit test sweeps 1K .. 1M, UnitTest/main.cpp:105).  Here one grid of co-resident
// CTAs (cooperative launch) does everything, with grid-wide barriers where the chain had launches: clear the scratch,
// histogram of all digits, digit starts (every CTA scans the four 256-bin histograms for itself: no barrier for that), then
// one pass per digit -- the same tile body as the scatter kernel (onesweep2_tile, 2048-element tiles), tile k of CTA c is
// tile c + k * grid, so a tile only waits on tiles of CTAs that are resident and move forward.
constexpr int MIDC_THREADS = 256, MIDC_IPT = 8;
template <typename ElemT>
struct MidSortSmem {
    typename Onesweep2Config<ElemT, MIDC_THREADS, MIDC_IPT, WO_ELEM>::Smem tile;
    unsigned long long start[MAX_PASSES][RADIX];  // digit starts of every pass (exclusive scans of the histograms)
    uint32_t hist[MAX_PASSES][RADIX];
    uint64_t scratch[RADIX / 32];
    uint32_t identity[MAX_PASSES];
};

template <typename ElemT>
__global__ void __launch_bounds__(MIDC_THREADS, 4)
mid_sort_kernel(ElemT* __restrict__ inout, ElemT* __restrict__ alt, uint32_t n, int sort_bits, unsigned long long* __restrict__ ghist /*[MAX_PASSES][RADIX]*/,
                uint32_t* __restrict__ lb_partial, uint64_t* __restrict__ lb_group, uint32_t num_tiles, uint32_t minus_one) {
    namespace cg = cooperative_groups;
    using Cfg = Onesweep2Config<ElemT, MIDC_THREADS, MIDC_IPT, WO_ELEM>;
    extern __shared__ __align__(128) unsigned char midc_smem_raw[];
    MidSortSmem<ElemT>& s = *reinterpret_cast<MidSortSmem<ElemT>*>(midc_smem_raw);
    cg::grid_group grid = cg::this_grid();
    const int tid = threadIdx.x;
    const int passes = (sort_bits + RADIX_BITS - 1) / RADIX_BITS;
    const uint32_t key_mask = sort_bits == 32 ? 0xffffffffu : ((1u << sort_bits) - 1u);

    // ---- clear the look-back tables and the global histograms ----
    const uint32_t num_groups = (num_tiles + LB_GROUP - 1) / LB_GROUP;
    for (uint32_t i = blockIdx.x * MIDC_THREADS + tid; i < num_tiles * RADIX; i += gridDim.x * MIDC_THREADS) lb_partial[i] = 0;
    for (uint32_t i = blockIdx.x * MIDC_THREADS + tid; i < num_groups * RADIX; i += gridDim.x * MIDC_THREADS) lb_group[i] = 0;
    if (blockIdx.x == 0)
        for (int i = tid; i < MAX_PASSES * RADIX; i += MIDC_THREADS) ghist[i] = 0;
    for (int i = tid; i < MAX_PASSES * RADIX; i += MIDC_THREADS) (&s.hist[0][0])[i] = 0;
    grid.sync();

    // ---- histogram of every digit ----
    for (uint32_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const uint32_t base = t * Cfg::TILE;
#pragma unroll
        for (int i = 0; i < MIDC_IPT; ++i) {
            const uint32_t j = base + i * MIDC_THREADS + tid;
            if (j < n) {
                const uint32_t key = Elem<ElemT>::key(inout[j]) & key_mask;
#pragma unroll
                for (int p = 0; p < MAX_PASSES; ++p)
                    if (p < passes) atomicAdd(&s.hist[p][(key >> (p * RADIX_BITS)) & (RADIX - 1)], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < passes * RADIX; i += MIDC_THREADS) {
        const uint32_t c = (&s.hist[0][0])[i];
        if (c) atomicAdd(&ghist[i], (unsigned long long)c);
    }
    grid.sync();

    // ---- digit starts and identity flags, computed by every CTA for itself ----
    for (int p = 0; p < passes; ++p) {
        const unsigned long long x = __ldcg(&ghist[p * RADIX + tid]);
        const int degenerate = __syncthreads_or(x == (unsigned long long)n);
        s.start[p][tid] = block_exclusive_scan_256<uint64_t>(x, s.scratch, tid);
        if (tid == 0) s.identity[p] = degenerate ? 1u : 0u;
    }
    __syncthreads();

    // ---- one pass per digit ----
    Lookback3 lb;
    lb.partial = lb_partial;
    lb.group = lb_group;
    ElemT* src = inout;
    ElemT* dst = alt;
    int moved = 0;
    for (int p = 0; p < passes; ++p) {
        if (s.identity[p]) continue;  // every element has the same digit: the pass would move nothing (CTA-uniform, same in every CTA)
        const int shift = p * RADIX_BITS;
        const int width = sort_bits - shift < RADIX_BITS ? sort_bits - shift : RADIX_BITS;
        const uint32_t digit_mask = (1u << width) - 1u;
        const uint32_t prmt_sel = 0x4440u | (uint32_t)(shift >> 3);
        for (uint32_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
#pragma unroll
            for (int i = tid; i < Cfg::WARPS * RADIX; i += MIDC_THREADS) (&s.tile.warp_offset[0][0])[i] = 0;
            __syncthreads();
            const uint64_t tile_base = (uint64_t)t * Cfg::TILE;
            const uint32_t valid = min((uint32_t)Cfg::TILE, n - (uint32_t)tile_base);
            if (valid == Cfg::TILE && digit_mask == (uint32_t)(RADIX - 1))
                onesweep2_tile<ElemT, MIDC_THREADS, MIDC_IPT, WO_ELEM, ORDER_LATE, LOAD_LDG, true, true, 0>(s.tile, src, dst, tile_base, valid, shift, digit_mask, prmt_sel, t, (uint32_t)p, s.start[p], lb, minus_one, false);
            else
                onesweep2_tile<ElemT, MIDC_THREADS, MIDC_IPT, WO_ELEM, ORDER_LATE, LOAD_LDG, false, false, 0>(s.tile, src, dst, tile_base, valid, shift, digit_mask, prmt_sel, t, (uint32_t)p, s.start[p], lb, minus_one, false);
            __syncthreads();  // the tile body's shared memory is reused by the next tile
        }
        ElemT* tmp = src; src = dst; dst = tmp;
        ++moved;
        grid.sync();  // pass p + 1 reads what every CTA wrote in pass p
    }
    if (moved & 1)  // the result sits in the alternate buffer
        for (uint32_t j = blockIdx.x * MIDC_THREADS + tid; j < n; j += gridDim.x * MIDC_THREADS) inout[j] = src[j];
}

// ---- host side -------------------------------------------------------------------------------------

// Scatter-pass kernels.  The production library contains exactly two per element type: the full-size tile and the 2048-element
// tile of the mid-size path.  Every other shape, the generation-1 / generation-3 bodies, the bulk (TMA) load and write-out
// variants and the measurement-only ranking modes (RANK_MATCH, RANK_ATOMIC_UNORDERED -- the latter is NOT stable by contract)
// are compiled only with -DB200RS_EXPERIMENTS (make EXPERIMENTS=1), where B200RS_KEYS_VARIANT / B200RS_PAIRS_VARIANT select
// them by index for tools/sweep.py.
struct Variant {
    const void* kernel;
    int threads, ipt;
    size_t smem;
    const char* name;
    int gen;  // 1: onesweep_kernel (tagged per-digit look-back), 2: onesweep2_kernel (two-level look-back), 3: onesweep3_kernel (persistent)
    int ctas_per_sm;  // generation 3: the grid is min(tiles, SMs x ctas_per_sm) persistent CTAs
};
#define B200RS_VARIANT(ElemT, THREADS, IPT, MODE, MIN_CTAS) \
    Variant{(const void*)onesweep_kernel<ElemT, THREADS, IPT, MODE, MIN_CTAS>, THREADS, IPT, sizeof(typename OnesweepConfig<ElemT, THREADS, IPT>::Smem), #THREADS "x" #IPT ":" #MODE "/" #MIN_CTAS, 1, 0}
#define B200RS_VARIANT_W(ElemT, THREADS, IPT, MODE, MIN_CTAS, W) \
    Variant{(const void*)onesweep_kernel<ElemT, THREADS, IPT, MODE, MIN_CTAS, false, false, W>, THREADS, IPT, sizeof(typename OnesweepConfig<ElemT, THREADS, IPT>::Smem), #THREADS "x" #IPT ":" #MODE "/" #MIN_CTAS "w" #W, 1, 0}
#define B200RS_VARIANT2(ElemT, THREADS, IPT, MIN_CTAS, WO, ORDER, LOAD) \
    Variant{(const void*)onesweep2_kernel<ElemT, THREADS, IPT, MIN_CTAS, WO, ORDER, LOAD>, THREADS, IPT, sizeof(typename Onesweep2Config<ElemT, THREADS, IPT, WO>::Smem), "v2:" #THREADS "x" #IPT "/" #MIN_CTAS ":" #WO ":" #ORDER ":" #LOAD, 2, 0}
#define B200RS_VARIANT2D(ElemT, THREADS, IPT, MIN_CTAS) \
    Variant{(const void*)onesweep2_kernel<ElemT, THREADS, IPT, MIN_CTAS, WO_ELEM, ORDER_LATE, LOAD_LDG, 0, 1>, THREADS, IPT, sizeof(typename Onesweep2Config<ElemT, THREADS, IPT, WO_ELEM>::Smem), "v2:" #THREADS "x" #IPT "/" #MIN_CTAS ":swizzled-when-regular", 2, 0}
#define B200RS_VARIANT2N(ElemT, THREADS, IPT, MIN_CTAS) \
    Variant{(const void*)onesweep2_kernel<ElemT, THREADS, IPT, MIN_CTAS, WO_ELEM, ORDER_LATE, LOAD_LDG, 0, 2>, THREADS, IPT, sizeof(typename Onesweep2Config<ElemT, THREADS, IPT, WO_ELEM>::Smem), "v2:" #THREADS "x" #IPT "/" #MIN_CTAS ":swizzled-when-regular:noinline", 2, 0}
#define B200RS_VARIANT2S(ElemT, THREADS, IPT, MIN_CTAS) \
    Variant{(const void*)onesweep2_kernel<ElemT, THREADS, IPT, MIN_CTAS, WO_ELEM, ORDER_LATE, LOAD_LDG, 1>, THREADS, IPT, sizeof(typename Onesweep2Config<ElemT, THREADS, IPT, WO_ELEM>::Smem), "v2:" #THREADS "x" #IPT "/" #MIN_CTAS ":swizzled", 2, 0}
#define B200RS_VARIANT3(ElemT, THREADS, IPT, MIN_CTAS, PF) \
    Variant{(const void*)onesweep3_kernel<ElemT, THREADS, IPT, MIN_CTAS, PF>, THREADS, IPT, sizeof(typename Onesweep2Config<ElemT, THREADS, IPT, WO_ELEM>::Smem), "v3:" #THREADS "x" #IPT "/" #MIN_CTAS ":persistent:pf" #PF, 3, MIN_CTAS}

template <typename ElemT> struct Variants;
template <> struct Variants<uint32_t> {
    // 256 threads x 35 keys, 4 CTAs/SM.  35 and not 32: with evenly spread digits (presorted / reversed / strided keys) every
    // run of the staged tile is TILE/256 slots long, and a multiple of 32 words puts all lanes of a scatter store into
    // one shared-memory bank (256x32: 1.30 ms per pass on presorted keys, 256x35: 0.58; uniform keys: 0.66 both); plus the
    // swizzled body for passes flagged PASS_REGULAR
    static const Variant& full_size() { static const Variant v = B200RS_VARIANT2D(uint32_t, 256, 35, 4); return v; }
    static const Variant& mid_size() { static const Variant v = B200RS_VARIANT2(uint32_t, 256, 8, 4, WO_ELEM, ORDER_LATE, LOAD_LDG); return v; }  // 2048-key tiles
#if defined(B200RS_EXPERIMENTS) && B200RS_EXPERIMENTS >= 2  // the full round-1 sweep list (make EXPERIMENTS=2; slow to compile)
    static const Variant* list(int* count) {
        static const Variant v[] = {
            B200RS_VARIANT(uint32_t, 512, 24, RANK_BALLOT, 3),  // default
            B200RS_VARIANT(uint32_t, 512, 20, RANK_BALLOT, 3),
            B200RS_VARIANT(uint32_t, 512, 16, RANK_BALLOT, 3),
            B200RS_VARIANT(uint32_t, 384, 24, RANK_BALLOT, 4),
            B200RS_VARIANT(uint32_t, 1024, 16, RANK_BALLOT, 1),
            B200RS_VARIANT(uint32_t, 512, 16, RANK_MATCH, 3),             // measurement only
            B200RS_VARIANT(uint32_t, 512, 20, RANK_ATOMIC_UNORDERED, 3),  // measurement only, not stable by contract
            B200RS_VARIANT_W(uint32_t, 512, 24, RANK_BALLOT, 3, 8),   // 7
            B200RS_VARIANT_W(uint32_t, 512, 24, RANK_BALLOT, 3, 16),  // 8
            B200RS_VARIANT_W(uint32_t, 512, 24, RANK_BALLOT, 3, 2),   // 9
            B200RS_VARIANT_W(uint32_t, 512, 24, RANK_BALLOT, 3, 1),   // 10
            B200RS_VARIANT2(uint32_t, 512, 24, 3, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 11
            B200RS_VARIANT2(uint32_t, 512, 24, 3, WO_ELEM, ORDER_LATE, LOAD_BULK),  // 12
            B200RS_VARIANT2(uint32_t, 512, 24, 3, WO_BULK, ORDER_EARLY, LOAD_LDG),  // 13  measurement: bulk write-out
            B200RS_VARIANT2(uint32_t, 512, 20, 3, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 14
            B200RS_VARIANT2(uint32_t, 512, 20, 3, WO_ELEM, ORDER_LATE, LOAD_BULK),  // 15
            B200RS_VARIANT2(uint32_t, 512, 16, 4, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 16
            B200RS_VARIANT2(uint32_t, 512, 16, 4, WO_ELEM, ORDER_LATE, LOAD_BULK),  // 17
            B200RS_VARIANT2(uint32_t, 384, 20, 4, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 18
            B200RS_VARIANT2(uint32_t, 384, 20, 4, WO_ELEM, ORDER_LATE, LOAD_BULK),  // 19
            B200RS_VARIANT2(uint32_t, 256, 24, 6, WO_ELEM, ORDER_LATE, LOAD_BULK),  // 20
            B200RS_VARIANT2(uint32_t, 640, 16, 3, WO_ELEM, ORDER_LATE, LOAD_BULK),  // 21
            B200RS_VARIANT2(uint32_t, 256, 24, 5, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 22  51 registers
            B200RS_VARIANT2(uint32_t, 384, 24, 3, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 23  56 registers
            B200RS_VARIANT2(uint32_t, 640, 16, 2, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 24  51 registers
            B200RS_VARIANT2(uint32_t, 256, 32, 4, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 25  default (48 registers, no spills)
            B200RS_VARIANT2(uint32_t, 320, 24, 4, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 26  51 registers
            B200RS_VARIANT2(uint32_t, 512, 20, 2, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 27  64 registers, 2 CTAs
            B200RS_VARIANT2(uint32_t, 256, 32, 5, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 28
            B200RS_VARIANT2(uint32_t, 256, 28, 5, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 29
            B200RS_VARIANT2(uint32_t, 256, 40, 4, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 30
            B200RS_VARIANT2(uint32_t, 256, 48, 3, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 31
            B200RS_VARIANT2(uint32_t, 256, 36, 4, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 32
            B200RS_VARIANT2(uint32_t, 256, 32, 4, WO_ELEM, ORDER_LATE, LOAD_BULK),  // 33
            B200RS_VARIANT3(uint32_t, 256, 32, 4, 1),  // 34  measurement: persistent CTAs + L2 prefetch of the next tile (0.75 ms/pass vs 0.66)
            B200RS_VARIANT3(uint32_t, 256, 32, 4, 0),  // 35  measurement: persistent CTAs only (0.80 ms/pass)
            B200RS_VARIANT2(uint32_t, 256, 33, 4, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 36  odd keys-per-digit average: see the note on structured inputs
            B200RS_VARIANT2(uint32_t, 256, 31, 4, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 37
            B200RS_VARIANT2(uint32_t, 256, 35, 4, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 38
            B200RS_VARIANT2(uint32_t, 256, 29, 4, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 39
            B200RS_VARIANT2S(uint32_t, 256, 32, 4),  // 40  staged tile swizzled
            B200RS_VARIANT2S(uint32_t, 256, 35, 4),  // 41
            B200RS_VARIANT2S(uint32_t, 256, 36, 4),  // 42
            B200RS_VARIANT2(uint32_t, 256, 8, 4, WO_ELEM, ORDER_LATE, LOAD_LDG),    // 43  mid-size inputs (see mid_index)
            B200RS_VARIANT2(uint32_t, 256, 16, 4, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 44
            B200RS_VARIANT2D(uint32_t, 256, 35, 4),  // 45  swizzled body in passes flagged PASS_REGULAR
        };
        *count = sizeof(v) / sizeof(v[0]);
        return v;
    }
    static const char* env() { return "B200RS_KEYS_VARIANT"; }
#endif
};
template <> struct Variants<uint2> {
    // 320 threads x 20 pairs = 25 pairs per digit on average (odd, see the keys note), 3 CTAs/SM
    // plus the swizzled body for passes flagged PASS_REGULAR, as a separate (noinline) function so that its registers and spills
    // stay out of the plain bodies: presorted pairs 4.86 -> 3.91 ms, uniform 4.253 -> 4.258 (inlined: 4.285)
    static const Variant& full_size() {
        static const Variant v = B200RS_VARIANT2N(uint2, 320, 20, 3);
#ifdef B200RS_EXPERIMENTS
        static const Variant plain = B200RS_VARIANT2(uint2, 320, 20, 3, WO_ELEM, ORDER_LATE, LOAD_LDG), d1 = B200RS_VARIANT2D(uint2, 320, 20, 3);
        const int i = b200rs_exp_env("B200RS_PAIRS_REGULAR", 2);
        if (i != 2) return i == 1 ? d1 : plain;
#endif
        return v;
    }
    static const Variant& mid_size() { static const Variant v = B200RS_VARIANT2(uint2, 256, 8, 4, WO_ELEM, ORDER_LATE, LOAD_LDG); return v; }  // 2048-pair tiles
#if defined(B200RS_EXPERIMENTS) && B200RS_EXPERIMENTS >= 2
    static const Variant* list(int* count) {
        static const Variant v[] = {
            B200RS_VARIANT(uint2, 384, 16, RANK_BALLOT, 3),  // default
            B200RS_VARIANT(uint2, 512, 16, RANK_BALLOT, 2),
            B200RS_VARIANT(uint2, 512, 12, RANK_BALLOT, 3),
            B200RS_VARIANT(uint2, 1024, 12, RANK_BALLOT, 1),
            B200RS_VARIANT(uint2, 512, 16, RANK_ATOMIC_UNORDERED, 2),  // measurement only, not stable by contract
            B200RS_VARIANT_W(uint2, 384, 16, RANK_BALLOT, 3, 8),   // 5
            B200RS_VARIANT_W(uint2, 384, 16, RANK_BALLOT, 3, 16),  // 6
            B200RS_VARIANT_W(uint2, 384, 16, RANK_BALLOT, 3, 2),   // 7
            B200RS_VARIANT2(uint2, 384, 16, 3, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 8  default
            B200RS_VARIANT2(uint2, 384, 16, 3, WO_ELEM, ORDER_LATE, LOAD_BULK),  // 9
            B200RS_VARIANT2(uint2, 384, 16, 3, WO_BULK, ORDER_EARLY, LOAD_LDG),  // 10  measurement: bulk write-out
            B200RS_VARIANT2(uint2, 512, 12, 3, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 11
            B200RS_VARIANT2(uint2, 512, 12, 3, WO_ELEM, ORDER_LATE, LOAD_BULK),  // 12
            B200RS_VARIANT2(uint2, 256, 24, 4, WO_ELEM, ORDER_LATE, LOAD_BULK),  // 13
            B200RS_VARIANT2(uint2, 256, 24, 4, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 14
            B200RS_VARIANT2(uint2, 512, 16, 2, WO_ELEM, ORDER_LATE, LOAD_BULK),  // 15
            B200RS_VARIANT2(uint2, 384, 20, 2, WO_ELEM, ORDER_LATE, LOAD_BULK),  // 16
            B200RS_VARIANT2(uint2, 256, 24, 3, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 17
            B200RS_VARIANT2(uint2, 256, 28, 3, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 18
            B200RS_VARIANT2(uint2, 256, 32, 3, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 19
            B200RS_VARIANT2(uint2, 256, 24, 4, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 20 (= 14)
            B200RS_VARIANT2(uint2, 320, 20, 3, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 21
            B200RS_VARIANT2(uint2, 288, 24, 3, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 22
            B200RS_VARIANT3(uint2, 384, 16, 3, 1),  // 23  measurement: persistent CTAs + L2 prefetch of the next tile (1.05 ms/pass vs 0.99)
            B200RS_VARIANT3(uint2, 384, 16, 3, 0),  // 24  measurement: persistent CTAs only (1.15 ms/pass)
            B200RS_VARIANT2(uint2, 384, 18, 3, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 25  27 pairs per digit on average (odd)
            B200RS_VARIANT2(uint2, 384, 14, 3, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 26  21
            B200RS_VARIANT2(uint2, 256, 23, 4, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 27  23
            B200RS_VARIANT2(uint2, 256, 25, 3, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 28  25
            B200RS_VARIANT2(uint2, 256, 21, 4, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 29  21
            B200RS_VARIANT2S(uint2, 384, 16, 3),  // 30  staged tile swizzled
            B200RS_VARIANT2S(uint2, 320, 20, 3),  // 31
            B200RS_VARIANT2(uint2, 256, 8, 4, WO_ELEM, ORDER_LATE, LOAD_LDG),    // 32  mid-size inputs (see mid_index)
            B200RS_VARIANT2(uint2, 256, 16, 4, WO_ELEM, ORDER_LATE, LOAD_LDG),   // 33
            B200RS_VARIANT2D(uint2, 320, 20, 3),  // 34  swizzled body in passes flagged PASS_REGULAR
        };
        *count = sizeof(v) / sizeof(v[0]);
        return v;
    }
    static const char* env() { return "B200RS_PAIRS_VARIANT"; }
#endif
};
// Below MID_N elements the sort is pure latency: a handful of tiles per pass, every CTA a serial chain of ticket, load,
// count, rank, look-back and write-out, on a mostly empty GPU (63 us for any n from 16K to 1M keys with the full-size tiles).
// Small tiles put the same elements on 4 x as many SMs and make each chain 4 x shorter.
constexpr uint64_t MID_N = 1ull << 19;  // measured: 2048-element tiles win up to 2^19 (44-48 us against 63), tie at 2^20
constexpr uint64_t MIDC_MAX_N = 1ull << 20;  // up to here the sort is one cooperative launch (mid_sort_kernel): at most 512 tiles of 2048
constexpr uint64_t MIN_TILE = 256 * 21;     // smallest tile among the variants used above MID_N: temp storage is sized for it
constexpr uint64_t MIN_TILE_MID = 256 * 8;  // ... and up to MID_N
inline uint64_t min_tile_for(uint64_t n) { return n <= MIDC_MAX_N ? MIN_TILE_MID : MIN_TILE; }  // (MIDC_MAX_N >= MID_N)

template <typename ElemT>
const Variant& pick_variant(uint64_t n) {
    const bool mid = n <= MID_N && !b200rs_exp_env("B200RS_NO_MID_PATH", 0);
    const Variant& production = mid ? Variants<ElemT>::mid_size() : Variants<ElemT>::full_size();
#if defined(B200RS_EXPERIMENTS) && B200RS_EXPERIMENTS >= 2
    if (const char* e = getenv(Variants<ElemT>::env())) {
        int count = 0;
        const Variant* v = Variants<ElemT>::list(&count);
        const int idx = atoi(e);
        // (a forced variant whose tiles are smaller than the plan's would overrun the look-back table)
        if (idx >= 0 && idx < count && (uint64_t)v[idx].threads * v[idx].ipt >= min_tile_for(n)) return v[idx];
    }
#endif
    return production;
}

struct SortPlan {
    int passes;
    size_t alt_off, hist_off, ticket_off, lookback_off, total_bytes;
    size_t clear_off, clear_bytes;  // histograms + tickets + look-back table are zeroed per call
    // key-only MSD path (b200rs_msd.cuh), present when msd_candidate(): [joint histogram | control words] are zeroed per call
    bool msd;
    size_t msd_joint_off, msd_ctl_off, msd_bucket_off_off, msd_cursor2_off, msd_cursor1_off, msd_tiles_off;
};

// The MSD path is tried for full 32-bit key sorts of this size range (below: per-bucket fixed costs of the counting step
// dominate; above: a uniform input's buckets exceed its shared memory and element indices are kept in 32 bits).
constexpr uint64_t MSD_MIN_N = 3ull << 26 /* measured: 2^27 keys 1.50 ms either way, 2^28 2.19 (MSD) against 2.85 (LSD) */, MSD_MAX_N = (1ull << 30) - 1;
constexpr uint32_t MSD_MIN_P_TILE = 4096;  // smallest partition tile among the compiled shapes: the tile table is sized for it
enum MsdMode { MSD_AUTO = 0, MSD_FORCED = 1 };  // MSD_FORCED: b200rs_sort_keys_u32_msd -- no lower size limit
template <typename ElemT>
inline bool msd_candidate(uint64_t n, int sort_bits, int mode) {
    return sizeof(ElemT) == 4 && sort_bits == 32 && n <= MSD_MAX_N && (mode == MSD_FORCED ? n >= 2 : n >= MSD_MIN_N);
}

template <typename ElemT>
int make_plan(uint64_t n, int sort_bits, SortPlan* p, int msd_mode = MSD_AUTO) {
    if (sort_bits < 0 || sort_bits > 32) return B200RS_ERR_INVALID_ARGUMENT;
    const uint64_t tiles = (n + min_tile_for(n) - 1) / min_tile_for(n);
    if (tiles > 0x7fffffffull) return B200RS_ERR_TOO_LARGE;
    p->passes = (sort_bits + RADIX_BITS - 1) / RADIX_BITS;
    size_t off = 0;
    p->msd = msd_candidate<ElemT>(n, sort_bits, msd_mode);
    // (MSD: the intermediate buffer pads every first-pass bucket to a 16-byte boundary, see msd_mid_start)
    p->alt_off = off;      off += b200rs_align_up(((size_t)n + (p->msd ? 1024 : 0)) * sizeof(ElemT), 256);
    p->clear_off = off;
    p->hist_off = off;     off += b200rs_align_up(sizeof(unsigned long long) * MAX_PASSES * RADIX, 256);
    p->ticket_off = off;   off += 256;
    // generation 1: tiles x 256 tagged u64 words; generation 2: partial[tiles][256] u32 | group[tiles/LB_GROUP][256] u64
    // (smaller); either way zeroed per sort
    p->lookback_off = off;
    {
        const size_t gen1 = (size_t)tiles * RADIX * sizeof(uint64_t);
        const size_t gen2 = b200rs_align_up((size_t)tiles * RADIX * sizeof(uint32_t), 256) + (size_t)((tiles + LB_GROUP - 1) / LB_GROUP) * RADIX * sizeof(uint64_t);
        off += b200rs_align_up(gen1 > gen2 ? gen1 : gen2, 256);
    }
    p->clear_bytes = off - p->clear_off;
    if (p->msd) {
        p->msd_joint_off = off;      off += (size_t)MSD_BUCKETS * sizeof(uint32_t);
        p->msd_ctl_off = off;        off += 256;
        p->msd_bucket_off_off = off; off += b200rs_align_up((size_t)(MSD_BUCKETS + 1) * sizeof(uint32_t), 256);
        p->msd_cursor2_off = off;    off += (size_t)MSD_BUCKETS * sizeof(uint32_t);
        p->msd_cursor1_off = off;    off += b200rs_align_up((size_t)2 * RADIX * sizeof(uint32_t), 256);  // cursor1[256] | hist3[256]
        p->msd_tiles_off = off;      off += b200rs_align_up(((size_t)(n / MSD_MIN_P_TILE) + RADIX + 1) * sizeof(MsdTile), 256);
    }
    p->total_bytes = off ? off : 256;
    return B200RS_OK;
}

// ---- key-only MSD path: host side ---------------------------------------------------------------------
struct MsdPartitionShape {
    const void* p1;
    const void* p2;
    int threads, tile;
    size_t smem;
    const char* name;
};
#define B200RS_MSD_P(THREADS, VPT, CAP, MIN_CTAS)                                                                                         \
    MsdPartitionShape{(const void*)msd_partition_kernel<THREADS, VPT, CAP, MIN_CTAS, false>, (const void*)msd_partition_kernel<THREADS, VPT, CAP, MIN_CTAS, true>, \
                      THREADS, THREADS * VPT * 4, sizeof(typename MsdPartitionConfig<THREADS, VPT, CAP>::Smem), #THREADS "x" #VPT "x4/cap" #CAP "/" #MIN_CTAS}
struct MsdBucketShape {
    const void* kernel;
    int threads, cap;
    size_t smem;
    const char* name;
};
#define B200RS_MSD_F(THREADS, IPT, MIN_CTAS) \
    MsdBucketShape{(const void*)msd_bucket_kernel<THREADS, IPT, MIN_CTAS>, THREADS, THREADS * IPT, MsdBucketConfig<THREADS, IPT>::SMEM_BYTES, #THREADS "x" #IPT "/" #MIN_CTAS}

inline const MsdPartitionShape& msd_partition_shape() {
    static const MsdPartitionShape production = B200RS_MSD_P(512, 6, 80, 2);  // 12288-key tiles, 80 KiB of bins, 2 CTAs/SM (first sweep: profiles/r2a_msd_first_shapes.txt)
#ifdef B200RS_EXPERIMENTS
    static const MsdPartitionShape v[] = {
        B200RS_MSD_P(256, 8, 64, 3),   // 0: 8192-key tiles, 64 KiB of bins
        B200RS_MSD_P(256, 6, 48, 4),   // 1: 6144-key tiles, 48 KiB
        B200RS_MSD_P(512, 4, 64, 3),   // 2
        B200RS_MSD_P(512, 3, 48, 4),   // 3
        B200RS_MSD_P(384, 4, 48, 4),   // 4
        B200RS_MSD_P(256, 4, 40, 5),   // 5: 4096-key tiles, 40 KiB
        B200RS_MSD_P(1024, 4, 112, 1), // 6: 16384-key tiles, one CTA per SM
        B200RS_MSD_P(512, 6, 80, 2),   // 7: 12288-key tiles
        B200RS_MSD_P(512, 8, 96, 2),   // 8: 16384-key tiles, 96 KiB
        B200RS_MSD_P(512, 5, 72, 3),   // 9: 10240-key tiles, 72 KiB
        B200RS_MSD_P(768, 4, 80, 2),   // 10: 12288-key tiles, 768 threads
        B200RS_MSD_P(1024, 3, 80, 2),  // 11: 12288-key tiles, 1024 threads
        B200RS_MSD_P(640, 5, 80, 2),   // 12: 12800-key tiles
    };
    const int idx = b200rs_exp_env("B200RS_MSD_P", 7);
    if (idx >= 0 && idx < (int)(sizeof(v) / sizeof(v[0]))) return v[idx];
#endif
    return production;
}
// the counting step keeps a whole bucket in registers + shared memory: the smallest shape that holds the largest bucket
inline const MsdBucketShape* msd_bucket_shape(uint32_t max_bucket) {
    static const MsdBucketShape v[] = {
        B200RS_MSD_F(256, 18, 3),  // 4608 keys
        B200RS_MSD_F(256, 24, 3),  // 6144 keys
        B200RS_MSD_F(256, 36, 2),  // 9216 keys
        B200RS_MSD_F(256, 48, 2),  // 12288 keys
    };
#ifdef B200RS_EXPERIMENTS
    static const MsdBucketShape x[] = {B200RS_MSD_F(512, 9, 3), B200RS_MSD_F(512, 10, 3), B200RS_MSD_F(512, 12, 2), B200RS_MSD_F(256, 18, 4), B200RS_MSD_F(256, 20, 3),
                                       B200RS_MSD_F(512, 9, 2)};
    const int idx = b200rs_exp_env("B200RS_MSD_F", -1);
    if (idx >= 4 && idx < 10 && max_bucket + 31u <= (uint32_t)x[idx - 4].cap) return &x[idx - 4];
    if (idx >= 0 && idx < 4 && max_bucket + 31u <= (uint32_t)v[idx].cap) return &v[idx];
#endif
    for (const MsdBucketShape& f : v)
        if (max_bucket + 31u <= (uint32_t)f.cap) return &f;  // (+31: the bucket is read through a 128-byte aligned window)
    return nullptr;
}
constexpr uint32_t MSD_MAX_BUCKET = 256 * 48 - 31;

// Returns B200RS_OK when the keys were sorted here, MSD_NOT_ELIGIBLE when the input is not eligible (the caller runs the LSD
// path), an error code otherwise.  ONE host round trip: the choice between the two paths depends on the joint histogram, and both
// paths are whole kernel chains, so the host waits for H + PL (about a tenth of the sort) and reads two words -- with the first
// partition pass already queued behind them, so the GPU is not idle meanwhile.
constexpr int MSD_NOT_ELIGIBLE = -1000;
// One launch of the chain.  `after_kernel`: the previous operation in the stream is a kernel of the chain, so this one may be
// scheduled while that one drains (programmatic dependent launch; see chain_wait_then_release).  With profiling on there is
// an event record between any two kernels and the attribute changes nothing.
inline cudaError_t msd_launch(const void* kernel, unsigned grid, unsigned threads, size_t smem, cudaStream_t stream, void** args, bool after_kernel) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = after_kernel ? 1 : 0;
    return cudaLaunchKernelExC(&cfg, kernel, args);
}

int sort_keys_msd(b200rs_device* dev, uint32_t* inout, uint64_t n, char* base, const SortPlan& plan) {
    if (((uintptr_t)inout & 15u) != 0 || b200rs_exp_env("B200RS_NO_MSD", 0)) return MSD_NOT_ELIGIBLE;
    if (!dev->pinned_word) B200RS_CUDA(cudaHostAlloc((void**)&dev->pinned_word, 64, cudaHostAllocDefault));
    uint32_t* alt = reinterpret_cast<uint32_t*>(base + plan.alt_off);
    uint32_t* joint = reinterpret_cast<uint32_t*>(base + plan.msd_joint_off);
    uint32_t* ctl = reinterpret_cast<uint32_t*>(base + plan.msd_ctl_off);
    uint32_t* bucket_off = reinterpret_cast<uint32_t*>(base + plan.msd_bucket_off_off);
    uint32_t* cursor2 = reinterpret_cast<uint32_t*>(base + plan.msd_cursor2_off);
    uint32_t* cursor1 = reinterpret_cast<uint32_t*>(base + plan.msd_cursor1_off);
    MsdTile* tiles = reinterpret_cast<MsdTile*>(base + plan.msd_tiles_off);
    const MsdPartitionShape& ps = msd_partition_shape();
    uint32_t n32 = (uint32_t)n;
    const bool pdl = !b200rs_exp_env("B200RS_MSD_NO_PDL", 0);
    // the verdict of H + PL arrives in two words of pinned host memory, written by the last plan kernel (no copy operation in
    // the stream); [0] holds a value no verdict can take until then
    constexpr uint32_t NO_VERDICT = 0xffffffffu;
    volatile uint32_t* verdict = dev->pinned_word + 8;
    verdict[0] = NO_VERDICT;

    B200RS_CUDA(cudaMemsetAsync(joint, 0, (size_t)MSD_BUCKETS * sizeof(uint32_t) + 256, dev->stream));  // joint histogram + control words
    {
        struct HShape { const void* kernel; int threads, vecs; };
        HShape h = {(const void*)msd_hist16_kernel<MSD_HIST_THREADS_DEFAULT, MSD_HIST_VECS_DEFAULT>, MSD_HIST_THREADS_DEFAULT, MSD_HIST_VECS_DEFAULT};
#ifdef B200RS_EXPERIMENTS
        static const HShape hs[] = {{(const void*)msd_hist16_kernel<1024, 8>, 1024, 8}, {(const void*)msd_hist16_kernel<1024, 2>, 1024, 2}, {(const void*)msd_hist16_kernel<1024, 4>, 1024, 4},
                                    {(const void*)msd_hist16_kernel<512, 8>, 512, 8},   {(const void*)msd_hist16_kernel<768, 4>, 768, 4},   {(const void*)msd_hist16_kernel<1024, 6>, 1024, 6}};
        const int hi = b200rs_exp_env("B200RS_MSD_H", 0);
        if (hi >= 0 && hi < (int)(sizeof(hs) / sizeof(hs[0]))) h = hs[hi];
#endif
        B200RS_TRY(b200rs_kernel_setup(dev, h.kernel, MSD_HIST_SMEM));
        b200rs_launch_scope scope(dev, "msd_hist16_keys", n, n * 4ull);
        const uint64_t per_block = (uint64_t)h.threads * h.vecs * 4;
        uint64_t blocks = (n + per_block - 1) / per_block;
        if (blocks > (uint64_t)dev->num_sms) blocks = (uint64_t)dev->num_sms;
        unsigned long long* joint2 = reinterpret_cast<unsigned long long*>(joint);
        uint64_t n64 = n;
        uint32_t cap = MSD_MAX_BUCKET;
        void* args[] = {(void*)&inout, (void*)&n64, (void*)&joint2, (void*)&ctl, (void*)&cap};
        B200RS_CUDA(msd_launch(h.kernel, (unsigned)blocks, (unsigned)h.threads, MSD_HIST_SMEM, dev->stream, args, false));
    }
    uint32_t* hist3 = cursor1 + RADIX;
    {
        b200rs_launch_scope scope(dev, "msd_plan_sums", MSD_BUCKETS, (uint64_t)MSD_BUCKETS * 4);
        void* args[] = {(void*)&joint, (void*)&hist3, (void*)&ctl};
        B200RS_CUDA(msd_launch((const void*)msd_plan_sums_kernel, RADIX, RADIX, 0, dev->stream, args, pdl));
    }
    {
        b200rs_launch_scope scope(dev, "msd_plan", MSD_BUCKETS, (uint64_t)MSD_BUCKETS * 12);
        uint32_t tile_keys = (uint32_t)ps.tile, cap = MSD_MAX_BUCKET;
        uint32_t* verdict_dev = dev->pinned_word + 8;  // (unified addressing: pinned host memory is addressable from the device as is)
        void* args[] = {(void*)&joint, (void*)&hist3, (void*)&n32, (void*)&tile_keys, (void*)&cap, (void*)&bucket_off, (void*)&cursor2, (void*)&cursor1, (void*)&tiles, (void*)&ctl, (void*)&verdict_dev};
        B200RS_CUDA(msd_launch((const void*)msd_plan_kernel, RADIX, RADIX, 0, dev->stream, args, pdl));
    }

    // P1 goes out BEFORE the host has the verdict: it reads the verdict itself (every CTA returns at once when the input is not
    // eligible, ~15 us), so the host's round trip hides behind it instead of leaving the GPU idle
    B200RS_TRY(b200rs_kernel_setup(dev, ps.p1, ps.smem));
    B200RS_TRY(b200rs_kernel_setup(dev, ps.p2, ps.smem));
    const uint32_t* in1 = inout;
    // L2 prefetch distance of the partition passes in tiles: about one wave of co-resident CTAs
    uint32_t pf_tiles = (uint32_t)b200rs_exp_env("B200RS_MSD_PPF", dev->num_sms * 2);
    const MsdTile* no_tiles = nullptr;
    int shift1 = 24, shift2 = 16;
    const uint32_t* ctl_c = ctl;
    {
        b200rs_launch_scope scope(dev, "msd_partition_keys_pass0", n, 8ull * n);
        void* args[] = {&in1, &alt, (void*)&n32, &shift1, &cursor1, &no_tiles, &ctl_c, &pf_tiles};
        B200RS_CUDA(msd_launch(ps.p1, (unsigned)((n + ps.tile - 1) / ps.tile), (unsigned)ps.threads, ps.smem, dev->stream, args, pdl));
    }
    // wait for the verdict (the plan kernels finish ~0.2 ms after the first launch; the stream is polled now and then so that a
    // failed launch or a fault cannot leave this loop spinning)
    for (uint32_t spins = 1; verdict[0] == NO_VERDICT; ++spins) {
        if ((spins & 0x3fffu) == 0) {
            const cudaError_t q = cudaStreamQuery(dev->stream);
            if (q != cudaErrorNotReady && verdict[0] == NO_VERDICT) return q == cudaSuccess ? (int)cudaErrorUnknown : (int)q;
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (verdict[0] != 0) return MSD_NOT_ELIGIBLE;
    const MsdBucketShape* fs = msd_bucket_shape(verdict[1]);
    if (!fs) return MSD_NOT_ELIGIBLE;  // (cannot happen: PL marks buckets above MSD_MAX_BUCKET ineligible; P1 only wrote the temp buffer)
    B200RS_TRY(b200rs_kernel_setup(dev, fs->kernel, fs->smem));
    {
        b200rs_launch_scope scope(dev, "msd_partition_keys_pass1", n, 8ull * n);
        const uint32_t* in2 = alt;
        const MsdTile* tiles_c = tiles;
        void* args[] = {&in2, &inout, (void*)&n32, &shift2, &cursor2, &tiles_c, &ctl_c, &pf_tiles};
        B200RS_CUDA(msd_launch(ps.p2, (unsigned)(n / ps.tile + RADIX), (unsigned)ps.threads, ps.smem, dev->stream, args, pdl));
    }
    {
        b200rs_launch_scope scope(dev, "msd_bucket_keys", n, 8ull * n);
        const uint32_t* off_c = bucket_off;
        // L2 prefetch distance in buckets: about twice the number of co-resident CTAs
        uint32_t pf_buckets = (uint32_t)b200rs_exp_env("B200RS_MSD_PF", dev->num_sms * 6);
        uint32_t minus_one = 0xffffffffu;
        void* args[] = {&inout, &off_c, &minus_one, &pf_buckets};
        B200RS_CUDA(msd_launch(fs->kernel, MSD_BUCKETS, (unsigned)fs->threads, fs->smem, dev->stream, args, pdl));
    }
    B200RS_CUDA(cudaGetLastError());
    return B200RS_OK;
}

template <typename ElemT>
int sort_impl(b200rs_device* dev, ElemT* inout, uint64_t n, int sort_bits, void* temp, size_t* temp_bytes, const char* what,
              const unsigned long long* n_dev = nullptr, int msd_mode = MSD_AUTO, int* msd_used = nullptr,
              const unsigned long long* pre_hist = nullptr /* device, [passes][256] digit counts of the input: the histogram launch is skipped */) {
    if (!dev || !temp_bytes) return B200RS_ERR_INVALID_ARGUMENT;
    if (msd_used) *msd_used = 0;
    SortPlan plan;
    B200RS_TRY(make_plan<ElemT>(n, sort_bits, &plan, msd_mode));
    if (!temp) {
        *temp_bytes = plan.total_bytes;
        return B200RS_OK;
    }
    if (*temp_bytes < plan.total_bytes) return B200RS_ERR_TEMP_TOO_SMALL;
    if (n && !inout) return B200RS_ERR_INVALID_ARGUMENT;
    if (((uintptr_t)inout & (sizeof(ElemT) - 1)) || ((uintptr_t)temp & 255u)) return B200RS_ERR_INVALID_ARGUMENT;
    if (n <= 1 || plan.passes == 0) return B200RS_OK;  // nothing to order

    b200rs_device_guard guard(dev);
    if (n <= (uint64_t)SMALL_CAP && !n_dev && !pre_hist && msd_mode != MSD_FORCED && !b200rs_exp_env("B200RS_NO_SMALL_PATH", 0)) {
        // one launch of one CTA: all passes in shared memory (no histogram, no tickets, no look-back, no temp storage)
        char small_label[48];
        snprintf(small_label, sizeof(small_label), "small_sort_%s", what);
        const size_t smem = sizeof(SmallSortSmem<ElemT>);
        B200RS_TRY(b200rs_kernel_setup(dev, (const void*)small_sort_kernel<ElemT>, smem));
        {
            b200rs_launch_scope scope(dev, small_label, n, 2ull * n * sizeof(ElemT));
            small_sort_kernel<ElemT><<<1, SMALL_THREADS, smem, dev->stream>>>(inout, (uint32_t)n, sort_bits, 0xffffffffu);
        }
        B200RS_CUDA(cudaGetLastError());
        return B200RS_OK;
    }
    char* base = static_cast<char*>(temp);
    if (n <= MIDC_MAX_N && !n_dev && !pre_hist && msd_mode != MSD_FORCED && !b200rs_exp_env("B200RS_NO_COOP_MID", 0)) {
        // one cooperative launch for the whole sort (see mid_sort_kernel); grid = what is co-resident, at most one CTA per tile
        using MCfg = Onesweep2Config<ElemT, MIDC_THREADS, MIDC_IPT, WO_ELEM>;
        const uint32_t tiles = (uint32_t)((n + MCfg::TILE - 1) / MCfg::TILE);
        const size_t smem = sizeof(MidSortSmem<ElemT>);
        int per_sm = 0;
        B200RS_TRY(b200rs_kernel_setup(dev, (const void*)mid_sort_kernel<ElemT>, smem, MIDC_THREADS, &per_sm));
        uint32_t grid = (uint32_t)(per_sm > 0 ? per_sm : 1) * (uint32_t)dev->num_sms;
        if (grid > tiles) grid = tiles;
        ElemT* alt = reinterpret_cast<ElemT*>(base + plan.alt_off);
        unsigned long long* ghist = reinterpret_cast<unsigned long long*>(base + plan.hist_off);
        uint32_t* lb_partial = reinterpret_cast<uint32_t*>(base + plan.lookback_off);
        uint64_t* lb_group = reinterpret_cast<uint64_t*>(base + plan.lookback_off + b200rs_align_up((size_t)tiles * RADIX * sizeof(uint32_t), 256));
        uint32_t n32 = (uint32_t)n, minus_one = 0xffffffffu, tiles_arg = tiles;
        void* args[] = {&inout, &alt, &n32, &sort_bits, &ghist, &lb_partial, &lb_group, &tiles_arg, &minus_one};
        char mid_label[48];
        snprintf(mid_label, sizeof(mid_label), "mid_sort_%s", what);
        {
            b200rs_launch_scope scope(dev, mid_label, n, (uint64_t)(1 + 2 * plan.passes) * n * sizeof(ElemT));
            B200RS_CUDA(cudaLaunchCooperativeKernel((const void*)mid_sort_kernel<ElemT>, dim3(grid), dim3(MIDC_THREADS), args, smem, dev->stream));
        }
        return B200RS_OK;
    }
    if (plan.msd && !n_dev) {
        const int r = sort_keys_msd(dev, reinterpret_cast<uint32_t*>(inout), n, base, plan);
        if (r == B200RS_OK && msd_used) *msd_used = 1;
        if (r != MSD_NOT_ELIGIBLE) return r;
    }
    const Variant& var = pick_variant<ElemT>(n);
    const uint64_t tile_elems = (uint64_t)var.threads * var.ipt;
    const uint32_t num_tiles = (uint32_t)((n + tile_elems - 1) / tile_elems);
    ElemT* alt = reinterpret_cast<ElemT*>(base + plan.alt_off);
    unsigned long long* ghist = reinterpret_cast<unsigned long long*>(base + plan.hist_off);
    uint32_t* tickets = reinterpret_cast<uint32_t*>(base + plan.ticket_off);
    uint64_t* lookback = reinterpret_cast<uint64_t*>(base + plan.lookback_off);
    // zero the histograms, the tickets and the part of the look-back table this tiling uses
    const size_t partial_bytes = b200rs_align_up((size_t)num_tiles * RADIX * sizeof(uint32_t), 256);
    const size_t num_groups = ((size_t)num_tiles + LB_GROUP - 1) / LB_GROUP;
    const size_t clear_bytes = (plan.lookback_off - plan.clear_off) +
                               (var.gen >= 2 ? partial_bytes + num_groups * RADIX * sizeof(uint64_t) : (size_t)num_tiles * RADIX * sizeof(uint64_t));
    B200RS_CUDA(cudaMemsetAsync(base + plan.clear_off, 0, clear_bytes, dev->stream));
    Lookback3 lb2;
    lb2.partial = reinterpret_cast<uint32_t*>(base + plan.lookback_off);
    lb2.group = reinterpret_cast<uint64_t*>(base + plan.lookback_off + partial_bytes);

    const uint32_t key_mask = sort_bits == 32 ? 0xffffffffu : ((1u << sort_bits) - 1u);
    char label[48];
    uint32_t* pass_ctl = tickets + 32;  // [passes], inside the zeroed ticket block
    // generation >= 2 kernels need the pre-scanned histograms + pass control words: the aligned histogram kernels do that in
    // their last CTA (fused_digit_start); only the unaligned fallback still needs the separate digit_start launch
    uint32_t* done_counter = var.gen >= 2 && ((uintptr_t)inout & 15u) == 0 && !pre_hist ? tickets + 48 : nullptr;
    if (pre_hist) {
        B200RS_CUDA(cudaMemcpyAsync(ghist, pre_hist, sizeof(unsigned long long) * (size_t)plan.passes * RADIX, cudaMemcpyDeviceToDevice, dev->stream));
    } else {
        snprintf(label, sizeof(label), "digit_histogram_%s", what);
        b200rs_launch_scope scope(dev, label, n, n * sizeof(ElemT));
        const uint64_t per_block = (uint64_t)HIST_THREADS * HIST_VEC_PER_THREAD * (16 / sizeof(ElemT));
        uint64_t blocks = (n + per_block - 1) / per_block;
        if (((uintptr_t)inout & 15u) == 0) {
            // measured at 2^28 (tools/quick_perf.py): keys, 4 digits 0.190 ms against 0.270 with the packed counters (2 digits:
            // 0.176 / 0.182); pairs, 4 digits 0.330 / 0.350, 2 digits 0.354 / 0.326.  B200RS_HIST_VARIANT=0|1 forces one.
            const int hist_variant = b200rs_exp_env("B200RS_HIST_VARIANT", (n > MID_N && (sizeof(ElemT) == 4 || plan.passes >= 3)) ? 1 : 0);  // small n: the 128 KiB table costs more to clear and drain than it saves
            if (hist_variant == 1) {
                // 32-bit lane-private counters, one 1024-thread CTA per SM (keys: issue-bound with the packed counters)
                const size_t smem = (size_t)plan.passes * RADIX * 32 * sizeof(uint32_t);
                const uint64_t per_block32 = (uint64_t)HIST32_THREADS * HIST_VEC_PER_THREAD * (16 / sizeof(ElemT));
                uint64_t blocks32 = (n + per_block32 - 1) / per_block32;
                if (blocks32 > (uint64_t)dev->num_sms) blocks32 = (uint64_t)dev->num_sms;
                auto kernel = plan.passes == 4 ? digit_histogram32_kernel<ElemT, 4> : plan.passes == 3 ? digit_histogram32_kernel<ElemT, 3>
                            : plan.passes == 2 ? digit_histogram32_kernel<ElemT, 2> : digit_histogram32_kernel<ElemT, 1>;
                B200RS_TRY(b200rs_kernel_setup(dev, (const void*)kernel, smem));
                kernel<<<(unsigned)blocks32, HIST32_THREADS, smem, dev->stream>>>(inout, n, key_mask, ghist, n_dev, done_counter, pass_ctl);
            } else {
                const uint64_t max_blocks = (uint64_t)dev->num_sms * 3;  // 3 x (512 threads, 64 KiB of counters) per SM, grid-stride beyond that
                if (blocks > max_blocks) blocks = max_blocks;
                auto kernel = plan.passes == 4 ? digit_histogram_kernel<ElemT, 4> : plan.passes == 3 ? digit_histogram_kernel<ElemT, 3>
                            : plan.passes == 2 ? digit_histogram_kernel<ElemT, 2> : digit_histogram_kernel<ElemT, 1>;
                B200RS_TRY(b200rs_kernel_setup(dev, (const void*)kernel, HIST_SMEM_BYTES));
                kernel<<<(unsigned)blocks, HIST_THREADS, HIST_SMEM_BYTES, dev->stream>>>(inout, n, key_mask, 0x101u, ghist, n_dev, done_counter, pass_ctl);
            }
        } else {
            const uint64_t max_blocks = (uint64_t)dev->num_sms * 4;
            if (blocks > max_blocks) blocks = max_blocks;
            digit_histogram_unaligned_kernel<ElemT><<<(unsigned)blocks, HIST_THREADS, 0, dev->stream>>>(inout, n, plan.passes, key_mask, ghist, n_dev);
        }
    }
    B200RS_CUDA(cudaGetLastError());
    if (var.gen >= 2 && !done_counter) {
        b200rs_launch_scope scope(dev, "digit_start", (uint64_t)plan.passes * RADIX, (uint64_t)plan.passes * RADIX * 16);
        digit_start_kernel<<<1, RADIX, 0, dev->stream>>>(ghist, plan.passes, n, n_dev, pass_ctl);
    }

    B200RS_TRY(b200rs_kernel_setup(dev, var.kernel, var.smem));
    // L2 prefetch distance of the scatter pass, in tiles.  Tickets are handed out at ~40 tiles per microsecond, so 64 tiles
    // ahead is ~1.5 us, longer than an HBM round trip; measured flat from 32 to 222 tiles (pairs 0.992 ms/pass against
    // 1.046 without), worse again from ~600 tiles on (the prefetched lines are evicted before use: 1.11 ms at 888).
    // B200RS_PF_TILES overrides (development knob; 0 switches the prefetch off).
    uint32_t pf_tiles = (uint32_t)b200rs_exp_env("B200RS_PF_TILES", var.gen == 2 ? 64 : 0);
    ElemT* src = inout;
    ElemT* dst = alt;
    for (int p = 0; p < plan.passes; ++p) {
        int shift = p * RADIX_BITS;
        const int width = sort_bits - shift < RADIX_BITS ? sort_bits - shift : RADIX_BITS;
        uint32_t digit_mask = (1u << width) - 1u;
        const unsigned long long* ghist_pass = ghist + (size_t)p * RADIX;
        uint32_t* ticket = tickets + p;
        uint32_t tag_base = (uint32_t)(2 * p);
        uint64_t n_arg = n;
        const uint8_t* no_lut = nullptr;
        uint32_t minus_one = 0xffffffffu;
        uint32_t pass = (uint32_t)p;
        void* args1[] = {&src, &dst, &n_arg, &shift, &digit_mask, &ghist_pass, &lookback, &ticket, &tag_base, &no_lut, &n_dev, &minus_one};
        void* args2[] = {&src, &dst, &n_arg, &shift, &digit_mask, &ghist_pass, &lb2, &ticket, &pass, &minus_one, &n_dev, &pass_ctl, &pf_tiles};  // ghist_pass: pre-scanned
        snprintf(label, sizeof(label), "onesweep_%s_pass%d", what, p);
        {
            b200rs_launch_scope scope(dev, label, n, 2ull * n * sizeof(ElemT));
            uint32_t grid = num_tiles;
            if (var.gen == 3 && grid > (uint32_t)(dev->num_sms * var.ctas_per_sm)) grid = (uint32_t)(dev->num_sms * var.ctas_per_sm);
            B200RS_CUDA(cudaLaunchKernel(var.kernel, dim3(grid), dim3(var.threads), var.gen >= 2 ? args2 : args1, var.smem, dev->stream));
        }
        ElemT* t = src; src = dst; dst = t;
    }
    // odd pass count: the result sits in the alternate buffer (the reference copies back too, Pprims.cpp:400-403)
    if (src != inout) B200RS_CUDA(cudaMemcpyAsync(inout, src, (size_t)n * sizeof(ElemT), cudaMemcpyDeviceToDevice, dev->stream));
    return B200RS_OK;
}

// ---- pieces of the multi-GPU sort (pairs) -----------------------------------------------------------

// Histogram of ONE digit (key >> shift) & mask, overwriting out[0..255].
template <typename ElemT>
__global__ void __launch_bounds__(HIST_THREADS)
single_digit_histogram_kernel(const ElemT* __restrict__ in, uint64_t n, int shift, uint32_t mask, unsigned long long* __restrict__ out) {
    __shared__ uint32_t s_hist[RADIX];
    for (int i = threadIdx.x; i < RADIX; i += HIST_THREADS) s_hist[i] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * HIST_THREADS;
    uint64_t i = (uint64_t)blockIdx.x * HIST_THREADS + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        uint32_t k[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) k[u] = Elem<ElemT>::key(in[i + u * stride]);
#pragma unroll
        for (int u = 0; u < 4; ++u) atomicAdd(&s_hist[(k[u] >> shift) & mask], 1u);
    }
    for (; i < n; i += stride) atomicAdd(&s_hist[(Elem<ElemT>::key(in[i]) >> shift) & mask], 1u);
    __syncthreads();
    for (int b = threadIdx.x; b < RADIX; b += HIST_THREADS)
        if (s_hist[b]) atomicAdd(&out[b], (unsigned long long)s_hist[b]);
}

constexpr int PART_THREADS = 384, PART_IPT = 16;  // same shape as the default pair scatter pass

}  // namespace

extern "C" int b200rs_digit_histogram_pairs(b200rs_device* dev, const b200rs_pair* in, uint64_t n, int shift, int bits, uint64_t* hist_out) {
    if (!dev || !hist_out || (n && !in) || shift < 0 || bits < 1 || bits > RADIX_BITS || shift + bits > 32) return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    B200RS_CUDA(cudaMemsetAsync(hist_out, 0, sizeof(uint64_t) * RADIX, dev->stream));
    if (n == 0) return B200RS_OK;
    uint64_t blocks = (n + (uint64_t)HIST_THREADS * 8 - 1) / ((uint64_t)HIST_THREADS * 8);
    if (blocks > (uint64_t)dev->num_sms * 4) blocks = (uint64_t)dev->num_sms * 4;
    {
        b200rs_launch_scope scope(dev, "single_digit_histogram_pairs", n, n * sizeof(uint2));
        single_digit_histogram_kernel<uint2><<<(unsigned)blocks, HIST_THREADS, 0, dev->stream>>>(
            reinterpret_cast<const uint2*>(in), n, shift, (1u << bits) - 1u, reinterpret_cast<unsigned long long*>(hist_out));
    }
    B200RS_CUDA(cudaGetLastError());
    return B200RS_OK;
}

extern "C" int b200rs_partition_pairs(b200rs_device* dev, const b200rs_pair* in, b200rs_pair* out, uint64_t n, int shift, int bits,
                                      const uint8_t* digit_to_part, const uint64_t* part_counts, void* temp, size_t* temp_bytes) {
    using Cfg = OnesweepConfig<uint2, PART_THREADS, PART_IPT>;
    if (!dev || !temp_bytes || shift < 0 || bits < 1 || bits > RADIX_BITS || shift + bits > 32) return B200RS_ERR_INVALID_ARGUMENT;
    const uint64_t tiles = (n + Cfg::TILE - 1) / Cfg::TILE;
    if (tiles > 0x7fffffffull) return B200RS_ERR_TOO_LARGE;
    const size_t need = 256 + b200rs_align_up((size_t)tiles * RADIX * sizeof(uint64_t), 256);
    if (!temp) {
        *temp_bytes = need;
        return B200RS_OK;
    }
    if (*temp_bytes < need) return B200RS_ERR_TEMP_TOO_SMALL;
    if (n == 0) return B200RS_OK;
    if (!in || !out || !digit_to_part || !part_counts || ((uintptr_t)temp & 255u) || (((uintptr_t)in | (uintptr_t)out) & 7u)) return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    B200RS_CUDA(cudaMemsetAsync(temp, 0, need, dev->stream));
    auto kernel = onesweep_kernel<uint2, PART_THREADS, PART_IPT, RANK_BALLOT, 3, true>;
    const size_t smem = sizeof(typename Cfg::Smem);
    B200RS_TRY(b200rs_kernel_setup(dev, (const void*)kernel, smem));
    uint32_t* ticket = reinterpret_cast<uint32_t*>(temp);
    uint64_t* lookback = reinterpret_cast<uint64_t*>(static_cast<char*>(temp) + 256);
    {
        b200rs_launch_scope scope(dev, "partition_pairs", n, 2ull * n * sizeof(uint2));
        kernel<<<(unsigned)tiles, PART_THREADS, smem, dev->stream>>>(reinterpret_cast<const uint2*>(in), reinterpret_cast<uint2*>(out), n, shift,
                                                                     (1u << bits) - 1u, reinterpret_cast<const unsigned long long*>(part_counts),
                                                                     lookback, ticket, 0u, digit_to_part, nullptr, 0xffffffffu);
    }
    B200RS_CUDA(cudaGetLastError());
    return B200RS_OK;
}

extern "C" int b200rs_scatter_pairs_to_parts(b200rs_device* dev, const b200rs_pair* in, uint64_t n, int shift, int bits,
                                             const uint8_t* digit_to_part, const uint64_t* part_base_addr, const uint64_t* n_dev,
                                             void* temp, size_t* temp_bytes) {
    using Cfg = OnesweepConfig<uint2, PART_THREADS, PART_IPT>;
    if (!dev || !temp_bytes || shift < 0 || bits < 1 || bits > RADIX_BITS || shift + bits > 32) return B200RS_ERR_INVALID_ARGUMENT;
    const uint64_t tiles = (n + Cfg::TILE - 1) / Cfg::TILE;
    if (tiles > 0x7fffffffull) return B200RS_ERR_TOO_LARGE;
    const size_t need = 256 + b200rs_align_up((size_t)tiles * RADIX * sizeof(uint64_t), 256);
    if (!temp) {
        *temp_bytes = need;
        return B200RS_OK;
    }
    if (*temp_bytes < need) return B200RS_ERR_TEMP_TOO_SMALL;
    if (n == 0) return B200RS_OK;
    if (!in || !digit_to_part || !part_base_addr || ((uintptr_t)temp & 255u) || ((uintptr_t)in & 7u)) return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    B200RS_CUDA(cudaMemsetAsync(temp, 0, need, dev->stream));
    auto kernel = onesweep_kernel<uint2, PART_THREADS, PART_IPT, RANK_BALLOT, 3, true, true>;
    const size_t smem = sizeof(typename Cfg::Smem);
    B200RS_TRY(b200rs_kernel_setup(dev, (const void*)kernel, smem));
    uint32_t* ticket = reinterpret_cast<uint32_t*>(temp);
    uint64_t* lookback = reinterpret_cast<uint64_t*>(static_cast<char*>(temp) + 256);
    {
        b200rs_launch_scope scope(dev, "scatter_pairs_to_parts", n, 2ull * n * sizeof(uint2));
        kernel<<<(unsigned)tiles, PART_THREADS, smem, dev->stream>>>(reinterpret_cast<const uint2*>(in), nullptr, n, shift, (1u << bits) - 1u,
                                                                     reinterpret_cast<const unsigned long long*>(part_base_addr), lookback, ticket,
                                                                     0u, digit_to_part, reinterpret_cast<const unsigned long long*>(n_dev), 0xffffffffu);
    }
    B200RS_CUDA(cudaGetLastError());
    return B200RS_OK;
}

namespace {

// ---- exchange kernel of the multi-GPU sort: stable partition into FEW parts + bulk stores into peer memory ------------
// Replaces the 256-digit scatter pass with a part table (onesweep_kernel<..., LUT, ABS>) for the exchange step: with at most
// 32 destinations a tile's run per destination is kilobytes long, so
//   * ranking needs no shared-memory atomics: the lanes holding the same part find each other with five ballots, and
//     lane p of every warp keeps the warp's running count of part p in a REGISTER (its own match mask comes from the same
//     five ballots); a lane fetches its group's base with one shuffle;
//   * the look-back table has one word per (tile, part) instead of 256 per tile;
//   * every run is staged at the same 16-byte phase as its destination and leaves the SM as ONE cp.async.bulk (TMA,
//     shared -> global) per part -- the destination may be another GPU's memory (CUDA IPC mapping, NVLink): the copy engine
//     streams it while the CTA's threads are already gone -- plus at most one 8-byte element at each ragged end.
// Same stability argument as the scatter passes: warp-striped load order, per-warp counts, ticket-ordered tiles.
// Shape: 7 worker warps x 32 lanes x 16 pairs (3584-pair tiles: 1.75 KB runs per destination at 16 parts) + 1 helper warp, 4 CTAs/SM.
//
// The helper warp (lane p = part p) owns everything that depends on other tiles, and does it WHILE the workers rank:
//   1. the workers load their pairs, look the part up and count it (one shared-memory RED per pair into the warp's row:
//      lanes with the same part are merged by the hardware) -- barrier;
//   2. helper: totals and the warps' first slots; the tile's counts go out at once (tile table).  Two-level look-back:
//      tiles form groups of XP_GROUP; a tile's exclusive prefix = [inclusive prefix of the previous group] + [counts of the
//      tiles before it in its own group] (independent loads, one round trip); the group's last tile also publishes the group
//      AGGREGATE at once; every tile walks the GROUP table backwards, XP_WINDOW groups = 32 tiles per step, adding aggregates until it
//      meets an INCLUSIVE word, and the group's last tile publishes its group's.  Then the layout of the staged tile (it depends on the 16-byte
//      phase of every destination).  Workers meanwhile: ranks (five ballots per pair) -- barrier;
//   3. workers stage, barrier, one bulk copy per part.
// History (2^28 pairs into 16 local parts, one B200; profiles/r2p_*, r2q_*, r2r_*): the first version counted inside the
// ranking loop and let lane p of warp 0 walk a per-tile table 4 tiles per step after it: 1.62 ms -- 8.25 dependent L2 round
// trips per tile (tiles start every 25 ns; an inclusive word appears a whole look-back later than the partial one), ~5 us
// during which the CTA's other warps sat at the barrier: 37 % of a CTA's life, 42 % of all warp samples "barrier".  Four
// ballots instead of five for <= 16 parts and the L2 prefetch of a later tile: 1.40 ms.  A lane-per-predecessor window (32
// uncoalesced words per part and step, every warp polling): 3.5 ms, the polls alone saturate L2.  Two-level table walked after
// the ranking: 1.30 ms, still 6.8 steps per tile -- the wait is for the slowest of the preceding tiles to FINISH RANKING.  Hence
// the early counts and the helper warp: 1.19 ms (the helper's wait for the slowest of the preceding ~16 tiles to finish COUNTING plus
// 3.8 steps of the group walk still outlast the ranking; 29 % of the warp samples are workers waiting for it).
constexpr int XP_IPT = 16, XP_MAX_PARTS = 32, XP_GROUP = 8, XP_WINDOW = 4;  // look-back: tiles per group, group words per step
constexpr uint32_t XP_VALID = 0x80000000u;  // tile table: bit 31 marks a published count
template <int XP_WORKERS>
struct XpSmem {
    static constexpr int XP_TILE = XP_WORKERS * 32 * XP_IPT;
    alignas(128) uint2 staged[XP_TILE + 2 * XP_MAX_PARTS];  // every run may start one element late and end one early
    uint32_t warp_base[XP_WORKERS][XP_MAX_PARTS];           // warp counts, then the warp's first slot inside the part's run
    uint32_t region[XP_MAX_PARTS];                          // staged slot of the part's first element
    uint32_t count[XP_MAX_PARTS];
    unsigned long long dst[XP_MAX_PARTS];                   // byte address of the part's run in its destination
    uint8_t lut[RADIX];
    unsigned long long splitter[XP_MAX_PARTS];
    uint32_t tile;
    unsigned long long n_eff;
};

// The helper warp's work (lane p = part p).  Inlined, groups of 8 tiles and 4 group words per step are what fits next to the
// workers' XP_IPT pairs per thread without spills: 1.19 ms at 2^28 pairs into 16 local parts; groups of 16 with 8 words per
// step 1.47 ms (64 B of spills in the workers' path), 8 words per step as a __noinline__ function 1.61 ms (the call spills).
template <int XP_WORKERS>
__device__ __forceinline__ void exchange_helper(XpSmem<XP_WORKERS>& s, int lane, int parts, uint32_t tile, uint32_t* lb_tile, uint64_t* lb_group,
                                             const unsigned long long* __restrict__ part_base) {
    uint32_t total = 0;
    if (lane < parts) {
#pragma unroll
        for (int w = 0; w < XP_WORKERS; ++w) {
            const uint32_t c = s.warp_base[w][lane];
            s.warp_base[w][lane] = total;
            total += c;
        }
    }
    unsigned long long exclusive = 0;
    if (lane < parts) {
        const uint32_t q = tile % XP_GROUP, g = tile / XP_GROUP;
        uint32_t* row = lb_tile + (uint64_t)tile * XP_MAX_PARTS + lane;
        uint64_t* grow = lb_group + (uint64_t)g * XP_MAX_PARTS + lane;
        st_relaxed_u32(row, XP_VALID | total);
        // chain: the first window goes out before the group mates are collected (both round trips overlap)
        int32_t ahead = (int32_t)g;  // groups before mine
        const uint64_t* p = grow - XP_MAX_PARTS;
        uint64_t w[XP_WINDOW];
#pragma unroll
        for (int j = 0; j < XP_WINDOW; ++j) w[j] = j < ahead ? ld_relaxed_u64(p - j * XP_MAX_PARTS) : (2ull << LB_TAG_SHIFT);  // below group 0: inclusive 0
        // counts of the tiles before this one in its group
        uint32_t mates = 0, got = 0;  // bit i of got: the count of the tile i + 1 places back has been added
        const uint32_t all = (1u << q) - 1u;
        while (got != all) {
            uint32_t x[XP_GROUP - 1];
#pragma unroll
            for (int i = 0; i < XP_GROUP - 1; ++i)
                x[i] = ((all & ~got) >> i) & 1u ? ld_relaxed_u32(row - (i + 1) * XP_MAX_PARTS) : 0u;
#pragma unroll
            for (int i = 0; i < XP_GROUP - 1; ++i)
                if (x[i] & XP_VALID) { mates += x[i] & ~XP_VALID; got |= 1u << i; }
        }
        if (q == XP_GROUP - 1) st_relaxed_u64(grow, (1ull << LB_TAG_SHIFT) | (uint64_t)(mates + total));
        bool done = false;
        for (;;) {
            int consumed = 0;
#pragma unroll
            for (int j = 0; j < XP_WINDOW; ++j) {
                const uint32_t tag = (uint32_t)(w[j] >> LB_TAG_SHIFT);
                if (!done && consumed == j && tag != 0) {
                    exclusive += w[j] & LB_VALUE_MASK;
                    consumed = j + 1;
                    done = tag == 2;
                }
            }
            if (done) break;
            p -= consumed * XP_MAX_PARTS;
            ahead -= consumed;
#pragma unroll
            for (int j = 0; j < XP_WINDOW; ++j) w[j] = j < ahead ? ld_relaxed_u64(p - j * XP_MAX_PARTS) : (2ull << LB_TAG_SHIFT);
        }
        exclusive += mates;
        if (q == XP_GROUP - 1) st_relaxed_u64(grow, (2ull << LB_TAG_SHIFT) | (exclusive + total));
    }
    const unsigned long long dst = lane < parts ? part_base[lane] + 8ull * exclusive : 0ull;
    const uint32_t phase = (uint32_t)(dst >> 3) & 1u;            // the run starts in the upper half of a 16-byte chunk
    const uint32_t padded = (phase + total + 1u) & ~1u;           // slots the run occupies, a whole number of chunks
    uint32_t inc = padded;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += y;
    }
    s.region[lane] = inc - padded + phase;
    s.count[lane] = total;
    s.dst[lane] = dst;
}

// SPLIT: the part of a pair is the number of thresholds (parts - 1 ascending 64-bit values, `splitters`) its key reaches,
// instead of a table lookup on one digit: destinations then own arbitrary key ranges (exact quantiles, dist.py's splitter plan).
template <int XP_WORKERS, int PART_BITS /* 4: at most 16 parts, 5: at most 32 */, bool BULK, bool SPLIT>
__global__ void __launch_bounds__((XP_WORKERS + 1) * 32, 1024 / ((XP_WORKERS + 1) * 32))
exchange_partition_kernel(const uint2* __restrict__ in, uint64_t n, int shift, uint32_t digit_mask, const uint8_t* __restrict__ digit_lut,
                          const unsigned long long* __restrict__ splitters,
                          const unsigned long long* __restrict__ part_base /*[parts] byte addresses, 8-byte aligned*/, int parts,
                          uint32_t* lb_tile /*[tiles][XP_MAX_PARTS] counts, zeroed*/, uint64_t* lb_group /*[tiles / XP_GROUP][XP_MAX_PARTS] tagged words, zeroed*/, uint32_t* ticket,
                          const unsigned long long* __restrict__ n_dev, uint32_t pf_tiles /* L2 prefetch distance in tiles, 0 = off */) {
    constexpr int XP_THREADS = (XP_WORKERS + 1) * 32, XP_TILE = XP_WORKERS * 32 * XP_IPT;
    extern __shared__ __align__(128) unsigned char xp_smem_raw[];
    XpSmem<XP_WORKERS>& s = *reinterpret_cast<XpSmem<XP_WORKERS>*>(xp_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool worker = warp < XP_WORKERS;
    if (tid == 0) {
        const uint32_t t = atomicAdd(ticket, 1u);
        const unsigned long long n_eff = n_dev ? min((unsigned long long)n, *n_dev) : (unsigned long long)n;
        s.tile = t;
        s.n_eff = n_eff;
        // the CTA that takes over this CTA's slot gets a ticket about pf_tiles higher: pull that tile into L2 now
        // (the prefetch starts at the 16-byte boundary below the tile and covers all of it but its last pair at most)
        if (pf_tiles) {
            const uint64_t pf_base = ((uint64_t)t + pf_tiles) * XP_TILE;
            if (pf_base + XP_TILE <= n_eff)
                bulk_prefetch_l2(reinterpret_cast<const void*>(reinterpret_cast<uintptr_t>(in + pf_base) & ~(uintptr_t)15), XP_TILE * (uint32_t)sizeof(uint2));
        }
    }
    if (!SPLIT)
        for (int i = tid; i < RADIX; i += XP_THREADS) s.lut[i] = digit_lut[i];
    if (SPLIT && tid < XP_MAX_PARTS) s.splitter[tid] = tid < parts - 1 ? splitters[tid] : ~0ull;
    for (int i = tid; i < XP_WORKERS * XP_MAX_PARTS; i += XP_THREADS) (&s.warp_base[0][0])[i] = 0;
    __syncthreads();
    n = s.n_eff;
    const uint32_t tile = s.tile;
    const uint64_t tile_base = (uint64_t)tile * XP_TILE;
    if (tile_base >= n) return;  // surplus CTAs of a grid sized for the upper bound
    const uint32_t valid = (uint32_t)min((uint64_t)XP_TILE, n - tile_base);
    const uint32_t slice = warp * (32 * XP_IPT) + lane;

    // ---- workers: load, part, early count ----
    uint2 elem[XP_IPT];
    uint32_t rec[XP_IPT / 2];  // per item: part | rank << 5, two per register
#pragma unroll
    for (int i = 0; i < XP_IPT / 2; ++i) rec[i] = 0;
    if (worker) {
        const uint2* __restrict__ src = in + tile_base + slice;
#pragma unroll
        for (int i = 0; i < XP_IPT; ++i)
            if (slice + i * 32 < valid) elem[i] = __ldg(src + i * 32);
        const uint32_t row = smem_addr(&s.warp_base[warp][0]);
#pragma unroll
        for (int i = 0; i < XP_IPT; ++i) {
            if (slice + i * 32 < valid) {
                uint32_t part = 0;
                if (SPLIT) {
                    for (int j = 0; j < parts - 1; ++j) part += (unsigned long long)elem[i].x >= s.splitter[j] ? 1u : 0u;
                } else {
                    part = (uint32_t)s.lut[(elem[i].x >> shift) & digit_mask];
                }
                red_add_shared(row + 4u * part, 1u);
                rec[i >> 1] |= part << (16 * (i & 1));
            }
        }
    }
    __syncthreads();

    if (!worker) {
        exchange_helper<XP_WORKERS>(s, lane, parts, tile, lb_tile, lb_group, part_base);
    } else {
        // ---- workers: ranking in registers ----
        const uint32_t lt = lanemask_lt();
        uint32_t xk[PART_BITS];  // lane p's selector: all-ones where bit k of p is CLEAR (its match mask takes the complement of that ballot)
#pragma unroll
        for (int k = 0; k < PART_BITS; ++k) xk[k] = ((lane >> k) & 1) ? 0u : 0xffffffffu;
        uint32_t cnt = 0;  // lane p: pairs of part p seen so far by this warp
#pragma unroll
        for (int i = 0; i < XP_IPT; ++i) {
            const bool live = slice + i * 32 < valid;
            const uint32_t part = (rec[i >> 1] >> (16 * (i & 1))) & 31u;
            const uint32_t live_lanes = __ballot_sync(0xffffffffu, live);
            uint32_t peers = live_lanes, owner = live_lanes;
#pragma unroll
            for (int k = 0; k < PART_BITS; ++k) {
                const uint32_t b = __ballot_sync(0xffffffffu, (part >> k) & 1u);
                peers &= ((part >> k) & 1u) ? b : ~b;   // lanes whose bit k equals mine
                owner &= b ^ xk[k];                      // lanes whose bit k equals bit k of MY LANE NUMBER
            }
            const uint32_t base = __shfl_sync(0xffffffffu, cnt, (int)part);  // count of my part before this item
            cnt += (uint32_t)__popc(owner);
            const uint32_t r = base + (uint32_t)__popc(peers & lt);
            rec[i >> 1] |= (r << 5) << (16 * (i & 1));
        }
    }
    __syncthreads();

    // ---- every pair to its staged slot ----
    const uint32_t staged = smem_addr(&s.staged[0]);
    if (worker) {
#pragma unroll
        for (int i = 0; i < XP_IPT; ++i) {
            if (slice + i * 32 < valid) {
                const uint32_t pr = (rec[i >> 1] >> (16 * (i & 1))) & 0xffffu;
                const uint32_t part = pr & 31u, r = pr >> 5;
                st_shared(staged + 8u * (s.region[part] + s.warp_base[warp][part] + r), elem[i]);
            }
        }
    }
    if (BULK) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the staged tile is read by the async proxy below
    __syncthreads();

    // ---- write-out ----
    if (BULK) {
        if (!worker && lane < parts) {
            uint32_t a = s.region[lane], c = s.count[lane];
            unsigned long long g = s.dst[lane];
            if (c && (a & 1u)) {  // ragged head: one element up to the 16-byte boundary
                *reinterpret_cast<uint2*>(g) = s.staged[a];
                ++a; --c; g += 8;
            }
            const uint32_t body = c & ~1u;
            if (body) bulk_copy_s2g(reinterpret_cast<void*>(g), staged + 8u * a, body * 8u);
            if (c & 1u) *reinterpret_cast<uint2*>(g + 8ull * body) = s.staged[a + body];
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must outlive the copy's reads
        }
    } else {
        for (int p = 0; p < parts; ++p) {
            const uint32_t a = s.region[p], c = s.count[p];
            uint2* g = reinterpret_cast<uint2*>(s.dst[p]);
            for (uint32_t j = tid; j < c; j += XP_THREADS) g[j] = s.staged[a + j];
        }
    }
}

// ---- histograms of the local sort, taken on the SENDING side while the exchange runs ---------------------------------
// The receiver's local sort starts with a histogram of all four digits of what it received (0.33 ms for 2^28 pairs).  The
// sender reads every pair it ships anyway and the exchange kernel is bound by NVLink, not by the SM: this kernel runs next
// to it (second stream) and counts, per destination, digits 0..2 of the pairs going there.  The tables are all-gathered
// (world x parts x 3 x 256 counters) and every rank adds up its own column; digit 3 comes from the top-digit histograms
// that were gathered for the plan.
constexpr int DH_THREADS = 512, DH_DIGITS = 3;
__global__ void __launch_bounds__(DH_THREADS)
dest_digit_histogram_kernel(const uint2* __restrict__ in, uint64_t n, const unsigned long long* __restrict__ n_dev, const uint8_t* __restrict__ digit_lut,
                            int parts, unsigned long long* __restrict__ out /*[parts][DH_DIGITS][RADIX], zeroed*/) {
    extern __shared__ uint32_t dh_hist[];  // [parts][DH_DIGITS][RADIX]
    __shared__ uint8_t s_lut[RADIX];
    if (n_dev) n = min(n, (uint64_t)*n_dev);
    const int cells = parts * DH_DIGITS * RADIX;
    for (int i = threadIdx.x; i < cells; i += DH_THREADS) dh_hist[i] = 0;
    if (threadIdx.x < RADIX) s_lut[threadIdx.x] = digit_lut[threadIdx.x];
    __syncthreads();
    const uint32_t table = smem_addr(dh_hist);
    const uint64_t nvec = n / 2;  // two pairs per 128-bit load
    const uint4* in4 = reinterpret_cast<const uint4*>(in);
    const uint64_t stride = (uint64_t)gridDim.x * DH_THREADS;
    auto count = [&](uint32_t key) {
        const uint32_t base = table + (uint32_t)s_lut[key >> 24] * (DH_DIGITS * RADIX * 4u);
        red_add_shared(base + 4u * (key & 255u), 1u);
        red_add_shared(base + 4u * (RADIX + ((key >> 8) & 255u)), 1u);
        red_add_shared(base + 4u * (2 * RADIX + ((key >> 16) & 255u)), 1u);
    };
    uint64_t v = (uint64_t)blockIdx.x * DH_THREADS + threadIdx.x;
    for (; v + stride < nvec; v += 2 * stride) {
        const uint4 a = __ldg(in4 + v), b = __ldg(in4 + v + stride);
        count(a.x); count(a.z); count(b.x); count(b.z);
    }
    for (; v < nvec; v += stride) {
        const uint4 a = __ldg(in4 + v);
        count(a.x); count(a.z);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && (n & 1)) count(in[n - 1].x);
    __syncthreads();
    for (int i = threadIdx.x; i < cells; i += DH_THREADS)
        if (dh_hist[i]) atomicAdd(&out[i], (unsigned long long)dh_hist[i]);
}

// ghist[p][d] of the local sort (4 x 256 u64): digits 0..2 = sum over sources of their tables' column `me`; digit 3 = the
// gathered top-digit histograms, restricted to the digits this rank owns.
__global__ void __launch_bounds__(RADIX)
recv_histogram_kernel(const unsigned long long* __restrict__ tables /*[world][parts][DH_DIGITS][RADIX]*/, const unsigned long long* __restrict__ top_all /*[world][RADIX]*/,
                      const uint8_t* __restrict__ digit_lut, int world, int me, const uint32_t* __restrict__ status, unsigned long long* __restrict__ ghist /*[4][RADIX]*/) {
    const int d = threadIdx.x;
    const bool aborted = status[0] != 0;
    for (int p = 0; p < DH_DIGITS; ++p) {
        unsigned long long sum = 0;
        for (int s = 0; s < world; ++s) sum += tables[(((size_t)s * world + me) * DH_DIGITS + p) * RADIX + d];
        ghist[p * RADIX + d] = aborted ? 0ull : sum;
    }
    unsigned long long top = 0;
    if (digit_lut[d] == me)
        for (int s = 0; s < world; ++s) top += top_all[(size_t)s * RADIX + d];
    ghist[DH_DIGITS * RADIX + d] = aborted ? 0ull : top;
}

// ---- on-device exchange plan (one CTA of 256 threads): keeps the multi-GPU sort free of host round trips ----
// hist_all[s][b] = pairs on source rank s with top digit b.  Digit ranges are contiguous per destination; the edge
// between rank r-1 and r is the digit boundary whose cumulative count is closest to r*N/P (exact integer compare,
// ties to the lower digit) -- the same rule as plan_exchange() in oclradixsort_b200/dist.py.
__global__ void __launch_bounds__(RADIX)
dist_plan_kernel(const unsigned long long* __restrict__ hist_all, int P, int me, const unsigned long long* __restrict__ peer_base,
                 unsigned long long capacity, int n_in_valid, uint8_t* __restrict__ lut_out, unsigned long long* __restrict__ part_base_out,
                 unsigned long long* __restrict__ counts_out /*[0]=n to scatter (0 when aborted), [1]=pairs this rank receives*/,
                 uint32_t* __restrict__ status_out, unsigned long long n_in) {
    __shared__ unsigned long long s_cum[RADIX + 1];
    __shared__ unsigned long long s_scan[RADIX / 32];
    __shared__ int s_edge[33];
    __shared__ unsigned long long s_best[RADIX / 32];
    __shared__ int s_best_b[RADIX / 32];
    const int b = threadIdx.x, lane = b & 31, warp = b >> 5;
    unsigned long long total = 0;
    for (int s = 0; s < P; ++s) total += hist_all[(size_t)s * RADIX + b];
    const unsigned long long excl = block_exclusive_scan_256<unsigned long long>(total, s_scan, b);
    s_cum[b] = excl;
    if (b == RADIX - 1) s_cum[RADIX] = excl + total;
    if (b == 0) { s_edge[0] = 0; s_edge[P] = RADIX; }
    __syncthreads();
    const unsigned long long N = s_cum[RADIX];
    for (int r = 1; r < P; ++r) {
        // candidate boundaries 0..256: thread b evaluates b (and thread 255 also 256)
        auto dist_to = [&](int edge) {
            const unsigned long long a = s_cum[edge] * (unsigned long long)P, t = (unsigned long long)r * N;
            return a > t ? a - t : t - a;
        };
        unsigned long long best = dist_to(b);
        int best_b = b;
        if (b == RADIX - 1 && dist_to(RADIX) < best) { best = dist_to(RADIX); best_b = RADIX; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long ob = __shfl_down_sync(0xffffffffu, best, o);
            const int obb = __shfl_down_sync(0xffffffffu, best_b, o);
            if (ob < best || (ob == best && obb < best_b)) { best = ob; best_b = obb; }
        }
        if (lane == 0) { s_best[warp] = best; s_best_b[warp] = best_b; }
        __syncthreads();
        if (b == 0) {
            for (int w = 1; w < RADIX / 32; ++w)
                if (s_best[w] < best || (s_best[w] == best && s_best_b[w] < best_b)) { best = s_best[w]; best_b = s_best_b[w]; }
            s_edge[r] = max(best_b, s_edge[r - 1]);
        }
        __syncthreads();
    }
    // owner of digit b
    int owner = 0;
    for (int r = 1; r < P; ++r) owner += (s_edge[r] <= b) ? 1 : 0;
    lut_out[b] = (uint8_t)owner;
    // part d (< P): pairs every source sends to d, my offset inside d's buffer, d's total
    if (b < P) {
        const int d = b;
        unsigned long long before_me = 0, all = 0;
        for (int s = 0; s < P; ++s) {
            unsigned long long c = 0;
            for (int k = s_edge[d]; k < s_edge[d + 1]; ++k) c += hist_all[(size_t)s * RADIX + k];
            if (s < me) before_me += c;
            all += c;
        }
        part_base_out[d] = peer_base[d] + 8ull * before_me;
        s_cum[d] = all;  // reuse: total received by d
    } else {
        part_base_out[b] = 0;
    }
    __syncthreads();
    if (b == 0) {
        bool ok = true;
        for (int d = 0; d < P; ++d) ok = ok && s_cum[d] <= capacity;
        status_out[0] = ok ? 0u : 1u;           // 1 = some rank's share exceeds its receive capacity: nothing is exchanged
        counts_out[0] = ok ? n_in : 0ull;
        counts_out[1] = ok ? s_cum[me] : 0ull;
    }
    (void)n_in_valid;
}

// ---- plan of the PIPELINED exchange (two halves per destination) --------------------------------------------------------
// Same digit ranges per destination as dist_plan_kernel; every destination's range is cut once more at the digit boundary
// closest to half of ITS pairs: half A = the lower digits, half B = the rest.  The exchange kernel then partitions into
// 2 x world parts: part 2d is stored straight into d's receive buffer (NVLink, as before), part 2d + 1 into a staging area
// in THIS GPU's memory, from where copy engines move it to d while the SMs already sort half A (own pairs of half B go
// straight to their final place).  d's receive buffer: [A: source 0, source 1, ... | B: source 0, source 1, ...] -- inside
// each half in (source rank, position) order, so the two stable local sorts leave exactly the one stable sort.
struct DistHalfPlan {
    unsigned long long status;                       // 1: some rank's share exceeds its receive capacity (nothing is exchanged)
    unsigned long long recv_total, recv_a, recv_b;   // pairs this rank receives
    unsigned long long stage_off[XP_MAX_PARTS / 2];  // pairs: start of destination d's half-B run in this rank's staging area
    unsigned long long stage_cnt[XP_MAX_PARTS / 2];  // pairs of half B this rank sends to d (0 for d == this rank)
    unsigned long long dst_addr[XP_MAX_PARTS / 2];   // byte address of that run in d's receive buffer
};
__global__ void __launch_bounds__(RADIX)
dist_plan_halves_kernel(const unsigned long long* __restrict__ hist_all, int P, int me, const unsigned long long* __restrict__ peer_base,
                        unsigned long long capacity, unsigned long long stage_base /* byte address of the staging area */,
                        uint8_t* __restrict__ lut_out /*[256] digit -> part (2 x destination + half)*/, unsigned long long* __restrict__ part_base_out /*[2P]*/,
                        unsigned long long* __restrict__ counts_out, uint32_t* __restrict__ status_out, unsigned long long n_in, DistHalfPlan* __restrict__ plan_out,
                        uint32_t a_permille /* share of a destination's pairs that half A should hold */) {
    __shared__ unsigned long long s_cum[RADIX + 1];
    __shared__ unsigned long long s_scan[RADIX / 32];
    __shared__ int s_edge[XP_MAX_PARTS / 2 + 1];
    __shared__ int s_mid[XP_MAX_PARTS / 2];
    __shared__ unsigned long long s_best[RADIX / 32];
    __shared__ int s_best_b[RADIX / 32];
    __shared__ unsigned long long s_all[XP_MAX_PARTS], s_before[XP_MAX_PARTS], s_mine[XP_MAX_PARTS];
    const int b = threadIdx.x, lane = b & 31, warp = b >> 5;
    unsigned long long total = 0;
    for (int s = 0; s < P; ++s) total += hist_all[(size_t)s * RADIX + b];
    const unsigned long long excl = block_exclusive_scan_256<unsigned long long>(total, s_scan, b);
    s_cum[b] = excl;
    if (b == RADIX - 1) s_cum[RADIX] = excl + total;
    if (b == 0) { s_edge[0] = 0; s_edge[P] = RADIX; }
    __syncthreads();
    const unsigned long long N = s_cum[RADIX];
    for (int r = 1; r < P; ++r) {  // (the rule of dist_plan_kernel: boundary whose cumulative count is closest to r N / P, ties to the lower digit)
        auto dist_to = [&](int edge) {
            const unsigned long long a = s_cum[edge] * (unsigned long long)P, t = (unsigned long long)r * N;
            return a > t ? a - t : t - a;
        };
        unsigned long long best = dist_to(b);
        int best_b = b;
        if (b == RADIX - 1 && dist_to(RADIX) < best) { best = dist_to(RADIX); best_b = RADIX; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long ob = __shfl_down_sync(0xffffffffu, best, o);
            const int obb = __shfl_down_sync(0xffffffffu, best_b, o);
            if (ob < best || (ob == best && obb < best_b)) { best = ob; best_b = obb; }
        }
        if (lane == 0) { s_best[warp] = best; s_best_b[warp] = best_b; }
        __syncthreads();
        if (b == 0) {
            for (int w = 1; w < RADIX / 32; ++w)
                if (s_best[w] < best || (s_best[w] == best && s_best_b[w] < best_b)) { best = s_best[w]; best_b = s_best_b[w]; }
            s_edge[r] = max(best_b, s_edge[r - 1]);
        }
        __syncthreads();
    }
    if (b < P) {  // the cut inside destination b's range
        const int lo = s_edge[b], hi = s_edge[b + 1];
        const unsigned long long tot = s_cum[hi] - s_cum[lo];
        int best_k = lo;
        const unsigned long long want = (unsigned long long)a_permille * tot;
        unsigned long long best = want;  // |1000 * 0 - want|
        for (int k = lo + 1; k <= hi; ++k) {
            const unsigned long long a = 1000ull * (s_cum[k] - s_cum[lo]);
            const unsigned long long dk = a > want ? a - want : want - a;
            if (dk < best) { best = dk; best_k = k; }
        }
        s_mid[b] = best_k;
    }
    __syncthreads();
    int owner = 0;
    for (int r = 1; r < P; ++r) owner += (s_edge[r] <= b) ? 1 : 0;
    lut_out[b] = (uint8_t)(2 * owner + (b >= s_mid[owner] ? 1 : 0));
    if (b < 2 * P) {  // part b: pairs every source sends, those of lower ranks, mine
        const int d = b >> 1;
        const int lo = (b & 1) ? s_mid[d] : s_edge[d], hi = (b & 1) ? s_edge[d + 1] : s_mid[d];
        unsigned long long before_me = 0, all = 0, mine = 0;
        for (int s = 0; s < P; ++s) {
            unsigned long long c = 0;
            for (int k = lo; k < hi; ++k) c += hist_all[(size_t)s * RADIX + k];
            if (s < me) before_me += c;
            if (s == me) mine = c;
            all += c;
        }
        s_all[b] = all; s_before[b] = before_me; s_mine[b] = mine;
    }
    __syncthreads();
    if (b == 0) {
        bool ok = true;
        for (int d = 0; d < P; ++d) ok = ok && s_all[2 * d] + s_all[2 * d + 1] <= capacity;
        unsigned long long running = 0;
        for (int d = 0; d < P; ++d) {
            const unsigned long long final_b = peer_base[d] + 8ull * (s_all[2 * d] + s_before[2 * d + 1]);
            part_base_out[2 * d] = peer_base[d] + 8ull * s_before[2 * d];
            if (d == me) {
                part_base_out[2 * d + 1] = final_b;
                plan_out->stage_off[d] = 0; plan_out->stage_cnt[d] = 0; plan_out->dst_addr[d] = final_b;
            } else {
                part_base_out[2 * d + 1] = stage_base + 8ull * running;
                plan_out->stage_off[d] = running; plan_out->stage_cnt[d] = s_mine[2 * d + 1]; plan_out->dst_addr[d] = final_b;
                running += (s_mine[2 * d + 1] + 15ull) & ~15ull;  // runs start on 128-byte boundaries of the staging area
            }
        }
        status_out[0] = ok ? 0u : 1u;
        counts_out[0] = ok ? n_in : 0ull;
        counts_out[1] = ok ? s_all[2 * me] + s_all[2 * me + 1] : 0ull;
        plan_out->status = ok ? 0ull : 1ull;
        plan_out->recv_a = s_all[2 * me];
        plan_out->recv_b = s_all[2 * me + 1];
        plan_out->recv_total = s_all[2 * me] + s_all[2 * me + 1];
    }
}

}  // namespace

namespace {
int exchange_impl(b200rs_device* dev, const b200rs_pair* in, uint64_t n, int shift, int bits, const uint8_t* digit_to_part, const uint64_t* splitters,
                  const uint64_t* part_base_addr, int parts, const uint64_t* n_dev, void* temp, size_t* temp_bytes) {
    if (!dev || !temp_bytes || parts < 1 || parts > XP_MAX_PARTS) return B200RS_ERR_INVALID_ARGUMENT;
    int workers = 7;
#ifdef B200RS_EXPERIMENTS
    workers = b200rs_exp_env("B200RS_XP_THREADS", 256) == 128 ? 3 : 7;
#endif
    const int threads = (workers + 1) * 32;
    // (the size query must not depend on the shape: the tables are sized for the smaller tile)
    const uint64_t tile = (uint64_t)workers * 32 * XP_IPT, tiles = (n + tile - 1) / tile, tiles_max = (n + 3 * 32 * XP_IPT - 1) / (3 * 32 * XP_IPT);
    if (tiles_max > 0x7fffffffull) return B200RS_ERR_TOO_LARGE;
    const size_t tile_table = b200rs_align_up((size_t)tiles_max * XP_MAX_PARTS * sizeof(uint32_t), 256);
    const size_t need = 256 + tile_table + b200rs_align_up((size_t)(tiles_max / XP_GROUP + 1) * XP_MAX_PARTS * sizeof(uint64_t), 256);
    if (!temp) {
        *temp_bytes = need;
        return B200RS_OK;
    }
    if (*temp_bytes < need) return B200RS_ERR_TEMP_TOO_SMALL;
    if (n == 0) return B200RS_OK;
    if (!in || (!digit_to_part && !splitters) || !part_base_addr || ((uintptr_t)temp & 255u) || ((uintptr_t)in & 7u)) return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    B200RS_CUDA(cudaMemsetAsync(temp, 0, 256 + (size_t)tiles * XP_MAX_PARTS * sizeof(uint32_t), dev->stream));
    B200RS_CUDA(cudaMemsetAsync(static_cast<char*>(temp) + 256 + tile_table, 0, (size_t)(tiles / XP_GROUP + 1) * XP_MAX_PARTS * sizeof(uint64_t), dev->stream));
    const bool split = splitters != nullptr, wide = parts > 16;
    const void* kernel = nullptr;
#ifdef B200RS_EXPERIMENTS  // element-wise write-out instead of the bulk copies: measurement only (8 GPUs: 3.20 against 2.88 ms for the exchange)
    const bool bulk = !b200rs_exp_env("B200RS_XP_NO_BULK", 0);
#define B200RS_XP_PICK(T) \
    kernel = split ? (wide ? (bulk ? (const void*)exchange_partition_kernel<T, 5, true, true> : (const void*)exchange_partition_kernel<T, 5, false, true>)   \
                           : (bulk ? (const void*)exchange_partition_kernel<T, 4, true, true> : (const void*)exchange_partition_kernel<T, 4, false, true>))  \
                   : (wide ? (bulk ? (const void*)exchange_partition_kernel<T, 5, true, false> : (const void*)exchange_partition_kernel<T, 5, false, false>) \
                           : (bulk ? (const void*)exchange_partition_kernel<T, 4, true, false> : (const void*)exchange_partition_kernel<T, 4, false, false>))
#else
#define B200RS_XP_PICK(T) \
    kernel = split ? (wide ? (const void*)exchange_partition_kernel<T, 5, true, true> : (const void*)exchange_partition_kernel<T, 4, true, true>)   \
                   : (wide ? (const void*)exchange_partition_kernel<T, 5, true, false> : (const void*)exchange_partition_kernel<T, 4, true, false>)
#endif
    size_t smem = 0;
#ifdef B200RS_EXPERIMENTS
    if (workers == 3) { B200RS_XP_PICK(3); smem = sizeof(XpSmem<3>); }
#endif
    if (!kernel) { B200RS_XP_PICK(7); smem = sizeof(XpSmem<7>); }
#undef B200RS_XP_PICK
    B200RS_TRY(b200rs_kernel_setup(dev, kernel, smem));
    uint32_t* ticket = reinterpret_cast<uint32_t*>(temp);
    uint32_t* lb_tile = reinterpret_cast<uint32_t*>(static_cast<char*>(temp) + 256);
    uint64_t* lb_group = reinterpret_cast<uint64_t*>(static_cast<char*>(temp) + 256 + tile_table);
    {
        b200rs_launch_scope scope(dev, splitters ? "exchange_pairs_splitters" : "exchange_pairs", n, 2ull * n * sizeof(uint2));
        const uint2* in2 = reinterpret_cast<const uint2*>(in);
        uint64_t n64 = n;
        uint32_t mask = (1u << bits) - 1u;
        // L2 prefetch distance: tickets are handed out at ~40-80 per microsecond, a tile about 1.5 us ahead
        uint32_t pf_tiles = (uint32_t)b200rs_exp_env("B200RS_XP_PF", 64);
        void* args[] = {(void*)&in2, (void*)&n64, (void*)&shift, (void*)&mask, (void*)&digit_to_part, (void*)&splitters, (void*)&part_base_addr, (void*)&parts,
                        (void*)&lb_tile, (void*)&lb_group, (void*)&ticket, (void*)&n_dev, (void*)&pf_tiles};
        B200RS_CUDA(cudaLaunchKernel(kernel, dim3((unsigned)tiles), dim3((unsigned)threads), args, smem, dev->stream));
    }
    B200RS_CUDA(cudaGetLastError());
    return B200RS_OK;
}

// Histograms of one key digit restricted to keys with given higher bits: hist_out[j][d] = pairs whose bits above the digit
// equal prefixes[j] and whose digit is d (the splitter plan refines its boundaries with these, one digit per round).
constexpr int FH_MAX = 31;
__global__ void __launch_bounds__(HIST_THREADS)
filtered_histogram_kernel(const uint2* __restrict__ in, uint64_t n, int shift, const uint32_t* __restrict__ prefixes, int count,
                          unsigned long long* __restrict__ out /*[count][RADIX]*/) {
    extern __shared__ uint32_t fh_hist[];  // [count][RADIX]
    __shared__ uint32_t s_prefix[FH_MAX];
    for (int i = threadIdx.x; i < count * RADIX; i += HIST_THREADS) fh_hist[i] = 0;
    if (threadIdx.x < count) s_prefix[threadIdx.x] = prefixes[threadIdx.x];
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * HIST_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * HIST_THREADS + threadIdx.x; i < n; i += stride) {
        const uint32_t key = in[i].x;
        const uint32_t high = shift + RADIX_BITS >= 32 ? 0u : key >> (shift + RADIX_BITS), d = (key >> shift) & (RADIX - 1);
        for (int j = 0; j < count; ++j)
            if (high == s_prefix[j]) atomicAdd(&fh_hist[j * RADIX + d], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < count * RADIX; i += HIST_THREADS)
        if (fh_hist[i]) atomicAdd(&out[i], (unsigned long long)fh_hist[i]);
}
}  // namespace

extern "C" int b200rs_exchange_pairs(b200rs_device* dev, const b200rs_pair* in, uint64_t n, int shift, int bits, const uint8_t* digit_to_part,
                                     const uint64_t* part_base_addr, int parts, const uint64_t* n_dev, void* temp, size_t* temp_bytes) {
    if (shift < 0 || bits < 1 || bits > RADIX_BITS || shift + bits > 32) return B200RS_ERR_INVALID_ARGUMENT;
    return exchange_impl(dev, in, n, shift, bits, digit_to_part, nullptr, part_base_addr, parts, n_dev, temp, temp_bytes);
}

extern "C" int b200rs_exchange_pairs_by_splitters(b200rs_device* dev, const b200rs_pair* in, uint64_t n, const uint64_t* splitters,
                                                  const uint64_t* part_base_addr, int parts, void* temp, size_t* temp_bytes) {
    if (temp && !splitters && parts > 1) return B200RS_ERR_INVALID_ARGUMENT;
    static const uint64_t none = 0;
    (void)none;
    return exchange_impl(dev, in, n, 0, 8, nullptr, splitters ? splitters : part_base_addr /* parts == 1: never read */, part_base_addr, parts, nullptr, temp, temp_bytes);
}

extern "C" int b200rs_filtered_histograms_pairs(b200rs_device* dev, const b200rs_pair* in, uint64_t n, int shift, const uint32_t* prefixes, int count,
                                                uint64_t* hist_out) {
    if (!dev || !hist_out || !prefixes || (n && !in) || shift < 0 || shift > 24 || (shift & 7) || count < 1 || count > FH_MAX) return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    B200RS_CUDA(cudaMemsetAsync(hist_out, 0, sizeof(uint64_t) * RADIX * count, dev->stream));
    if (n == 0) return B200RS_OK;
    uint64_t blocks = (n + (uint64_t)HIST_THREADS * 8 - 1) / ((uint64_t)HIST_THREADS * 8);
    if (blocks > (uint64_t)dev->num_sms * 4) blocks = (uint64_t)dev->num_sms * 4;
    {
        b200rs_launch_scope scope(dev, "filtered_histograms_pairs", n, n * sizeof(uint2));
        filtered_histogram_kernel<<<(unsigned)blocks, HIST_THREADS, (size_t)count * RADIX * sizeof(uint32_t), dev->stream>>>(
            reinterpret_cast<const uint2*>(in), n, shift, prefixes, count, reinterpret_cast<unsigned long long*>(hist_out));
    }
    B200RS_CUDA(cudaGetLastError());
    return B200RS_OK;
}

extern "C" int b200rs_dist_plan(b200rs_device* dev, const uint64_t* hist_all, int world, int rank, const uint64_t* peer_base, uint64_t capacity,
                                uint64_t n_in, uint8_t* lut_out, uint64_t* part_base_out, uint64_t* counts_out, uint32_t* status_out) {
    if (!dev || !hist_all || !peer_base || !lut_out || !part_base_out || !counts_out || !status_out || world < 1 || world > 32 || rank < 0 || rank >= world)
        return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    {
        b200rs_launch_scope scope(dev, "dist_plan", (uint64_t)world * RADIX, (uint64_t)world * RADIX * 8);
        dist_plan_kernel<<<1, RADIX, 0, dev->stream>>>(reinterpret_cast<const unsigned long long*>(hist_all), world, rank,
                                                       reinterpret_cast<const unsigned long long*>(peer_base), capacity, 0, lut_out,
                                                       reinterpret_cast<unsigned long long*>(part_base_out),
                                                       reinterpret_cast<unsigned long long*>(counts_out), status_out, n_in);
    }
    B200RS_CUDA(cudaGetLastError());
    return B200RS_OK;
}

extern "C" int b200rs_dist_plan_halves(b200rs_device* dev, const uint64_t* hist_all, int world, int rank, const uint64_t* peer_base, uint64_t capacity,
                                       uint64_t stage_base, uint64_t n_in, int a_permille, uint8_t* lut_out, uint64_t* part_base_out, uint64_t* counts_out,
                                       uint32_t* status_out, uint64_t* plan_out) {
    static_assert(sizeof(DistHalfPlan) == (4 + 3 * (XP_MAX_PARTS / 2)) * sizeof(uint64_t), "plan_out layout documented in include/b200rs.h");
    if (!dev || !hist_all || !peer_base || !lut_out || !part_base_out || !counts_out || !status_out || !plan_out || world < 1 || world > XP_MAX_PARTS / 2 || rank < 0 ||
        rank >= world || a_permille < 0 || a_permille > 1000)
        return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    {
        b200rs_launch_scope scope(dev, "dist_plan_halves", (uint64_t)world * RADIX, (uint64_t)world * RADIX * 8);
        dist_plan_halves_kernel<<<1, RADIX, 0, dev->stream>>>(reinterpret_cast<const unsigned long long*>(hist_all), world, rank, reinterpret_cast<const unsigned long long*>(peer_base),
                                                              capacity, stage_base, lut_out, reinterpret_cast<unsigned long long*>(part_base_out),
                                                              reinterpret_cast<unsigned long long*>(counts_out), status_out, n_in, reinterpret_cast<DistHalfPlan*>(plan_out),
                                                              (uint32_t)a_permille);
    }
    B200RS_CUDA(cudaGetLastError());
    return B200RS_OK;
}

extern "C" int b200rs_sort_pairs_u32_devn(b200rs_device* dev, b200rs_pair* inout, uint64_t n_max, const uint64_t* n_dev, int sort_bits,
                                          void* temp, size_t* temp_bytes) {
    return sort_impl<uint2>(dev, reinterpret_cast<uint2*>(inout), n_max, sort_bits, temp, temp_bytes, "pairs",
                            reinterpret_cast<const unsigned long long*>(n_dev));
}

namespace {
int sort_pairs_devn_with_histogram(b200rs_device* dev, b200rs_pair* inout, uint64_t n_max, const uint64_t* n_dev, const uint64_t* hist4x256, void* temp,
                                   size_t* temp_bytes) {
    return sort_impl<uint2>(dev, reinterpret_cast<uint2*>(inout), n_max, 32, temp, temp_bytes, "pairs", reinterpret_cast<const unsigned long long*>(n_dev),
                            MSD_AUTO, nullptr, reinterpret_cast<const unsigned long long*>(hist4x256));
}
}  // namespace

namespace {
// The pipelined form of the partitioned sort (from DIST_PIPELINE_MIN_CAPACITY pairs of receive capacity and up to 16 ranks):
//   top-digit histogram -> allgather -> plan with two halves per destination (dist_plan_halves_kernel), read back by the host
//   (the one host round trip of this path: the sizes of the two local sorts and of the copies) -> ONE exchange kernel: half A of
//   every destination over NVLink, half B into the local staging area -> [copy engines: half B to the peers] || [barrier,
//   local sort of half A on a second stream] -> barrier -> local sort of half B -> join.
// The SMs sort while the copy engines keep NVLink busy; the unpipelined form leaves the SMs idle for the whole exchange.
constexpr uint64_t DIST_PIPELINE_MIN_CAPACITY = 1ull << 22;
constexpr uint64_t DIST_PIPELINE_SORT_EXTRA = 1ull << 21;  // the second sort's fixed temp overhead is covered by a plan for this many pairs
constexpr int DIST_PIPELINE_MAX_WORLD = XP_MAX_PARTS / 2;
constexpr int DIST_PIPELINE_A_PERMILLE = 500;  // half A's share of every destination's pairs

int dist_sort_pipelined(b200rs_device* dev, const b200rs_dist_comm* comm, const uint64_t* recv_base, uint64_t recv_capacity_pairs, const b200rs_pair* in, uint64_t n,
                        uint64_t* counts_dev, uint32_t* status_dev, void* temp, size_t* temp_bytes) {
    const int world = comm->world, rank = comm->rank;
    size_t xp_bytes = 0, sort_bytes = 0, extra_bytes = 0;
    B200RS_TRY(b200rs_exchange_pairs(dev, nullptr, n, 24, 8, nullptr, nullptr, 2 * world, nullptr, nullptr, &xp_bytes));
    B200RS_TRY(b200rs_sort_pairs_u32(dev, nullptr, recv_capacity_pairs, 32, nullptr, &sort_bytes));
    B200RS_TRY(b200rs_sort_pairs_u32(dev, nullptr, DIST_PIPELINE_SORT_EXTRA, 32, nullptr, &extra_bytes));
    const size_t sort_reserved = b200rs_align_up(sort_bytes, 256) + b200rs_align_up(extra_bytes, 256) + 512;
    const size_t stage_bytes = b200rs_align_up((size_t)(n + 16ull * world) * sizeof(b200rs_pair), 256);
    // temp: [own top histogram][gathered world x 256][peer bases 256][part bases 256][lut 256 B][plan][exchange temp][staging][sort temp A | sort temp B]
    const size_t hist_off = 0, gathered_off = hist_off + RADIX * 8, peers_off = gathered_off + (size_t)world * RADIX * 8, parts_off = peers_off + RADIX * 8,
                 lut_off = parts_off + RADIX * 8, plan_off = lut_off + 256, xp_off = plan_off + b200rs_align_up(sizeof(DistHalfPlan), 256),
                 stage_off = xp_off + b200rs_align_up(xp_bytes, 256), sort_off = stage_off + stage_bytes, need = sort_off + sort_reserved;
    if (!temp) {
        *temp_bytes = need;
        return B200RS_OK;
    }
    if (*temp_bytes < need) return B200RS_ERR_TEMP_TOO_SMALL;
    if (!recv_base || !counts_dev || !status_dev || (n && !in) || !comm->allgather || !comm->barrier || ((uintptr_t)temp & 255u)) return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    char* base = static_cast<char*>(temp);
    uint64_t* hist = reinterpret_cast<uint64_t*>(base + hist_off);
    uint64_t* gathered = reinterpret_cast<uint64_t*>(base + gathered_off);
    uint64_t* peers = reinterpret_cast<uint64_t*>(base + peers_off);
    uint64_t* part_base = reinterpret_cast<uint64_t*>(base + parts_off);
    uint8_t* lut = reinterpret_cast<uint8_t*>(base + lut_off);
    DistHalfPlan* plan_dev = reinterpret_cast<DistHalfPlan*>(base + plan_off);
    char* stage = base + stage_off;
    if (!dev->aux) {
        B200RS_CUDA(cudaStreamCreateWithFlags(&dev->aux, cudaStreamNonBlocking));
        B200RS_CUDA(cudaEventCreateWithFlags(&dev->ev_aux[0], cudaEventDisableTiming));
        B200RS_CUDA(cudaEventCreateWithFlags(&dev->ev_aux[1], cudaEventDisableTiming));
    }
    constexpr int COPY_STREAMS = 8;  // one copy per peer in flight (up to 8): 2 streams moved ~450 GB/s per GPU, see profiles/r2u_*
    if (!dev->copy[0]) {
        for (int i = 0; i < COPY_STREAMS; ++i) {
            B200RS_CUDA(cudaStreamCreateWithFlags(&dev->copy[i], cudaStreamNonBlocking));
            B200RS_CUDA(cudaEventCreateWithFlags(&dev->ev_copied[i], cudaEventDisableTiming));
        }
        for (cudaEvent_t& e : dev->ev_pipe) B200RS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        B200RS_CUDA(cudaHostAlloc(&dev->pinned_plan, sizeof(DistHalfPlan), cudaHostAllocDefault));
    }
    cudaEvent_t ev_exchanged = dev->ev_pipe[0], ev_a_landed = dev->ev_pipe[3], ev_a_sorted = dev->ev_pipe[4];
    int copy_streams = b200rs_exp_env("B200RS_DIST_COPY_STREAMS", COPY_STREAMS);
    copy_streams = copy_streams < 1 ? 1 : (copy_streams > COPY_STREAMS ? COPY_STREAMS : copy_streams);

    B200RS_CUDA(cudaMemcpyAsync(peers, recv_base, (size_t)world * 8, cudaMemcpyHostToDevice, dev->stream));  // (pageable source: staged by the runtime before the call returns)
    B200RS_TRY(b200rs_digit_histogram_pairs(dev, in, n, 24, 8, hist));
    // the all-gather also orders this step after every rank's previous local sorts: nobody overwrites a receive buffer that is still being read
    {
        const int rc = comm->allgather(comm->user, hist, gathered, RADIX * 8);
        if (rc != 0) return rc;
    }
    B200RS_TRY(b200rs_dist_plan_halves(dev, gathered, world, rank, peers, recv_capacity_pairs, (uint64_t)(uintptr_t)stage, n,
                                       b200rs_exp_env("B200RS_DIST_A_PERMILLE", DIST_PIPELINE_A_PERMILLE), lut, part_base, counts_dev, status_dev,
                                       reinterpret_cast<uint64_t*>(plan_dev)));
    B200RS_CUDA(cudaMemcpyAsync(dev->pinned_plan, plan_dev, sizeof(DistHalfPlan), cudaMemcpyDeviceToHost, dev->stream));
    B200RS_CUDA(cudaStreamSynchronize(dev->stream));
    const DistHalfPlan plan = *static_cast<const DistHalfPlan*>(dev->pinned_plan);
    if (plan.status != 0) return B200RS_OK;  // status_dev[0] = 1, counts_dev = 0: the caller re-plans (same contract as the unpipelined form)

    size_t have = xp_bytes;
    B200RS_TRY(b200rs_exchange_pairs(dev, in, n, 24, 8, lut, part_base, 2 * world, counts_dev, base + xp_off, &have));
    B200RS_CUDA(cudaEventRecord(ev_exchanged, dev->stream));
    // half B: one copy per peer, farthest-first rotation so that no destination is everybody's first target
    for (int c = 0; c < copy_streams; ++c) B200RS_CUDA(cudaStreamWaitEvent(dev->copy[c], ev_exchanged, 0));
    for (int i = 1; i < world; ++i) {
        const int d = (rank + i) % world;
        if (plan.stage_cnt[d] == 0) continue;
        dev->launches++;
        B200RS_CUDA(cudaMemcpyAsync(reinterpret_cast<void*>((uintptr_t)plan.dst_addr[d]), stage + 8ull * plan.stage_off[d], 8ull * plan.stage_cnt[d], cudaMemcpyDeviceToDevice,
                                    dev->copy[(i - 1) % copy_streams]));
    }
    for (int c = 0; c < copy_streams; ++c) B200RS_CUDA(cudaEventRecord(dev->ev_copied[c], dev->copy[c]));
    {
        const int rc = comm->barrier(comm->user);  // every rank's exchange kernel is done: all halves A have landed
        if (rc != 0) return rc;
    }
    B200RS_CUDA(cudaEventRecord(ev_a_landed, dev->stream));

    b200rs_pair* mine = reinterpret_cast<b200rs_pair*>(recv_base[rank]);
    size_t bytes_a = 0, bytes_b = 0;
    B200RS_TRY(b200rs_sort_pairs_u32(dev, nullptr, plan.recv_a, 32, nullptr, &bytes_a));
    B200RS_TRY(b200rs_sort_pairs_u32(dev, nullptr, plan.recv_b, 32, nullptr, &bytes_b));
    const size_t temp_b_off = b200rs_align_up(bytes_a, 256);
    const bool concurrent = temp_b_off + b200rs_align_up(bytes_b, 256) <= sort_reserved && !b200rs_exp_env("B200RS_DIST_SEQUENTIAL_SORTS", 0);
    int rc_a = B200RS_OK;
    if (concurrent) {  // half A is sorted on the second stream while the copy engines deliver half B
        B200RS_CUDA(cudaStreamWaitEvent(dev->aux, ev_a_landed, 0));
        cudaStream_t main_stream = dev->stream;
        dev->stream = dev->aux;
        size_t h = bytes_a;
        rc_a = b200rs_sort_pairs_u32(dev, mine, plan.recv_a, 32, base + sort_off, &h);
        dev->stream = main_stream;
        if (rc_a != B200RS_OK) return rc_a;
        B200RS_CUDA(cudaEventRecord(ev_a_sorted, dev->aux));
    }
    for (int c = 0; c < copy_streams; ++c) B200RS_CUDA(cudaStreamWaitEvent(dev->stream, dev->ev_copied[c], 0));
    {
        const int rc = comm->barrier(comm->user);  // every rank's copies are done: all halves B have landed
        if (rc != 0) return rc;
    }
    if (!concurrent) {
        size_t h = bytes_a;
        B200RS_TRY(b200rs_sort_pairs_u32(dev, mine, plan.recv_a, 32, base + sort_off, &h));
    }
    {
        size_t h = bytes_b;
        B200RS_TRY(b200rs_sort_pairs_u32(dev, mine + plan.recv_a, plan.recv_b, 32, base + sort_off + (concurrent ? temp_b_off : 0), &h));
    }
    if (concurrent) B200RS_CUDA(cudaStreamWaitEvent(dev->stream, ev_a_sorted, 0));
    return B200RS_OK;
}
}  // namespace

// ---- the whole partitioned sort of one rank (see include/b200rs.h) -------------------------------------------------------
extern "C" int b200rs_dist_sort_pairs_u32(b200rs_device* dev, const b200rs_dist_comm* comm, const uint64_t* recv_base, uint64_t recv_capacity_pairs,
                                          const b200rs_pair* in, uint64_t n, uint64_t* counts_dev, uint32_t* status_dev, void* temp, size_t* temp_bytes) {
    if (!dev || !comm || !temp_bytes || comm->world < 1 || comm->world > XP_MAX_PARTS || comm->rank < 0 || comm->rank >= comm->world) return B200RS_ERR_INVALID_ARGUMENT;
    const int world = comm->world;
    // (the choice depends only on what is the same on every rank: the ranks must run the same sequence of collectives)
    if (world >= 2 && world <= DIST_PIPELINE_MAX_WORLD && recv_capacity_pairs >= DIST_PIPELINE_MIN_CAPACITY && !b200rs_exp_env("B200RS_DIST_NO_PIPELINE", 0))
        return dist_sort_pipelined(dev, comm, recv_base, recv_capacity_pairs, in, n, counts_dev, status_dev, temp, temp_bytes);
    // temp: [own top histogram 256 x u64][gathered world x 256 x u64][peer bases 256 x u64][part bases 256 x u64][lut 256]
    //       [own destination tables world x 3 x 256 x u64][gathered world x that][histograms of the local sort 4 x 256 x u64][exchange temp][local sort temp]
    size_t xp_bytes = 0, sort_bytes = 0;
    B200RS_TRY(b200rs_exchange_pairs(dev, nullptr, n, 24, 8, nullptr, nullptr, world, nullptr, nullptr, &xp_bytes));
    B200RS_TRY(b200rs_sort_pairs_u32_devn(dev, nullptr, recv_capacity_pairs, nullptr, 32, nullptr, &sort_bytes));
    const size_t table_bytes = (size_t)world * DH_DIGITS * RADIX * 8;
    const size_t hist_off = 0, gathered_off = hist_off + RADIX * 8, peers_off = gathered_off + (size_t)world * RADIX * 8, parts_off = peers_off + RADIX * 8,
                 lut_off = parts_off + RADIX * 8, table_off = lut_off + 256, tables_off = table_off + table_bytes, ghist_off = tables_off + world * table_bytes,
                 xp_off = ghist_off + 4 * RADIX * 8, sort_off = xp_off + b200rs_align_up(xp_bytes, 256), need = sort_off + b200rs_align_up(sort_bytes, 256);
    if (!temp) {
        *temp_bytes = need;
        return B200RS_OK;
    }
    if (*temp_bytes < need) return B200RS_ERR_TEMP_TOO_SMALL;
    if (!recv_base || !counts_dev || !status_dev || (n && !in) || !comm->allgather || !comm->barrier || ((uintptr_t)temp & 255u)) return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    char* base = static_cast<char*>(temp);
    uint64_t* hist = reinterpret_cast<uint64_t*>(base + hist_off);
    uint64_t* gathered = reinterpret_cast<uint64_t*>(base + gathered_off);
    uint64_t* peers = reinterpret_cast<uint64_t*>(base + peers_off);
    uint64_t* part_base = reinterpret_cast<uint64_t*>(base + parts_off);
    uint8_t* lut = reinterpret_cast<uint8_t*>(base + lut_off);
    unsigned long long* table = reinterpret_cast<unsigned long long*>(base + table_off);
    unsigned long long* tables = reinterpret_cast<unsigned long long*>(base + tables_off);
    uint64_t* ghist = reinterpret_cast<uint64_t*>(base + ghist_off);
    if (!dev->aux) {  // second stream + events for the histogram that runs next to the exchange
        B200RS_CUDA(cudaStreamCreateWithFlags(&dev->aux, cudaStreamNonBlocking));
        B200RS_CUDA(cudaEventCreateWithFlags(&dev->ev_aux[0], cudaEventDisableTiming));
        B200RS_CUDA(cudaEventCreateWithFlags(&dev->ev_aux[1], cudaEventDisableTiming));
    }
    B200RS_CUDA(cudaMemcpyAsync(peers, recv_base, (size_t)world * 8, cudaMemcpyHostToDevice, dev->stream));  // (pageable source: staged by the runtime before the call returns)
    B200RS_CUDA(cudaMemsetAsync(table, 0, table_bytes, dev->stream));
    B200RS_TRY(b200rs_digit_histogram_pairs(dev, in, n, 24, 8, hist));
    // the all-gather also orders this step after every rank's previous local sort: nobody overwrites a receive buffer that is still being read
    {
        const int rc = comm->allgather(comm->user, hist, gathered, RADIX * 8);
        if (rc != 0) return rc;
    }
    B200RS_TRY(b200rs_dist_plan(dev, gathered, world, comm->rank, peers, recv_capacity_pairs, n, lut, part_base, counts_dev, status_dev));
    // second stream: per-destination histograms of digits 0..2 of what this rank ships, hidden behind the exchange where that
    // is bound by NVLink (from 4 ranks on: 3/4 and more of the pairs leave the GPU; with 2 ranks half of them stay and the two
    // kernels compete for the SMs: measured 2.05 against 1.70 ms for the exchange, nothing gained)
    const bool overlap = b200rs_exp_env("B200RS_DIST_OVERLAP", world >= 4 ? 1 : 0) != 0;
    B200RS_CUDA(cudaEventRecord(dev->ev_aux[0], dev->stream));
    B200RS_CUDA(cudaStreamWaitEvent(dev->aux, dev->ev_aux[0], 0));
    if (n && overlap) {
        dev->launches++;
        const size_t smem = (size_t)world * DH_DIGITS * RADIX * sizeof(uint32_t);
        B200RS_TRY(b200rs_kernel_setup(dev, (const void*)dest_digit_histogram_kernel, smem));
        uint64_t blocks = (n / 2 + (uint64_t)DH_THREADS * 4 - 1) / ((uint64_t)DH_THREADS * 4);
        if (blocks > (uint64_t)dev->num_sms) blocks = (uint64_t)dev->num_sms;
        if (blocks == 0) blocks = 1;
        dest_digit_histogram_kernel<<<(unsigned)blocks, DH_THREADS, smem, dev->aux>>>(reinterpret_cast<const uint2*>(in), n, reinterpret_cast<const unsigned long long*>(counts_dev),
                                                                                     lut, world, table);
        B200RS_CUDA(cudaGetLastError());
    }
    B200RS_CUDA(cudaEventRecord(dev->ev_aux[1], dev->aux));
    size_t have = xp_bytes;
    B200RS_TRY(b200rs_exchange_pairs(dev, in, n, 24, 8, lut, part_base, world, counts_dev, base + xp_off, &have));
    B200RS_CUDA(cudaStreamWaitEvent(dev->stream, dev->ev_aux[1], 0));
    if (overlap) {
        const int rc = comm->allgather(comm->user, table, tables, table_bytes);
        if (rc != 0) return rc;
    }
    {
        const int rc = comm->barrier(comm->user);  // every rank's stores have landed before anyone sorts
        if (rc != 0) return rc;
    }
    have = sort_bytes;
    b200rs_pair* mine = reinterpret_cast<b200rs_pair*>(recv_base[comm->rank]);
    if (!overlap) return b200rs_sort_pairs_u32_devn(dev, mine, recv_capacity_pairs, counts_dev + 1, 32, base + sort_off, &have);
    {
        b200rs_launch_scope scope(dev, "recv_histogram", 4 * RADIX, (uint64_t)world * table_bytes);
        recv_histogram_kernel<<<1, RADIX, 0, dev->stream>>>(tables, reinterpret_cast<const unsigned long long*>(gathered), lut, world, comm->rank, status_dev,
                                                            reinterpret_cast<unsigned long long*>(ghist));
    }
    B200RS_CUDA(cudaGetLastError());
    return sort_pairs_devn_with_histogram(dev, mine, recv_capacity_pairs, counts_dev + 1, ghist, base + sort_off, &have);
}

extern "C" int b200rs_enable_peer_access(b200rs_device* dev, int peer_device_idx) {
    if (!dev || peer_device_idx < 0) return B200RS_ERR_INVALID_ARGUMENT;
    if (peer_device_idx == dev->device_idx) return B200RS_OK;
    b200rs_device_guard guard(dev);
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device_idx, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        (void)cudaGetLastError();
        return B200RS_OK;
    }
    B200RS_CUDA(e);
    return B200RS_OK;
}
extern "C" int b200rs_ipc_export(b200rs_device* dev, void* ptr, unsigned char handle_out[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    if (!dev || !ptr || !handle_out) return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    cudaIpcMemHandle_t h;
    B200RS_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle_out, &h, 64);
    return B200RS_OK;
}
extern "C" int b200rs_ipc_import(b200rs_device* dev, const unsigned char handle[64], void** ptr) {
    if (!dev || !handle || !ptr) return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    B200RS_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return B200RS_OK;
}
extern "C" int b200rs_ipc_release(b200rs_device* dev, void* ptr) {
    if (!dev || !ptr) return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    B200RS_CUDA(cudaIpcCloseMemHandle(ptr));
    return B200RS_OK;
}

extern "C" int b200rs_sort_keys_u32(b200rs_device* dev, uint32_t* inout, uint64_t n, int sort_bits, void* temp, size_t* temp_bytes) {
    return sort_impl<uint32_t>(dev, inout, n, sort_bits, temp, temp_bytes, "keys");
}

extern "C" int b200rs_sort_keys_u32_msd(b200rs_device* dev, uint32_t* inout, uint64_t n, void* temp, size_t* temp_bytes, int* used) {
    return sort_impl<uint32_t>(dev, inout, n, 32, temp, temp_bytes, "keys", nullptr, MSD_FORCED, used);
}

extern "C" int b200rs_sort_pairs_u32(b200rs_device* dev, b200rs_pair* inout, uint64_t n, int sort_bits, void* temp, size_t* temp_bytes) {
    return sort_impl<uint2>(dev, reinterpret_cast<uint2*>(inout), n, sort_bits, temp, temp_bytes, "pairs");
}
