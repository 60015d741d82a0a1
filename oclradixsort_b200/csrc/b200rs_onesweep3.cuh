// b200rs_onesweep3.cuh -- third-generation scatter pass: the generation-2 tile (two-level look-back, branch-free
// leader atomics, element-wise write-out; b200rs_onesweep2.cuh) inside a PERSISTENT CTA with a software pipeline
// across tiles.  Included by b200rs_sort.cu inside its anonymous namespace.
//
// What the generation-2 profile showed (profiles/r1b_phase_summary.txt): every tile pays, with nothing to overlap it
// inside its own CTA, (a) the CTA launch + an L2 round trip for its ticket before the first load can be issued
// (the barrier behind the ticket is ~7 % of all warp samples), and (b) the full HBM latency of its loads at the head of
// the counting phase (long-scoreboard stalls: 38 % of that phase's samples for keys, 55 % for pairs).  Here a CTA
// keeps taking tickets until the input is exhausted, and
//   * the ticket of tile k+1 is fetched at the very start of tile k (thread 0, result parked in a register and
//     handed over through shared memory in the digit-thread phase), and
//   * tile k+1 is pulled into L2 by ONE cp.async.bulk.prefetch.L2 (thread 0, as soon as it knows the ticket), so the
//     counting phase of tile k+1 sees L2 latency instead of HBM latency.  (Loading tile k+1 into the registers that
//     the ranking loop frees was tried first: the write-out keeps IPT staged elements in flight, so 2 x IPT values
//     are live and every shape spilled 130-700 bytes per thread.)
// Tile ids still come from the atomic ticket, so a tile only ever waits for tiles that are held by running CTAs.
#pragma once

// NOT inlined on purpose: inside the persistent loop the inlined body shares one register allocation with everything that
// lives across iterations, and ptxas then spills the tile's elements (130-700 bytes per thread for every shape tried);
// as a separate function the body compiles like the generation-2 kernel (no spills) and the call costs a few
// instructions per tile.  Returns the next tile id.
template <typename ElemT, int THREADS, int IPT, bool FULL, bool BYTE_DIGIT, int PF>
__device__ __noinline__ uint32_t onesweep3_tile(typename Onesweep2Config<ElemT, THREADS, IPT, WO_ELEM>::Smem& s, const ElemT* __restrict__ in,
                                                ElemT* __restrict__ out, uint64_t n, uint64_t tile_base, uint32_t valid, int shift,
                                                uint32_t digit_mask, uint32_t prmt_sel, uint32_t tile, uint32_t pass,
                                                const unsigned long long* __restrict__ digit_start, Lookback3 lb, uint32_t minus_one,
                                                uint32_t next_ticket /* thread 0 only */) {
    using Cfg = Onesweep2Config<ElemT, THREADS, IPT, WO_ELEM>;
    constexpr uint32_t E = (uint32_t)sizeof(ElemT);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t slice = warp * Cfg::WARP_SLICE + lane;  // tile-local index of item 0
    const uint32_t my_offset = smem_addr(&s.warp_offset[warp][0]);
    const uint32_t staged = smem_addr(&s.staged[0]);

    // ---- 1. warp-striped load + per-warp digit counts ----
    ElemT elem[IPT];
    {
        const ElemT* __restrict__ src = in + tile_base + slice;
#pragma unroll
        for (int i = 0; i < IPT; ++i)
            if (FULL || slice + i * 32 < valid) elem[i] = __ldg(src + i * 32);
    }
#pragma unroll
    for (int i = 0; i < IPT; ++i)
        if (FULL || slice + i * 32 < valid) {
            const uint32_t d = digit_of_opaque<BYTE_DIGIT>(Elem<ElemT>::key(elem[i]), shift, digit_mask, prmt_sel);
            red_add_shared(my_offset + 4u * d, 1u);
        }
    __syncthreads();  // A

    // ---- 2. one thread per digit: tile totals, publication, layout of the staged tile ----
    uint32_t total = 0, sbase = 0, mates = 0;
    if (tid < RADIX) {
#pragma unroll
        for (int w = 0; w < Cfg::WARPS; ++w) total += s.warp_offset[w][tid];
        mates = lookback3_publish(lb, tile, pass, total, tid);
        sbase = block_exclusive_scan_256<uint32_t>(total, reinterpret_cast<uint32_t*>(s.scan_scratch), tid);
        uint32_t run = staged + E * sbase - E;  // byte address biased by one element (see onesweep2_tile)
#pragma unroll
        for (int w = 0; w < Cfg::WARPS; ++w) {
            const uint32_t c = s.warp_offset[w][tid];
            s.warp_offset[w][tid] = run;
            run += E * c;
        }
        if (tid == 0) {
            s.tile = next_ticket;  // the atomic was issued a whole counting phase ago
            if (PF) {
                const uint64_t next_base = (uint64_t)next_ticket * Cfg::TILE;
                if (next_base + Cfg::TILE <= n && (((uintptr_t)in & 15u) == 0)) bulk_prefetch_l2(in + next_base, Cfg::TILE * E);
            }
        }
    }
    __syncthreads();  // B

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
            peers = live ? (peers & live_lanes) : (1u << lane);
        }
        const uint32_t upto = E * (uint32_t)__popc(peers & le);
        const bool leader = (peers & gt) == 0 && live;
        uint32_t base = atom_add_shared(leader ? my_offset + 4u * digit : dummy, upto);
        base = __shfl_sync(0xffffffffu, base, 31 - __clz(peers));
        if (live) st_shared(base + upto, elem[i]);
    }

    const uint32_t next_tile = s.tile;

    // ---- 4. look-back (digit threads), then the element-wise write-out ----
    if (tid < RADIX) {
        const uint64_t exclusive = (uint64_t)digit_start[tid] + lookback3_resolve(lb, tile, pass, total, mates, tid);
        s.run_ptr[tid] = (uint64_t)(uintptr_t)(out + exclusive) - (uint64_t)sbase * E;
    }
    __syncthreads();  // C: every warp is done with its counter row
#pragma unroll
    for (int i = tid; i < Cfg::WARPS * RADIX; i += THREADS) (&s.warp_offset[0][0])[i] = 0;  // for the next tile's counts
    {
        const uint64_t my_bytes = (uint64_t)tid * E;
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const uint32_t j = (uint32_t)tid + (uint32_t)k * THREADS;
            if (FULL || j < valid) {
                const ElemT e = s.staged[j];
                const uint32_t d = digit_of<BYTE_DIGIT>(Elem<ElemT>::key(e), shift, digit_mask, prmt_sel);
                *reinterpret_cast<ElemT*>(s.run_ptr[d] + my_bytes + (uint64_t)k * THREADS * E) = e;  // generic store on purpose, see onesweep2_tile
            }
        }
    }
    __syncthreads();  // D: staged tile, run pointers and counter rows may be reused
    return next_tile;
}

template <typename ElemT, int THREADS, int IPT, int MIN_CTAS, int PF>
__global__ void __launch_bounds__(THREADS, MIN_CTAS)
onesweep3_kernel(const ElemT* __restrict__ in, ElemT* __restrict__ out, uint64_t n, int shift, uint32_t digit_mask,
                 const unsigned long long* __restrict__ digit_start, Lookback3 lb, uint32_t* ticket, uint32_t pass, uint32_t minus_one,
                 const unsigned long long* __restrict__ n_dev, const uint32_t* __restrict__ pass_ctl, uint32_t /*pf_tiles: unused*/) {
    using Cfg = Onesweep2Config<ElemT, THREADS, IPT, WO_ELEM>;
    static_assert(THREADS >= RADIX && THREADS % 32 == 0, "one thread per digit is needed");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typename Cfg::Smem& s = *reinterpret_cast<typename Cfg::Smem*>(smem_raw);
    const int tid = threadIdx.x;
    if (tid == 0) {
        const uint32_t ctl = pass_ctl[pass];
        const uint64_t n_eff = n_dev ? min(n, (uint64_t)*n_dev) : n;
        s.tile = atomicAdd(ticket, 1u);
        s.ctl = ctl;
        s.n_eff = n_eff;
    }
#pragma unroll
    for (int i = tid; i < Cfg::WARPS * RADIX; i += THREADS) (&s.warp_offset[0][0])[i] = 0;
    __syncthreads();
    n = s.n_eff;
    uint32_t tile = s.tile;
    const bool identity = (s.ctl & PASS_IDENTITY) != 0;
    __syncthreads();  // s.tile is rewritten below

    if (identity) {
        // one digit holds every element: the pass moves nothing, tiles are copied straight across
        const bool vec = ((((uintptr_t)in | (uintptr_t)out) & 15u) == 0);
        while ((uint64_t)tile * Cfg::TILE < n) {
            const uint64_t tile_base = (uint64_t)tile * Cfg::TILE;
            const uint32_t valid = (uint32_t)min((uint64_t)Cfg::TILE, n - tile_base);
            if (tid == 0) s.tile = atomicAdd(ticket, 1u);
            const ElemT* __restrict__ src = in + tile_base;
            ElemT* __restrict__ dst = out + tile_base;
            if (valid == Cfg::TILE && vec) {
                constexpr int VECS = Cfg::TILE * (int)sizeof(ElemT) / 16;
                const uint4* s4 = reinterpret_cast<const uint4*>(src);
                uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll 4
                for (int v = tid; v < VECS; v += THREADS) d4[v] = __ldg(s4 + v);
            } else {
                for (uint32_t j = tid; j < valid; j += THREADS) dst[j] = src[j];
            }
            __syncthreads();
            tile = s.tile;
            __syncthreads();
        }
        return;
    }

    const bool byte_digit = digit_mask == (uint32_t)(RADIX - 1);
    const uint32_t prmt_sel = 0x4440u | (uint32_t)(shift >> 3);
    while ((uint64_t)tile * Cfg::TILE < n) {
        const uint64_t tile_base = (uint64_t)tile * Cfg::TILE;
        const uint32_t valid = (uint32_t)min((uint64_t)Cfg::TILE, n - tile_base);
        uint32_t next_ticket = 0;
        if (tid == 0) next_ticket = atomicAdd(ticket, 1u);  // consumed in the digit-thread phase of this tile
        // whole tiles of whole-byte digits take the specialised body; ragged last tile and 1..7-bit top digits the generic one
        if (valid == Cfg::TILE && byte_digit)
            tile = onesweep3_tile<ElemT, THREADS, IPT, true, true, PF>(s, in, out, n, tile_base, valid, shift, digit_mask, prmt_sel, tile, pass, digit_start, lb, minus_one, next_ticket);
        else
            tile = onesweep3_tile<ElemT, THREADS, IPT, false, false, PF>(s, in, out, n, tile_base, valid, shift, digit_mask, prmt_sel, tile, pass, digit_start, lb, minus_one, next_ticket);
    }
}
