// tools/microbench_rank.cu -- SM cycles per 32 keys for candidate warp-ranking loops (8-bit digits),
// one CTA per SM.  Each variant ranks ITEMS keys per thread against per-warp counters in shared memory,
// exactly as the sort kernel would.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 ...
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITEMS = 16;
constexpr int REPS = 64;

__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

// peers via 8 ballots, plain C++
__device__ __forceinline__ uint32_t match8_cpp(uint32_t d) {
    uint32_t peers = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const bool bit = (d >> k) & 1u;
        const uint32_t b = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? b : ~b;
    }
    return peers;
}
// peers via 8 ballots, PTX with predicate reuse
__device__ __forceinline__ uint32_t match8_ptx(uint32_t d) {
    uint32_t peers = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint32_t b;
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %1, %2;\n\tsetp.ne.u32 p, t, 0;\n\tvote.sync.ballot.b32 %0, p, 0xffffffff;\n\t@!p not.b32 %0, %0;\n\t}"
                     : "=r"(b) : "r"(d), "r"(1u << k));
        peers &= b;
    }
    return peers;
}
// mask from sign-extension, single LOP3 combine: peers &= ~(b ^ m), m = all-ones when bit set
__device__ __forceinline__ uint32_t match8_sext(uint32_t d) {
    uint32_t peers = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t b = __ballot_sync(0xffffffffu, (d >> k) & 1u);
        const uint32_t m = (uint32_t)((int32_t)(d << (31 - k)) >> 31);
        peers &= ~(b ^ m);
    }
    return peers;
}

template <int V>
__global__ void __launch_bounds__(1024) bench(uint32_t* out, long long* cycles, int shift) {
    __shared__ uint32_t cnt[32][256];
    __shared__ uint32_t bitmap[V == 4 ? 32 : 1][V == 4 ? 128 : 1];  // V4 uses 7-bit digits to fit static smem
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 32 * 256; i += blockDim.x) { (&cnt[0][0])[i] = 0; if (V == 4 && i < 32 * 128) (&bitmap[0][0])[i] = 0; }
    uint32_t key[ITEMS];
    uint32_t s = threadIdx.x * 2654435761u + blockIdx.x * 977u + 12345u;
    for (int i = 0; i < ITEMS; ++i) { s = s * 1664525u + 1013904223u; key[i] = s; }
    uint32_t* my = cnt[warp];
    uint32_t* bm = bitmap[warp];
    const uint32_t lt = lanemask_lt();
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int r = 0; r < REPS; ++r) {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t d = (key[i] >> shift) & 255u;
            uint32_t rank = 0;
            if (V == 1) {  // HW match + leader LDS/STS + shfl
                const uint32_t peers = __match_any_sync(0xffffffffu, d);
                const uint32_t lower = peers & lt;
                uint32_t before = 0;
                if (lower == 0) { before = my[d]; my[d] = before + __popc(peers); }
                __syncwarp();
                rank = __shfl_sync(0xffffffffu, before, __ffs(peers) - 1) + __popc(lower);
            }
            if (V == 2 || V == 5 || V == 6) {  // ballots + leader ATOMS(ret) + shfl
                const uint32_t peers = V == 2 ? match8_cpp(d) : V == 5 ? match8_ptx(d) : match8_sext(d);
                const uint32_t lower = peers & lt;
                uint32_t before = 0;
                if (lower == 0) before = atomicAdd(&my[d], __popc(peers));
                rank = __shfl_sync(0xffffffffu, before, __ffs(peers) - 1) + __popc(lower);
            }
            if (V == 3) {  // ballots + every lane reads the counter, leader stores it back
                const uint32_t peers = match8_cpp(d);
                const uint32_t lower = peers & lt;
                const uint32_t before = my[d];
                __syncwarp();
                if (lower == 0) my[d] = before + __popc(peers);
                __syncwarp();
                rank = before + __popc(lower);
            }
            if (V == 4) {  // atomicOr bitmap match + counter read + leader store/clear
                atomicOr(&bm[d & 127], 1u << lane);
                __syncwarp();
                const uint32_t peers = bm[d & 127];
                const uint32_t before = my[d];
                const uint32_t lower = peers & lt;
                __syncwarp();
                if (lower == 0) { my[d] = before + __popc(peers); bm[d & 127] = 0; }
                __syncwarp();
                rank = before + __popc(lower);
            }
            if (V == 7) {  // count only: RED.ADD (what an early-count phase costs)
                atomicAdd(&my[d], 1u);
            }
            if (V == 8) {  // ballots only (no counters)
                rank = match8_cpp(d);
            }
            if (V == 9) {  // ballots (ptx) only
                rank = match8_ptx(d);
            }
            if (V == 10) {  // unstable: plain ATOMS with return (NOT usable -- order within the warp undefined); cost reference
                rank = atomicAdd(&my[d], 1u);
            }
            acc += rank;
        }
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) key[i] = key[i] * 1664525u + 1013904223u + (acc & 1);
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + my[lane];
}

template <int V>
void run(const char* name, int warps, int shift = 13) {
    const int sms = 148;
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, sizeof(uint32_t) * sms * 1024);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    bench<V><<<sms, warps * 32>>>(out, cyc, shift);
    bench<V><<<sms, warps * 32>>>(out, cyc, shift);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
    printf("%-52s warps=%2d : %7.2f SM-cycles per 32 keys (%s)\n", name, warps, avg / ((double)REPS * ITEMS * warps), cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {16, 32}) {
        run<1>("V1 HW match + LDS/STS + shfl", w);
        run<2>("V2 ballots(c++) + leader ATOMS + shfl", w);
        run<5>("V5 ballots(ptx) + leader ATOMS + shfl", w);
        run<6>("V6 ballots(sext) + leader ATOMS + shfl", w);
        run<3>("V3 ballots(c++) + LDS all + leader STS", w);
        run<4>("V4 atomicOr bitmap + LDS + leader STS", w);
        run<7>("V7 count only RED.ADD", w);
        run<8>("V8 ballots(c++) only", w);
        run<9>("V9 ballots(ptx) only", w);
        run<10>("V10 plain ATOMS+ret (unstable; cost reference)", w);
    }
    printf("-- skewed digits (shift=30 -> 4 distinct values)\n");
    run<2>("V2 ballots + leader ATOMS + shfl, skewed", 32, 30);
    run<4>("V4 atomicOr bitmap, skewed", 32, 30);
    run<7>("V7 RED.ADD, skewed", 32, 30);
    return 0;
}
