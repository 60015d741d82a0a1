// tools/microbench.cu -- per-SM throughput of the warp/shared-memory primitives a radix-sort ranking
// loop can be built from, on the real part.  One CTA per SM, W warps, every warp executes the same
// unrolled loop; reports SM cycles per warp-instruction (lower = faster).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 256;
constexpr int UNROLL = 8;

__device__ __forceinline__ uint32_t rnd(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 16; }

template <int OP>
__global__ void bench(uint32_t* out, long long* cycles, int table_words) {
    extern __shared__ uint32_t tab[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) tab[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t seed = threadIdx.x * 2654435761u + blockIdx.x;
    uint32_t acc = 0;
    uint32_t idx[UNROLL];
    for (int u = 0; u < UNROLL; ++u) idx[u] = rnd(seed) & (table_words - 1);
    uint32_t* mytab = tab + (OP >= 100 ? 0 : 0);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            uint32_t d = idx[u];
            if (OP == 0) acc += __match_any_sync(0xffffffffu, d + acc);                       // MATCH.ANY (dependent on acc: latency chain per warp)
            if (OP == 1) acc ^= __match_any_sync(0xffffffffu, d);                             // MATCH.ANY independent
            if (OP == 2) acc += __ballot_sync(0xffffffffu, (d >> (it & 7)) & 1);              // VOTE
            if (OP == 3) atomicAdd(&mytab[d], 1u);                                            // RED.shared random
            if (OP == 4) acc += atomicAdd(&mytab[d], 1u);                                     // ATOMS w/ return random
            if (OP == 5) acc += mytab[(d + acc) & (table_words - 1)];                                // LDS random (dependent)
            if (OP == 6) acc += mytab[d];                                                     // LDS random independent
            if (OP == 7) mytab[d] = acc + u;                                                  // STS random
            if (OP == 8) acc += __shfl_sync(0xffffffffu, d, d & 31);                          // SHFL idx
            if (OP == 9) atomicOr(&mytab[d], 1u << lane);                                     // RED.OR random
            if (OP == 10) acc += mytab[lane + 32 * (d & 63)];                                  // LDS conflict-free
            if (OP == 11) mytab[lane + 32 * (d & 63)] = acc;                                   // STS conflict-free
            if (OP == 12) atomicAdd(&mytab[lane + 32 * (d & 63)], 1u);                         // RED conflict-free
            if (OP == 13) { if ((d & 31) == 0 || lane == 0) acc += atomicAdd(&mytab[d], 3u); } // ATOMS few lanes
            if (OP == 14) acc += __popc(d + acc);                                             // POPC chain
            if (OP == 15) { uint32_t m = __match_any_sync(0xffffffffu, d & 15); acc ^= m; }   // MATCH.ANY few distinct
            if (OP == 16) atomicAdd(&mytab[d & 1], 1u);                                        // RED same-address heavy
            if (OP == 17) acc += __shfl_up_sync(0xffffffffu, d + acc, 1);                      // SHFL up chain
            if (OP == 18) { uint2 v = *reinterpret_cast<uint2*>(&mytab[2 * d]); acc += v.x + v.y; } // LDS.64 random
            if (OP == 19) { *reinterpret_cast<uint2*>(&mytab[2 * d]) = make_uint2(acc, u); }   // STS.64 random
        }
        if (OP != 0 && OP != 5 && OP != 14 && OP != 17)
            for (int u = 0; u < UNROLL; ++u) idx[u] = (idx[u] * 5 + 1) & (table_words - 1);  // new addresses (full-period LCG), 2 ALU ops
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + tab[threadIdx.x];
}

template <int OP>
void run(const char* name, int warps, int table_words) {
    int sms = 148;
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, sizeof(uint32_t) * sms * 1024);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    cudaFuncSetAttribute(bench<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    bench<OP><<<sms, warps * 32, 65536>>>(out, cyc, table_words);
    bench<OP><<<sms, warps * 32, 65536>>>(out, cyc, table_words);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
    double per = avg / ((double)ITERS * UNROLL * warps);
    printf("%-34s warps=%2d table=%4d : %7.2f cyc/warp-instr/SM  (%s)\n", name, warps, table_words, per, cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {4, 16, 32}) {
        run<1>("MATCH.ANY 8-bit random", w, 256);
        run<15>("MATCH.ANY 4-bit values", w, 256);
        run<2>("VOTE.BALLOT", w, 256);
        run<8>("SHFL.IDX", w, 256);
        run<3>("RED.ADD smem random256", w, 256);
        run<4>("ATOMS.ADD+ret random256", w, 256);
        run<13>("ATOMS.ADD+ret sparse lanes", w, 256);
        run<9>("RED.OR smem random256", w, 256);
        run<16>("RED.ADD 2 addresses", w, 256);
        run<6>("LDS random256", w, 256);
        run<6>("LDS random8192", w, 8192);
        run<7>("STS random8192", w, 8192);
        run<18>("LDS.64 random4096", w, 4096);
        run<19>("STS.64 random4096", w, 4096);
        run<10>("LDS conflict-free", w, 256);
        run<11>("STS conflict-free", w, 256);
        run<12>("RED.ADD conflict-free", w, 256);
    }
    run<0>("MATCH.ANY dependent chain", 1, 256);
    run<5>("LDS dependent chain", 1, 256);
    run<14>("POPC dependent chain", 1, 256);
    run<17>("SHFL dependent chain", 1, 256);
    return 0;
}
