// tools/launch_cost.cu -- what does a launch cost whose CTAs all exit at once?  (Decides whether the LSD kernels can be
// launched unconditionally behind a device-side "already sorted by the MSD path" flag instead of a host round trip.)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 4) early_exit(const unsigned* flag, unsigned* out) {
    extern __shared__ unsigned s[];
    if (*flag) return;
    s[threadIdx.x] = threadIdx.x;
    __syncthreads();
    out[blockIdx.x * 256 + threadIdx.x] = s[255 - threadIdx.x];
}
int main() {
    unsigned *flag, *out;
    cudaMalloc(&flag, 4); cudaMalloc(&out, 40000 * 256 * 4);
    unsigned one = 1; cudaMemcpy(flag, &one, 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(early_exit, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int smem : {0, 36 * 1024}) for (int grid : {148, 1184, 8192, 30000}) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(a);
            for (int i = 0; i < 4; ++i) early_exit<<<grid, 256, smem>>>(flag, out);
            cudaEventRecord(b); cudaEventSynchronize(b);
        }
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("grid %6d smem %6d: %.2f us per launch (4 back-to-back early-exit launches)\n", grid, smem, ms * 1000 / 4);
    }
    // host round trip: kernel -> 8-byte D2H -> sync -> next kernel
    unsigned* pinned; cudaHostAlloc((void**)&pinned, 64, cudaHostAllocDefault);
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        early_exit<<<148, 256>>>(flag, out);
        cudaMemcpyAsync(pinned, flag, 8, cudaMemcpyDeviceToHost, 0);
        cudaStreamSynchronize(0);
        early_exit<<<148, 256>>>(flag, out);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("kernel + D2H + sync + kernel: %.2f us\n", ms * 1000);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
