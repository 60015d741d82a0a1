// tools/cub_compare.cu -- EXTERNAL COMPARISON ONLY (north_star: "CUB DeviceRadixSort is timed as an external
// comparison only").  A stand-alone binary; nothing in libb200rs.so, the tests or bench.py's product path links
// or calls it.  Times, on the same sizes and byte accounting as bench.py:
//   cub::DeviceRadixSort::SortKeys<uint32_t>                 36 B/key
//   cub::DeviceRadixSort::SortKeys<uint64_t>(bits 0..32)     AoS pair {key, value} packed as one u64, key = low half:
//                                                            the same stable key-value semantics as Pprims::radixSort
//                                                            on Buffer<uint2> (Pprims.cpp:200-302); 72 B/pair
//   cub::DeviceRadixSort::SortPairs<uint32_t, uint32_t>      SoA pairs (what CUB users normally run); 72 B/pair
//   cub::DeviceScan::ExclusiveSum<uint32_t>                  8 B/element
// Usage: cub_compare [log2n=28] [reps=10] [peak_GBps=6547.5]   -> one JSON line on stdout.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

// splitmix64 of the element index: uniform keys, reproducible, generated on the device
__global__ void fill_u32(uint32_t* p, uint64_t n, uint64_t seed) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t z = (i + seed) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        p[i] = (uint32_t)(z ^ (z >> 31));
    }
}
__global__ void fill_pairs(uint2* p, uint64_t n, uint64_t seed) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t z = (i + seed) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        p[i] = make_uint2((uint32_t)(z ^ (z >> 31)), (uint32_t)i);
    }
}

template <typename F, typename G>
static float time_best(int reps, F&& restore, G&& run) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < reps + 2; ++r) {
        restore();
        CK(cudaEventRecord(e0));
        run();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 2) best = std::min(best, ms);
    }
    return best;
}

int main(int argc, char** argv) {
    const int log2n = argc > 1 ? atoi(argv[1]) : 28;
    const int reps = argc > 2 ? atoi(argv[2]) : 10;
    const double peak = argc > 3 ? atof(argv[3]) : 6547.5;
    const uint64_t n = 1ull << log2n;
    const int blocks = 148 * 8;

    void *src, *a, *b, *temp = nullptr;
    CK(cudaMalloc(&src, n * 8));
    CK(cudaMalloc(&a, n * 8));
    CK(cudaMalloc(&b, n * 8));
    size_t temp_bytes = 0, need = 0;

    struct Row { const char* name; float ms; double bytes; };
    std::vector<Row> rows;

    // ---- SortKeys u32 ----
    fill_u32<<<blocks, 256>>>((uint32_t*)src, n, 1);
    {
        cub::DoubleBuffer<uint32_t> db((uint32_t*)a, (uint32_t*)b);
        CK(cub::DeviceRadixSort::SortKeys(nullptr, need, db, (int)n));
        temp_bytes = std::max(temp_bytes, need);
    }
    {
        cub::DoubleBuffer<uint64_t> db((uint64_t*)a, (uint64_t*)b);
        CK(cub::DeviceRadixSort::SortKeys(nullptr, need, db, (int)n, 0, 32));
        temp_bytes = std::max(temp_bytes, need);
        CK(cub::DeviceScan::ExclusiveSum(nullptr, need, (uint32_t*)a, (uint32_t*)b, (int)n));
        temp_bytes = std::max(temp_bytes, need);
    }
    void *va, *vb;  // SoA values
    CK(cudaMalloc(&va, n * 4));
    CK(cudaMalloc(&vb, n * 4));
    {
        cub::DoubleBuffer<uint32_t> dk((uint32_t*)a, (uint32_t*)b), dv((uint32_t*)va, (uint32_t*)vb);
        CK(cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, (int)n));
        temp_bytes = std::max(temp_bytes, need);
    }
    CK(cudaMalloc(&temp, temp_bytes));

    auto nop = [] {};
    for (int bits : {32, 16}) {
        float ms = time_best(reps, [&] { CK(cudaMemcpyAsync(a, src, n * 4, cudaMemcpyDeviceToDevice)); },
                             [&] {
                                 cub::DoubleBuffer<uint32_t> db((uint32_t*)a, (uint32_t*)b);
                                 size_t tb = temp_bytes;
                                 CK(cub::DeviceRadixSort::SortKeys(temp, tb, db, (int)n, 0, bits));
                             });
        rows.push_back({bits == 32 ? "cub_sort_keys_u32" : "cub_sort_keys_u32_bits16", ms, (double)n * (4 + 8 * (bits / 8))});
    }
    // ---- AoS pairs as packed u64, bits [0,32) ----
    fill_pairs<<<blocks, 256>>>((uint2*)src, n, 2);
    {
        float ms = time_best(reps, [&] { CK(cudaMemcpyAsync(a, src, n * 8, cudaMemcpyDeviceToDevice)); },
                             [&] {
                                 cub::DoubleBuffer<uint64_t> db((uint64_t*)a, (uint64_t*)b);
                                 size_t tb = temp_bytes;
                                 CK(cub::DeviceRadixSort::SortKeys(temp, tb, db, (int)n, 0, 32));
                             });
        rows.push_back({"cub_sort_keys_u64_as_aos_pairs_bits0_32", ms, (double)n * 72});
    }
    // ---- SoA pairs ----
    fill_u32<<<blocks, 256>>>((uint32_t*)src, n, 3);
    {
        float ms = time_best(reps, [&] { CK(cudaMemcpyAsync(a, src, n * 4, cudaMemcpyDeviceToDevice)); },
                             [&] {
                                 cub::DoubleBuffer<uint32_t> dk((uint32_t*)a, (uint32_t*)b), dv((uint32_t*)va, (uint32_t*)vb);
                                 size_t tb = temp_bytes;
                                 CK(cub::DeviceRadixSort::SortPairs(temp, tb, dk, dv, (int)n));
                             });
        rows.push_back({"cub_sort_pairs_u32_u32_soa", ms, (double)n * 72});
    }
    // ---- ExclusiveSum ----
    {
        float ms = time_best(reps, nop, [&] {
            size_t tb = temp_bytes;
            CK(cub::DeviceScan::ExclusiveSum(temp, tb, (uint32_t*)src, (uint32_t*)b, (int)n));
        });
        rows.push_back({"cub_exclusive_sum_u32", ms, (double)n * 8});
    }
    // ---- plain copy as the yardstick (read + write bytes) ----
    {
        float ms = time_best(reps, nop, [&] { CK(cudaMemcpyAsync(b, src, n * 8, cudaMemcpyDeviceToDevice)); });
        rows.push_back({"cudaMemcpy_d2d", ms, (double)n * 16});
    }

    printf("{\"tool\": \"cub_compare\", \"cub_version\": %d, \"log2n\": %d, \"reps\": %d, \"peak_gbs\": %.1f, \"results\": {", CUB_VERSION, log2n, reps, peak);
    for (size_t i = 0; i < rows.size(); ++i)
        printf("%s\"%s\": {\"ms\": %.4f, \"gelem_s\": %.2f, \"gbs\": %.1f, \"roofline_frac\": %.4f}", i ? ", " : "", rows[i].name, rows[i].ms,
               n / rows[i].ms / 1e6, rows[i].bytes / rows[i].ms / 1e6, rows[i].bytes / rows[i].ms / 1e6 / peak);
    printf("}}\n");
    return 0;
}
