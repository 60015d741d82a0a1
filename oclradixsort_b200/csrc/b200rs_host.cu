// b200rs_host.cu -- HOST-buffer entry points: copy in, run the device path, copy out, synchronise.
// This is the call a CPU-side user of the reference makes in place of the map / fill / unmap /
// radixSort / map / read sequence of UnitTest/main.cpp:118-139.  There is no CPU implementation
// behind these: they fail if the CUDA path fails.
#include "b200rs_internal.h"

namespace {

template <typename SortFn>
int sort_host(b200rs_device* dev, void* host_inout, uint64_t n, size_t elem_bytes, int sort_bits, SortFn sort_fn) {
    if (!dev || (n && !host_inout)) return B200RS_ERR_INVALID_ARGUMENT;
    if (sort_bits < 0 || sort_bits > 32) return B200RS_ERR_INVALID_ARGUMENT;
    if (n == 0) return B200RS_OK;
    b200rs_device_guard guard(dev);
    size_t temp_bytes = 0;
    B200RS_TRY(sort_fn(nullptr, nullptr, &temp_bytes));
    const size_t data_bytes = (size_t)n * elem_bytes;
    B200RS_TRY(b200rs_reserve(dev, &dev->scratch_data, &dev->scratch_data_bytes, data_bytes));
    B200RS_TRY(b200rs_reserve(dev, &dev->scratch_temp, &dev->scratch_temp_bytes, temp_bytes));
    B200RS_CUDA(cudaMemcpyAsync(dev->scratch_data, host_inout, data_bytes, cudaMemcpyHostToDevice, dev->stream));
    size_t have = dev->scratch_temp_bytes;
    B200RS_TRY(sort_fn(dev->scratch_data, dev->scratch_temp, &have));
    B200RS_CUDA(cudaMemcpyAsync(host_inout, dev->scratch_data, data_bytes, cudaMemcpyDeviceToHost, dev->stream));
    B200RS_CUDA(cudaStreamSynchronize(dev->stream));
    return B200RS_OK;
}

}  // namespace

extern "C" int b200rs_sort_keys_u32_host(b200rs_device* dev, uint32_t* host_inout, uint64_t n, int sort_bits) {
    return sort_host(dev, host_inout, n, sizeof(uint32_t), sort_bits, [&](void* data, void* temp, size_t* temp_bytes) {
        return b200rs_sort_keys_u32(dev, static_cast<uint32_t*>(data), n, sort_bits, temp, temp_bytes);
    });
}

extern "C" int b200rs_sort_pairs_u32_host(b200rs_device* dev, b200rs_pair* host_inout, uint64_t n, int sort_bits) {
    return sort_host(dev, host_inout, n, sizeof(b200rs_pair), sort_bits, [&](void* data, void* temp, size_t* temp_bytes) {
        return b200rs_sort_pairs_u32(dev, static_cast<b200rs_pair*>(data), n, sort_bits, temp, temp_bytes);
    });
}

extern "C" int b200rs_exclusive_scan_u32_host(b200rs_device* dev, uint32_t* host_dst, const uint32_t* host_src, uint64_t n,
                                              uint32_t* host_total_out) {
    if (!dev || (n && (!host_dst || !host_src))) return B200RS_ERR_INVALID_ARGUMENT;
    if (n == 0) {
        if (host_total_out) *host_total_out = 0;
        return B200RS_OK;
    }
    b200rs_device_guard guard(dev);
    size_t temp_bytes = 0;
    B200RS_TRY(b200rs_exclusive_scan_u32(dev, nullptr, nullptr, n, nullptr, nullptr, &temp_bytes));
    const size_t data_bytes = (size_t)n * sizeof(uint32_t);
    // one extra word behind the data receives the total
    B200RS_TRY(b200rs_reserve(dev, &dev->scratch_data, &dev->scratch_data_bytes, data_bytes + 256));
    B200RS_TRY(b200rs_reserve(dev, &dev->scratch_temp, &dev->scratch_temp_bytes, temp_bytes));
    uint32_t* d = static_cast<uint32_t*>(dev->scratch_data);
    uint32_t* d_total = reinterpret_cast<uint32_t*>(static_cast<char*>(dev->scratch_data) + b200rs_align_up(data_bytes, 256) - 0);
    if (b200rs_align_up(data_bytes, 256) + sizeof(uint32_t) > dev->scratch_data_bytes) {
        B200RS_TRY(b200rs_reserve(dev, &dev->scratch_data, &dev->scratch_data_bytes, b200rs_align_up(data_bytes, 256) + 256));
        d = static_cast<uint32_t*>(dev->scratch_data);
        d_total = reinterpret_cast<uint32_t*>(static_cast<char*>(dev->scratch_data) + b200rs_align_up(data_bytes, 256));
    }
    B200RS_CUDA(cudaMemcpyAsync(d, host_src, data_bytes, cudaMemcpyHostToDevice, dev->stream));
    size_t have = dev->scratch_temp_bytes;
    B200RS_TRY(b200rs_exclusive_scan_u32(dev, d, d, n, d_total, dev->scratch_temp, &have));  // in place on the device
    B200RS_CUDA(cudaMemcpyAsync(host_dst, d, data_bytes, cudaMemcpyDeviceToHost, dev->stream));
    uint32_t total = 0;
    if (host_total_out) B200RS_CUDA(cudaMemcpyAsync(&total, d_total, sizeof(uint32_t), cudaMemcpyDeviceToHost, dev->stream));
    B200RS_CUDA(cudaStreamSynchronize(dev->stream));
    if (host_total_out) *host_total_out = total;
    return B200RS_OK;
}
