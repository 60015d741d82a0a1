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

// Batch form: `count` independent host arrays of n elements each, every one sorted in place.  Three streams: the
// host -> device copy of array i+1 and the device -> host copy of array i-1 run while array i is being sorted (two
// device buffers), so both directions of the host link are busy at once; the sorts themselves stay on the handle's
// stream and share one temp block.
template <typename SortFn>
int sort_host_batch(b200rs_device* dev, void* const* host_inout, int count, uint64_t n, size_t elem_bytes, int sort_bits, SortFn sort_fn) {
    if (!dev || count < 0 || (count && !host_inout)) return B200RS_ERR_INVALID_ARGUMENT;
    if (sort_bits < 0 || sort_bits > 32) return B200RS_ERR_INVALID_ARGUMENT;
    if (n == 0 || count == 0) return B200RS_OK;
    for (int i = 0; i < count; ++i)
        if (!host_inout[i]) return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    if (!dev->copy_in) {
        B200RS_CUDA(cudaStreamCreateWithFlags(&dev->copy_in, cudaStreamNonBlocking));
        B200RS_CUDA(cudaStreamCreateWithFlags(&dev->copy_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            B200RS_CUDA(cudaEventCreateWithFlags(&dev->ev_in[i], cudaEventDisableTiming));
            B200RS_CUDA(cudaEventCreateWithFlags(&dev->ev_sorted[i], cudaEventDisableTiming));
            B200RS_CUDA(cudaEventCreateWithFlags(&dev->ev_out[i], cudaEventDisableTiming));
        }
        B200RS_CUDA(cudaEventCreateWithFlags(&dev->ev_start, cudaEventDisableTiming));
    }
    size_t temp_bytes = 0;
    B200RS_TRY(sort_fn(nullptr, nullptr, &temp_bytes));
    const size_t data_bytes = (size_t)n * elem_bytes;
    B200RS_TRY(b200rs_reserve(dev, &dev->scratch_data, &dev->scratch_data_bytes, data_bytes));
    if (count > 1) B200RS_TRY(b200rs_reserve(dev, &dev->scratch_data2, &dev->scratch_data2_bytes, data_bytes));
    B200RS_TRY(b200rs_reserve(dev, &dev->scratch_temp, &dev->scratch_temp_bytes, temp_bytes));
    void* buf[2] = {dev->scratch_data, dev->scratch_data2};
    // whatever the handle's stream was doing with the scratch buffers comes first
    B200RS_CUDA(cudaEventRecord(dev->ev_start, dev->stream));
    B200RS_CUDA(cudaStreamWaitEvent(dev->copy_in, dev->ev_start, 0));
    for (int i = 0; i < count; ++i) {
        const int b = i & 1;
        if (i >= 2) B200RS_CUDA(cudaStreamWaitEvent(dev->copy_in, dev->ev_out[b], 0));  // array i-2 has left buffer b
        B200RS_CUDA(cudaMemcpyAsync(buf[b], host_inout[i], data_bytes, cudaMemcpyHostToDevice, dev->copy_in));
        B200RS_CUDA(cudaEventRecord(dev->ev_in[b], dev->copy_in));
        B200RS_CUDA(cudaStreamWaitEvent(dev->stream, dev->ev_in[b], 0));
        size_t have = dev->scratch_temp_bytes;
        B200RS_TRY(sort_fn(buf[b], dev->scratch_temp, &have));
        B200RS_CUDA(cudaEventRecord(dev->ev_sorted[b], dev->stream));
        B200RS_CUDA(cudaStreamWaitEvent(dev->copy_out, dev->ev_sorted[b], 0));
        B200RS_CUDA(cudaMemcpyAsync(host_inout[i], buf[b], data_bytes, cudaMemcpyDeviceToHost, dev->copy_out));
        B200RS_CUDA(cudaEventRecord(dev->ev_out[b], dev->copy_out));
    }
    B200RS_CUDA(cudaStreamSynchronize(dev->copy_out));
    B200RS_CUDA(cudaStreamSynchronize(dev->stream));
    return B200RS_OK;
}

}  // namespace

extern "C" int b200rs_sort_keys_u32_host_batch(b200rs_device* dev, uint32_t* const* host_inout, int count, uint64_t n, int sort_bits) {
    return sort_host_batch(dev, reinterpret_cast<void* const*>(host_inout), count, n, sizeof(uint32_t), sort_bits, [&](void* data, void* temp, size_t* temp_bytes) {
        return b200rs_sort_keys_u32(dev, static_cast<uint32_t*>(data), n, sort_bits, temp, temp_bytes);
    });
}

extern "C" int b200rs_sort_pairs_u32_host_batch(b200rs_device* dev, b200rs_pair* const* host_inout, int count, uint64_t n, int sort_bits) {
    return sort_host_batch(dev, reinterpret_cast<void* const*>(host_inout), count, n, sizeof(b200rs_pair), sort_bits, [&](void* data, void* temp, size_t* temp_bytes) {
        return b200rs_sort_pairs_u32(dev, static_cast<b200rs_pair*>(data), n, sort_bits, temp, temp_bytes);
    });
}

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
