// b200rs_device.cu -- device / stream / memory shim of the C ABI (include/b200rs.h).
// Stands in for the reference's OpenCL backend (Adl/CL/AdlCL.inl): one device, one in-order
// stream, cudaMalloc buffers, async copies.  No kernels here.
#include "b200rs_internal.h"

#include <new>

extern "C" {

int b200rs_version(void) { return B200RS_VERSION; }

const char* b200rs_error_string(int code) {
    if (code == B200RS_OK) return "ok";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    switch (code) {
        case B200RS_ERR_INVALID_ARGUMENT: return "b200rs: invalid argument";
        case B200RS_ERR_TEMP_TOO_SMALL: return "b200rs: temp storage too small";
        case B200RS_ERR_NO_DEVICE: return "b200rs: no such CUDA device";
        case B200RS_ERR_UNSUPPORTED_ARCH: return "b200rs: device is not sm_100 (library is built for sm_100a only)";
        case B200RS_ERR_TOO_LARGE: return "b200rs: problem size not supported by this entry point";
        case B200RS_ERR_OUT_OF_MEMORY: return "b200rs: out of device memory";
        case B200RS_ERR_CAPACITY: return "b200rs: receive capacity exceeded";
        default: return "b200rs: unknown error";
    }
}

int b200rs_device_count(int* count) {
    if (!count) return B200RS_ERR_INVALID_ARGUMENT;
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        *count = 0;
        return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? B200RS_ERR_NO_DEVICE : (int)e;
    }
    return B200RS_OK;
}

static int create_common(int device_idx, cudaStream_t borrowed, bool borrow, b200rs_device** out) {
    if (!out) return B200RS_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int count = 0;
    B200RS_TRY(b200rs_device_count(&count));
    if (device_idx < 0 || device_idx >= count) return B200RS_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    B200RS_CUDA(cudaGetDeviceProperties(&prop, device_idx));
    if (prop.major != 10) return B200RS_ERR_UNSUPPORTED_ARCH;  // sm_100a cubin only: fail loudly, no fallback

    b200rs_device* d = new (std::nothrow) b200rs_device();
    if (!d) return B200RS_ERR_OUT_OF_MEMORY;
    d->device_idx = device_idx;
    d->num_sms = prop.multiProcessorCount;
    strncpy(d->name, prop.name, sizeof(d->name) - 1);
    b200rs_device_guard guard(d);
    if (borrow) {
        d->stream = borrowed;
        d->owns_stream = false;
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete d;
            return (int)e;
        }
        d->owns_stream = true;
    }
    *out = d;
    return B200RS_OK;
}

int b200rs_device_create(int device_idx, b200rs_device** dev) { return create_common(device_idx, nullptr, false, dev); }

int b200rs_device_create_on_stream(int device_idx, void* cuda_stream, b200rs_device** dev) {
    return create_common(device_idx, (cudaStream_t)cuda_stream, true, dev);
}

int b200rs_device_release_scratch(b200rs_device* dev) {
    if (!dev) return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    B200RS_CUDA(cudaStreamSynchronize(dev->stream));
    if (dev->scratch_data) cudaFree(dev->scratch_data);
    if (dev->scratch_data2) cudaFree(dev->scratch_data2);
    if (dev->scratch_temp) cudaFree(dev->scratch_temp);
    dev->scratch_data = dev->scratch_data2 = dev->scratch_temp = nullptr;
    dev->scratch_data_bytes = dev->scratch_data2_bytes = dev->scratch_temp_bytes = 0;
    return B200RS_OK;
}

int b200rs_device_destroy(b200rs_device* dev) {
    if (!dev) return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    int rc = b200rs_device_release_scratch(dev);
    for (auto& s : dev->spans) {
        cudaEventDestroy(s.start);
        cudaEventDestroy(s.stop);
    }
    if (dev->pinned_word) cudaFreeHost(dev->pinned_word);
    for (int i = 0; i < 2; ++i) {
        if (dev->ev_in[i]) cudaEventDestroy(dev->ev_in[i]);
        if (dev->ev_sorted[i]) cudaEventDestroy(dev->ev_sorted[i]);
        if (dev->ev_out[i]) cudaEventDestroy(dev->ev_out[i]);
    }
    if (dev->ev_start) cudaEventDestroy(dev->ev_start);
    if (dev->ev_aux[0]) cudaEventDestroy(dev->ev_aux[0]);
    if (dev->ev_aux[1]) cudaEventDestroy(dev->ev_aux[1]);
    for (cudaEvent_t e : dev->ev_pipe) if (e) cudaEventDestroy(e);
    for (cudaStream_t c : dev->copy) if (c) cudaStreamDestroy(c);
    for (cudaEvent_t e : dev->ev_copied) if (e) cudaEventDestroy(e);
    if (dev->pinned_plan) cudaFreeHost(dev->pinned_plan);
    if (dev->aux) cudaStreamDestroy(dev->aux);
    if (dev->copy_in) cudaStreamDestroy(dev->copy_in);
    if (dev->copy_out) cudaStreamDestroy(dev->copy_out);
    if (dev->owns_stream && dev->stream) cudaStreamDestroy(dev->stream);
    delete dev;
    return rc;
}

int b200rs_device_sync(b200rs_device* dev) {
    if (!dev) return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    B200RS_CUDA(cudaStreamSynchronize(dev->stream));
    return B200RS_OK;
}

int b200rs_device_num_sms(const b200rs_device* dev, int* num_sms) {
    if (!dev || !num_sms) return B200RS_ERR_INVALID_ARGUMENT;
    *num_sms = dev->num_sms;
    return B200RS_OK;
}

int b200rs_device_name(const b200rs_device* dev, char name_out[128]) {
    if (!dev || !name_out) return B200RS_ERR_INVALID_ARGUMENT;
    memcpy(name_out, dev->name, 128);
    return B200RS_OK;
}

int b200rs_device_index(const b200rs_device* dev, int* device_idx) {
    if (!dev || !device_idx) return B200RS_ERR_INVALID_ARGUMENT;
    *device_idx = dev->device_idx;
    return B200RS_OK;
}

int b200rs_device_mem_info(const b200rs_device* dev, size_t* free_bytes, size_t* total_bytes) {
    if (!dev || !free_bytes || !total_bytes) return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    B200RS_CUDA(cudaMemGetInfo(free_bytes, total_bytes));
    return B200RS_OK;
}

void* b200rs_device_stream(const b200rs_device* dev) { return dev ? (void*)dev->stream : nullptr; }

int b200rs_device_launch_count(const b200rs_device* dev, uint64_t* launches) {
    if (!dev || !launches) return B200RS_ERR_INVALID_ARGUMENT;
    *launches = dev->launches;
    return B200RS_OK;
}

// ---- memory ------------------------------------------------------------------------------------

int b200rs_malloc(b200rs_device* dev, size_t bytes, void** ptr) {
    if (!dev || !ptr) return B200RS_ERR_INVALID_ARGUMENT;
    *ptr = nullptr;
    if (bytes == 0) return B200RS_OK;
    b200rs_device_guard guard(dev);
    B200RS_CUDA(cudaMalloc(ptr, bytes));
    return B200RS_OK;
}

int b200rs_free(b200rs_device* dev, void* ptr) {
    if (!dev) return B200RS_ERR_INVALID_ARGUMENT;
    if (!ptr) return B200RS_OK;
    b200rs_device_guard guard(dev);
    B200RS_CUDA(cudaFree(ptr));  // implicit device sync: in-flight work on `ptr` finishes first
    return B200RS_OK;
}

int b200rs_host_alloc(b200rs_device* dev, size_t bytes, void** host_ptr) {
    if (!dev || !host_ptr) return B200RS_ERR_INVALID_ARGUMENT;
    *host_ptr = nullptr;
    if (bytes == 0) return B200RS_OK;
    b200rs_device_guard guard(dev);
    B200RS_CUDA(cudaHostAlloc(host_ptr, bytes, cudaHostAllocDefault));
    return B200RS_OK;
}

int b200rs_host_free(b200rs_device* dev, void* host_ptr) {
    if (!dev) return B200RS_ERR_INVALID_ARGUMENT;
    if (!host_ptr) return B200RS_OK;
    b200rs_device_guard guard(dev);
    B200RS_CUDA(cudaFreeHost(host_ptr));
    return B200RS_OK;
}

static int copy_async(b200rs_device* dev, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
    if (!dev || (bytes && (!dst || !src))) return B200RS_ERR_INVALID_ARGUMENT;
    if (bytes == 0) return B200RS_OK;
    b200rs_device_guard guard(dev);
    B200RS_CUDA(cudaMemcpyAsync(dst, src, bytes, kind, dev->stream));
    return B200RS_OK;
}

int b200rs_memcpy_h2d(b200rs_device* dev, void* dst, const void* host_src, size_t bytes) {
    return copy_async(dev, dst, host_src, bytes, cudaMemcpyHostToDevice);
}
int b200rs_memcpy_d2h(b200rs_device* dev, void* host_dst, const void* src, size_t bytes) {
    return copy_async(dev, host_dst, src, bytes, cudaMemcpyDeviceToHost);
}
int b200rs_memcpy_d2d(b200rs_device* dev, void* dst, const void* src, size_t bytes) {
    return copy_async(dev, dst, src, bytes, cudaMemcpyDeviceToDevice);
}

int b200rs_memset(b200rs_device* dev, void* ptr, int byte_value, size_t bytes) {
    if (!dev || (bytes && !ptr)) return B200RS_ERR_INVALID_ARGUMENT;
    if (bytes == 0) return B200RS_OK;
    b200rs_device_guard guard(dev);
    B200RS_CUDA(cudaMemsetAsync(ptr, byte_value, bytes, dev->stream));
    return B200RS_OK;
}

// ---- profiling -----------------------------------------------------------------------------------

int b200rs_profile_enable(b200rs_device* dev, int enable) {
    if (!dev) return B200RS_ERR_INVALID_ARGUMENT;
    dev->profiling = enable != 0;
    return B200RS_OK;
}

int b200rs_profile_read(b200rs_device* dev, b200rs_profile_entry* out, int capacity, int* count) {
    if (!dev || !count || (capacity > 0 && !out)) return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    B200RS_CUDA(cudaStreamSynchronize(dev->stream));
    int n = 0;
    for (auto& s : dev->spans) {
        if (n >= capacity) break;  // the rest stays queued for the next read
        float ms = 0.f;
        cudaEventElapsedTime(&ms, s.start, s.stop);
        s.entry.ms = ms;
        out[n++] = s.entry;
        cudaEventDestroy(s.start);
        cudaEventDestroy(s.stop);
    }
    dev->spans.erase(dev->spans.begin(), dev->spans.begin() + n);
    *count = n;
    return B200RS_OK;
}

// ---- events: device-side interval timing for adl::Stopwatch (Adl/AdlStopwatch.h:27-83) -------------

int b200rs_event_create(b200rs_device* dev, void** event_out) {
    if (!dev || !event_out) return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    cudaEvent_t e = nullptr;
    B200RS_CUDA(cudaEventCreate(&e));
    *event_out = (void*)e;
    return B200RS_OK;
}

int b200rs_event_record(b200rs_device* dev, void* event) {
    if (!dev || !event) return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    B200RS_CUDA(cudaEventRecord((cudaEvent_t)event, dev->stream));
    return B200RS_OK;
}

int b200rs_event_elapsed_ms(b200rs_device* dev, void* start_event, void* stop_event, float* ms_out) {
    if (!dev || !start_event || !stop_event || !ms_out) return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    B200RS_CUDA(cudaEventSynchronize((cudaEvent_t)stop_event));
    B200RS_CUDA(cudaEventElapsedTime(ms_out, (cudaEvent_t)start_event, (cudaEvent_t)stop_event));
    return B200RS_OK;
}

int b200rs_event_query(b200rs_device* dev, void* event, int* done) {
    if (!dev || !event || !done) return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    const cudaError_t e = cudaEventQuery((cudaEvent_t)event);
    if (e == cudaErrorNotReady) {
        (void)cudaGetLastError();
        *done = 0;
        return B200RS_OK;
    }
    B200RS_CUDA(e);
    *done = 1;
    return B200RS_OK;
}

int b200rs_event_synchronize(b200rs_device* dev, void* event) {
    if (!dev || !event) return B200RS_ERR_INVALID_ARGUMENT;
    b200rs_device_guard guard(dev);
    B200RS_CUDA(cudaEventSynchronize((cudaEvent_t)event));
    return B200RS_OK;
}

int b200rs_event_destroy(b200rs_device* dev, void* event) {
    if (!dev) return B200RS_ERR_INVALID_ARGUMENT;
    if (!event) return B200RS_OK;
    b200rs_device_guard guard(dev);
    B200RS_CUDA(cudaEventDestroy((cudaEvent_t)event));
    return B200RS_OK;
}

}  // extern "C"

// Grow-only scratch slot (role of uArray::setSize on Pprims' work buffers, Pprims.cpp:226-229).
int b200rs_reserve(b200rs_device* dev, void** slot, size_t* slot_bytes, size_t bytes) {
    if (bytes <= *slot_bytes) return B200RS_OK;
    if (*slot) {
        B200RS_CUDA(cudaStreamSynchronize(dev->stream));
        B200RS_CUDA(cudaFree(*slot));
        *slot = nullptr;
        *slot_bytes = 0;
    }
    cudaError_t e = cudaMalloc(slot, bytes);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        *slot = nullptr;
        return e == cudaErrorMemoryAllocation ? B200RS_ERR_OUT_OF_MEMORY : (int)e;
    }
    *slot_bytes = bytes;
    return B200RS_OK;
}
