// b200rs_internal.h -- shared by the .cu translation units of libb200rs.so (not installed).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "b200rs.h"

struct b200rs_profile_span {
    cudaEvent_t start, stop;
    b200rs_profile_entry entry;
};

// One CUDA device + one in-order stream (the reference's cl_context + cl_command_queue,
// Adl/CL/AdlCL.inl:284-303).
struct b200rs_device {
    int device_idx = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    int num_sms = 0;
    char name[128] = {0};
    uint64_t launches = 0;

    // grow-only scratch for the *_host entry points (role of Pprims' work buffers)
    void* scratch_data = nullptr;
    size_t scratch_data_bytes = 0;
    void* scratch_data2 = nullptr;
    size_t scratch_data2_bytes = 0;
    void* scratch_temp = nullptr;
    size_t scratch_temp_bytes = 0;
    uint32_t* pinned_word = nullptr;  // 1-word pinned mailbox (scan total)

    // copy streams + events of the pipelined host-batch entry points (created on first use)
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_sorted[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr}, ev_start = nullptr;

    // second stream of the partitioned sort (histograms next to the exchange kernel), created on first use
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_aux[2] = {nullptr, nullptr};
    // pipelined partitioned sort: copy-engine streams for the second half of the exchange, events that tie them to `stream`
    cudaStream_t copy[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_copied[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_pipe[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    void* pinned_plan = nullptr;  // plan of the pipelined partitioned sort, read back once per sort

    bool profiling = false;
    std::vector<b200rs_profile_span> spans;

    // kernels whose dynamic shared-memory limit has been raised on this device, with their cached occupancy
    // (cudaFuncSetAttribute / cudaOccupancyMaxActiveBlocksPerMultiprocessor are paid once per handle, not per call)
    struct kernel_setup { const void* func; int smem; int threads; int ctas_per_sm; };
    std::vector<kernel_setup> kernel_setups;
};

#define B200RS_CUDA(expr)                                   \
    do {                                                    \
        cudaError_t e_ = (expr);                            \
        if (e_ != cudaSuccess) {                            \
            (void)cudaGetLastError();                       \
            return (int)e_;                                 \
        }                                                   \
    } while (0)

#define B200RS_TRY(expr)                                    \
    do {                                                    \
        int r_ = (expr);                                    \
        if (r_ != B200RS_OK) return r_;                     \
    } while (0)

// RAII: binds the handle's device for the duration of a call and restores the previous one.
struct b200rs_device_guard {
    int prev = -1;
    bool changed = false;
    explicit b200rs_device_guard(const b200rs_device* d) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != d->device_idx) {
            changed = cudaSetDevice(d->device_idx) == cudaSuccess;
        }
    }
    ~b200rs_device_guard() {
        if (changed) cudaSetDevice(prev);
    }
};

// Brackets one kernel launch with events when profiling is on; always counts the launch.
struct b200rs_launch_scope {
    b200rs_device* dev;
    bool active;
    b200rs_profile_span span;
    b200rs_launch_scope(b200rs_device* d, const char* kernel, uint64_t elements, uint64_t bytes) : dev(d), active(d->profiling) {
        d->launches++;
        if (!active) return;
        memset(&span, 0, sizeof(span));
        strncpy(span.entry.kernel, kernel, sizeof(span.entry.kernel) - 1);
        span.entry.elements = elements;
        span.entry.bytes = bytes;
        if (cudaEventCreate(&span.start) != cudaSuccess || cudaEventCreate(&span.stop) != cudaSuccess) {
            active = false;
            return;
        }
        cudaEventRecord(span.start, d->stream);
    }
    ~b200rs_launch_scope() {
        if (!active) return;
        cudaEventRecord(span.stop, dev->stream);
        dev->spans.push_back(span);
    }
};

// Raises `func`'s dynamic shared-memory limit to `smem` bytes the first time the handle sees it; optionally returns the
// number of co-resident CTAs per SM (cached).  Called before every launch that needs more than 48 KiB.
static inline int b200rs_kernel_setup(b200rs_device* dev, const void* func, size_t smem, int threads = 0, int* ctas_per_sm = nullptr) {
    for (auto& k : dev->kernel_setups)
        if (k.func == func && k.smem == (int)smem && (!ctas_per_sm || k.threads == threads)) {
            if (ctas_per_sm) *ctas_per_sm = k.ctas_per_sm;
            return B200RS_OK;
        }
    if (smem > 48 * 1024) B200RS_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    if (ctas_per_sm) {
        B200RS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, func, threads, smem));
        *ctas_per_sm = per_sm;
    }
    dev->kernel_setups.push_back({func, (int)smem, ctas_per_sm ? threads : 0, per_sm});
    return B200RS_OK;
}

// Development knobs (environment variables) exist only in builds made with -DB200RS_EXPERIMENTS (make EXPERIMENTS=1, used
// by tools/sweep.py); the production library reads no environment variable, so nothing outside the call can change a result.
static inline int b200rs_exp_env(const char* name, int fallback) {
#ifdef B200RS_EXPERIMENTS
    const char* e = getenv(name);
    return e ? atoi(e) : fallback;
#else
    (void)name;
    return fallback;
#endif
}

static inline size_t b200rs_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// implemented in b200rs_device.cu
int b200rs_reserve(b200rs_device* dev, void** slot, size_t* slot_bytes, size_t bytes);
