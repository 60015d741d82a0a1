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

    bool profiling = false;
    std::vector<b200rs_profile_span> spans;
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

static inline size_t b200rs_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// implemented in b200rs_device.cu
int b200rs_reserve(b200rs_device* dev, void** slot, size_t* slot_bytes, size_t bytes);
