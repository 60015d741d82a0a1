/*
 * b200rs.h -- C ABI of libb200rs.so: the B200 (sm_100a) implementation of OCLRadixSort's
 * data-parallel hot path (LSD radix sort of u32 keys and u32/u32 pairs, exclusive u32 scan).
 *
 * The reference has no FFI layer; its boundary is the header-level C++ API
 *     adl::DeviceUtils / adl::Device / adl::Buffer<T>      (Adl/Adl.h:71-222)
 *     Tahoe::Pprims::radixSort / Pprims::scan               (Tahoe/ParallelPrimitives/Pprims.h:35-41)
 * (paths relative to the reference tree).  The drop-in C++ headers under include/Adl and
 * include/Tahoe keep those spellings and forward to the entry points declared here; each entry
 * point cites the reference interface it stands in for.  Plain pointers and sizes only.
 *
 * Conventions
 *   - Every function returns int: 0 = B200RS_OK, > 0 = a cudaError_t value, < 0 = a B200RS_ERR_* code.
 *     Nothing throws; nothing falls back to the CPU.
 *   - A b200rs_device owns (or borrows) ONE CUDA stream; all work is enqueued on it in order and
 *     is asynchronous with respect to the host unless the function name ends in _sync or _host
 *     (reference contract: one in-order queue, visibility after DeviceUtils::waitForCompletion,
 *     Adl/CL/AdlCL.inl:303,567-570).
 *   - Not thread-safe per handle (neither is the reference: Pprims scratch, KernelManager map).
 *   - Device pointers must be 4-byte (keys, scan) / 8-byte (pairs) aligned; 16-byte alignment
 *     (any cudaMalloc pointer) enables the 128-bit paths.
 */
#ifndef B200RS_H
#define B200RS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200RS_VERSION 100 /* 0.1.0 */

#define B200RS_OK 0
#define B200RS_ERR_INVALID_ARGUMENT (-1) /* null handle/pointer, sort_bits outside [0,32], misaligned pointer */
#define B200RS_ERR_TEMP_TOO_SMALL (-2)   /* *temp_bytes smaller than the size query returned */
#define B200RS_ERR_NO_DEVICE (-3)        /* no CUDA device / device index out of range */
#define B200RS_ERR_UNSUPPORTED_ARCH (-4) /* device is not compute capability 10.x (binary is sm_100a only) */
#define B200RS_ERR_TOO_LARGE (-5)        /* n beyond what the entry point supports */
#define B200RS_ERR_OUT_OF_MEMORY (-6)    /* internal scratch allocation failed */
#define B200RS_ERR_CAPACITY (-7)         /* distributed sort: a rank's receive buffer is too small */

typedef struct b200rs_device b200rs_device; /* opaque */

/* 8-byte AoS pair, key first: Tahoe::SortData (Tahoe/Algorithm/Sort/RadixSort.h:10-21) and
 * Tahoe::uint2 {x=key,y=value} (Tahoe/Math/Math.h:175-188; Pprims.h:37). */
typedef struct b200rs_pair { uint32_t key; uint32_t value; } b200rs_pair;

int b200rs_version(void);
const char* b200rs_error_string(int code);

/* ---- device: replaces adl::DeviceUtils / adl::DeviceCL -------------------------------------- */

/* DeviceUtils::getNDevices(TYPE_CL), Adl/Adl.inl:8-23. */
int b200rs_device_count(int* count);
/* DeviceUtils::allocate(TYPE_CL, cfg) -> DeviceCL::initialize, Adl/Adl.inl:73-98, Adl/CL/AdlCL.inl:148-345.
 * Creates one in-order non-blocking stream (the cl_command_queue of AdlCL.inl:303). */
int b200rs_device_create(int device_idx, b200rs_device** dev);
/* Same, but enqueue on a stream the caller owns (cudaStream_t; e.g. torch's current stream). */
int b200rs_device_create_on_stream(int device_idx, void* cuda_stream, b200rs_device** dev);
/* DeviceUtils::deallocate, Adl/Adl.inl:100-105 (the used-memory assert lives in the C++ header). */
int b200rs_device_destroy(b200rs_device* dev);
/* DeviceUtils::waitForCompletion(device) = clFinish, Adl/CL/AdlCL.inl:567-570. */
int b200rs_device_sync(b200rs_device* dev);
/* DeviceUtils::getNCUs, Adl/Adl.inl:45-71 / AdlCL.inl:704-709 (compute units -> SMs). */
int b200rs_device_num_sms(const b200rs_device* dev, int* num_sms);
/* Device::getDeviceName, Adl/Adl.h:135. */
int b200rs_device_name(const b200rs_device* dev, char name_out[128]);
int b200rs_device_index(const b200rs_device* dev, int* device_idx);
int b200rs_device_mem_info(const b200rs_device* dev, size_t* free_bytes, size_t* total_bytes);
/* The cudaStream_t work is enqueued on (for event timing by the caller). */
void* b200rs_device_stream(const b200rs_device* dev);

/* ---- memory: replaces DeviceCL::allocate/deallocate/copy, Adl/CL/AdlCL.inl:356-510 ------------ */

int b200rs_malloc(b200rs_device* dev, size_t bytes, void** ptr);           /* clCreateBuffer, AdlCL.inl:356-410 */
int b200rs_free(b200rs_device* dev, void* ptr);                            /* clReleaseMemObject, :412-439 */
int b200rs_host_alloc(b200rs_device* dev, size_t bytes, void** host_ptr);  /* pinned staging for map/unmap, :544-565 */
int b200rs_host_free(b200rs_device* dev, void* host_ptr);
int b200rs_memcpy_h2d(b200rs_device* dev, void* dst, const void* host_src, size_t bytes); /* Buffer::write, :489-510 */
int b200rs_memcpy_d2h(b200rs_device* dev, void* host_dst, const void* src, size_t bytes); /* Buffer::read,  :466-487 */
int b200rs_memcpy_d2d(b200rs_device* dev, void* dst, const void* src, size_t bytes);      /* Buffer::write(Buffer&), :441-464 */
int b200rs_memset(b200rs_device* dev, void* ptr, int byte_value, size_t bytes);           /* Buffer::clear, :512-542 */

/* ---- the hot path: replaces Pprims::radixSort / Pprims::scan ---------------------------------- */

/*
 * Temp storage is CUB-style: call with temp == NULL to get the required size in *temp_bytes, then
 * call again with a device allocation of at least that size.  (In the reference the scratch is
 * the Pprims-owned uArray work buffers, Pprims.cpp:226-229,332-333.)
 *
 * Stable ascending unsigned LSD radix sort of the low sort_bits bits of each key, result left in
 * `inout`.  sort_bits in [0,32]; the reference's GPU path accepts multiples of 4 (Pprims.cpp:330),
 * every width is accepted here.  Any n >= 0 (the reference's key-only kernels need n % 256 == 0,
 * Pprims.cpp:327).  Replaces Pprims::radixSort(device, Buffer<u32>&, n, sortBits), Pprims.cpp:304-406.
 * Asynchronous like every entry point, with one exception: a 32-bit sort of 201 326 592 to 2^30 - 1 keys first takes a joint
 * histogram of the top 16 bits to choose between the key-only MSD pipeline and the LSD passes (the results are the same bits
 * either way: without a payload nothing distinguishes equal keys); the call returns once that histogram has run on the
 * device (a tenth of the sort) and the remaining kernels are queued -- it waits for work queued earlier on the stream too.
 */
int b200rs_sort_keys_u32(b200rs_device* dev, uint32_t* inout, uint64_t n, int sort_bits, void* temp, size_t* temp_bytes);
/*
 * The same sort (all 32 bits) with the algorithm chosen by the caller: the most-significant-digit-first pipeline that
 * b200rs_sort_keys_u32 picks by itself for large inputs (joint histogram of the top 16 bits, two unstable 8-bit partition
 * passes, counting sort of every bucket in shared memory; csrc/b200rs_msd.cuh) is tried at ANY n >= 2.  *used = 1 when it
 * sorted the keys, 0 when the input was not eligible (more than 12257 keys share their top 16 bits, inout not 16-byte
 * aligned, n >= 2^30) and the LSD path did.  Same result either way: key-only, so nothing depends
 * on stability.  Size query as above (this entry may need more temp than b200rs_sort_keys_u32 for the same n).
 * Blocks the host once (the eligibility test).  Pprims::radixSort(device, Buffer<u32>&, n), Pprims.cpp:304-406.
 */
int b200rs_sort_keys_u32_msd(b200rs_device* dev, uint32_t* inout, uint64_t n, void* temp, size_t* temp_bytes, int* used);
/* Same on AoS pairs; equal keys keep their input order.  Replaces
 * Pprims::radixSort(device, Buffer<uint2>&, n, sortBits), Pprims.cpp:200-302. */
int b200rs_sort_pairs_u32(b200rs_device* dev, b200rs_pair* inout, uint64_t n, int sort_bits, void* temp, size_t* temp_bytes);
/*
 * Exclusive prefix sum in u32 arithmetic (wraps mod 2^32): dst[0] = 0, dst[i] = src[0]+...+src[i-1].
 * dst may equal src.  total_out (device pointer, may be NULL) receives the sum of all n inputs.
 * Correct for every n (the reference returns without doing anything for n >= 1048576, Pprims.cpp:132-138).
 * Replaces Pprims::scan(device, dst, src, n, sumOut), Pprims.cpp:122-179.
 */
int b200rs_exclusive_scan_u32(b200rs_device* dev, uint32_t* dst, const uint32_t* src, uint64_t n, uint32_t* total_out,
                              void* temp, size_t* temp_bytes);

/*
 * Element-wise primitives of Tahoe/ParallelPrimitives: Pprims::copy / Pprims::fill,
 * Pprims.cpp:31-121 (dormant in the reference: commented out on the host, kernels still shipped as
 * CopyIntKernel / CopyF4Kernel / FillIntKernel / FillU32Kernel / FillF4Kernel, PprimsKernels.cl:9-48).
 * dst[i] = src[i] / dst[i] = value for i in [0, n).  u32 forms: 4-byte aligned pointers, any n; u128 forms
 * (float4 elements): 16-byte aligned pointers.  copy: dst and src must not partially overlap.
 */
int b200rs_copy_u32(b200rs_device* dev, uint32_t* dst, const uint32_t* src, uint64_t n);
int b200rs_copy_u128(b200rs_device* dev, void* dst, const void* src, uint64_t n);
int b200rs_fill_u32(b200rs_device* dev, uint32_t* dst, uint32_t value, uint64_t n);
int b200rs_fill_u128(b200rs_device* dev, void* dst, const uint32_t value[4], uint64_t n);

/*
 * Building blocks of the multi-GPU partitioned sort (new capability; the reference is single-device, SURVEY.md
 * section 8e).  Histogram of one key digit, and a STABLE partition of pairs by a 256-entry digit -> part table:
 * `out` receives part 0, then part 1, ... each in input order; part_counts[p] (device, 256 x u64, zero for unused
 * parts) must hold the number of elements of part p, e.g. summed from b200rs_digit_histogram_pairs.
 */
int b200rs_digit_histogram_pairs(b200rs_device* dev, const b200rs_pair* in, uint64_t n, int shift, int bits, uint64_t* hist_out);
int b200rs_partition_pairs(b200rs_device* dev, const b200rs_pair* in, b200rs_pair* out, uint64_t n, int shift, int bits,
                           const uint8_t* digit_to_part, const uint64_t* part_counts, void* temp, size_t* temp_bytes);

/*
 * Fused partition + exchange: like b200rs_partition_pairs, but part p is written starting at the absolute byte
 * address part_base_addr[p] (device array, 256 x u64, 8-byte aligned addresses).  The addresses may belong to OTHER
 * GPUs' buffers mapped with b200rs_ipc_import: the kernel then stores straight into peer memory over NVLink and
 * no separate all-to-all is needed.  Elements of a part keep their input order.
 */
int b200rs_scatter_pairs_to_parts(b200rs_device* dev, const b200rs_pair* in, uint64_t n, int shift, int bits,
                                  const uint8_t* digit_to_part, const uint64_t* part_base_addr, const uint64_t* n_dev /* optional: device-side
                                  element count <= n (0 = do nothing) */, void* temp, size_t* temp_bytes);
/*
 * The exchange step of the multi-GPU sort for FEW destinations (parts <= 32): the same stable partition as
 * b200rs_scatter_pairs_to_parts -- digit -> part through digit_to_part[256], part p's pairs appended in input order at byte
 * address part_base_addr[p] (any memory this device can store to: local, or a peer's, mapped through b200rs_ipc_import) -- by
 * a kernel built for long runs: ranking in registers, one look-back word per (tile, part), one bulk copy (TMA) per part and
 * tile.  n_dev (device pointer, may be NULL): the number of pairs is min(n, *n_dev).  Temp: size query as usual.
 * New capability (the reference is single-device, Adl/CL/AdlCL.inl:284-303); SURVEY.md section 8e step 3.
 */
int b200rs_exchange_pairs(b200rs_device* dev, const b200rs_pair* in, uint64_t n, int shift, int bits, const uint8_t* digit_to_part,
                          const uint64_t* part_base_addr, int parts, const uint64_t* n_dev, void* temp, size_t* temp_bytes);
/*
 * The same exchange with destinations that own arbitrary KEY RANGES: the part of a pair is the number of thresholds
 * splitters[0 .. parts - 2] (device, ascending 64-bit values; 2^32 = "no key reaches it") its key is >= to.  Used by the
 * splitter plan of oclradixsort_b200/dist.py, which finds exact quantile keys -- and splits a key that alone exceeds a
 * rank's share by source rank, through per-source thresholds -- so that any distribution is balanced.
 */
int b200rs_exchange_pairs_by_splitters(b200rs_device* dev, const b200rs_pair* in, uint64_t n, const uint64_t* splitters,
                                       const uint64_t* part_base_addr, int parts, void* temp, size_t* temp_bytes);
/*
 * hist_out[j][d] (device, count x 256 x u64) = pairs whose key bits above the digit at `shift` equal prefixes[j] (device,
 * count <= 31 values) and whose digit (key >> shift) & 255 is d; shift in {0, 8, 16, 24} (24: the prefixes must be 0).  One
 * refinement round of the splitter plan.
 */
int b200rs_filtered_histograms_pairs(b200rs_device* dev, const b200rs_pair* in, uint64_t n, int shift, const uint32_t* prefixes, int count,
                                     uint64_t* hist_out);
/*
 * Exchange plan computed on the device (the unpipelined multi-GPU sort needs no host round trip): from the all-gathered
 * top-digit histograms hist_all[world][256] it derives contiguous digit ranges per rank (about N/world pairs each),
 * lut_out[256] (digit -> destination rank), part_base_out[256] (where THIS rank's pairs for destination d start:
 * peer_base[d] + 8 * pairs sent to d by lower ranks), counts_out[0] = n_in (or 0 if aborted), counts_out[1] = pairs this
 * rank will receive, status_out[0] = 1 if some rank's share exceeds `capacity` pairs (then nothing is exchanged).
 */
int b200rs_dist_plan(b200rs_device* dev, const uint64_t* hist_all, int world, int rank, const uint64_t* peer_base, uint64_t capacity,
                     uint64_t n_in, uint8_t* lut_out, uint64_t* part_base_out, uint64_t* counts_out, uint32_t* status_out);
/*
 * Plan of the PIPELINED partitioned sort (see b200rs_dist_sort_pairs_u32): the digit ranges of b200rs_dist_plan, every
 * destination's range cut once more into halves A | B at the digit boundary where A holds closest to a_permille / 1000 of
 * the destination's pairs.  lut_out[256]: digit -> part = 2 x destination + half; part_base_out[2 x world]: where THIS rank's
 * pairs of each part go -- (d, A): inside d's receive buffer; (d, B), d != rank: inside this rank's staging area at
 * stage_base (runs on 128-byte boundaries); (rank, B): its final place; counts_out / status_out as b200rs_dist_plan.
 * plan_out (device, 52 x u64, what the host reads back): [0] status, [1] pairs this rank receives, [2] of them in half A,
 * [3] in half B, [4 + d] staging offset (pairs) of destination d's half-B run, [20 + d] its length, [36 + d] the byte
 * address it is copied to in d's receive buffer.  world <= 16.  Host mirror: plan_exchange_halves() in oclradixsort_b200/dist.py.
 */
int b200rs_dist_plan_halves(b200rs_device* dev, const uint64_t* hist_all, int world, int rank, const uint64_t* peer_base, uint64_t capacity,
                            uint64_t stage_base, uint64_t n_in, int a_permille, uint8_t* lut_out, uint64_t* part_base_out, uint64_t* counts_out,
                            uint32_t* status_out, uint64_t* plan_out);
/* b200rs_sort_pairs_u32 whose element count min(n_max, *n_dev) is read on the device (temp is sized for n_max).  Only the
 * first *n_dev elements are sorted; with an ODD number of passes (sort_bits <= 8 or in 17..24) the final copy back covers
 * n_max elements, so inout[*n_dev .. n_max) then holds unspecified values. */
int b200rs_sort_pairs_u32_devn(b200rs_device* dev, b200rs_pair* inout, uint64_t n_max, const uint64_t* n_dev, int sort_bits,
                               void* temp, size_t* temp_bytes);
/*
 * The whole partitioned sort of one rank as ONE call (replaces nothing in the reference, which is
 * single-device: Adl/CL/AdlCL.inl:284-303; this is what Tahoe::Pprims::radixSortDistributed calls).  One rank per GPU;
 * rank r's `in` slices, in rank order, are the global input order.  The caller supplies the two collectives the path
 * needs -- the library itself links no communication library:
 *   allgather(user, send_dev, recv_dev, bytes)  every rank contributes `bytes` from send_dev; recv_dev receives world x bytes
 *                                               in rank order.  Device pointers.  Must be ordered with the handle's stream
 *                                               (enqueue on that stream, or synchronise it before and return when done).
 *   barrier(user)                               returns (or is stream-ordered) after every rank's work enqueued so far on its
 *                                               handle's stream has completed: peer stores have landed.
 * Steps: top-digit histogram -> allgather -> on-device plan (contiguous digit ranges of about N / world pairs per rank) ->
 * exchange straight into the ranks' receive buffers recv_base[0 .. world) (device-visible addresses of EVERY rank's receive
 * buffer, own included; peers' through b200rs_ipc_import or peer access) -> barrier -> local stable sort of what arrived.
 * The sorted pairs of this rank are left at recv_base[rank]; counts_dev[1] (device, 2 x u64) = how many; status_dev[0] = 1
 * when a rank's share exceeds recv_capacity_pairs (nothing is exchanged then; the caller re-plans, e.g.
 * oclradixsort_b200/dist.py's splitter path).  Equal keys keep global input order: the concatenation of the ranks' outputs
 * is the stable sort of the concatenated input.
 * Two forms, chosen from recv_capacity_pairs and world only (the same on every rank, so all ranks run the same collectives):
 *   - below 2^22 pairs of capacity, or beyond 16 ranks: one exchange kernel (b200rs_exchange_pairs), one barrier, one local
 *     sort; no host round trip inside (collectives per call: 1 or 2 allgathers, 1 barrier);
 *   - otherwise PIPELINED: every destination's digit range is cut in two halves A | B.  One exchange kernel stores half A over
 *     NVLink and stages half B in this GPU's memory; copy engines then move half B to the peers WHILE half A is sorted on a
 *     second stream; half B is sorted when it has landed.  The host reads the plan back once (sizes of the copies and of the
 *     two sorts: the call blocks until the plan kernel has run, the rest is stream-ordered as before; 1 allgather, 2 barriers).
 * Temp: size query as usual (recv_capacity_pairs must be the same on every rank).
 */
typedef int (*b200rs_allgather_fn)(void* user, const void* send_dev, void* recv_dev, size_t bytes);
typedef int (*b200rs_barrier_fn)(void* user);
typedef struct b200rs_dist_comm {
    int rank, world;
    b200rs_allgather_fn allgather;
    b200rs_barrier_fn barrier;
    void* user;
} b200rs_dist_comm;
int b200rs_dist_sort_pairs_u32(b200rs_device* dev, const b200rs_dist_comm* comm, const uint64_t* recv_base /* host array [world] */,
                               uint64_t recv_capacity_pairs, const b200rs_pair* in, uint64_t n, uint64_t* counts_dev, uint32_t* status_dev,
                               void* temp, size_t* temp_bytes);
/* One process driving several GPUs (e.g. one host thread per rank): lets `dev` load from / store to memory of CUDA device
 * peer_device_idx directly (cudaDeviceEnablePeerAccess; already-enabled is not an error). */
int b200rs_enable_peer_access(b200rs_device* dev, int peer_device_idx);
/* CUDA IPC for one-process-per-GPU jobs: export a b200rs_malloc'ed buffer, map a peer's buffer, unmap it. */
int b200rs_ipc_export(b200rs_device* dev, void* ptr, unsigned char handle_out[64]);
int b200rs_ipc_import(b200rs_device* dev, const unsigned char handle[64], void** ptr);
int b200rs_ipc_release(b200rs_device* dev, void* ptr);

/*
 * HOST-buffer forms: what a caller holding CPU arrays uses in place of the reference's
 * "getHostPtr / fill / returnHostPtr / radixSort / getHostPtr / read" sequence
 * (UnitTest/main.cpp:118-139).  Host -> device copy, sort/scan, device -> host copy, stream sync;
 * device buffers and temp storage are grow-only scratch owned by the handle (like Pprims' work
 * buffers) and released by b200rs_device_release_scratch / b200rs_device_destroy.
 * Pinned host memory (b200rs_host_alloc) makes the copies run at full link speed.
 */
int b200rs_sort_keys_u32_host(b200rs_device* dev, uint32_t* host_inout, uint64_t n, int sort_bits);
int b200rs_sort_pairs_u32_host(b200rs_device* dev, b200rs_pair* host_inout, uint64_t n, int sort_bits);
int b200rs_exclusive_scan_u32_host(b200rs_device* dev, uint32_t* host_dst, const uint32_t* host_src, uint64_t n,
                                   uint32_t* host_total_out);
/* Batch forms: `count` independent host arrays of n elements each, every one sorted in place -- the size sweep of
 * UnitTest/main.cpp:105-171 run as one call.  Pipelined over three streams and two device buffers: array i+1 is copied
 * in and array i-1 copied out while array i is sorted, so both directions of the host link are busy at once.  Use pinned
 * host memory (b200rs_host_alloc); with pageable memory the copies serialise. */
int b200rs_sort_keys_u32_host_batch(b200rs_device* dev, uint32_t* const* host_inout, int count, uint64_t n, int sort_bits);
int b200rs_sort_pairs_u32_host_batch(b200rs_device* dev, b200rs_pair* const* host_inout, int count, uint64_t n, int sort_bits);
int b200rs_device_release_scratch(b200rs_device* dev);

/* ---- per-launch timing: replaces Device::toggleProfiling, Adl/Adl.h:142 + AdlKernelUtilsCL.inl:654-677 */

typedef struct b200rs_profile_entry {
    char kernel[48];   /* kernel name, e.g. "onesweep_keys_pass2" */
    float ms;          /* device time between CUDA events on the handle's stream */
    uint64_t elements; /* elements the launch processed */
    uint64_t bytes;    /* algorithmic bytes of the launch */
} b200rs_profile_entry;

/* While enabled, every kernel the library launches is bracketed by CUDA events. */
int b200rs_profile_enable(b200rs_device* dev, int enable);
/* Synchronises the stream and moves up to `capacity` of the oldest recorded entries to `out`; *count = how many
 * (== capacity means more may be queued: call again).  capacity 0 just synchronises. */
int b200rs_profile_read(b200rs_device* dev, b200rs_profile_entry* out, int capacity, int* count);
/* Number of kernels launched through this handle since creation (bench.py's gpu_launches). */
int b200rs_device_launch_count(const b200rs_device* dev, uint64_t* launches);

/* ---- device-side interval timing: what adl::Stopwatch measures with (Adl/AdlStopwatch.h:27-83; the reference's
 * CL stopwatch is a host clock, Adl/AdlStopwatch.inl:16-19 -- asynchronous launches need events on the stream) ---- */
int b200rs_event_create(b200rs_device* dev, void** event_out);  /* a cudaEvent_t */
int b200rs_event_record(b200rs_device* dev, void* event);       /* on the handle's stream */
/* Waits for stop_event, then *ms_out = device time between the two events. */
int b200rs_event_elapsed_ms(b200rs_device* dev, void* start_event, void* stop_event, float* ms_out);
int b200rs_event_destroy(b200rs_device* dev, void* event);
/* *done = 1 when everything enqueued before the event's last record has finished (never blocks); b200rs_event_synchronize
 * blocks until then.  Used by adl::Buffer's non-blocking unmap (the pinned stage of returnHostPtr is recycled when its
 * event has passed; reference: non-blocking clEnqueueUnmapMemObject, AdlCL.inl:557-565). */
int b200rs_event_query(b200rs_device* dev, void* event, int* done);
int b200rs_event_synchronize(b200rs_device* dev, void* event);

#ifdef __cplusplus
}
#endif
#endif /* B200RS_H */
