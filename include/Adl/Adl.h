// Adl/Adl.h -- the reference's device / buffer wrapper (Adl/Adl.h:39-274, Adl/Adl.inl, Adl/AdlKernel.h)
// re-implemented as a thin header over the C ABI of libb200rs.so (include/b200rs.h).
//
// What is kept: every spelling UnitTest/main.cpp, Pprims and uArray use -- namespace adl, DeviceType
// {TYPE_CL, TYPE_DX11, TYPE_HOST}, DeviceUtils::{Config, allocate, deallocate, waitForCompletion, getNCUs,
// getNDevices, flush}, Device, Buffer<T> with its public fields, HostBuffer<T>, Launcher::BufferInfo,
// SyncObject, and `extern char s_cacheDirectory[128]` which the application defines (main.cpp:74).
// What is different by design: one backend.  TYPE_CL creates a CUDA device (one stream = the
// reference's one in-order command queue); TYPE_HOST / TYPE_DX11 are refused -- there is no CPU
// device and no multi-backend dispatch.  Kernels are compiled ahead of time into the library, so
// Device::getKernel / Launcher::launch* exist only so that dependent code compiles; calling them asserts.
#ifndef ADL_H
#define ADL_H

#include <limits.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>

#include <Adl/AdlConfig.h>
#include <Tahoe/Math/Error.h>
#include <b200rs.h>

namespace adl {

typedef unsigned long long u64;

extern char s_cacheDirectory[128];  // defined by the application; unused (no run-time kernel cache)

#define ADL_SUCCESS 0
#define ADL_FAILURE 1

template <typename T> inline T max2(const T& a, const T& b) { return a > b ? a : b; }
template <typename T> inline T min2(const T& a, const T& b) { return a < b ? a : b; }

enum DeviceType {
    TYPE_CL = 0,  // the GPU device: CUDA / sm_100a here
    TYPE_DX11 = 1,
    TYPE_HOST,
};

struct Device;
struct SyncObject;
struct Kernel;

struct BufferBase {
    enum BufferType {
        BUFFER,
        BUFFER_CONST, BUFFER_STAGING, BUFFER_APPEND, BUFFER_RAW, BUFFER_W_COUNTER, BUFFER_INDEX, BUFFER_VERTEX,  // DX11 kinds, unused
        BUFFER_ZERO_COPY,  // CL_MEM_ALLOC_HOST_PTR of the CL backend (AdlCL.inl:381-382): PINNED HOST memory the GPU reads and writes in place
                           // over the host link (same pointer on both sides, unified addressing); getHostPtr returns it without a copy
    };
};

// Prints the library's message for a non-zero C-ABI return code and records the failure.
inline bool adlCheck(int code, const char* where) {
    if (code == B200RS_OK) return true;
    fprintf(stderr, "adl: %s failed: %d (%s)\n", where, code, b200rs_error_string(code));
    ADLASSERT(code == B200RS_OK);
    return false;
}

class DeviceUtils {
public:
    struct Config {
        enum DeviceType { DEVICE_GPU, DEVICE_CPU };
        enum DeviceVendor { VD_AMD, VD_INTEL, VD_NV };
        Config() : m_type(DEVICE_GPU), m_deviceIdx(0), m_vendor(VD_NV), m_clContextProperties(0) {}
        DeviceType m_type;
        int m_deviceIdx;
        DeviceVendor m_vendor;         // ignored
        void* m_clContextProperties;   // ignored
    };

    static inline int getNDevices(DeviceType type);
    static inline int getNCUs(const Device* device);
    static inline Device* allocate(DeviceType type, Config cfg = Config());
    static inline void deallocate(Device* device);
    static inline void waitForCompletion(const Device* device);
    static inline void waitForCompletion(const SyncObject* syncObj);
    static inline bool isComplete(const SyncObject* syncObj);
    static inline void flush(const Device* device);
};

// One CUDA device + one in-order stream.  The reference's Device is an abstract base with CL / DX11 /
// Host subclasses; with a single backend it is a concrete struct holding the C-ABI handle.
struct Device {
    typedef DeviceUtils::Config Config;

    explicit Device(DeviceType type)
        : m_type(type), m_procType(Config::DEVICE_GPU), m_memoryUsage(0), m_interopAvailable(false), m_enableProfiling(false),
          m_binaryFileVersion(0), m_handle(0) {
        for (int i = 0; i < NUM_STAGES; ++i) { m_stages[i].p = 0; m_stages[i].bytes = 0; m_stages[i].event = 0; m_stages[i].mapped = false; m_stages[i].inFlight = false; }
    }
    virtual ~Device() {}

    virtual void* getContext() const { return m_handle; }
    virtual void initialize(const Config& cfg) {
        int count = 0;
        b200rs_device_count(&count);
        const int idx = count > 0 ? min2(count - 1, max2(cfg.m_deviceIdx, 0)) : 0;  // clamped like AdlCL.inl:244
        adlCheck(b200rs_device_create(idx, &m_handle), "b200rs_device_create");
        m_procType = Config::DEVICE_GPU;
    }
    virtual void release() {
        if (m_handle) b200rs_device_sync(m_handle);  // copies out of the stages may still be in flight
        for (int i = 0; i < NUM_STAGES; ++i) {
            if (m_stages[i].p) b200rs_host_free(m_handle, m_stages[i].p);
            if (m_stages[i].event) b200rs_event_destroy(m_handle, m_stages[i].event);
            m_stages[i].p = 0; m_stages[i].bytes = 0; m_stages[i].event = 0; m_stages[i].mapped = false; m_stages[i].inFlight = false;
        }
        if (m_handle) adlCheck(b200rs_device_destroy(m_handle), "b200rs_device_destroy");
        m_handle = 0;
    }
    virtual void waitForCompletion() const {
        if (m_handle) adlCheck(b200rs_device_sync(m_handle), "b200rs_device_sync");
    }
    virtual void waitForCompletion(const SyncObject*) const { waitForCompletion(); }
    virtual bool isComplete(const SyncObject*) const { waitForCompletion(); return true; }
    virtual void flush() const {}  // work is submitted eagerly; nothing to flush
    virtual void getDeviceName(char nameOut[128]) const {
        nameOut[0] = 0;
        if (m_handle) b200rs_device_name(m_handle, nameOut);
    }
    virtual void getDeviceVendor(char nameOut[128]) const { strncpy(nameOut, "NVIDIA Corporation", 128); }
    // Kernels are compiled ahead of time into libb200rs.so: there is nothing to look up at run time.
    virtual Kernel* getKernel(const char*, const char*, const char* = NULL, const char** = NULL, int = 0, bool = true) const {
        ADLASSERT(0);
        return 0;
    }
    virtual u64 getUsedMemory() const { return m_memoryUsage; }
    virtual u64 getMaxAllocationSize() const { return ULLONG_MAX; }

    // per-launch device timing (reference: toggleProfiling + ProfileCL.*.csv, AdlKernelUtilsCL.inl:654-677);
    // read the log with b200rs_profile_read(getHandle(), ...)
    void toggleProfiling(bool enable) {
        m_enableProfiling = enable;
        if (m_handle) b200rs_profile_enable(m_handle, enable ? 1 : 0);
    }
    // Appends the per-launch log recorded since the last call to `path` (default "ProfileB200.<device name>.csv"),
    // one line per launch: "kernel","ms","elements","bytes","GB/s" -- the reference appends
    // "kernel","ms","global size x","y","y" to ProfileCL.<device>.<driver>.csv per launch (AdlKernelUtilsCL.inl:663-676).
    // Waits for the stream.  Returns the number of launches written.
    int writeProfileCsv(const char* path = 0) {
        if (!m_handle) return 0;
        char name[128], file[256];
        getDeviceName(name);
        for (char* c = name; *c; ++c) if (*c == ' ' || *c == '/') *c = '_';
        snprintf(file, sizeof(file), "ProfileB200.%s.csv", name);
        FILE* f = fopen(path ? path : file, "a");
        if (!f) return 0;
        int total = 0, got = 0;
        b200rs_profile_entry e[64];
        do {
            if (!adlCheck(b200rs_profile_read(m_handle, e, 64, &got), "b200rs_profile_read")) break;
            for (int i = 0; i < got; ++i)
                fprintf(f, "\"%s\",\"%g\",\"%llu\",\"%llu\",\"%g\"\n", e[i].kernel, e[i].ms, (unsigned long long)e[i].elements,
                        (unsigned long long)e[i].bytes, e[i].ms > 0.f ? (double)e[i].bytes / e[i].ms * 1e-6 : 0.0);
            total += got;
        } while (got == 64);
        fclose(f);
        return total;
    }
    void setBinaryFileVersion(unsigned int ver) { m_binaryFileVersion = ver; }
    unsigned int getBinaryFileVersion() const { return m_binaryFileVersion; }
    DeviceType getType() const { return m_type; }
    Config::DeviceType getProcType() const { return m_procType; }
    b200rs_device* getHandle() const { return m_handle; }

    // Pinned staging for Buffer::getHostPtr / returnHostPtr: a ring of NUM_STAGES grow-only blocks per device.
    //   map   (getHostPtr)    takes a block nobody holds; the device -> host copy into it is stream-ordered, the caller
    //                         waits (DeviceUtils::waitForCompletion) before reading, as with the reference's non-blocking
    //                         clEnqueueMapBuffer (AdlCL.inl:544-555)
    //   unmap (returnHostPtr) enqueues the host -> device copy out of the block and RETURNS (non-blocking
    //                         clEnqueueUnmapMemObject, AdlCL.inl:557-565): the block carries an event and is handed out
    //                         again only after that event has passed, so the host never overwrites bytes a copy still reads
    // A mapping that finds every block held by other live mappings gets a block of its own.
    void* acquireStage(size_t bytes, int* slot) const {
        int pick = -1;
        for (int pass = 0; pass < 2 && pick < 0; ++pass)
            for (int i = 0; i < NUM_STAGES && pick < 0; ++i) {
                Stage& st = m_stages[i];
                if (st.mapped) continue;
                if (st.inFlight) {
                    int done = 0;
                    if (pass == 0) b200rs_event_query(m_handle, st.event, &done);
                    else done = b200rs_event_synchronize(m_handle, st.event) == B200RS_OK;  // second pass: wait for the first block that is only in flight
                    if (!done) continue;
                    st.inFlight = false;
                }
                pick = i;
            }
        if (pick < 0) {  // every block is mapped by somebody else
            void* p = 0;
            adlCheck(b200rs_host_alloc(m_handle, bytes, &p), "b200rs_host_alloc");
            *slot = -1;
            return p;
        }
        Stage& st = m_stages[pick];
        if (bytes > st.bytes) {
            if (st.p) b200rs_host_free(m_handle, st.p);
            st.p = 0;
            st.bytes = 0;
            if (!adlCheck(b200rs_host_alloc(m_handle, bytes, &st.p), "b200rs_host_alloc")) return 0;
            st.bytes = bytes;
        }
        st.mapped = true;
        *slot = pick;
        return st.p;
    }
    // `copyEnqueued`: a stream-ordered copy that reads the block was just enqueued (unmap); the block is recycled after it
    void releaseStage(void* p, int slot, bool copyEnqueued) const {
        if (slot < 0) {  // a private block: nothing else can wait for it
            if (copyEnqueued) waitForCompletion();
            b200rs_host_free(m_handle, p);
            return;
        }
        Stage& st = m_stages[slot];
        st.mapped = false;
        if (copyEnqueued) {
            if (!st.event) adlCheck(b200rs_event_create(m_handle, &st.event), "b200rs_event_create");
            st.inFlight = st.event && b200rs_event_record(m_handle, st.event) == B200RS_OK;
            if (!st.inFlight) waitForCompletion();
        }
    }

    DeviceType m_type;
    Config::DeviceType m_procType;
    mutable u64 m_memoryUsage;  // bytes held by Buffers; must be 0 at DeviceUtils::deallocate
    bool m_interopAvailable;
    bool m_enableProfiling;
    unsigned int m_binaryFileVersion;

private:
    enum { NUM_STAGES = 2 };
    struct Stage { void* p; size_t bytes; void* event; bool mapped; bool inFlight; };
    b200rs_device* m_handle;
    mutable Stage m_stages[NUM_STAGES];
};

// Typed device allocation.  Public fields and their order follow the reference (Adl.h:201-220) so that
// Buffer<SortData> / Buffer<uint2> / Buffer<u32> can be reinterpreted into each other (main.cpp:157, Pprims).
template <typename T>
struct Buffer : public BufferBase {
    inline Buffer();
    inline Buffer(const Device* device, u64 nElems, BufferType type = BUFFER);
    inline virtual ~Buffer();

    inline void setRawPtr(const Device* device, T* ptr, u64 size, BufferType type = BUFFER);  // wrap, do not own
    inline void allocate(const Device* device, u64 nElems, BufferType type = BUFFER);
    inline void write(const T* hostSrcPtr, u64 nElems, u64 dstOffsetNElems = 0, SyncObject* syncObj = 0);
    inline void read(T* hostDstPtr, u64 nElems, u64 srcOffsetNElems = 0, SyncObject* syncObj = 0) const;
    inline void write(const Buffer<T>& src, u64 nElems, SyncObject* syncObj = 0);
    inline void read(Buffer<T>& dst, u64 nElems, u64 offsetNElems = 0, SyncObject* syncObj = 0) const;
    inline void clear();
    inline void fill(void* pattern, int patternSize);
    // map / unmap with read+write semantics (reference: clEnqueueMapBuffer, AdlCL.inl:544-565): the
    // returned host view is valid after DeviceUtils::waitForCompletion; returnHostPtr writes it back.
    inline T* getHostPtr(u64 size = (u64)-1) const;
    inline void returnHostPtr(T* ptr) const;
    inline void setSize(u64 size);  // grow-only; contents are not preserved
    u64 getSize() const { return m_size; }
    DeviceType getType() const { ADLASSERT(m_device != 0); return m_device->m_type; }

    const Device* m_device;
    u64 m_size;  // elements
    T* m_ptr;    // device pointer
    union {
        struct { void* m_uav; void* m_srv; } m_dx11;  // layout filler, unused
        struct { char* m_hostPtr; } m_cl;             // pinned host view while mapped
    };
    bool m_allocated;

    // Contents are defined on the device (something was written through adl / Pprims / uArray since the allocation).  A
    // buffer that never was is mapped without the device -> host copy: the reference caller's first step is "map a fresh
    // buffer, fill it on the host, unmap" (UnitTest/main.cpp:118-126).  Code that writes through the raw m_ptr calls this.
    void markDeviceWritten() const { m_deviceWritten = true; }
    bool isZeroCopy() const { return m_zeroCopy; }

private:
    mutable u64 m_mappedElems;
    mutable int m_mapSlot;
    mutable bool m_deviceWritten;
    bool m_zeroCopy;
    inline void releaseStorage();
};

class BufferUtils {
public:
    // With one backend a buffer is always "native": map returns it unchanged, unmap does nothing.
    template <DeviceType TYPE, bool COPY, typename T>
    static Buffer<T>* map(const Device*, const Buffer<T>* in, int = -1) { return const_cast<Buffer<T>*>(in); }
    template <bool COPY, typename T>
    static void unmap(Buffer<T>*, const Buffer<T>*, int = -1) {}
};

// Device-resident buffer with HostBuffer's name: element access maps one element at a time and is
// meant for debugging only.
template <typename T>
struct HostBuffer : public Buffer<T> {
    HostBuffer() : Buffer<T>() {}
    HostBuffer(const Device* device, int nElems, BufferBase::BufferType type = BufferBase::BUFFER) : Buffer<T>(device, nElems, type) {}
};

struct Kernel {
    DeviceType m_type;
    void* m_kernel;
    const char* m_funcName;
};

struct SyncObject {
    explicit SyncObject(const Device* device) : m_device(device), m_ptr(0) {}
    const Device* m_device;
    void* m_ptr;
};

// Argument-binding launcher of the reference (AdlKernel.h:59-143).  Kept as a type because uArray names
// it; there are no run-time kernels to launch, so launch1D / launch2D assert.
class Launcher {
public:
    struct BufferInfo {
        BufferInfo() : m_buffer(0), m_isReadOnly(false) {}
        template <typename T> BufferInfo(Buffer<T>* buff, bool isReadOnly = false) : m_buffer(buff), m_isReadOnly(isReadOnly) {}
        template <typename T> BufferInfo(const Buffer<T>* buff, bool isReadOnly = false) : m_buffer((void*)buff), m_isReadOnly(isReadOnly) {}
        void* m_buffer;
        bool m_isReadOnly;
    };
    enum { MAX_ARG_SIZE = 16, MAX_ARG_COUNT = 64 };

    Launcher(const Device* dd, const Kernel* kernel) : m_deviceData(dd), m_kernel(kernel), m_idx(0), m_idxRw(0) {}
    void setBuffers(BufferInfo* buffInfo, int n) { (void)buffInfo; m_idx += n; }
    template <typename T> void setConst(const T&) { ++m_idx; }
    void launch1D(int, int = 64, SyncObject* = 0) { ADLASSERT(0); }
    void launch2D(int, int, int = 8, int = 8, SyncObject* = 0) { ADLASSERT(0); }

    const Device* m_deviceData;
    const Kernel* m_kernel;
    int m_idx;
    int m_idxRw;
};

// ---- DeviceUtils ---------------------------------------------------------------------------------

int DeviceUtils::getNDevices(DeviceType type) {
    if (type != TYPE_CL) return 0;
    int n = 0;
    b200rs_device_count(&n);
    return n;
}

int DeviceUtils::getNCUs(const Device* device) {
    int n = 0;
    if (device && device->getHandle()) b200rs_device_num_sms(device->getHandle(), &n);
    return n;
}

Device* DeviceUtils::allocate(DeviceType type, Config cfg) {
    if (type != TYPE_CL || cfg.m_type != Config::DEVICE_GPU) {
        // no Host (CPU) or DX11 device: fail loudly instead of silently computing on the CPU
        fprintf(stderr, "adl: only the GPU device exists in this build (TYPE_CL + DEVICE_GPU -> CUDA sm_100a); requested type %d\n", (int)type);
        ADLASSERT(0);
        return 0;
    }
    Device* d = new Device(type);
    d->m_interopAvailable = false;
    d->initialize(cfg);
    if (!d->getHandle()) {
        delete d;
        return 0;
    }
    return d;
}

void DeviceUtils::deallocate(Device* device) {
    if (!device) return;
    ADLASSERT(device->getUsedMemory() == 0);  // every Buffer (and Pprims scratch) must be gone by now
    device->release();
    delete device;
}

void DeviceUtils::waitForCompletion(const Device* device) {
    if (device) device->waitForCompletion();
}
void DeviceUtils::waitForCompletion(const SyncObject* syncObj) {
    if (syncObj) syncObj->m_device->waitForCompletion(syncObj);
}
bool DeviceUtils::isComplete(const SyncObject* syncObj) { return syncObj ? syncObj->m_device->isComplete(syncObj) : true; }
void DeviceUtils::flush(const Device* device) {
    if (device) device->flush();
}

// ---- Buffer<T> -------------------------------------------------------------------------------------

template <typename T>
Buffer<T>::Buffer() : m_device(0), m_size(0), m_ptr(0), m_allocated(false), m_mappedElems(0), m_mapSlot(-1), m_deviceWritten(false), m_zeroCopy(false) {
    m_dx11.m_uav = 0;
    m_dx11.m_srv = 0;
}

template <typename T>
Buffer<T>::Buffer(const Device* device, u64 nElems, BufferType type)
    : m_device(0), m_size(0), m_ptr(0), m_allocated(false), m_mappedElems(0), m_mapSlot(-1), m_deviceWritten(false), m_zeroCopy(false) {
    m_dx11.m_uav = 0;
    m_dx11.m_srv = 0;
    allocate(device, nElems, type);
}

template <typename T>
void Buffer<T>::releaseStorage() {
    if (m_allocated && m_ptr && m_device) {
        if (m_zeroCopy) {
            m_device->waitForCompletion();  // kernels may still be using the host block
            adlCheck(b200rs_host_free(m_device->getHandle(), m_ptr), "b200rs_host_free");
        } else {
            adlCheck(b200rs_free(m_device->getHandle(), m_ptr), "b200rs_free");
        }
        m_device->m_memoryUsage -= m_size * sizeof(T);
    }
    m_ptr = 0;
    m_size = 0;
    m_allocated = false;
    m_zeroCopy = false;
    m_deviceWritten = false;
}

template <typename T>
Buffer<T>::~Buffer() {
    releaseStorage();
    m_device = 0;
}

template <typename T>
void Buffer<T>::setRawPtr(const Device* device, T* ptr, u64 size, BufferType type) {
    ADLASSERT(type == BUFFER);
    ADLASSERT(!m_allocated);
    if (m_device) ADLASSERT(m_device == device);
    m_device = device;
    m_ptr = ptr;
    m_size = size;
    m_deviceWritten = true;  // foreign memory: assume it holds data
}

template <typename T>
void Buffer<T>::allocate(const Device* device, u64 nElems, BufferType type) {
    ADLASSERT(m_device == 0 || m_device == device);
    ADLASSERT(!m_allocated);
    m_device = device;
    m_size = 0;
    m_ptr = 0;
    if (nElems == 0 || device == 0) return;
    void* p = 0;
    m_zeroCopy = type == BUFFER_ZERO_COPY;
    if (m_zeroCopy) {  // pinned host memory, device-accessible under the same address
        if (!adlCheck(b200rs_host_alloc(device->getHandle(), nElems * sizeof(T), &p), "b200rs_host_alloc")) { m_zeroCopy = false; return; }
    } else if (!adlCheck(b200rs_malloc(device->getHandle(), nElems * sizeof(T), &p), "b200rs_malloc")) {
        return;  // m_ptr = 0, m_size = 0 like AdlCL.inl:390-406
    }
    m_ptr = (T*)p;
    m_size = nElems;
    m_allocated = true;
    m_deviceWritten = false;
    device->m_memoryUsage += nElems * sizeof(T);
}

template <typename T>
void Buffer<T>::write(const T* hostSrcPtr, u64 nElems, u64 dstOffsetNElems, SyncObject*) {
    if (nElems == 0) return;
    ADLASSERT(nElems + dstOffsetNElems <= m_size);
    m_deviceWritten = true;
    adlCheck(b200rs_memcpy_h2d(m_device->getHandle(), m_ptr + dstOffsetNElems, hostSrcPtr, nElems * sizeof(T)), "b200rs_memcpy_h2d");
}

template <typename T>
void Buffer<T>::read(T* hostDstPtr, u64 nElems, u64 srcOffsetNElems, SyncObject*) const {
    if (nElems == 0) return;
    ADLASSERT(nElems + srcOffsetNElems <= m_size);
    adlCheck(b200rs_memcpy_d2h(m_device->getHandle(), hostDstPtr, m_ptr + srcOffsetNElems, nElems * sizeof(T)), "b200rs_memcpy_d2h");
}

template <typename T>
void Buffer<T>::write(const Buffer<T>& src, u64 nElems, SyncObject*) {
    if (nElems == 0) return;
    ADLASSERT(nElems <= m_size && nElems <= src.m_size);
    m_deviceWritten = true;
    adlCheck(b200rs_memcpy_d2d(m_device->getHandle(), m_ptr, src.m_ptr, nElems * sizeof(T)), "b200rs_memcpy_d2d");
}

template <typename T>
void Buffer<T>::read(Buffer<T>& dst, u64 nElems, u64 offsetNElems, SyncObject*) const {
    ADLASSERT(offsetNElems == 0);
    if (nElems == 0) return;
    ADLASSERT(nElems <= m_size && nElems <= dst.m_size);
    dst.markDeviceWritten();
    adlCheck(b200rs_memcpy_d2d(m_device->getHandle(), dst.m_ptr, m_ptr, nElems * sizeof(T)), "b200rs_memcpy_d2d");
}

template <typename T>
void Buffer<T>::clear() {
    if (m_size == 0) return;
    m_deviceWritten = true;
    adlCheck(b200rs_memset(m_device->getHandle(), m_ptr, 0, m_size * sizeof(T)), "b200rs_memset");
}

template <typename T>
void Buffer<T>::fill(void* pattern, int patternSize) {
    // repeats `pattern` over the whole buffer ON THE DEVICE (reference: clEnqueueFillBuffer, AdlCL.inl:523-542;
    // AdlHost.inl:132-145): patterns of 1, 2, 4, 8 or 16 bytes are widened to 16 (or 4) bytes and written by the library's fill
    // kernel; other pattern sizes are expanded on the host once and copied.
    const u64 bytes = m_size * sizeof(T);
    ADLASSERT(patternSize > 0 && bytes % (u64)patternSize == 0);
    if (bytes == 0) return;
    m_deviceWritten = true;
    const bool pow2 = patternSize <= 16 && (patternSize & (patternSize - 1)) == 0;
    if (pow2 && ((uintptr_t)m_ptr % 16 == 0) && bytes % 16 == 0) {
        uint32_t wide[4];
        for (int off = 0; off < 16; off += patternSize) memcpy((char*)wide + off, pattern, (size_t)patternSize);
        adlCheck(b200rs_fill_u128(m_device->getHandle(), m_ptr, wide, bytes / 16), "b200rs_fill_u128");
        return;
    }
    if (pow2 && patternSize <= 4 && ((uintptr_t)m_ptr % 4 == 0) && bytes % 4 == 0) {
        uint32_t word = 0;
        for (int off = 0; off < 4; off += patternSize) memcpy((char*)&word + off, pattern, (size_t)patternSize);
        adlCheck(b200rs_fill_u32(m_device->getHandle(), (uint32_t*)m_ptr, word, bytes / 4), "b200rs_fill_u32");
        return;
    }
    char* host = new char[bytes];
    for (u64 off = 0; off < bytes; off += (u64)patternSize) memcpy(host + off, pattern, (size_t)patternSize);
    adlCheck(b200rs_memcpy_h2d(m_device->getHandle(), m_ptr, host, bytes), "b200rs_memcpy_h2d");
    m_device->waitForCompletion();
    delete[] host;
}

template <typename T>
T* Buffer<T>::getHostPtr(u64 size) const {
    ADLASSERT(m_cl.m_hostPtr == 0);  // one mapping at a time per buffer
    const u64 n = (size == (u64)-1 || size > m_size) ? m_size : size;
    if (n == 0) return 0;
    Buffer<T>* self = const_cast<Buffer<T>*>(this);
    if (m_zeroCopy) {  // the buffer IS host memory: valid once the device work on it has completed (the caller waits, as for any map)
        self->m_cl.m_hostPtr = (char*)m_ptr;
        m_mappedElems = n;
        return m_ptr;
    }
    int slot = -1;
    void* stage = m_device->acquireStage(n * sizeof(T), &slot);
    if (!stage) return 0;
    self->m_cl.m_hostPtr = (char*)stage;
    m_mappedElems = n;
    m_mapSlot = slot;
    // a buffer nothing was ever written to holds no data worth the copy (the caller is about to fill the view)
    if (m_deviceWritten) adlCheck(b200rs_memcpy_d2h(m_device->getHandle(), stage, m_ptr, n * sizeof(T)), "b200rs_memcpy_d2h");  // async; caller waits
    return (T*)stage;
}

template <typename T>
void Buffer<T>::returnHostPtr(T* ptr) const {
    if (ptr == 0) return;
    ADLASSERT((char*)ptr == m_cl.m_hostPtr);
    Buffer<T>* self = const_cast<Buffer<T>*>(this);
    self->m_cl.m_hostPtr = 0;
    if (m_zeroCopy) {
        m_mappedElems = 0;
        m_deviceWritten = true;
        return;
    }
    // non-blocking: the copy out of the stage is stream-ordered in front of whatever uses the buffer next; the stage is
    // recycled when the event recorded behind the copy has passed (Device::releaseStage)
    const bool ok = adlCheck(b200rs_memcpy_h2d(m_device->getHandle(), m_ptr, ptr, m_mappedElems * sizeof(T)), "b200rs_memcpy_h2d");
    m_deviceWritten = true;
    m_device->releaseStage(ptr, m_mapSlot, ok);
    m_mappedElems = 0;
    m_mapSlot = -1;
}

template <typename T>
void Buffer<T>::setSize(u64 size) {
    ADLASSERT(m_device != 0);
    if (!m_allocated) {
        ADLASSERT(m_ptr == 0);
        allocate(m_device, size, BUFFER);
    } else if (m_size < size) {
        const Device* d = m_device;
        d->waitForCompletion();
        releaseStorage();
        allocate(d, size, BUFFER);
    }
}

}  // namespace adl

#include <Adl/AdlStopwatch.h>

#endif  // ADL_H
