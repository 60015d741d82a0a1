// Tahoe/Base/Memory/AllocatorBase.h -- class-scope allocation hook with the reference's names
// (reference: Tahoe/Base/Memory/AllocatorBase.h:13-85): a malloc/free DefaultAllocator singleton,
// TH_MEM_ALLOCATOR, and TH_DECLARE_ALLOCATOR(Class) which routes the class's operator new/delete
// through it (used by Pprims and uArray).
#pragma once

#include <stdlib.h>
#include <new>

#include <Tahoe/Math/Math.h>
#include <Tahoe/Base/Config.h>

namespace Tahoe {

class AllocatorBase {
public:
    virtual ~AllocatorBase() {}
    virtual void* allocate(size_t size, const char* tag, u32 line) = 0;
    virtual void deallocate(void* p) = 0;
};

class DefaultAllocator : public AllocatorBase {
public:
    static DefaultAllocator& getInstance() {
        static DefaultAllocator s_instance;
        return s_instance;
    }
    virtual void* allocate(size_t size, const char* /*tag*/, u32 /*line*/) {
        void* p = malloc(size ? size : 1);
        ADLASSERT(p != 0);
        return p;
    }
    virtual void deallocate(void* p) { free(p); }
    bool checkConsistency() { return true; }
    u64 getCurrentUsage() const { return 0; }
    u64 getPeakUsage() const { return 0; }
};

}  // namespace Tahoe

#define TH_MEM_ALLOCATOR DefaultAllocator

#define TH_DECLARE_ALLOCATOR(x)                                                                                          \
    inline void* operator new(size_t size) { return TH_MEM_ALLOCATOR::getInstance().allocate(size, #x, __LINE__); }       \
    inline void* operator new(size_t, void* where) { return where; }                                                     \
    inline void operator delete(void* p) { TH_MEM_ALLOCATOR::getInstance().deallocate(p); }                              \
    inline void operator delete(void*, void*) {}
