/*
 * oracle/ref_wrap.cpp -- TEST INFRASTRUCTURE.  C entry points around the UNMODIFIED reference,
 * compiled where it lies under /root/reference into oracle/_ref/libref_oclradixsort.so
 * (see oracle/Makefile).  Used to pin oracle/radixsort_oracle.c, to generate tests/golden/,
 * and as bench.py's cpu_baseline (kind "reference").
 */
#include <Adl/Adl.h>
#include <Tahoe/ParallelPrimitives/Pprims.h>
#include <Tahoe/Algorithm/Sort/RadixSort.h>
#include <stdint.h>
#include <chrono>

char adl::s_cacheDirectory[128] = "cache"; /* the application defines it: UnitTest/main.cpp:74 */

extern "C" {

/* Tahoe::RadixSort::sort(u32*, int), RadixSort.cpp:58-104 */
void ref_radixsort_u32(uint32_t* data, int n) { Tahoe::RadixSort::sort((Tahoe::u32*)data, n); }

/* Tahoe::RadixSort::sort(SortData*, int), RadixSort.cpp:10-56 */
void ref_radixsort_pairs(void* data, int n) { Tahoe::RadixSort::sort((Tahoe::SortData*)data, n); }

/* The Adl Host-backend route: DeviceUtils::allocate(TYPE_HOST) -> Buffer -> Pprims::radixSort
 * (Pprims.cpp:306-316 -> RadixSort::sort).  Data goes in and out through Buffer::write/read. */
void ref_hostbackend_sort_u32(uint32_t* data, int n) {
    adl::Device* d = adl::DeviceUtils::allocate(adl::TYPE_HOST);
    {
        Tahoe::Pprims p;
        adl::Buffer<Tahoe::u32> buf(d, n);
        buf.write((const Tahoe::u32*)data, n);
        adl::DeviceUtils::waitForCompletion(d);
        p.radixSort(d, buf, n);
        buf.read((Tahoe::u32*)data, n);
        adl::DeviceUtils::waitForCompletion(d);
    }
    adl::DeviceUtils::deallocate(d);
}

/* Pprims.cpp:202-212 -> RadixSort::sort(SortData*) */
void ref_hostbackend_sort_pairs(void* data, int n) {
    adl::Device* d = adl::DeviceUtils::allocate(adl::TYPE_HOST);
    {
        Tahoe::Pprims p;
        adl::Buffer<Tahoe::uint2> buf(d, n);
        buf.write((const Tahoe::uint2*)data, n);
        adl::DeviceUtils::waitForCompletion(d);
        p.radixSort(d, buf, n);
        buf.read((Tahoe::uint2*)data, n);
        adl::DeviceUtils::waitForCompletion(d);
    }
    adl::DeviceUtils::deallocate(d);
}

/* Timed forms for bench.py's cpu_baseline / --impl reference legs: the clock brackets only
 * Pprims::radixSort on the Host device (map + RadixSort::sort + unmap), not the staging copies. */
double ref_hostbackend_sort_u32_timed(uint32_t* data, int n) {
    adl::Device* d = adl::DeviceUtils::allocate(adl::TYPE_HOST);
    double seconds;
    {
        Tahoe::Pprims p;
        adl::Buffer<Tahoe::u32> buf(d, n);
        buf.write((const Tahoe::u32*)data, n);
        std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        p.radixSort(d, buf, n);
        adl::DeviceUtils::waitForCompletion(d);
        seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        buf.read((Tahoe::u32*)data, n);
    }
    adl::DeviceUtils::deallocate(d);
    return seconds;
}

double ref_hostbackend_sort_pairs_timed(void* data, int n) {
    adl::Device* d = adl::DeviceUtils::allocate(adl::TYPE_HOST);
    double seconds;
    {
        Tahoe::Pprims p;
        adl::Buffer<Tahoe::uint2> buf(d, n);
        buf.write((const Tahoe::uint2*)data, n);
        std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        p.radixSort(d, buf, n);
        adl::DeviceUtils::waitForCompletion(d);
        seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        buf.read((Tahoe::uint2*)data, n);
    }
    adl::DeviceUtils::deallocate(d);
    return seconds;
}

} /* extern "C" */
