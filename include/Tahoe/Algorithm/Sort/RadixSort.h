// Tahoe/Algorithm/Sort/RadixSort.h -- declarations of the reference's CPU sort
// (reference: Tahoe/Algorithm/Sort/RadixSort.h:8-46).  SortData is the 8-byte {key, value} element
// of the key-value sort (same layout as Tahoe::uint2 / b200rs_pair).
//
// RadixSort::sort is what the reference's unit test checks the device result against
// (UnitTest/main.cpp:128,158).  It is NOT part of libb200rs.so and nothing in the CUDA path calls it:
// the definitions live with the test infrastructure (oracle/RadixSort.cpp, or the reference's own
// RadixSort.cpp when the drop-in unit test is built in the container).
#pragma once

#include <Tahoe/Math/Math.h>

namespace Tahoe {

struct SortData {
    union {
        u32 m_key;
        struct { u16 m_key16[2]; };
    };
    u32 m_value;

    SortData() {}
    SortData(u32 key, u32 value) : m_key(key), m_value(value) {}

    friend bool operator<(const SortData& a, const SortData& b) { return a.m_key < b.m_key; }
};

class RadixSort {
public:
    enum { BITS_PER_PASS = 8, NUM_TABLES = (1 << BITS_PER_PASS) };

    static void sort(SortData* data, int n);  // stable, ascending by m_key
    static void sort(u32* data, int n);       // ascending
};

}  // namespace Tahoe
