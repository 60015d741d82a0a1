// Tahoe/Math/Math.h -- the integer typedefs, small POD vectors and helpers the sort / scan path
// and its callers use (reference: Tahoe/Math/Math.h:19,53-60,90-93,114-128,175-201,230-242,324-330).
// The reference header also carries float4 / Matrix3x3 renderer algebra; none of it is on the
// path, so it is not provided.
#pragma once

#include <stdlib.h>
#include <math.h>
#include <algorithm>

#include <Tahoe/Math/Error.h>

// smallest multiple of `alignment` that is >= num
#define NEXTMULTIPLEOF(num, alignment) ((((num) + (alignment) - 1) / (alignment)) * (alignment))

namespace Tahoe {

typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned short u16;
typedef unsigned char u8;

// {x = key, y = value} pair of Pprims::radixSort(Buffer<uint2>); 8 bytes, key first
struct uint2 {
    union {
        struct { u32 x, y; };
        u32 s[2];
    };
};

struct int2 {
    int2() {}
    int2(int a, int b) : x(a), y(b) {}
    union {
        struct { int x, y; };
        int s[2];
    };
};

// 16-byte POD {x, y, z, w}: the element type of Pprims::copy / Pprims::fill on float4 data (reference:
// Math.h:95-111; the renderer algebra on it -- operators, dot3, cross3 ... -- is not on the path and not provided)
struct __attribute__((aligned(16))) float4 {
    union {
        struct { float x, y, z, w; };
        float s[4];
    };
};
inline float4 make_float4(float x, float y, float z, float w = 0.f) { float4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }

struct int4 {
    union {
        struct { int x, y, z, w; };
        int s[4];
    };
};

template <typename T> inline T max2(const T& a, const T& b) { return a > b ? a : b; }
template <typename T> inline T min2(const T& a, const T& b) { return a < b ? a : b; }
template <typename T> inline T clamp(T v, T lo, T hi) { return max2(min2(hi, v), lo); }
template <typename T> inline void swap2(T& a, T& b) { T t = a; a = b; b = t; }

template <class T> inline T nextPowerOf2(T n) {
    T p = 1;
    while (p < n) p <<= 1;
    return p;
}
template <class T> inline T roundUpToMultiple(const T x, const T m) { return ((x + m - 1) / m) * m; }

}  // namespace Tahoe
