// Tahoe/Math/Error.h -- assertion / debug-print macros with the reference's names
// (reference: Tahoe/Math/Error.h:24-58).  Behaviour kept:
//   _DEBUG                    ADLASSERT(x) aborts when x is false
//   TH_UNIT_TEST (release)    ADLASSERT(x) is gtest's EXPECT_TRUE(x)  -> UnitTest/main.cpp relies on this
//   otherwise                 ADLASSERT(x) evaluates x and does nothing
#pragma once

#include <stdarg.h>
#include <stdio.h>

#if defined(_DEBUG)
#include <assert.h>
#endif
#if defined(TH_UNIT_TEST)
#include <gtest/gtest.h>
#endif

#include <Tahoe/Base/Config.h>

#if defined(_DEBUG)
#define ADLASSERT(x) do { if (!(x)) { assert(0); } } while (0)
#define ADLWARN(x) do { printf(x); } while (0)
#elif defined(TH_UNIT_TEST)
#define ADLASSERT(x) EXPECT_TRUE(x)
#define ADLWARN(x) do { x; } while (0)
#else
#define ADLASSERT(x) do { if (x) {} } while (0)
#define ADLWARN(x) do { x; } while (0)
#endif

#define ADLCOMPILEASSERT(x) static_assert(x, "CompileTimeAssert")
#define ADLCOMPILEASSERT1(x, msg) static_assert(x, msg)

inline void thDebugPrintf(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vprintf(fmt, ap);
    va_end(ap);
}

#if defined(_DEBUG)
#define debugPrintf(...) do { thDebugPrintf(__VA_ARGS__); Tahoe::TH_LOG_DEBUG(__VA_ARGS__); } while (0)
#else
#define debugPrintf(...) Tahoe::TH_LOG_DEBUG(__VA_ARGS__)
#endif
