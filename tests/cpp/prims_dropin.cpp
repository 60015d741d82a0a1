// tests/cpp/prims_dropin.cpp -- C++ caller of the drop-in headers for the rows next to the hot path
// (SURVEY.md section 8f): Pprims::copy / Pprims::fill on uArray and Buffer (reference: Pprims.cpp:31-121), adl::Stopwatch
// (AdlStopwatch.h:27-83), the per-launch profile CSV (AdlKernelUtilsCL.inl:654-677) and the uArray dirty-state machine
// (uArray.h:13-228).  Written the way UnitTest/main.cpp is: same device set-up, CPU loop as the checker.
// Built by `make prims_test` into tools/_build/prims_dropin; run on the GPU box by tests/test_gpu_prims.py.
#include <Adl/Adl.h>
#include <Tahoe/ParallelPrimitives/Pprims.h>

#include <stdio.h>
#include <string.h>

using namespace adl;
using namespace Tahoe;

char adl::s_cacheDirectory[128];

static int g_failures = 0;
#define CHECK(cond)                                                         \
    do {                                                                    \
        if (!(cond)) {                                                      \
            printf("CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);  \
            ++g_failures;                                                   \
        }                                                                   \
    } while (0)

int main() {
    DeviceUtils::Config cfg;
    cfg.m_type = DeviceUtils::Config::DEVICE_GPU;
    Device* d = DeviceUtils::allocate(TYPE_CL, cfg);
    if (!d) { printf("no device\n"); return 2; }
    {
        Pprims p;
        const int sizes[] = {1, 5, 1024, 4099, 1 << 20, (1 << 22) + 3};
        for (unsigned s = 0; s < sizeof(sizes) / sizeof(sizes[0]); ++s) {
            const int n = sizes[s];
            // ---- copy / fill on uArray<int> (the reference's signatures) ----
            uArray<int> a(n), b(n);
            for (int i = 0; i < n; ++i) { a[i] = i * 7 - 3; b[i] = -1; }
            p.copy(d, b, a, n);
            DeviceUtils::waitForCompletion(d);
            bool same = true;
            for (int i = 0; i < n; ++i) same = same && (b[i] == i * 7 - 3);  // reading b pulls the device copy back
            CHECK(same);
            p.fill(d, b, 42, n / 2);
            DeviceUtils::waitForCompletion(d);
            same = true;
            for (int i = 0; i < n; ++i) same = same && (b[i] == (i < n / 2 ? 42 : i * 7 - 3));
            CHECK(same);
            // ---- u32 fill + float4 copy / fill on Buffers ----
            Buffer<u32> bu(d, n);
            p.fill(d, bu, 0xdeadbeefu, n);
            Array<u32> hu(n);
            bu.read(hu.begin(), n);
            DeviceUtils::waitForCompletion(d);
            same = true;
            for (int i = 0; i < n; ++i) same = same && (hu[i] == 0xdeadbeefu);
            CHECK(same);
            uArray<float4> fa(n), fb(n);
            for (int i = 0; i < n; ++i) { fa[i] = make_float4((float)i, 0.5f * i, -1.f * i, 1.f); fb[i] = make_float4(0, 0, 0, 0); }
            p.copy(d, fb, fa, n);
            DeviceUtils::waitForCompletion(d);
            same = true;
            for (int i = 0; i < n; ++i) same = same && memcmp(&fa[i], &fb[i], sizeof(float4)) == 0;
            CHECK(same);
            const float4 v = make_float4(1.5f, -2.25f, 3e-9f, 7.f);
            p.fill(d, fb, v, n);
            DeviceUtils::waitForCompletion(d);
            same = true;
            for (int i = 0; i < n; ++i) same = same && memcmp(&fb[i], &v, sizeof(float4)) == 0;
            CHECK(same);
        }
        // ---- Stopwatch (device time) around a sort, split marks; profile CSV of the same launches ----
        const int n = 1 << 22;
        Buffer<u32> keys(d, n);
        Array<u32> h(n);
        srand(123);
        for (int i = 0; i < n; ++i) h[i] = ((u32)rand() << 16) ^ (u32)rand();
        keys.write(h.begin(), n);
        DeviceUtils::waitForCompletion(d);
        d->toggleProfiling(true);
        Stopwatch sw(d);
        sw.start();
        p.radixSort(d, keys, n);
        sw.split();
        p.radixSort(d, keys, n, 16);
        sw.stop();
        float t[4];
        sw.getMs(t, 4);
        CHECK(sw.getNIntervals() == 2);
        CHECK(t[0] > 0.f && t[1] > 0.f && t[2] == 0.f && sw.getMs() == t[0]);
        remove("gpurun_out_profile_test.csv");
        const int launches = d->writeProfileCsv("gpurun_out_profile_test.csv");
        CHECK(launches >= 1 + 4 + 1 + 2);  // histogram (digit starts are computed by its last CTA) + passes, twice
        d->toggleProfiling(false);
        Stopwatch host;  // no device: the host clock
        host.start();
        host.stop();
        CHECK(host.getMs() >= 0.f && host.getNIntervals() == 1);
        keys.read(h.begin(), n);
        DeviceUtils::waitForCompletion(d);
        // sorted on all 32 bits, then stably on the low 16: ordered by the low half, ties still ascending
        bool sorted = true;
        for (int i = 1; i < n; ++i) {
            const u32 a = h[i - 1] & 0xffffu, b = h[i] & 0xffffu;
            sorted = sorted && (a < b || (a == b && h[i - 1] <= h[i]));
        }
        CHECK(sorted);
        printf("sort 4M keys: %.3f ms (32 bits), %.3f ms (16 bits), %d launches logged\n", t[0], t[1], launches);
    }
    CHECK(d->getUsedMemory() == 0);
    DeviceUtils::deallocate(d);
    printf(g_failures ? "PRIMS DROPIN FAILED (%d)\n" : "PRIMS DROPIN OK\n", g_failures);
    return g_failures ? 1 : 0;
}
