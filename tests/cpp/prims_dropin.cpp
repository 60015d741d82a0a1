// tests/cpp/prims_dropin.cpp -- C++ caller of the drop-in headers for the rows next to the hot path
// (SURVEY.md section 8f): Pprims::copy / Pprims::fill on uArray and Buffer (reference: Pprims.cpp:31-121), adl::Stopwatch
// (AdlStopwatch.h:27-83), the per-launch profile CSV (AdlKernelUtilsCL.inl:654-677) and the uArray dirty-state machine
// (uArray.h:13-228).  Written the way UnitTest/main.cpp is: same device set-up, CPU loop as the checker.
// Built by `make prims_test` into tools/_build/prims_dropin; run on the GPU box by tests/test_gpu_prims.py.
#include <Adl/Adl.h>
#include <Tahoe/ParallelPrimitives/Pprims.h>

#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <vector>

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
    {
        // ---- map / unmap fast path (reference: AdlCL.inl:381-386,512-565) ----
        Pprims p;
        const int n = 1 << 20;
        std::vector<u32> want(n);
        srand(7);
        // (1) the reference caller's sequence (UnitTest/main.cpp:118-139): map a FRESH buffer (no device -> host copy), fill, unmap
        //     (returns without waiting), sort, map, wait, read
        Buffer<u32> keys(d, n);
        u32* m = keys.getHostPtr(n);
        for (int i = 0; i < n; ++i) want[i] = m[i] = ((u32)rand() << 16) ^ (u32)rand();
        keys.returnHostPtr(m);
        p.radixSort(d, keys, n);
        m = keys.getHostPtr(n);
        DeviceUtils::waitForCompletion(d);
        std::sort(want.begin(), want.end());
        CHECK(memcmp(m, &want[0], sizeof(u32) * n) == 0);
        keys.returnHostPtr(m);
        // (2) unmap is non-blocking: two buffers mapped, filled and unmapped back to back must not share a stage that a copy still reads
        Buffer<u32> a(d, n), b(d, n);
        for (int round = 0; round < 3; ++round) {
            u32* ma = a.getHostPtr(n);
            if (round) DeviceUtils::waitForCompletion(d);  // (a written buffer is copied into the view: wait before touching it)
            for (int i = 0; i < n; ++i) ma[i] = 0xA0000000u + (u32)i + (u32)round;
            a.returnHostPtr(ma);
            u32* mb = b.getHostPtr(n);  // fresh in round 0: nothing to wait for, the host writes at once
            if (round) DeviceUtils::waitForCompletion(d);
            for (int i = 0; i < n; ++i) mb[i] = 0xB0000000u + (u32)i + (u32)round;
            b.returnHostPtr(mb);
        }
        std::vector<u32> ra(n), rb(n);
        a.read(&ra[0], n);
        b.read(&rb[0], n);
        DeviceUtils::waitForCompletion(d);
        bool ok = true;
        for (int i = 0; i < n; ++i) ok = ok && ra[i] == 0xA0000002u + (u32)i && rb[i] == 0xB0000002u + (u32)i;
        CHECK(ok);
        // (3) BUFFER_ZERO_COPY: pinned host memory the kernels sort in place; map returns the buffer itself
        Buffer<u32> z(d, 1 << 16, BufferBase::BUFFER_ZERO_COPY);
        CHECK(z.isZeroCopy());
        u32* mz = z.getHostPtr();
        CHECK(mz == z.m_ptr);
        std::vector<u32> wz(1 << 16);
        for (int i = 0; i < (1 << 16); ++i) wz[i] = mz[i] = ((u32)rand() << 16) ^ (u32)rand();
        z.returnHostPtr(mz);
        p.radixSort(d, z, 1 << 16);
        Buffer<int> zs(d, 1 << 16, BufferBase::BUFFER_ZERO_COPY);
        p.scan(d, zs, reinterpret_cast<Buffer<int>&>(z), 1 << 16);
        DeviceUtils::waitForCompletion(d);
        std::sort(wz.begin(), wz.end());
        CHECK(memcmp(z.m_ptr, &wz[0], sizeof(u32) << 16) == 0);
        u32 run = 0;
        ok = true;
        for (int i = 0; i < (1 << 16); ++i) { ok = ok && (u32)zs.m_ptr[i] == run; run += wz[i]; }
        CHECK(ok);
        // (4) Buffer::fill / clear run on the device for 1-, 2-, 4-, 8- and 16-byte patterns; other sizes are expanded on the host
        const int psizes[] = {1, 2, 4, 8, 16, 12, 3};
        for (unsigned k = 0; k < sizeof(psizes) / sizeof(psizes[0]); ++k) {
            const int ps = psizes[k];
            unsigned char pat[16];
            for (int j = 0; j < ps; ++j) pat[j] = (unsigned char)(0x11 * (j + 1) + k);
            Buffer<u32> f(d, 3 * 4 * 1024);  // bytes divisible by every pattern size above
            f.fill(pat, ps);
            std::vector<u32> hf(f.getSize());
            f.read(&hf[0], f.getSize());
            DeviceUtils::waitForCompletion(d);
            const unsigned char* hb = (const unsigned char*)&hf[0];
            ok = true;
            for (size_t j = 0; j < hf.size() * 4; ++j) ok = ok && hb[j] == pat[j % ps];
            CHECK(ok);
            f.clear();
            f.read(&hf[0], f.getSize());
            DeviceUtils::waitForCompletion(d);
            ok = true;
            for (size_t j = 0; j < hf.size(); ++j) ok = ok && hf[j] == 0;
            CHECK(ok);
        }
        // (5) time: the Buffer-API sequence against the raw host-buffer C entry (copies included), 1 Mi pairs
        const int np = 1 << 20;
        Buffer<uint2> kv(d, np);
        std::vector<uint2> src(np);
        for (int i = 0; i < np; ++i) { src[i].x = ((u32)rand() << 16) ^ (u32)rand(); src[i].y = (u32)i; }
        double best_api = 1e30, best_raw = 1e30;
        void* pinned = 0;
        b200rs_host_alloc(d->getHandle(), sizeof(uint2) * np, &pinned);
        for (int rep = 0; rep < 6; ++rep) {
            auto t0 = std::chrono::steady_clock::now();
            uint2* mk = kv.getHostPtr(np);
            if (rep) DeviceUtils::waitForCompletion(d);
            memcpy(mk, &src[0], sizeof(uint2) * np);
            kv.returnHostPtr(mk);
            p.radixSort(d, kv, np);
            mk = kv.getHostPtr(np);
            DeviceUtils::waitForCompletion(d);
            volatile u32 first = mk[0].x;
            (void)first;
            kv.returnHostPtr(mk);
            auto t1 = std::chrono::steady_clock::now();
            best_api = std::min(best_api, std::chrono::duration<double, std::milli>(t1 - t0).count());
            memcpy(pinned, &src[0], sizeof(uint2) * np);
            t0 = std::chrono::steady_clock::now();
            b200rs_sort_pairs_u32_host(d->getHandle(), (b200rs_pair*)pinned, np, 32);
            t1 = std::chrono::steady_clock::now();
            best_raw = std::min(best_raw, std::chrono::duration<double, std::milli>(t1 - t0).count() + 0.0);
        }
        // the Buffer-API figure includes the 8 MiB host memcpy into the mapped view; time that alone to compare like with like
        double best_cpy = 1e30;
        for (int rep = 0; rep < 6; ++rep) {
            auto t0 = std::chrono::steady_clock::now();
            memcpy(pinned, &src[0], sizeof(uint2) * np);
            auto t1 = std::chrono::steady_clock::now();
            best_cpy = std::min(best_cpy, std::chrono::duration<double, std::milli>(t1 - t0).count());
        }
        DeviceUtils::waitForCompletion(d);
        b200rs_host_free(d->getHandle(), pinned);
        printf("MAP_UNMAP 1Mi pairs: buffer_api_ms %.3f (of which host memcpy %.3f) raw_host_entry_ms %.3f\n", best_api, best_cpy, best_raw);
    }
    CHECK(d->getUsedMemory() == 0);
    DeviceUtils::deallocate(d);
    printf(g_failures ? "PRIMS DROPIN FAILED (%d)\n" : "PRIMS DROPIN OK\n", g_failures);
    return g_failures ? 1 : 0;
}
