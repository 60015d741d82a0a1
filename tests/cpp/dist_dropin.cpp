// tests/cpp/dist_dropin.cpp -- a C++ caller of the partitioned sort (include/b200rs.h
// b200rs_dist_sort_pairs_u32 through Tahoe::Pprims::radixSortDistributed): ONE process, one host thread per GPU, the two
// collectives implemented with a pthread barrier and peer copies -- no communication library.  Every rank generates its slice
// of the input, the ranks' results in rank order are compared with std::stable_sort of the whole input (value = global
// index, so the stable result is unique).  Written the way UnitTest/main.cpp sets a device up.  Needs >= 2 GPUs.
//   usage: dist_dropin [world] [pairs per rank]
#include <Adl/Adl.h>
#include <Tahoe/ParallelPrimitives/Pprims.h>

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

using namespace adl;
using namespace Tahoe;

char adl::s_cacheDirectory[128];

struct Shared {
    int world, n;
    pthread_barrier_t bar;
    const void* send[32];
    u64 recvBase[32];
    std::vector<uint2> in[32], out[32];
    int failed;
};
struct RankCtx {
    Shared* sh;
    int rank;
    Device* dev;
};

static int allgatherCb(void* user, const void* send, void* recv, size_t bytes) {
    RankCtx* c = (RankCtx*)user;
    c->dev->waitForCompletion();  // my contribution is complete
    c->sh->send[c->rank] = send;
    pthread_barrier_wait(&c->sh->bar);
    for (int s = 0; s < c->sh->world; ++s)
        if (b200rs_memcpy_d2d(c->dev->getHandle(), (char*)recv + (size_t)s * bytes, c->sh->send[s], bytes) != B200RS_OK) return -1;
    c->dev->waitForCompletion();
    pthread_barrier_wait(&c->sh->bar);  // nobody reuses its send buffer before everyone has copied it
    return 0;
}
static int barrierCb(void* user) {
    RankCtx* c = (RankCtx*)user;
    c->dev->waitForCompletion();  // my peer stores have landed
    pthread_barrier_wait(&c->sh->bar);
    return 0;
}

static void* rankMain(void* arg) {
    RankCtx* c = (RankCtx*)arg;
    Shared* sh = c->sh;
    const int r = c->rank, n = sh->n;
    DeviceUtils::Config cfg;
    cfg.m_type = DeviceUtils::Config::DEVICE_GPU;
    cfg.m_deviceIdx = r;
    Device* d = DeviceUtils::allocate(TYPE_CL, cfg);
    c->dev = d;
    if (!d) { sh->failed = 1; return 0; }
    for (int p = 0; p < sh->world; ++p) b200rs_enable_peer_access(d->getHandle(), p);
    {
        Pprims p;
        const u64 cap = (u64)n * 5 / 4 + 1024;
        Buffer<uint2> in(d, n), recv(d, cap);
        srand(1000 + r);
        sh->in[r].resize(n);
        for (int i = 0; i < n; ++i) {
            u32 k = ((u32)rand() << 16) ^ (u32)rand();
            if (i & 1) k &= 0xff0000ffu;  // half the keys from a small set: equal keys across ranks
            sh->in[r][i].x = k;
            sh->in[r][i].y = (u32)(r * n + i);
        }
        in.write(&sh->in[r][0], n);
        sh->recvBase[r] = (u64)(uintptr_t)recv.m_ptr;
        d->waitForCompletion();
        pthread_barrier_wait(&sh->bar);
        b200rs_dist_comm comm;
        comm.rank = r; comm.world = sh->world; comm.allgather = allgatherCb; comm.barrier = barrierCb; comm.user = c;
        long long m = 0;
        for (int rep = 0; rep < 2; ++rep) m = p.radixSortDistributed(d, comm, sh->recvBase, cap, in, n);  // twice: the receive buffers are reused
        if (m < 0) sh->failed = 1;
        else {
            sh->out[r].resize((size_t)m);
            if (m) recv.read(&sh->out[r][0], (u64)m);
            d->waitForCompletion();
        }
        pthread_barrier_wait(&sh->bar);  // peers may still be reading my buffers through their callbacks until here
    }
    DeviceUtils::deallocate(d);
    return 0;
}

static bool byKey(const uint2& a, const uint2& b) { return a.x < b.x; }

int main(int argc, char** argv) {
    Shared sh;
    sh.world = argc > 1 ? atoi(argv[1]) : 2;
    sh.n = argc > 2 ? atoi(argv[2]) : 300007;
    sh.failed = 0;
    if (DeviceUtils::getNDevices(TYPE_CL) < sh.world) { printf("DIST DROPIN SKIPPED: %d GPUs, %d ranks wanted\n", DeviceUtils::getNDevices(TYPE_CL), sh.world); return 0; }
    pthread_barrier_init(&sh.bar, 0, sh.world);
    std::vector<pthread_t> th(sh.world);
    std::vector<RankCtx> ctx(sh.world);
    for (int r = 0; r < sh.world; ++r) { ctx[r].sh = &sh; ctx[r].rank = r; ctx[r].dev = 0; pthread_create(&th[r], 0, rankMain, &ctx[r]); }
    for (int r = 0; r < sh.world; ++r) pthread_join(th[r], 0);
    std::vector<uint2> whole, got;
    for (int r = 0; r < sh.world; ++r) { whole.insert(whole.end(), sh.in[r].begin(), sh.in[r].end()); got.insert(got.end(), sh.out[r].begin(), sh.out[r].end()); }
    std::stable_sort(whole.begin(), whole.end(), byKey);
    bool same = !sh.failed && whole.size() == got.size();
    for (size_t i = 0; same && i < whole.size(); ++i) same = whole[i].x == got[i].x && whole[i].y == got[i].y;
    printf("ranks %d, %d pairs each, received:", sh.world, sh.n);
    for (int r = 0; r < sh.world; ++r) printf(" %zu", sh.out[r].size());
    printf("\n%s\n", same ? "DIST DROPIN OK" : "DIST DROPIN FAILED");
    return same ? 0 : 1;
}
