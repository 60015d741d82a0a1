/*
 * oracle/radixsort_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see radixsort_oracle.h).
 *
 * Serial CPU restatement of the reference Host path.  Parity status: PINNED (header).
 */
#include "radixsort_oracle.h"

#include <stdlib.h>
#include <string.h>

enum { DIGIT_BITS = 8, NUM_BINS = 1 << DIGIT_BITS }; /* RadixSort.h:39-43 */

/*
 * One stable counting-sort pass per digit, ping-ponging between the caller's array and a
 * scratch array -- the structure of RadixSort.cpp:20-52 / :68-100 (count, exclusive scan of
 * the 256 counters, distribute in input order, swap).  Differences from the reference,
 * none of which change the result: size_t counters (so n >= 2^31 works), a digit schedule
 * that stops at sort_bits, and a copy-back when the pass count is odd (the reference always
 * runs 4 passes, so its result always lands back in `data`).
 */
#define DEFINE_LSD_SORT(NAME, ELEM_T, KEY_OF)                                               \
    int NAME(ELEM_T* data, size_t n, int sort_bits) {                                       \
        if (n == 0 || sort_bits <= 0) return 0;                                             \
        if (sort_bits > 32) sort_bits = 32;                                                 \
        ELEM_T* scratch = (ELEM_T*)malloc(n * sizeof(ELEM_T));                              \
        if (!scratch) return -1;                                                            \
        ELEM_T* from = data;                                                                \
        ELEM_T* to = scratch;                                                               \
        for (int lo = 0; lo < sort_bits; lo += DIGIT_BITS) {                                \
            const int width = (sort_bits - lo < DIGIT_BITS) ? (sort_bits - lo) : DIGIT_BITS;\
            const uint32_t mask = (1u << width) - 1u;                                       \
            size_t next_slot[NUM_BINS];                                                     \
            memset(next_slot, 0, sizeof(next_slot));                                        \
            for (size_t i = 0; i < n; ++i) next_slot[(KEY_OF(from[i]) >> lo) & mask]++;     \
            size_t running = 0;                                                             \
            for (int b = 0; b < NUM_BINS; ++b) {                                            \
                const size_t c = next_slot[b];                                              \
                next_slot[b] = running;                                                     \
                running += c;                                                               \
            }                                                                               \
            for (size_t i = 0; i < n; ++i)                                                  \
                to[next_slot[(KEY_OF(from[i]) >> lo) & mask]++] = from[i];                  \
            ELEM_T* t = from; from = to; to = t;                                            \
        }                                                                                   \
        if (from != data) memcpy(data, from, n * sizeof(ELEM_T));                           \
        free(scratch);                                                                      \
        return 0;                                                                           \
    }

#define KEY_OF_U32(x) (x)
#define KEY_OF_PAIR(x) ((x).key)

DEFINE_LSD_SORT(oracle_sort_u32, uint32_t, KEY_OF_U32)       /* RadixSort.cpp:58-104 */
DEFINE_LSD_SORT(oracle_sort_pairs, oracle_pair_t, KEY_OF_PAIR) /* RadixSort.cpp:10-56  */

void oracle_scan_u32(uint32_t* dst, const uint32_t* src, size_t n, uint32_t* total_out) {
    /* UnitTest/main.cpp:193-199: dst[i] must equal the running sum before adding src[i]. */
    uint32_t running = 0;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t v = src[i]; /* read first: dst may alias src */
        dst[i] = running;
        running += v; /* unsigned: wraps mod 2^32 like the u32 kernels */
    }
    if (total_out) *total_out = running;
}

/* getRandom<T>(minV,maxV), UnitTest/main.cpp:79-86:
 *   double r = min2((double)RAND_MAX-1, (double)rand())/RAND_MAX;  T range = maxV-minV;
 *   return (T)(minV + r*range);                                                         */
static double unit_rand(void) {
    double v = (double)rand();
    const double cap = (double)RAND_MAX - 1.0;
    if (v > cap) v = cap;
    return v / (double)RAND_MAX;
}

void oracle_gen_sort32(uint32_t* out, size_t n, unsigned seed) {
    srand(seed);
    for (size_t i = 0; i < n; ++i) out[i] = (uint32_t)(0u + unit_rand() * 0xffffffffu);
}

void oracle_gen_keyvalue(oracle_pair_t* out, size_t n, unsigned seed) {
    srand(seed);
    for (size_t i = 0; i < n; ++i) {
        out[i].key = (uint32_t)(0u + unit_rand() * 0xffffffffu);
        out[i].value = (uint32_t)i;
    }
}

void oracle_gen_scan(int32_t* out, size_t n, unsigned seed) {
    srand(seed);
    for (size_t i = 0; i < n; ++i) out[i] = (int32_t)(0 + unit_rand() * 0xf);
}

uint64_t oracle_fnv1a64(const void* bytes, size_t nbytes) {
    const unsigned char* p = (const unsigned char*)bytes;
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < nbytes; ++i) {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}
