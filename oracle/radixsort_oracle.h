/*
 * oracle/radixsort_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the reference's Host-backend algorithms for the
 * hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may link or call anything in oracle/.  The product
 * (libb200rs.so, include/, oclradixsort_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_pin.py checks this restatement
 *   (1) against the known-answer table of SURVEY.md section 8c (produced by the
 *       reference's own Host path), committed as tests/golden/reference_hashes.json;
 *   (2) in the build container, against oracle/_ref/libref_oclradixsort.so, which
 *       oracle/Makefile compiles from the UNMODIFIED reference sources
 *       (/root/reference/Tahoe/Algorithm/Sort/RadixSort.cpp and the Adl Host
 *       backend path through /root/reference/Tahoe/ParallelPrimitives/Pprims.cpp).
 *
 * Reference citations (relative to /root/reference):
 *   sort keys   Tahoe/Algorithm/Sort/RadixSort.cpp:58-104
 *   sort pairs  Tahoe/Algorithm/Sort/RadixSort.cpp:10-56, layout RadixSort.h:10-27
 *   scan        UnitTest/main.cpp:193-199 (serial exclusive sum),
 *               u32 wrap per Tahoe/ClKernels/PrefixScanKernels.cl:26-67,
 *               total (sumOut) per PrefixScanKernels.cl:139-142 / Pprims.cpp:164-167
 *   generators  UnitTest/main.cpp:76-86,109,122,152,183
 */
#ifndef RADIXSORT_ORACLE_H
#define RADIXSORT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 8-byte AoS pair; key at byte offset 0, value at 4 (RadixSort.h:10-21, Math.h:175-188). */
typedef struct oracle_pair { uint32_t key; uint32_t value; } oracle_pair_t;

/* Stable ascending unsigned LSD sort on bits [0,sort_bits) of the key; 8-bit digits,
 * the last digit narrower when sort_bits%8 != 0.  sort_bits==32 is exactly
 * RadixSort::sort.  Returns 0, or -1 if scratch allocation failed. */
int oracle_sort_u32(uint32_t* data, size_t n, int sort_bits);
int oracle_sort_pairs(oracle_pair_t* data, size_t n, int sort_bits);

/* Exclusive prefix sum in u32 arithmetic (wraps mod 2^32); *total_out (optional) = sum of
 * all n inputs.  dst may alias src. */
void oracle_scan_u32(uint32_t* dst, const uint32_t* src, size_t n, uint32_t* total_out);

/* Input generators of the reference's unit test: srand(seed) then one glibc rand() per
 * element through getRandom() (UnitTest/main.cpp:79-86). */
void oracle_gen_sort32(uint32_t* out, size_t n, unsigned seed);        /* main.cpp:122 */
void oracle_gen_keyvalue(oracle_pair_t* out, size_t n, unsigned seed); /* main.cpp:152 */
void oracle_gen_scan(int32_t* out, size_t n, unsigned seed);           /* main.cpp:183 */

/* FNV-1a-64 over raw bytes (the hash of SURVEY.md section 8c). */
uint64_t oracle_fnv1a64(const void* bytes, size_t nbytes);

#ifdef __cplusplus
}
#endif
#endif
