// oracle/RadixSort.cpp -- TEST INFRASTRUCTURE.  Definitions of Tahoe::RadixSort::sort (declared in
// include/Tahoe/Algorithm/Sort/RadixSort.h) in terms of the C oracle, for building drop-in test programs
// where /root/reference is not available.  Not linked into libb200rs.so.
#include <Tahoe/Algorithm/Sort/RadixSort.h>

#include "radixsort_oracle.h"

namespace Tahoe {
void RadixSort::sort(SortData* data, int n) { oracle_sort_pairs((oracle_pair_t*)data, (size_t)n, 32); }
void RadixSort::sort(u32* data, int n) { oracle_sort_u32((uint32_t*)data, (size_t)n, 32); }
}  // namespace Tahoe
