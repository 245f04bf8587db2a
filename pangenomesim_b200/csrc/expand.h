// expand.h -- host-side 2-bit -> ASCII expansion (PGS_OUT_HOST_EXPAND), see expand.cpp
#pragma once
#include <cstdint>

namespace pgs {

// `sites` bases of a packed row (site s = bits 2(s%4).. of byte s/4, A=0 C=1 G=2 T=3) -> dst
void expand_bases(const void *packed_row, int64_t sites, char *dst);
const char *expand_kind(); // "avx2", "ssse3" or "scalar"

} // namespace pgs
