// charset.hpp -- character_coverage keep-set from the 256-bin byte histogram (host side; 256 values).
//
// Reference: shredword/csrc/bpe/bpe.cpp:156-171 with histogram.cpp:30-53.  The histogram counts every byte of every
// UNIQUE word once (not weighted by the word's count).  The distinct bytes are ordered by count descending; ties keep
// the iteration order of the reference's 256-bucket StrMap of one-byte strings, whose djb2 bucket is
// (5381*33 + b) & 255 = (b + 165) & 255 (hash.cpp:35-39) -- glibc's qsort is a stable merge sort for this input, so
// that order survives.  keep = (size_t)((float)distinct * coverage) in float32.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>

namespace shred {

inline void charset_keep(const uint64_t hist[256], float coverage, uint8_t keep[256], uint32_t* n_distinct, uint32_t* n_keep) {
  struct CC { uint64_t count; uint32_t rank; uint32_t byte; } cc[256];
  uint32_t c = 0;
  for (uint32_t b = 0; b < 256; b++) if (hist[b]) { cc[c].count = hist[b]; cc[c].rank = (b + 165u) & 255u; cc[c].byte = b; c++; }
  std::sort(cc, cc + c, [](const CC& x, const CC& y) { return x.count != y.count ? x.count > y.count : x.rank < y.rank; });
  volatile float prod = static_cast<float>(c) * coverage;  // bpe.cpp:169: size_t * float evaluates in float32
  size_t k = static_cast<size_t>(prod);
  std::memset(keep, 0, 256);
  for (size_t i = 0; i < k && i < c; i++) keep[cc[i].byte] = 1;
  if (n_distinct) *n_distinct = c;
  if (n_keep) *n_keep = static_cast<uint32_t>(k < c ? k : c);
}

}  // namespace shred
