// shard.hpp -- how the unique-word table is split across ranks.
//
// Words stay in reference order and every rank owns one CONTIGUOUS range of them, balanced by symbol slots.  A rank's
// flat positions are then monotone in the reference's scan order inside the rank, and rank r's words all precede rank
// r+1's, so a globally comparable sequence number is simply (rank << 44) | local_sequence.  Every rank replays the same
// heap from the same globally aggregated records, so no winner broadcast is needed.
#pragma once
#include <cstdint>

namespace shred {

constexpr int kSeqRankShift = 44;  // local sequences are < 2^36 (flat position * 4 + slot); rank < 256 keeps the whole key below 2^52

inline uint64_t seq_base(int rank) { return static_cast<uint64_t>(rank) << kSeqRankShift; }

// first word of rank r's range: the first word whose slot offset is >= r * total_slots / world.
// `off` has n_words + 1 entries (exclusive prefix sums of len + 1).
template <class OffT>
inline uint64_t shard_begin(const OffT* off, uint64_t n_words, int r, int world) {
  if (r <= 0) return 0;
  if (r >= world) return n_words;
  const uint64_t total = static_cast<uint64_t>(off[n_words]);
  const uint64_t want = static_cast<uint64_t>((static_cast<unsigned __int128>(total) * static_cast<unsigned>(r)) / static_cast<unsigned>(world));
  uint64_t lo = 0, hi = n_words;  // first index with off[i] >= want
  while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (static_cast<uint64_t>(off[mid]) >= want) hi = mid; else lo = mid + 1; }
  return lo;
}

}  // namespace shred
