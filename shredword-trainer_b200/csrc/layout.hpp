// layout.hpp -- the symbol array's encoding and the per-occurrence logic of a merge, shared VERBATIM by the CUDA kernels
// (cuda/kernels_merge.cuh) and by the CPU stand-in tests/hostsim/engine_listsim.cpp, so that the layout rules can be
// checked against the reference on machines without a GPU.  Pure functions over an abstract load/store of ids[].
//
// Position-stable layout.  Every unique word owns a fixed slot range  [HDR|wi] b0 b1 ... b(len-1)  of the flat int32
// array ids[], one slot per byte of the word, words back to back in reference order.
// A token that covers bytes [s, e) of its word lives at slot s (the span's first slot) and never moves:
//     ids[s]   = token id (>= 0; unk symbols carry unk_code)
//     ids[s+1] = SKIP(e-s)  and  ids[e-1] = SKIP(e-s)      when e-s >= 2 (forward / backward navigation)
//     every other slot of the span < 0 (DEAD or a stale SKIP): a slot holds a value >= 0 iff a token starts there.
// A merge (A,B)->N at slot p therefore rewrites 4 slots and moves nothing (reference relinks list nodes, bpe.cpp:291-294):
// "flat position of the token's first byte" stays monotone in the reference's (word, position) scan order and is both the
// sequence number the host needs (Appendix A14) and what the per-pair occurrence lists store.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SHRED_HD __host__ __device__ __forceinline__
#else
#define SHRED_HD inline
#endif

namespace shred {
namespace lay {

constexpr uint32_t HDR_TAG = 0x80000000u, SKIP_TAG = 0xC0000000u, TAG_MASK = 0xC0000000u, LOW30 = 0x3FFFFFFFu;
constexpr int32_t DEAD = -1;                   // == SKIP_TAG | LOW30: never a valid SKIP (span lengths stay below 2^30 - 1)
constexpr int32_t UNK_CODE_NEG = 0x7FFFFFFF;   // stored code of unk symbols when unk_id < 0
constexpr uint32_t NONE32 = 0xFFFFFFFFu;

SHRED_HD bool is_tok(int32_t v) { return v >= 0; }
SHRED_HD bool is_hdr(int32_t v) { return (static_cast<uint32_t>(v) & TAG_MASK) == HDR_TAG; }
SHRED_HD bool is_skip(int32_t v) { return (static_cast<uint32_t>(v) & TAG_MASK) == SKIP_TAG && v != DEAD; }
SHRED_HD uint32_t skip_len(int32_t v) { return static_cast<uint32_t>(v) & LOW30; }
SHRED_HD int32_t make_skip(uint32_t len) { return static_cast<int32_t>(SKIP_TAG | len); }
SHRED_HD int32_t make_hdr(uint32_t wi) { return static_cast<int32_t>(HDR_TAG | wi); }
SHRED_HD uint32_t hdr_word(int32_t v) { return static_cast<uint32_t>(v) & LOW30; }

struct Params {
  int32_t unk_id, unk_code;
  uint64_t min_freq;
};
SHRED_HD int32_t code_to_id(int32_t code, const Params& P) { return (P.unk_id < 0 && code == P.unk_code) ? P.unk_id : code; }
SHRED_HD uint64_t fc_key(int32_t a, int32_t b) {  // bpe.cpp:277-278: both operands sign-extend
  return (static_cast<uint64_t>(static_cast<int64_t>(a)) << 32) | static_cast<uint64_t>(static_cast<int64_t>(b));
}
SHRED_HD bool key_has_unk(uint64_t key, const Params& P) {  // decoded as bpe.cpp:301 does
  return static_cast<int32_t>(key >> 32) == P.unk_id || static_cast<int32_t>(key & 0xFFFFFFFFu) == P.unk_id;
}

// first slot of the token after the one starting at p (may hold a header / terminator)
template <class Ld>
SHRED_HD uint64_t next_start(Ld ld, uint64_t p) {
  const int32_t m = ld(p + 1);
  return is_skip(m) ? p + skip_len(m) : p + 1;
}

// One entry p of the occurrence list of (A,B): is it (still) an occurrence that this merge rewrites, and if so which four
// count deltas does the reference's left-to-right pass produce for it (bpe.cpp:265-296)?
//   left neighbour  = the id standing there when the pass reaches p: N if the two tokens before p are themselves merged in
//                     this pass, else the raw id;   right neighbour = the raw id after B (before any later merge of the pass).
// A and B never equal unk_id (such pairs are never merged, bpe.cpp:53), so their span lengths lenA/lenB are exact.
// Reads only the pre-merge state of ids[]; every occurrence can be probed independently and in parallel.
struct Occ {
  uint32_t pl;        // slot where the left neighbour starts AFTER this pass (start of the previous occurrence if that one merges)
  int32_t lid, rid;   // ids as the reference's delta keys see them (unk_code translated back to unk_id)
  bool has_l, has_r;
};
template <class Ld>
SHRED_HD bool probe_occurrence(Ld ld, uint64_t p, int32_t A, int32_t B, uint32_t lenA, uint32_t lenB, int32_t N, const Params& P, Occ* o) {
  if (ld(p) != A) return false;             // stale entry: the token at p was merged away or into something else
  if (ld(p + lenA) != B) return false;      // the token after it is no longer B
  bool left_merged;
  uint64_t pl = 0;
  int32_t l1 = DEAD;
  if (A != B) {
    const int32_t m = ld(p - 1);            // p >= 1: slot 0 is the header of word 0
    if (is_skip(m)) { pl = p - skip_len(m); l1 = ld(pl); } else { pl = p - 1; l1 = m; }
    // (A,B) pairs cannot overlap when A != B: the left neighbour merges iff it is B preceded by A
    left_merged = l1 == B && pl >= lenA && ld(pl - lenA) == A;
    if (left_merged) pl -= lenA;
  } else {
    uint64_t q = p;                         // start of the run of A's: pairs are taken greedily from there (bpe.cpp:268-295)
    while (q >= lenA && ld(q - lenA) == A) q -= lenA;
    const uint64_t k = (p - q) / lenA;
    if (k & 1ull) return false;             // second half of a merged pair, not an occurrence
    left_merged = k > 0;
    if (left_merged) { pl = p - 2ull * lenA; l1 = A; }
    else {
      const int32_t m = ld(p - 1);
      if (is_skip(m)) { pl = p - skip_len(m); l1 = ld(pl); } else { pl = p - 1; l1 = m; }
    }
  }
  const int32_t r = ld(p + lenA + lenB);
  o->has_l = l1 >= 0;
  o->has_r = r >= 0;
  o->lid = left_merged ? N : code_to_id(l1, P);
  o->rid = code_to_id(r, P);
  o->pl = static_cast<uint32_t>(pl);
  return true;
}

// the rewrite of one occurrence (bpe.cpp:291-294): 4 stores, nothing moves
template <class St>
SHRED_HD void rewrite_occurrence(St st, uint64_t p, uint32_t lenA, uint32_t lenB, int32_t N) {
  const uint64_t q = p + lenA, e = q + lenB;
  const int32_t mark = make_skip(lenA + lenB);
  st(q, DEAD);       // the slot where B started never reads as a token again (occurrence lists are validated lazily)
  st(p + 1, mark);
  st(e - 1, mark);
  st(p, N);
}

}  // namespace lay
}  // namespace shred
