// kernels_tokenize.cuh -- whitespace tokeniser + unique-word table (bpe.cpp:131-153, hash.cpp:29-53); shared by the trainer's
// ingest (engine_cuda.cu) and the encoder (encoder_cuda.cu).  Fragment: included inside `namespace shred { namespace {` after common.cuh.
#pragma once

struct WordTable {
  ull* tag;       // 0 = empty
  ull* first;     // smallest byte offset of an occurrence
  ull* count;
  uint32_t* len;
  uint32_t* bucket;  // djb2 & 4095
  uint64_t mask, cap;
};

// Walk the token that starts at `off`: its length, the reference bucket hash djb2 (hash.cpp:35-39) and the 64-bit placement
// tag (two 32-bit multiplicative hashes, mixed; never 0).  text is padded with >= 32 spaces, so the walk terminates.
__device__ __forceinline__ uint64_t token_walk(const uint8_t* __restrict__ text, uint64_t off, uint32_t seed, uint32_t* len_out, uint32_t* dj_out) {
  uint32_t h1 = 2166136261u ^ seed, h2 = 0x9E3779B9u + seed, dj = 5381u, len = 0;
  for (;;) {
    const uint32_t c = text[off + len];
    if (is_delim(c)) break;
    h1 = (h1 ^ c) * 16777619u;
    h2 = (h2 + c) * 0x85EBCA6Bu; h2 ^= h2 >> 15;
    dj = dj * 33u + c;
    ++len;
  }
  *len_out = len; *dj_out = dj;
  return mix64((static_cast<uint64_t>(h1) << 32) | h2) | 1ull;
}

// token-start mask of 16 corpus bytes (bit i = a token starts at byte i): not a delimiter, preceded by one
__device__ __forceinline__ uint32_t start_mask16(const uint4& v, uint32_t prev) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t dm = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) { uint32_t c = (w[i >> 2] >> ((i & 3) * 8)) & 255u; dm |= (is_delim(c) ? 1u : 0u) << i; }
  return ~dm & ((dm << 1) | (is_delim(prev) ? 1u : 0u)) & 0xFFFFu;
}

// Find or claim the word-table slot of the token at `off` (tag, len, dj from token_walk): first offset by look-then-atomicMin,
// occurrence count (COUNT; the encoder does not need it), and a byte comparison with the word's representative -- two
// words with the same 64-bit tag raise ERR_WT_COLLISION and the caller starts over with another seed.  Returns the slot.
// Claims are counted in the caller's register (*n_claimed; one atomic per warp at the end of the kernel instead of one per
// new word on a single address); LIST appends the claimed slot to new_list (one atomic per converged group of claimers).
template <bool COUNT, bool LIST>
__device__ __forceinline__ uint64_t word_insert(const uint8_t* __restrict__ text, const WordTable& wt, DevCounters* ctr, uint64_t off, uint64_t tag, uint32_t len,
                                                uint32_t dj, uint32_t* n_claimed, uint32_t* new_list = nullptr, uint32_t* new_n = nullptr) {
  uint64_t slot = tag & wt.mask;
  for (uint32_t probe = 0; probe < 8192u; ++probe) {
    ull cur = wt.tag[slot];
    if (cur == 0ull) {
      ull prevt = atomicCAS(&wt.tag[slot], 0ull, static_cast<ull>(tag));
      if (prevt == 0ull) {  // claimed: publish the immutable facts
        wt.len[slot] = len;
        wt.bucket[slot] = dj & 4095u;
        ++*n_claimed;
        if (LIST) {
          const unsigned grp = __activemask();
          const unsigned lane = threadIdx.x & 31u;
          const int leader = __ffs(grp) - 1;
          uint32_t at = 0;
          if (lane == static_cast<unsigned>(leader)) at = atomicAdd(new_n, static_cast<uint32_t>(__popc(grp)));
          at = __shfl_sync(grp, at, leader);
          new_list[at + __popc(grp & ((1u << lane) - 1u))] = static_cast<uint32_t>(slot);
        }
        cur = tag;
      } else cur = prevt;
    }
    if (cur == tag) {
      // first occurrence: most tokens come after the word's first sighting, so look before paying for an atomic
      ull old = *reinterpret_cast<volatile ull*>(&wt.first[slot]);
      if (off < old) old = atomicMin(&wt.first[slot], static_cast<ull>(off));
      if (COUNT) {  // lanes of this warp that hit the same slot right now add once (hot words are most of a Zipf corpus)
        const unsigned am = __activemask();
        const unsigned grp = __match_any_sync(am, slot);
        if ((threadIdx.x & 31u) == static_cast<unsigned>(__ffs(grp) - 1)) atomicAdd(&wt.count[slot], static_cast<ull>(__popc(grp)));
      }
      if (old != SEQ_MAX && old != off) {  // same tag: must be the same bytes
        bool same = is_delim(text[old + len]);
        for (uint32_t j = 0; j < len && same; j++) same = text[old + j] == text[off + j];
        if (!same) atomicOr(&ctr->err, ERR_WT_COLLISION);
      }
      return slot;
    }
    slot = (slot + 1) & wt.mask;
  }
  atomicOr(&ctr->err, ERR_WT_FULL);
  return ~0ull;
}

// Each thread owns 16 consecutive corpus bytes (one uint4 load) and inserts every token that STARTS inside them at an offset in
// [lo, n).  text is padded with >= 32 spaces, so token walks terminate.  Ranges let the ingest tokenise a file while it is still
// arriving over PCIe: a range ends right after a delimiter, so every token that starts inside it also ends inside it, and the
// not-yet-landed rest of the buffer is pre-filled with spaces (no token start, no NUL).
__global__ void __launch_bounds__(256) k_tokenize(const uint8_t* __restrict__ text, uint64_t lo, uint64_t n, WordTable wt, DevCounters* ctr, uint32_t seed) {
  const uint64_t n16 = (n + 15) >> 4;
  uint32_t my_tokens = 0, my_claims = 0;
  for (uint64_t t = (lo >> 4) + blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; t < n16; t += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint64_t base = t << 4;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(text + base));
    uint32_t prev = base ? text[base - 1] : 32u;
    // a NUL byte hides the rest of its line in the reference (fgets + strlen, bpe.cpp:131-147): report it, the host
    // blanks the hidden spans and loads again
    if (((v.x - 0x01010101u) & ~v.x & 0x80808080u) | ((v.y - 0x01010101u) & ~v.y & 0x80808080u) | ((v.z - 0x01010101u) & ~v.z & 0x80808080u) |
        ((v.w - 0x01010101u) & ~v.w & 0x80808080u))
      atomicOr(&ctr->err, ERR_HAS_NUL);
    uint32_t starts = start_mask16(v, prev);
    while (starts) {
      const int i = __ffs(starts) - 1;
      starts &= starts - 1;
      const uint64_t off = base + i;
      if (off >= n) break;
      if (off < lo) continue;
      uint32_t len, dj;
      const uint64_t tag = token_walk(text, off, seed, &len, &dj);
      ++my_tokens;
      word_insert<true, false>(text, wt, ctr, off, tag, len, dj, &my_claims);
    }
  }
  // token and new-word counts: warp reduce, one atomic per warp
  for (int o = 16; o; o >>= 1) { my_tokens += __shfl_down_sync(0xFFFFFFFFu, my_tokens, o); my_claims += __shfl_down_sync(0xFFFFFFFFu, my_claims, o); }
  if ((threadIdx.x & 31) == 0 && my_tokens) atomicAdd(&ctr->n_tokens, static_cast<ull>(my_tokens));
  if ((threadIdx.x & 31) == 0 && my_claims) atomicAdd(&ctr->n_unique, my_claims);
}
