// kernels_encode.cuh -- BPE encoder / decoder kernels for the trainer's model file.
// Fragment of encoder_cuda.cu: included inside `namespace shred { namespace {` after common.cuh, kernels_tokenize.cuh, kernels_scan.cuh.
//
// Reference algorithm: BPETokenizer._encode_chunk, shredword/utils/bpe.py:191-203 (with get_stats :10-21 and merge :23-38):
// while the word has two or more ids, take the adjacent pair with the smallest merge id, stop if none is in the merge
// dict, replace every occurrence left to right without overlap.  Every DISTINCT word is encoded once (one warp per word),
// then every occurrence copies its word's ids (k_expand), which is the HBM-bound part.
#pragma once

struct MergeEnt { uint64_t key; int32_t val; int32_t pad; };  // key = (a << 32 | b) + 1, 0 = empty; one 16-byte load
struct MergeTable {
  const MergeEnt* ent;   // open addressing, load <= 1/4
  uint64_t mask;
};
constexpr int32_t NO_MERGE = 0x7FFFFFFF;

__device__ __forceinline__ int32_t merge_lookup(const MergeTable& mt, int32_t a, int32_t b) {
  const uint64_t k = ((static_cast<uint64_t>(static_cast<uint32_t>(a)) << 32) | static_cast<uint32_t>(b)) + 1ull;
  for (uint64_t s = mix64(k) & mt.mask;; s = (s + 1) & mt.mask) {
    const uint4 e = __ldg(reinterpret_cast<const uint4*>(mt.ent + s));
    const uint64_t cur = (static_cast<uint64_t>(e.y) << 32) | e.x;
    if (cur == k) return static_cast<int32_t>(e.z);
    if (cur == 0ull) return NO_MERGE;
  }
}

// byte lengths of the words claimed by the last k_enc_tokenize (= upper bounds of their encoded lengths), for the pool offsets
__global__ void k_enc_newlens(WordTable wt, const uint32_t* __restrict__ new_list, uint32_t n_new, ull* __restrict__ u_len) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_new; i += gridDim.x * blockDim.x) u_len[i] = wt.len[new_list[i]];
}

// One warp per distinct word.  The ids live in shared memory (words up to ENC_SM_WORD symbols) or in the word's own range of
// the pool (longer words) and are rewritten in place, 32 positions per step:
//   pass 1  every lane looks up its adjacent pairs, warp-min of the merge ids
//   pass 2  match flags by ballot; for a pair (a, a) the leftmost non-overlapping choice inside a run of matches is
//           "even distance from the start of the run" (run parity is carried across steps); survivors are left-packed
constexpr uint32_t ENC_SM_WORD = 128, ENC_WARPS = 8;

__global__ void __launch_bounds__(ENC_WARPS * 32) k_enc_words(const uint8_t* __restrict__ text, WordTable wt, const uint32_t* __restrict__ u_slot,
                                                               const ull* __restrict__ u_off, uint32_t n_unique, MergeTable mt, int32_t* __restrict__ pool,
                                                               ull pool_base, uint32_t* __restrict__ enc_len, ull* __restrict__ enc_off) {
  __shared__ int32_t sm[ENC_WARPS][ENC_SM_WORD];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const uint32_t n_warps = gridDim.x * ENC_WARPS;
  for (uint32_t u = blockIdx.x * ENC_WARPS + warp; u < n_unique; u += n_warps) {
    const uint32_t slot = u_slot[u];
    const uint32_t len = wt.len[slot];
    const ull first = wt.first[slot], base = pool_base + u_off[u];
    int32_t* w = len <= ENC_SM_WORD ? sm[warp] : pool + base;
    for (uint32_t p = lane; p < len; p += 32) w[p] = text[first + p];
    __syncwarp();
    uint32_t n = len;
    while (n >= 2) {
      int32_t mine = NO_MERGE;
      int2 ab = make_int2(-1, -1);  // the pair behind `mine`
      for (uint32_t p = lane; p + 1 < n; p += 32) {
        const int32_t x = w[p], y = w[p + 1];
        const int32_t r = merge_lookup(mt, x, y);
        if (r < mine) { mine = r; ab = make_int2(x, y); }
      }
      const int32_t best = __reduce_min_sync(0xFFFFFFFFu, mine);  // redux.sync: one instruction
      if (best == NO_MERGE) break;
      {  // a merge id belongs to exactly one pair: take it from any lane that found it
        const int src = __ffs(__ballot_sync(0xFFFFFFFFu, mine == best)) - 1;
        ab.x = __shfl_sync(0xFFFFFFFFu, ab.x, src);
        ab.y = __shfl_sync(0xFFFFFFFFu, ab.y, src);
      }
      uint32_t out = 0, carry = 0, prev_taken = 0;
      for (uint32_t c = 0; c < n; c += 32) {
        const uint32_t p = c + lane;
        const int32_t id = p < n ? w[p] : -1;
        const int32_t nx = p + 1 < n ? w[p + 1] : -2;
        __syncwarp();  // every read of this step happens before its writes
        const bool m = id == ab.x && nx == ab.y;
        const uint32_t ones = __ballot_sync(0xFFFFFFFFu, m);
        bool take = m;
        if (ab.x == ab.y) {
          const uint32_t below = ~ones & lt;  // lanes below me that do not match
          const uint32_t run = below ? lane - (31u - __clz(below)) - 1u : lane + carry;
          take = m && !(run & 1u);
          const uint32_t top = __clz(~ones);  // matches at the top end of this step (32 when all match)
          carry = (top == 32u ? carry : top) & 1u;
        }
        const uint32_t tk = __ballot_sync(0xFFFFFFFFu, take);
        const bool drop = lane ? (tk >> (lane - 1)) & 1u : prev_taken;  // second element of a merged pair
        prev_taken = tk >> 31;
        const bool keep = p < n && !drop;
        const uint32_t kb = __ballot_sync(0xFFFFFFFFu, keep);
        if (keep) w[out + __popc(kb & lt)] = take ? best : id;
        out += __popc(kb);
        __syncwarp();
      }
      n = out;
    }
    if (len <= ENC_SM_WORD) for (uint32_t p = lane; p < n; p += 32) pool[base + p] = w[p];
    if (lane == 0) { enc_len[slot] = n; enc_off[slot] = base; }
    __syncwarp();
  }
}

// ---- occurrences.  A "unit" is the 4096 corpus bytes one 256-thread block handles per step (16 bytes per thread); units are
// aligned in the whole text, a call works on the occurrences that START in the window [lo, hi) (one piece of the text).
constexpr uint32_t UNIT_BYTES = 4096;

__device__ __forceinline__ uint32_t unit_starts(const uint8_t* __restrict__ text, uint64_t lo, uint64_t hi, uint64_t unit, uint64_t* base_out) {
  const uint64_t base = unit * UNIT_BYTES + static_cast<uint64_t>(threadIdx.x) * 16u;
  *base_out = base;
  if (base >= hi || base + 16 <= lo) return 0;
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(text + base));
  uint32_t starts = start_mask16(v, base ? text[base - 1] : 32u);
  if (base + 16 > hi) starts &= (1u << (hi - base)) - 1u;
  if (base < lo) starts &= ~((1u << (lo - base)) - 1u);
  return starts;
}

__global__ void __launch_bounds__(256) k_enc_count_starts(const uint8_t* __restrict__ text, uint64_t lo, uint64_t hi, uint64_t u0, uint64_t n_units,
                                                          ull* __restrict__ unit_cnt) {
  __shared__ uint32_t ws[8];
  for (uint64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    uint64_t base;
    uint32_t c = __popc(unit_starts(text, lo, hi, u0 + unit, &base));
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31u) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t s = 0; for (int i = 0; i < 8; i++) s += ws[i]; unit_cnt[unit] = s; }
    __syncthreads();
  }
}

// Tokenise for the encoder: every occurrence of the window (in text order; unit_base = occurrences before each unit) finds or
// claims the word-table slot of its word and records it; newly claimed words are listed.  Same table as the trainer's
// k_tokenize, minus the occurrence counts.  The table persists across the pieces of one call.
__global__ void __launch_bounds__(256) k_enc_tokenize(const uint8_t* __restrict__ text, uint64_t lo, uint64_t hi, uint64_t u0, uint64_t n_units,
                                                      const ull* __restrict__ unit_base, WordTable wt, DevCounters* ctr, uint32_t seed, uint32_t* __restrict__ tok_slot,
                                                      uint32_t* __restrict__ new_list, uint32_t* __restrict__ new_n) {
  __shared__ uint32_t ws[8];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t my_claims = 0;
  for (uint64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    uint64_t base;
    uint32_t starts = unit_starts(text, lo, hi, u0 + unit, &base);
    const uint32_t c = __popc(starts);
    uint32_t incl = c;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= static_cast<uint32_t>(o)) incl += y; }
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    uint32_t before = 0;
    for (uint32_t i = 0; i < warp; i++) before += ws[i];
    ull t = unit_base[unit] + before + incl - c;
    __syncthreads();
    while (starts) {
      const int i = __ffs(starts) - 1;
      starts &= starts - 1;
      uint32_t len, dj;
      const uint64_t tag = token_walk(text, base + i, seed, &len, &dj);
      tok_slot[t++] = static_cast<uint32_t>(word_insert<false, true>(text, wt, ctr, base + i, tag, len, dj, &my_claims, new_list, new_n));
    }
  }
  for (int o = 16; o; o >>= 1) my_claims += __shfl_down_sync(0xFFFFFFFFu, my_claims, o);
  if (lane == 0 && my_claims) atomicAdd(&ctr->n_unique, my_claims);
}

// encoded length of every occurrence (scanned into the CSR offsets of the output afterwards)
__global__ void k_enc_toklen(const uint32_t* __restrict__ tok_slot, uint64_t n_tok, const uint32_t* __restrict__ enc_len, ull* __restrict__ tok_len) {
  for (uint64_t t = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; t < n_tok; t += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    tok_len[t] = enc_len[tok_slot[t]];
}

// One warp per 32 consecutive items (word occurrences when encoding, token ids when decoding): item i owns out[off[i], off[i+1])
// and copies it from src_pool[src ...], src = src_of[i] (decode) or slot_src[slot_of[i]] (encode: the word's place in the pool).
// Lanes walk the OUTPUT positions, so stores are fully coalesced; the owning item of a position is found by a 5-step binary
// search over the 32 offsets held in the lanes.
template <typename T>
__global__ void __launch_bounds__(256) k_expand(const ull* __restrict__ off, const ull* __restrict__ src_of, const uint32_t* __restrict__ slot_of,
                                                const ull* __restrict__ slot_src, uint64_t n_items, const T* __restrict__ src_pool, T* __restrict__ out) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t n_warps = static_cast<uint64_t>(gridDim.x) * (blockDim.x >> 5);
  for (uint64_t i0 = (static_cast<uint64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32u; i0 < n_items; i0 += n_warps * 32u) {
    const uint64_t i = i0 + lane;
    const ull my_off = off[i < n_items ? i : n_items];
    const ull my_src = i < n_items ? (slot_of ? slot_src[slot_of[i]] : src_of[i]) : 0ull;
    const ull first = __shfl_sync(0xFFFFFFFFu, my_off, 0);
    const uint64_t last_item = i0 + 32 < n_items ? i0 + 32 : n_items;
    const uint32_t total = static_cast<uint32_t>(off[last_item] - first);
    const uint32_t rel = static_cast<uint32_t>(my_off - first);
    for (uint32_t j0 = 0; j0 < total; j0 += 32) {
      const uint32_t j = j0 + lane;
      uint32_t lo = 0, hi = 31;  // last lane whose range starts at or before j
#pragma unroll
      for (int s = 0; s < 5; s++) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        const uint32_t v = __shfl_sync(0xFFFFFFFFu, rel, mid);
        if (v <= j) lo = mid; else hi = mid - 1;
      }
      const ull s0 = __shfl_sync(0xFFFFFFFFu, my_src, lo);
      const uint32_t r0 = __shfl_sync(0xFFFFFFFFu, rel, lo);
      if (j < total) out[first + j] = src_pool[s0 + (j - r0)];
    }
  }
}

// Small results the host needs between launches go to mapped pinned memory: a copy-engine readback would queue behind the
// result copy of the previous piece (2 ms at 64 MB pieces).
__global__ void k_mirror(const ull* __restrict__ src, volatile ull* dst_mapped, uint32_t n_words) {
  for (uint32_t i = threadIdx.x; i < n_words; i += blockDim.x) dst_mapped[i] = src[i];
  __threadfence_system();
}

// streamed encode: the CSR offsets of a piece become global
__global__ void k_add_u64(ull* __restrict__ a, uint64_t n, ull base) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) a[i] += base;
}

// decode: byte length and source of every id; an id outside [0, vocab) raises the error flag
__global__ void k_dec_lens(const int32_t* __restrict__ ids, uint64_t n, const ull* __restrict__ tok_off, uint32_t vocab, ull* __restrict__ len, ull* __restrict__ src,
                           uint32_t* __restrict__ bad) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const int32_t id = ids[i];
    if (id < 0 || static_cast<uint32_t>(id) >= vocab) { atomicOr(bad, 1u); len[i] = 0; src[i] = 0; }
    else { const ull a = tok_off[id]; len[i] = tok_off[id + 1] - a; src[i] = a; }
  }
}
