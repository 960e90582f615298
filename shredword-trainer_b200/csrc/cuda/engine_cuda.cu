// engine_cuda.cu -- the B200 (sm_100a) device engine of the BPE trainer: corpus ingest, pair counting, merge
// application and pair-table maintenance as hand-written CUDA kernels.  Implements shred::Engine (../engine.hpp).
//
// Data layout in HBM
//   ids[]   int32, one flat array holding every unique word back to back in reference word order (A3):
//             [HDR|wi] s0 s1 ... s(len-1) [DEAD ...]          HDR|wi < -1, symbols >= 0, DEAD == -1
//           A word keeps its slot between compactions; merges left-pack its live symbols and fill the tail with DEAD.
//           Because words are stored in reference scan order, "flat position" is monotone in the reference's
//           (word index, position) order and serves as the sequence number the host needs (Appendix A14).
//   wcnt[]  uint64 word counts, woff[] uint64 slot offsets (N+1), wlen[] uint32 live lengths.
//   pair table   open addressing, uint64 key (first<<32|second) -> uint64 freq        (reference BIMap, hash.cpp:104-130)
//   delta table  open addressing scratch, key -> (sum of +/-count, min sequence)      (reference FreqChangeMap, bpe.cpp:9-38)
//
// Kernels (reference loop each one replaces)
//   k_tokenize ........ bpe.cpp:131-153 + hash.cpp:29-53   tokenise on \t\r\n space, unique-word table insert
//   k_hist ............ histogram.cpp:30-36                unweighted byte histogram over unique words
//   k_scatter/k_sort_buckets  hash.cpp:61-72               word order = (djb2 & 4095, first occurrence)
//   k_symbolize ....... histogram.cpp:7-27                 bytes -> ids with unk substitution
//   k_count ........... bpe.cpp:187-218                    adjacent pair counts
//   k_merge ........... bpe.cpp:265-318                    one cooperative launch per merge: HBM-bound scan of the candidate tiles
//                                                          with per-occurrence count deltas | grid barrier | deltas folded into the
//                                                          pair table + records published | in-place rewrite of the touched words
//   k_token_freq ...... bpe.cpp:409-415                    final token frequencies
#include <cuda_runtime.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../../include/shred_abi.h"
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>

#include "../charset.hpp"
#include "../engine.hpp"
#include "../shard.hpp"

namespace shred {
namespace {

#define CK(call)                                                                                              \
  do {                                                                                                        \
    cudaError_t e_ = (call);                                                                                  \
    if (e_ != cudaSuccess) {                                                                                  \
      std::fprintf(stderr, "[ERROR]\t CUDA: %s -> %s (%s:%d)\n", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return -1;                                                                                              \
    }                                                                                                         \
  } while (0)

#define RC(call)            \
  do {                      \
    int rc_ = (call);       \
    if (rc_ != 0) return rc_; \
  } while (0)

typedef unsigned long long ull;

constexpr int32_t DEAD = -1;
constexpr uint32_t HDR_BIT = 0x80000000u;
constexpr int32_t UNK_CODE_NEG = 0x7FFFFFFF;  // stored code of unk symbols when unk_id < 0
constexpr uint64_t PT_EMPTY = ~0ull;
constexpr uint64_t SEQ_MAX = ~0ull;
constexpr int N_SM_FALLBACK = 148;

enum : uint32_t { ERR_DT_FULL = 1, ERR_PT_FULL = 2, ERR_WT_FULL = 4, ERR_WT_COLLISION = 8, ERR_REC_FULL = 16, ERR_HAS_NUL = 32, ERR_BARRIER = 64,
                  ERR_PEER_TIMEOUT = 128, ERR_INBOX_FULL = 256 };

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
__device__ __forceinline__ bool is_delim(uint32_t c) { return c <= 32u && ((0x100002600ull >> c) & 1ull); }  // \t \n \r space
__device__ __forceinline__ uint64_t fc_key(int32_t a, int32_t b) {  // bpe.cpp:277-278: both operands sign-extend
  return (static_cast<uint64_t>(static_cast<int64_t>(a)) << 32) | static_cast<uint64_t>(static_cast<int64_t>(b));
}

struct Ctrl {  // mapped pinned host memory, written by finalize_block
  volatile uint64_t flag;
  uint64_t n_recs, occ, occ_local, pt_n, n_leaders, n_keys, cand_tiles;
  uint32_t err, pad;
};

struct DevCounters {  // device memory
  uint32_t wl_n, dt_n, rec_n, blocks_done;
  ull occ;
  ull pt_n;
  uint32_t err, pad;
  ull n_tokens;
  uint32_t n_unique, blocks_done2;
  uint32_t cand_tiles, bar;
  uint32_t sent_ctas, pad4;
};

struct DeltaTable {
  uint64_t* keys; ull* delta; ull* seq; uint32_t* list; uint64_t* klist;  // list/klist: slot and key of every used slot, dense
  uint64_t mask; uint64_t empty; uint32_t cap;
};
struct PairEnt { uint64_t key; uint64_t freq; };  // one 16-byte load fetches both
struct PairTable {
  PairEnt* ent;
  uint32_t* serial;  // dense id per entry = number of entries that existed when it was created (the host indexes by it)
  uint64_t mask; uint64_t cap;
};

// ------------------------------------------------------------------------------------------------ hash-table helpers

__device__ __forceinline__ void dt_add(const DeltaTable& dt, DevCounters* ctr, uint64_t key, int64_t delta, uint64_t seq) {
  uint64_t slot = mix64(key) & dt.mask;
  for (uint32_t probe = 0; probe < dt.cap; ++probe) {
    uint64_t cur = dt.keys[slot];
    if (cur == dt.empty) {
      uint64_t prev = atomicCAS(reinterpret_cast<ull*>(&dt.keys[slot]), static_cast<ull>(dt.empty), static_cast<ull>(key));
      if (prev == dt.empty) {
        uint32_t idx = atomicAdd(&ctr->dt_n, 1u);
        if (idx < dt.cap) { dt.list[idx] = static_cast<uint32_t>(slot); dt.klist[idx] = key; }
        cur = key;
      } else cur = prev;
    }
    if (cur == key) {
      atomicAdd(&dt.delta[slot], static_cast<ull>(delta));
      atomicMin(&dt.seq[slot], static_cast<ull>(seq));
      return;
    }
    slot = (slot + 1) & dt.mask;
  }
  atomicOr(&ctr->err, ERR_DT_FULL);
}

__device__ __forceinline__ ull gtime() { ull t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ ulonglong2 ld_ent(const PairEnt* e) { return *reinterpret_cast<const ulonglong2*>(e); }

// Finds `key` (inserting it if absent) and returns its slot; *old_freq = its frequency (0 for a new entry).
// `first` is the already-loaded entry at the home slot (lets the caller issue several home loads back to back).
// Only finalize_block calls this, with distinct keys per pass, so a claimed entry has exactly one writer.
__device__ __forceinline__ uint64_t pt_find_or_insert(const PairTable& pt, DevCounters* ctr, uint64_t key, ulonglong2 first, uint64_t* old_freq) {
  uint64_t slot = mix64(key) & pt.mask;
  ulonglong2 e = first;
  for (uint64_t probe = 0; probe < pt.cap; ++probe) {
    if (e.x == key) { *old_freq = e.y; return slot; }
    if (e.x == PT_EMPTY) {
      const uint64_t prev = atomicCAS(reinterpret_cast<ull*>(&pt.ent[slot].key), static_cast<ull>(PT_EMPTY), static_cast<ull>(key));
      if (prev == PT_EMPTY) { pt.serial[slot] = static_cast<uint32_t>(atomicAdd(&ctr->pt_n, 1ull)); *old_freq = 0; return slot; }
      if (prev == key) { *old_freq = pt.ent[slot].freq; return slot; }
    }
    slot = (slot + 1) & pt.mask;
    e = ld_ent(&pt.ent[slot]);
  }
  atomicOr(&ctr->err, ERR_PT_FULL);
  *old_freq = 0;
  return 0;
}

// ------------------------------------------------------------------------------------------------------------ ingest

struct WordTable {
  ull* tag;       // 0 = empty
  ull* first;     // smallest byte offset of an occurrence
  ull* count;
  uint32_t* len;
  uint32_t* bucket;  // djb2 & 4095
  uint64_t mask, cap;
};

// Each thread owns 16 consecutive corpus bytes (one uint4 load) and inserts every token that STARTS inside them.
// text is padded with >= 32 spaces, so token walks terminate.
__global__ void __launch_bounds__(256) k_tokenize(const uint8_t* __restrict__ text, uint64_t n, WordTable wt, DevCounters* ctr, uint32_t seed) {
  const uint64_t n16 = (n + 15) >> 4;
  uint32_t my_tokens = 0;
  for (uint64_t t = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; t < n16; t += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint64_t base = t << 4;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(text + base));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t prev = base ? text[base - 1] : 32u;
    // a NUL byte hides the rest of its line in the reference (fgets + strlen, bpe.cpp:131-147): report it, the host
    // blanks the hidden spans and loads again
    if (((v.x - 0x01010101u) & ~v.x & 0x80808080u) | ((v.y - 0x01010101u) & ~v.y & 0x80808080u) | ((v.z - 0x01010101u) & ~v.z & 0x80808080u) |
        ((v.w - 0x01010101u) & ~v.w & 0x80808080u))
      atomicOr(&ctr->err, ERR_HAS_NUL);
    // delimiter mask of my 16 bytes
    uint32_t dm = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) { uint32_t c = (w[i >> 2] >> ((i & 3) * 8)) & 255u; dm |= (is_delim(c) ? 1u : 0u) << i; }
    uint32_t starts = ~dm & ((dm << 1) | (is_delim(prev) ? 1u : 0u)) & 0xFFFFu;
    while (starts) {
      const int i = __ffs(starts) - 1;
      starts &= starts - 1;
      const uint64_t off = base + i;
      if (off >= n) break;
      // walk the token: two 32-bit multiplicative hashes (placement tag) + djb2 (reference bucket, hash.cpp:35-39)
      uint32_t h1 = 2166136261u ^ seed, h2 = 0x9E3779B9u + seed, dj = 5381u, len = 0;
      for (;;) {
        const uint32_t c = text[off + len];
        if (is_delim(c)) break;
        h1 = (h1 ^ c) * 16777619u;
        h2 = (h2 + c) * 0x85EBCA6Bu; h2 ^= h2 >> 15;
        dj = dj * 33u + c;
        ++len;
      }
      ++my_tokens;
      const uint64_t tag = mix64((static_cast<uint64_t>(h1) << 32) | h2 | 0) | 1ull;
      uint64_t slot = tag & wt.mask;
      bool done = false;
      for (uint32_t probe = 0; probe < 8192u && !done; ++probe) {
        ull cur = wt.tag[slot];
        if (cur == 0ull) {
          ull prevt = atomicCAS(&wt.tag[slot], 0ull, static_cast<ull>(tag));
          if (prevt == 0ull) {  // claimed: publish the immutable facts
            wt.len[slot] = len;
            wt.bucket[slot] = dj & 4095u;
            atomicAdd(&ctr->n_unique, 1u);
            cur = tag;
          } else cur = prevt;
        }
        if (cur == tag) {
          // first occurrence: most tokens come after the word's first sighting, so look before paying for an atomic
          ull old = *reinterpret_cast<volatile ull*>(&wt.first[slot]);
          if (off < old) old = atomicMin(&wt.first[slot], static_cast<ull>(off));
          {  // count: lanes of this warp that hit the same slot right now add once (hot words are most of a Zipf corpus)
            const unsigned am = __activemask();
            const unsigned grp = __match_any_sync(am, slot);
            if ((threadIdx.x & 31u) == static_cast<unsigned>(__ffs(grp) - 1)) atomicAdd(&wt.count[slot], static_cast<ull>(__popc(grp)));
          }
          if (old != SEQ_MAX && old != off) {  // same tag: must be the same bytes, else retry ingest with a new seed
            bool same = is_delim(text[old + len]);
            for (uint32_t j = 0; j < len && same; j++) same = text[old + j] == text[off + j];
            if (!same) atomicOr(&ctr->err, ERR_WT_COLLISION);
          }
          done = true;
        } else slot = (slot + 1) & wt.mask;
      }
      if (!done) atomicOr(&ctr->err, ERR_WT_FULL);
    }
  }
  // token count: warp reduce, one atomic per warp
  for (int o = 16; o; o >>= 1) my_tokens += __shfl_down_sync(0xFFFFFFFFu, my_tokens, o);
  if ((threadIdx.x & 31) == 0 && my_tokens) atomicAdd(&ctr->n_tokens, static_cast<ull>(my_tokens));
}

// unique slots -> dense list + per-bucket population
__global__ void k_collect(WordTable wt, uint32_t* u_slot, uint32_t* u_n, uint32_t* bucket_cnt) {
  for (uint64_t s = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; s < wt.cap; s += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    if (wt.tag[s] != 0ull) {
      uint32_t idx = atomicAdd(u_n, 1u);
      u_slot[idx] = static_cast<uint32_t>(s);
      atomicAdd(&bucket_cnt[wt.bucket[s]], 1u);
    }
  }
}

// exclusive scan of 4096 bucket counts by one block of 1024 threads
__global__ void k_scan4096(const uint32_t* cnt, uint32_t* start) {
  __shared__ uint32_t part[1024];
  const int t = threadIdx.x;
  uint32_t c[4], s = 0;
  for (int i = 0; i < 4; i++) { c[i] = cnt[t * 4 + i]; s += c[i]; }
  part[t] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) { uint32_t v = t >= o ? part[t - o] : 0; __syncthreads(); part[t] += v; __syncthreads(); }
  uint32_t run = part[t] - s;
  for (int i = 0; i < 4; i++) { start[t * 4 + i] = run; run += c[i]; }
  if (t == 1023) start[4096] = run;
}

__global__ void k_scatter(WordTable wt, const uint32_t* u_slot, uint32_t n, const uint32_t* bstart, uint32_t* cursor, uint32_t* tmp_slot, ull* tmp_first) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t s = u_slot[i], b = wt.bucket[s];
    const uint32_t pos = bstart[b] + atomicAdd(&cursor[b], 1u);
    tmp_slot[pos] = s;
    tmp_first[pos] = wt.first[s];
  }
}

// order of the words inside a djb2 bucket = first occurrence (all first offsets are distinct): wi = bucket_start + rank.
// One CTA per bucket: bitonic sort of (first_offset << 24 | index) in shared memory.  Buckets larger than SORT_CAP
// (more than ~30 M unique words) are left to k_rank_big.
constexpr uint32_t SORT_CAP = 8192, SORT_THREADS = 512;
__global__ void __launch_bounds__(SORT_THREADS) k_sort_buckets(const uint32_t* __restrict__ tmp_slot, const ull* __restrict__ tmp_first, const uint32_t* __restrict__ bstart,
                                                               uint32_t* __restrict__ order_slot) {
  extern __shared__ ull sk[];
  const uint32_t b = blockIdx.x, bs = bstart[b], n = bstart[b + 1] - bs;
  if (n == 0 || n > SORT_CAP) return;
  uint32_t np = 1;
  while (np < n) np <<= 1;
  for (uint32_t i = threadIdx.x; i < np; i += blockDim.x) sk[i] = i < n ? ((tmp_first[bs + i] << 24) | i) : ~0ull;
  __syncthreads();
  for (uint32_t k = 2; k <= np; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = threadIdx.x; i < np; i += blockDim.x) {
        const uint32_t ixj = i ^ j;
        if (ixj > i) {
          const ull a = sk[i], c = sk[ixj];
          const bool up = (i & k) == 0;
          if ((a > c) == up) { sk[i] = c; sk[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) order_slot[bs + i] = tmp_slot[bs + static_cast<uint32_t>(sk[i] & 0xFFFFFFull)];
}

// fallback for oversized buckets: each element counts the smaller first offsets in its bucket (O(n_b^2), L1 broadcast)
__global__ void k_rank_big(WordTable wt, const uint32_t* tmp_slot, const ull* tmp_first, uint32_t n, const uint32_t* bstart, uint32_t* order_slot) {
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const uint32_t s = tmp_slot[e], b = wt.bucket[s];
    const uint32_t bs = bstart[b], be = bstart[b + 1];
    if (be - bs <= SORT_CAP) continue;
    const ull mine = tmp_first[e];
    uint32_t rank = 0;
    for (uint32_t j = bs; j < be; j++) rank += tmp_first[j] < mine ? 1u : 0u;
    order_slot[bs + rank] = s;
  }
}

// unweighted byte histogram over unique words (histogram.cpp:30-36) + per-word facts in reference order
__global__ void __launch_bounds__(256) k_hist_words(const uint8_t* __restrict__ text, WordTable wt, const uint32_t* order_slot, uint32_t n,
                                                    ull* hist, ull* wcnt, uint32_t* wlen, ull* len1) {
  __shared__ uint32_t sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  for (uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x; wi < n; wi += gridDim.x * blockDim.x) {
    const uint32_t s = order_slot[wi];
    const uint32_t len = wt.len[s];
    const ull first = wt.first[s];
    wcnt[wi] = wt.count[s];
    wlen[wi] = len;
    len1[wi] = static_cast<ull>(len) + 1ull;
    for (uint32_t j = 0; j < len; j++) atomicAdd(&sh[text[first + j]], 1u);
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], static_cast<ull>(sh[threadIdx.x]));
}

__global__ void __launch_bounds__(256) k_symbolize(const uint8_t* __restrict__ text, WordTable wt, const uint32_t* order_slot, uint32_t n,
                                                   const ull* woff, const uint8_t* keep, int32_t unk_code, int32_t* ids, uint32_t* wid) {
  __shared__ uint8_t sk[256];
  sk[threadIdx.x] = keep[threadIdx.x];
  __syncthreads();
  for (uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x; wi < n; wi += gridDim.x * blockDim.x) {
    const uint32_t s = order_slot[wi];
    const uint32_t len = wt.len[s];
    const ull first = wt.first[s];
    const ull base = woff[wi];
    ids[base] = static_cast<int32_t>(HDR_BIT | wi);
    wid[base] = wi;
    for (uint32_t j = 0; j < len; j++) { const uint32_t c = text[first + j]; ids[base + 1 + j] = sk[c] ? static_cast<int32_t>(c) : unk_code; wid[base + 1 + j] = wi; }
  }
}

__global__ void k_fill_i32(int32_t* p, uint64_t from, uint64_t to, int32_t v) {
  for (uint64_t i = from + blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < to; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) p[i] = v;
}
__global__ void k_fill_u64(ull* p, uint64_t n, ull v) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) p[i] = v;
}

// ---- device-wide exclusive scan of uint64 (three passes; 2048 items per block) used at load and at compaction
constexpr int SCAN_ITEMS = 8, SCAN_THREADS = 256, SCAN_TILE = SCAN_ITEMS * SCAN_THREADS;

__device__ __forceinline__ ull block_excl_scan(ull v, ull* total) {  // 256 threads
  __shared__ ull wsum[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  ull x = v;
  for (int o = 1; o < 32; o <<= 1) { ull y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  if (warp == 0) {
    ull s = lane < 8 ? wsum[lane] : 0;
    for (int o = 1; o < 8; o <<= 1) { ull y = __shfl_up_sync(0xFFFFFFFFu, s, o); if (lane >= o) s += y; }
    if (lane < 8) wsum[lane] = s;
  }
  __syncthreads();
  const ull before = warp ? wsum[warp - 1] : 0;
  *total = wsum[7];
  __syncthreads();
  return before + x - v;
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(const ull* in, uint64_t n, ull* sums) {
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * SCAN_TILE + static_cast<uint64_t>(threadIdx.x) * SCAN_ITEMS;
  ull s = 0;
  for (int i = 0; i < SCAN_ITEMS; i++) if (base + i < n) s += in[base + i];
  ull total;
  block_excl_scan(s, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_top(ull* sums, uint32_t nb, ull* grand_total) {  // one block
  ull carry = 0;
  for (uint32_t base = 0; base < nb; base += SCAN_THREADS) {
    const uint32_t i = base + threadIdx.x;
    ull v = i < nb ? sums[i] : 0, total;
    ull ex = block_excl_scan(v, &total);
    if (i < nb) sums[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) *grand_total = carry;
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const ull* in, uint64_t n, const ull* sums, ull* out) {
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * SCAN_TILE + static_cast<uint64_t>(threadIdx.x) * SCAN_ITEMS;
  ull v[SCAN_ITEMS], s = 0;
  for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = base + i < n ? in[base + i] : 0; s += v[i]; }
  ull total;
  ull run = sums[blockIdx.x] + block_excl_scan(s, &total);
  for (int i = 0; i < SCAN_ITEMS; i++) { if (base + i < n) out[base + i] = run; run += v[i]; }
}

// -------------------------------------------------------------------------------------------------------------- count

struct Params {
  int32_t unk_id, unk_code;
  uint64_t min_freq;
};

__device__ __forceinline__ int32_t code_to_id(int32_t code, const Params& P) { return (P.unk_id < 0 && code == P.unk_code) ? P.unk_id : code; }

// bpe.cpp:197-214: every adjacent pair without unk adds the word's count; first sighting = flat position.
// Flat, coalesced pass over the symbol array (int4 of ids + int4 of word indices per thread); the handful of distinct
// pairs of a fresh corpus would serialise on global atomics, so each CTA first aggregates into a shared-memory hash
// table (sum of counts, min position) and flushes one delta-table update per distinct pair at the end.
constexpr uint32_t CNT_SLOTS = 2048, CNT_PROBES = 12;
__device__ __forceinline__ void cnt_add(ull* s_key, ull* s_sum, ull* s_seq, const DeltaTable& dt, DevCounters* ctr, uint64_t key, uint64_t c, uint64_t seq) {
  uint32_t slot = static_cast<uint32_t>((key * 0x9E3779B97F4A7C15ull) >> 53) & (CNT_SLOTS - 1);
  for (uint32_t probe = 0; probe < CNT_PROBES; ++probe) {
    ull cur = s_key[slot];
    if (cur == ~0ull) { const ull prev = atomicCAS(&s_key[slot], ~0ull, static_cast<ull>(key)); cur = prev == ~0ull ? key : prev; }
    if (cur == key) {
      atomicAdd(&s_sum[slot], static_cast<ull>(c));
      if (seq < *reinterpret_cast<volatile ull*>(&s_seq[slot])) atomicMin(&s_seq[slot], static_cast<ull>(seq));  // positions grow along the grid-stride loop: rarely taken
      return;
    }
    slot = (slot + 1) & (CNT_SLOTS - 1);
  }
  dt_add(dt, ctr, key, static_cast<int64_t>(c), seq);  // shared table crowded: straight to the global one
}
__global__ void __launch_bounds__(256) k_count(const int4* __restrict__ ids4, const uint4* __restrict__ wid4, uint32_t n4, const ull* __restrict__ wcnt, Params P,
                                               DeltaTable dt, DevCounters* ctr, uint64_t seq_base) {
  __shared__ ull s_key[CNT_SLOTS], s_sum[CNT_SLOTS], s_seq[CNT_SLOTS];
  for (uint32_t i = threadIdx.x; i < CNT_SLOTS; i += blockDim.x) { s_key[i] = ~0ull; s_sum[i] = 0ull; s_seq[i] = SEQ_MAX; }
  __syncthreads();
  const int32_t* ids = reinterpret_cast<const int32_t*>(ids4);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t n4_ceil = (n4 + 31u) & ~31u;  // whole warps stay in the loop so the shuffle below is full-width
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4_ceil; i += gridDim.x * blockDim.x) {
    const bool in = i < n4;
    const int4 v = in ? __ldg(ids4 + i) : make_int4(DEAD, DEAD, DEAD, DEAD);
    int32_t nxt = __shfl_down_sync(0xFFFFFFFFu, v.x, 1);
    if (lane == 31) nxt = (i + 1 < n4) ? __ldg(ids + 4 * (static_cast<uint64_t>(i) + 1)) : DEAD;
    const int32_t s[5] = {v.x, v.y, v.z, v.w, nxt};
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) m |= (s[k] >= 0 && s[k + 1] >= 0 && s[k] != P.unk_code && s[k + 1] != P.unk_code) ? (1u << k) : 0u;
    if (m) {
      const uint4 w = __ldg(wid4 + i);
      const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int k = 0; k < 4; k++) if (m & (1u << k))
        cnt_add(s_key, s_sum, s_seq, dt, ctr, fc_key(s[k], s[k + 1]), wcnt[ws[k]], seq_base | (4ull * i + k));
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < CNT_SLOTS; i += blockDim.x)
    if (s_key[i] != ~0ull) dt_add(dt, ctr, s_key[i], static_cast<int64_t>(s_sum[i]), s_seq[i]);
}

// -------------------------------------------------------------------------------------------------------------- merge

// ---- tile occurrence index ------------------------------------------------------------------------------------
// planes[id * W + (t >> 5)] bit (t & 31) is set if token `id` occurs in slots [512 t, 512 t + 512] (the first slot of
// the next tile included, so a pair that straddles the boundary is found from tile t).  Bits are only ever added between
// two compactions, so the index is a superset of the truth: a merge scans exactly the tiles whose bit is set for both A
// and B and provably misses nothing.  Late merges touch a few hundred of tens of thousands of tiles.
constexpr uint32_t TILE_SHIFT = 9, TILE_SLOTS = 1u << TILE_SHIFT, TILE_I4 = TILE_SLOTS / 4, MAX_TILES_PER_CTA = 1024;
static_assert(TILE_I4 % (32 * 4) == 0, "a tile is a whole number of warp chunks");

__device__ __forceinline__ void plane_set(uint32_t* planes, uint32_t W, uint32_t id_cap, int32_t id, uint64_t slot) {
  if (id < 0 || static_cast<uint32_t>(id) >= id_cap) return;
  uint32_t t = static_cast<uint32_t>(slot >> TILE_SHIFT);
  uint32_t* wp = planes + static_cast<uint64_t>(id) * W + (t >> 5);
  uint32_t bit = 1u << (t & 31);
  if (!(*wp & bit)) atomicOr(wp, bit);
  if ((slot & (TILE_SLOTS - 1)) == 0 && t > 0) {
    --t;
    wp = planes + static_cast<uint64_t>(id) * W + (t >> 5);
    bit = 1u << (t & 31);
    if (!(*wp & bit)) atomicOr(wp, bit);
  }
}

__global__ void __launch_bounds__(256) k_build_planes(const int32_t* __restrict__ ids, uint64_t n_slots, uint32_t* planes, uint32_t W, uint32_t id_cap) {
  for (uint64_t p = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; p < n_slots; p += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    plane_set(planes, W, id_cap, ids[p], p);
}

// Shared by the count pass (COUNT=true, bpe.cpp:219-227) and the merge pass (bpe.cpp:297-318): fold the aggregated deltas
// into the pair table and write one record per touched key for the host.  Runs in ONE block (any size): the stand-alone
// k_finalize_count kernel, or the last block of k_scan_merge to finish.  Ends by publishing the counters to the host and
// re-arming them.
template <bool COUNT>
__device__ __forceinline__ void finalize_block(const DeltaTable& dt, const PairTable& pt, DevCounters* ctr, Rec* recs, uint32_t rec_cap, Ctrl* ctrl,
                                               int32_t A, int32_t B, const Params& P, uint64_t flag_value, ull* dbg = nullptr) {
  __shared__ uint32_t s_rec_n;
  if (dbg && threadIdx.x == 0) dbg[2] = gtime();
  constexpr int ILP = 4;  // keys in flight per thread: the pass is a chain of dependent DRAM/L2 round trips
  const uint32_t n = ctr->dt_n < dt.cap ? ctr->dt_n : dt.cap;
  if (threadIdx.x == 0) s_rec_n = 0;
  __syncthreads();
  if (!COUNT && threadIdx.x == blockDim.x - 1) {  // bpe.cpp:315: the merged pair's frequency becomes 0
    const uint64_t k = fc_key(A, B);
    uint64_t old;
    const uint64_t s = pt_find_or_insert(pt, ctr, k, ld_ent(&pt.ent[mix64(k) & pt.mask]), &old);
    pt.ent[s].freq = 0ull;
  }
  for (uint32_t base = 0; base < n; base += blockDim.x * ILP) {
    uint64_t key[ILP]; uint32_t ds[ILP]; ulonglong2 home[ILP]; int64_t d[ILP]; uint64_t sq[ILP]; bool ok[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) {
      const uint32_t i = base + j * blockDim.x + threadIdx.x;
      ok[j] = i < n;
      if (ok[j]) { key[j] = dt.klist[i]; ds[j] = dt.list[i]; }
    }
#pragma unroll
    for (int j = 0; j < ILP; j++) if (ok[j]) {
      home[j] = ld_ent(&pt.ent[mix64(key[j]) & pt.mask]);
      d[j] = static_cast<int64_t>(dt.delta[ds[j]]);
      sq[j] = dt.seq[ds[j]];
    }
#pragma unroll
    for (int j = 0; j < ILP; j++) if (ok[j]) {
      dt.keys[ds[j]] = dt.empty; dt.delta[ds[j]] = 0ull; dt.seq[ds[j]] = SEQ_MAX;  // re-arm the scratch slot
      const int32_t pa = static_cast<int32_t>(key[j] >> 32), pb = static_cast<int32_t>(key[j] & 0xFFFFFFFFu);  // bpe.cpp:301
      Rec out; out.key = key[j]; out.seq = sq[j]; out.serial = REC_NO_SERIAL; out.kind = REC_PUSH; out.val = 0;
      bool emit = false;
      if (!COUNT && pa == A && pb == B) continue;  // bpe.cpp:302
      if (!COUNT && (pa == P.unk_id || pb == P.unk_id)) {  // phantom pair: tracked by the host (Appendix A12)
        out.kind = REC_PHANTOM; out.val = static_cast<uint64_t>(d[j]); emit = true;
      } else {
        uint64_t old;
        const uint64_t s = pt_find_or_insert(pt, ctr, key[j], home[j], &old);
        uint64_t nf;
        if (d[j] < 0) { const uint64_t ad = static_cast<uint64_t>(-d[j]); nf = old >= ad ? old - ad : 0; } else nf = old + static_cast<uint64_t>(d[j]);  // bpe.cpp:303-307
        pt.ent[s].freq = nf;
        if (nf >= P.min_freq) { out.kind = REC_PUSH; out.val = nf; emit = true; }            // bpe.cpp:308-311
        else if (!COUNT && old >= P.min_freq) { out.kind = REC_DEMOTE; out.val = nf; emit = true; }
        if (emit) out.serial = pt.serial[s];
      }
      if (emit) {
        const uint32_t idx = atomicAdd(&s_rec_n, 1u);
        if (idx < rec_cap) recs[idx] = out; else atomicOr(&ctr->err, ERR_REC_FULL);
      }
    }
  }
  __syncthreads();
  if (dbg && threadIdx.x == 0) dbg[3] = gtime();
  if (threadIdx.x == 0) {  // counters for the host, then re-arm them for the next pass
    ctrl->n_recs = s_rec_n < rec_cap ? s_rec_n : rec_cap;
    ctrl->occ = ctr->occ;
    ctrl->occ_local = ctr->occ;
    ctrl->pt_n = ctr->pt_n;
    ctrl->n_leaders = ctr->wl_n;
    ctrl->n_keys = ctr->dt_n;
    ctrl->cand_tiles = ctr->cand_tiles;
    ctrl->err = ctr->err;
    ctr->dt_n = 0; ctr->rec_n = 0; ctr->blocks_done = 0; ctr->occ = 0ull; ctr->cand_tiles = 0;
  }
  __threadfence_system();  // every thread's records (and thread 0's counters) are visible to the host ...
  __syncthreads();
  if (dbg && threadIdx.x == 0) dbg[4] = gtime();
  if (threadIdx.x == 0) ctrl->flag = flag_value;  // ... before the flag it spins on
}

__global__ void __launch_bounds__(256) k_finalize_count(DeltaTable dt, PairTable pt, DevCounters* ctr, Rec* recs, uint32_t rec_cap, Ctrl* ctrl, Params P,
                                                        uint64_t flag_value) {
  finalize_block<true>(dt, pt, ctr, recs, rec_cap, ctrl, 0, 0, P, flag_value);
}

// One occurrence of (A,B) at flat position p: the four count deltas of bpe.cpp:274-290, computed independently per
// occurrence.  Left neighbour = the id that stands there when the reference's left-to-right pass reaches p (N if the two
// symbols before p were themselves merged in this pass), right neighbour = the raw id two slots on.
__device__ __forceinline__ void emit_occurrence(const int32_t* ids, uint64_t p, const uint32_t* __restrict__ wid, const ull* __restrict__ wcnt,
                                                int32_t A, int32_t B, int32_t N, const Params& P, const DeltaTable& dt, DevCounters* ctr, uint32_t* ml,
                                                uint32_t& my_occ, uint64_t seq_base) {
  const int32_t l1 = ids[p - 1];
  const uint32_t wi = wid[p];  // independent loads first: wid -> wcnt is the longest chain
  const int32_t r2 = ids[p + 2];
  bool left_merged;
  if (A != B) {
    left_merged = l1 == B && ids[p - 2] == A;  // (A,B) pairs cannot overlap when A != B
  } else {
    uint64_t q = p;  // start of the run of A's: pairs are taken greedily from there (bpe.cpp:268-295)
    while (ids[q - 1] == A) --q;
    if ((p - q) & 1ull) return;  // second half of a merged pair, not an occurrence
    left_merged = p > q;
  }
  const int64_t c = static_cast<int64_t>(wcnt[wi]);
  const uint64_t seq = seq_base | (p * 4ull);
  if (l1 >= 0) {
    const int32_t lid = left_merged ? N : code_to_id(l1, P);
    dt_add(dt, ctr, fc_key(lid, A), -c, seq + 0);
    dt_add(dt, ctr, fc_key(lid, N), c, seq + 1);
  }
  if (r2 >= 0) {
    const int32_t rid = code_to_id(r2, P);
    dt_add(dt, ctr, fc_key(B, rid), -c, seq + 2);
    dt_add(dt, ctr, fc_key(N, rid), c, seq + 3);
  }
  ml[atomicAdd(&ctr->wl_n, 1u)] = static_cast<uint32_t>(p);
  ++my_occ;
}

// ---- multi-GPU exchange over NVLink peer memory ---------------------------------------------------------------
// Every rank owns a contiguous range of the unique words and a full replica of the pair table and of the host heap.
// Per pass (count, merge, token frequencies) each rank's aggregated (key, delta, sequence) list is the only thing that
// crosses GPUs: the kernel STORES it straight into every peer's inbox (memory mapped with CUDA IPC, NVLink/NVSwitch),
// raises a sequence flag there, waits for the peers' flags in its own inbox, and folds their entries into its own delta
// table.  No host round trip and no NCCL call sits between the scan and the pair-table update.
constexpr int MAX_RANKS = 8;
constexpr uint64_t INBOX_ENTRIES = 1ull << 20, INBOX_HDR = 64, INBOX_BYTES = INBOX_HDR + INBOX_ENTRIES * 24;
struct InboxHdr { ull seq; ull n; ull aux; };
struct DistArgs {
  int rank, world;
  uint8_t* peer[MAX_RANKS];  // inbox base of every rank (peer[rank] is local memory)
  ull xseq;                  // exchange number (>= 1); its parity selects the inbox half
};
__device__ __forceinline__ uint8_t* inbox_region(uint8_t* base, int world, ull xseq, int src) {
  return base + ((xseq & 1ull) * static_cast<ull>(world) + static_cast<ull>(src)) * INBOX_BYTES;
}

// Software grid barrier for the cooperative per-merge kernel (all CTAs are co-resident: cudaLaunchCooperativeKernel).
// `bar` only ever grows; `target` = value it reaches when every CTA of this launch has arrived at this barrier.
__device__ __forceinline__ void grid_barrier(uint32_t* bar, uint32_t target, uint32_t* err) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    const long long t0 = clock64();
    while (static_cast<int32_t>(*reinterpret_cast<volatile uint32_t*>(bar) - target) < 0) {
      if (clock64() - t0 > 4000000000ll) { atomicOr(err, ERR_BARRIER); break; }  // ~2 s: never hang the GPU on a host-side accounting bug
    }
    __threadfence();
  }
  __syncthreads();
}

// Cooperative exchange of the delta table's dense list (klist/list/delta/seq) between ranks.  Must be entered after a
// grid barrier (the local list is complete); ends with one grid barrier (number `barrier_no` of this launch).
// On return the local delta table holds the GLOBAL aggregate and *occ_global the global occurrence count.
__device__ __forceinline__ void exchange_deltas(const DeltaTable& dt, DevCounters* ctr, const DistArgs& D, uint32_t bar_base, int barrier_no, ull occ_local,
                                                ull* occ_global) {
  __shared__ bool last_sender;
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  const uint32_t n_local = ctr->dt_n < dt.cap ? ctr->dt_n : dt.cap;
  if (n_local > INBOX_ENTRIES && gtid == 0) atomicOr(&ctr->err, ERR_INBOX_FULL);
  const uint32_t n_send = n_local < INBOX_ENTRIES ? n_local : static_cast<uint32_t>(INBOX_ENTRIES);
  for (uint32_t i = gtid; i < n_send; i += gthreads) {  // P2P stores into every peer's inbox
    const uint32_t ds = dt.list[i];
    const ull k = dt.klist[i], d = dt.delta[ds], sq = dt.seq[ds];
    for (int dst = 0; dst < D.world; dst++) if (dst != D.rank) {
      ull* e = reinterpret_cast<ull*>(inbox_region(D.peer[dst], D.world, D.xseq, D.rank) + INBOX_HDR) + 3ull * i;
      e[0] = k; e[1] = d; e[2] = sq;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last_sender = atomicAdd(&ctr->sent_ctas, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last_sender && threadIdx.x == 0) {  // every CTA's stores are out: announce the list to the peers
    __threadfence();
    ctr->sent_ctas = 0;
    for (int dst = 0; dst < D.world; dst++) if (dst != D.rank) {
      InboxHdr* h = reinterpret_cast<InboxHdr*>(inbox_region(D.peer[dst], D.world, D.xseq, D.rank));
      h->n = n_send; h->aux = occ_local;
    }
    __threadfence_system();
    for (int dst = 0; dst < D.world; dst++) if (dst != D.rank)
      *reinterpret_cast<volatile ull*>(&reinterpret_cast<InboxHdr*>(inbox_region(D.peer[dst], D.world, D.xseq, D.rank))->seq) = D.xseq;
  }
  if (threadIdx.x == 0) {  // every CTA waits for the peers' lists to land in MY inbox (local memory)
    const long long t0 = clock64();
    for (int src = 0; src < D.world; src++) if (src != D.rank) {
      volatile ull* f = &reinterpret_cast<InboxHdr*>(inbox_region(D.peer[D.rank], D.world, D.xseq, src))->seq;
      while (*f != D.xseq) {
        __nanosleep(64);  // hundreds of CTAs poll this line while the peer's NVLink write has to get in
        if (clock64() - t0 > 8000000000ll) { atomicOr(&ctr->err, ERR_PEER_TIMEOUT); break; }  // ~4 s: never hang the GPU
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  ull occ = occ_local;
  for (int src = 0; src < D.world; src++) if (src != D.rank) {  // fold the peers' entries into my delta table
    const uint8_t* reg = inbox_region(D.peer[D.rank], D.world, D.xseq, src);
    const InboxHdr* h = reinterpret_cast<const InboxHdr*>(reg);
    const ull n_src = __ldcv(&h->n);
    occ += __ldcv(&h->aux);
    const ull* e = reinterpret_cast<const ull*>(reg + INBOX_HDR);
    for (ull i = gtid; i < n_src && i < INBOX_ENTRIES; i += gthreads)
      dt_add(dt, ctr, __ldcv(e + 3 * i), static_cast<int64_t>(__ldcv(e + 3 * i + 1)), __ldcv(e + 3 * i + 2));
  }
  *occ_global = occ;
  grid_barrier(&ctr->bar, bar_base + barrier_no * gridDim.x, &ctr->err);
}

// count pass, sharded: exchange the local pair counts, then block 0 folds the global aggregate and publishes
__global__ void __launch_bounds__(256) k_dist_count_finalize(DeltaTable dt, PairTable pt, DevCounters* ctr, Rec* recs, uint32_t rec_cap, Ctrl* ctrl, Params P,
                                                             uint64_t flag_value, DistArgs D, uint32_t bar_base) {
  ull occ;
  exchange_deltas(dt, ctr, D, bar_base, 1, 0ull, &occ);  // entered at kernel start: k_count has completed
  if (blockIdx.x == 0) finalize_block<true>(dt, pt, ctr, recs, rec_cap, ctrl, 0, 0, P, flag_value);
}

// token frequencies, sharded: sum of the ranks' partial arrays (T x uint64), same inbox protocol
__global__ void __launch_bounds__(256) k_dist_sum_u64(ull* vals, uint64_t T, DevCounters* ctr, DistArgs D, uint32_t bar_base) {
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  for (uint64_t i = gtid; i < T; i += gthreads) {
    const ull v = vals[i];
    for (int dst = 0; dst < D.world; dst++) if (dst != D.rank)
      reinterpret_cast<ull*>(inbox_region(D.peer[dst], D.world, D.xseq, D.rank) + INBOX_HDR)[i] = v;
  }
  __threadfence_system();
  grid_barrier(&ctr->bar, bar_base + 1 * gridDim.x, &ctr->err);
  if (gtid == 0) {
    for (int dst = 0; dst < D.world; dst++) if (dst != D.rank)
      *reinterpret_cast<volatile ull*>(&reinterpret_cast<InboxHdr*>(inbox_region(D.peer[dst], D.world, D.xseq, D.rank))->seq) = D.xseq;
    const long long t0 = clock64();
    for (int src = 0; src < D.world; src++) if (src != D.rank) {
      volatile ull* f = &reinterpret_cast<InboxHdr*>(inbox_region(D.peer[D.rank], D.world, D.xseq, src))->seq;
      while (*f != D.xseq) if (clock64() - t0 > 8000000000ll) { atomicOr(&ctr->err, ERR_PEER_TIMEOUT); break; }
    }
    __threadfence_system();
  }
  grid_barrier(&ctr->bar, bar_base + 2 * gridDim.x, &ctr->err);
  for (uint64_t i = gtid; i < T; i += gthreads) {
    ull v = vals[i];
    for (int src = 0; src < D.world; src++) if (src != D.rank)
      v += __ldcv(reinterpret_cast<const ull*>(inbox_region(D.peer[D.rank], D.world, D.xseq, src) + INBOX_HDR) + i);
    vals[i] = v;
  }
}

// The per-merge kernel (cooperative launch, persistent grid = SM count x resident CTAs).
//   phase 1  HBM-bound scan of the candidate tiles: every thread streams int4 (4 symbols) and tests the 4 adjacent pairs
//            that start in it; an occurrence emits its count deltas straight into the delta table and is remembered
//   barrier
//   phase 2  every thread folds a share of the touched keys into the pair table and writes the records (bpe.cpp:297-318);
//            the last CTA to finish publishes the counters and the flag the host spins on
//   phase 3  in-place left-packed rewrite of the touched words (bpe.cpp:291-296), off the host's critical path: the
//            first occurrence to claim a word (claimed[wi] = merge number) rewrites it; the last CTA re-arms the counters
template <int UNROLL, bool DIST>
__global__ void __launch_bounds__(256, 6) k_merge(int4* ids4, uint32_t n4, uint32_t n_tiles, uint32_t tiles_per_cta,
                                               const uint32_t* __restrict__ planeA, const uint32_t* __restrict__ planeB, uint32_t* planes, uint32_t W, uint32_t id_cap,
                                               const uint32_t* __restrict__ wid, const ull* __restrict__ wcnt, const ull* __restrict__ woff, uint32_t* wlen,
                                               uint32_t* claimed, uint32_t merge_no, int32_t A, int32_t B, int32_t N, Params P, DeltaTable dt, PairTable pt,
                                               DevCounters* ctr, uint32_t* __restrict__ ml, Rec* recs, uint32_t rec_cap, Ctrl* ctrl, uint64_t flag_value,
                                               uint32_t bar_base, ull* dbg, DistArgs D) {
  __shared__ uint32_t cand[MAX_TILES_PER_CTA];
  __shared__ uint32_t n_cand;
  __shared__ bool last;
  int32_t* ids = reinterpret_cast<int32_t*>(ids4);
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  constexpr uint32_t CHUNK = 32u * UNROLL;
  if (dbg && blockIdx.x == 0 && threadIdx.x == 0) dbg[0] = gtime();
  // ---- phase 1: this CTA's contiguous tile range -> candidate tiles (both tokens present); no planes = every tile
  const uint32_t t0 = blockIdx.x * tiles_per_cta, t1 = min(t0 + tiles_per_cta, n_tiles);
  uint32_t my_occ = 0, nc_total = 0;
  for (uint32_t ts = t0; ts < t1; ts += MAX_TILES_PER_CTA) {  // (one round unless the array exceeds ~900 M slots)
  const uint32_t te_round = min(ts + MAX_TILES_PER_CTA, t1);
  __syncthreads();
  if (threadIdx.x == 0) n_cand = 0;
  __syncthreads();
  for (uint32_t t = ts + threadIdx.x; t < te_round; t += blockDim.x) {
    const bool c = planeA == nullptr || (((planeA[t >> 5] & planeB[t >> 5]) >> (t & 31)) & 1u);
    if (c) cand[atomicAdd(&n_cand, 1u)] = t;
  }
  __syncthreads();
  const uint32_t nc = n_cand;
  nc_total += nc;
  for (uint32_t ci = warp; ci < nc; ci += warps) {
    const uint64_t tb = static_cast<uint64_t>(cand[ci]) * TILE_I4;
    const uint64_t te = min(tb + TILE_I4, static_cast<uint64_t>(n4));
    for (uint64_t base = tb; base < te; base += CHUNK) {
      int4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; u++) {
        const uint64_t i = base + u * 32u + lane;
        v[u] = i < n4 ? __ldcv(ids4 + i) : make_int4(DEAD, DEAD, DEAD, DEAD);
      }
      int32_t after = DEAD;  // first symbol after this chunk (needed by lane 31 of the last row)
      if (lane == 31) { const uint64_t i = base + CHUNK; if (i < n4) after = __ldcv(ids + 4 * i); }
#pragma unroll
      for (int u = 0; u < UNROLL; u++) {
        int32_t nxt = __shfl_down_sync(0xFFFFFFFFu, v[u].x, 1);
        const int32_t row_next = (u + 1 < UNROLL) ? __shfl_sync(0xFFFFFFFFu, v[(u + 1 < UNROLL) ? u + 1 : u].x, 0) : after;
        if (lane == 31) nxt = row_next;
        uint32_t m = 0;
        m |= (v[u].x == A && v[u].y == B) ? 1u : 0u;
        m |= (v[u].y == A && v[u].z == B) ? 2u : 0u;
        m |= (v[u].z == A && v[u].w == B) ? 4u : 0u;
        m |= (v[u].w == A && nxt == B) ? 8u : 0u;
        if (__any_sync(0xFFFFFFFFu, m != 0)) {
          const uint64_t p0 = (base + u * 32u + lane) * 4u;
          while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            emit_occurrence(ids, p0 + k, wid, wcnt, A, B, N, P, dt, ctr, ml, my_occ, DIST ? (static_cast<uint64_t>(D.rank) << kSeqRankShift) : 0ull);
          }
        }
      }
    }
  }
  }
  for (int o = 16; o; o >>= 1) my_occ += __shfl_down_sync(0xFFFFFFFFu, my_occ, o);
  if (lane == 0 && my_occ) atomicAdd(&ctr->occ, static_cast<ull>(my_occ));
  if (threadIdx.x == 0 && nc_total) atomicAdd(&ctr->cand_tiles, nc_total);
  grid_barrier(&ctr->bar, bar_base + gridDim.x, &ctr->err);
  if (dbg && blockIdx.x == 0 && threadIdx.x == 0) dbg[1] = gtime();
  const ull occ_local = ctr->occ;
  ull occ_global = occ_local;
  if (DIST) exchange_deltas(dt, ctr, D, bar_base, 2, occ_local, &occ_global);

  // ---- phase 2: fold the aggregated deltas into the pair table, one key per thread
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  const uint32_t n_keys = ctr->dt_n < dt.cap ? ctr->dt_n : dt.cap;
  if (gtid == gthreads - 1) {  // bpe.cpp:315: the merged pair's frequency becomes 0
    const uint64_t k = fc_key(A, B);
    uint64_t old;
    const uint64_t sl = pt_find_or_insert(pt, ctr, k, ld_ent(&pt.ent[mix64(k) & pt.mask]), &old);
    pt.ent[sl].freq = 0ull;
  }
  bool wrote = false;
  for (uint32_t i = gtid; i < n_keys; i += gthreads) {
    const uint64_t key = dt.klist[i];
    const uint32_t ds = dt.list[i];
    const ulonglong2 home = ld_ent(&pt.ent[mix64(key) & pt.mask]);
    const int64_t d = static_cast<int64_t>(dt.delta[ds]);
    const uint64_t sq = dt.seq[ds];
    dt.keys[ds] = dt.empty; dt.delta[ds] = 0ull; dt.seq[ds] = SEQ_MAX;  // re-arm the scratch slot
    const int32_t pa = static_cast<int32_t>(key >> 32), pb = static_cast<int32_t>(key & 0xFFFFFFFFu);  // bpe.cpp:301
    if (pa == A && pb == B) continue;  // bpe.cpp:302
    Rec out; out.key = key; out.seq = sq; out.serial = REC_NO_SERIAL; out.kind = REC_PUSH; out.val = 0;
    bool emit = false;
    if (pa == P.unk_id || pb == P.unk_id) {  // phantom pair: tracked by the host (Appendix A12)
      out.kind = REC_PHANTOM; out.val = static_cast<uint64_t>(d); emit = true;
    } else {
      uint64_t old;
      const uint64_t sl = pt_find_or_insert(pt, ctr, key, home, &old);
      uint64_t nf;
      if (d < 0) { const uint64_t ad = static_cast<uint64_t>(-d); nf = old >= ad ? old - ad : 0; } else nf = old + static_cast<uint64_t>(d);  // bpe.cpp:303-307
      pt.ent[sl].freq = nf;
      if (nf >= P.min_freq) { out.kind = REC_PUSH; out.val = nf; emit = true; }            // bpe.cpp:308-311
      else if (old >= P.min_freq) { out.kind = REC_DEMOTE; out.val = nf; emit = true; }
      if (emit) out.serial = pt.serial[sl];
    }
    if (emit) {
      const uint32_t idx = atomicAdd(&ctr->rec_n, 1u);
      if (idx < rec_cap) recs[idx] = out; else atomicOr(&ctr->err, ERR_REC_FULL);
      wrote = true;
    }
  }
  if (wrote) __threadfence_system();  // my records are visible to the host before I count myself done
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&ctr->blocks_done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x == 0) {  // publish
    __threadfence();
    const uint32_t nr = *reinterpret_cast<volatile uint32_t*>(&ctr->rec_n);
    ctrl->n_recs = nr < rec_cap ? nr : rec_cap;
    ctrl->occ = occ_global;
    ctrl->occ_local = occ_local;
    ctrl->pt_n = *reinterpret_cast<volatile ull*>(&ctr->pt_n);
    ctrl->n_leaders = *reinterpret_cast<volatile uint32_t*>(&ctr->wl_n);
    ctrl->n_keys = n_keys;
    ctrl->cand_tiles = *reinterpret_cast<volatile uint32_t*>(&ctr->cand_tiles);
    ctrl->err = *reinterpret_cast<volatile uint32_t*>(&ctr->err);
    __threadfence_system();
    ctrl->flag = flag_value;
    if (dbg) dbg[2] = gtime();
  }

  // ---- phase 3: rewrite the touched words in place (left-packed); the host is already replaying its heap
  const uint32_t n_match = ctr->wl_n;
  for (uint32_t i = gtid; i < n_match; i += gthreads) {
    const uint32_t wi = wid[ml[i]];
    if (atomicMax(&claimed[wi], merge_no) >= merge_no) continue;
    const uint64_t q = woff[wi] + 1;
    asm volatile("prefetch.global.L1 [%0];" ::"l"(ids + q));        // the walk below is a chain of dependent loads:
    asm volatile("prefetch.global.L1 [%0];" ::"l"(ids + q + 32));   // pull the word's lines into L1 first
    uint64_t r = q, w = q;
    int32_t cur = ids[r];
    while (cur >= 0) {
      const int32_t nxt = ids[r + 1];
      if (cur == A && nxt == B) {
        const int32_t nn = ids[r + 2];
        ids[w] = N;
        if (planes) plane_set(planes, W, id_cap, N, w);
        ++w; r += 2;
        cur = nn;
      } else {
        if (w != r) { ids[w] = cur; if (planes) plane_set(planes, W, id_cap, cur, w); }  // a moved symbol may enter another tile
        ++w; ++r;
        cur = nxt;
      }
    }
    for (uint64_t k = w; k < r; k++) ids[k] = DEAD;
    wlen[wi] = static_cast<uint32_t>(w - q);
  }
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&ctr->blocks_done2, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x == 0) {  // every CTA has read the counters: re-arm them for the next merge
    ctr->wl_n = 0; ctr->dt_n = 0; ctr->rec_n = 0; ctr->blocks_done = 0; ctr->blocks_done2 = 0; ctr->occ = 0ull; ctr->cand_tiles = 0;
    if (dbg) dbg[3] = gtime();
  }
}

__global__ void k_rehash(PairTable oldt, PairTable newt, DevCounters* ctr) {
  for (uint64_t s = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; s < oldt.cap; s += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const ulonglong2 e = ld_ent(&oldt.ent[s]);
    if (e.x == PT_EMPTY) continue;
    uint64_t slot = mix64(e.x) & newt.mask;
    for (;;) {
      uint64_t prev = atomicCAS(reinterpret_cast<ull*>(&newt.ent[slot].key), static_cast<ull>(PT_EMPTY), static_cast<ull>(e.x));
      if (prev == PT_EMPTY) { newt.ent[slot].freq = e.y; newt.serial[slot] = oldt.serial[s]; break; }
      slot = (slot + 1) & newt.mask;
    }
  }
}

// ------------------------------------------------------------------------------------------------ compaction / save

__global__ void k_rebase(const ull* in, uint32_t n, ull base, ull* out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = in[i] - base;
}
__global__ void k_len1(const uint32_t* wlen, uint32_t n, ull* len1) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) len1[i] = static_cast<ull>(wlen[i]) + 1ull;
}
__global__ void k_compact(const int32_t* __restrict__ src, const ull* __restrict__ old_off, const ull* __restrict__ new_off, const uint32_t* __restrict__ wlen,
                          uint32_t n, int32_t* dst, uint32_t* dst_wid) {
  for (uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x; wi < n; wi += gridDim.x * blockDim.x) {
    const ull so = old_off[wi], d = new_off[wi];
    const uint32_t len = wlen[wi];
    for (uint32_t j = 0; j <= len; j++) { dst[d + j] = src[so + j]; dst_wid[d + j] = wi; }
  }
}
__global__ void k_token_freq(const int32_t* __restrict__ ids, const ull* __restrict__ woff, const uint32_t* __restrict__ wlen, const ull* __restrict__ wcnt,
                             uint32_t n, Params P, ull* freq, uint64_t T) {
  for (uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x; wi < n; wi += gridDim.x * blockDim.x) {
    const ull base = woff[wi] + 1, c = wcnt[wi];
    const uint32_t len = wlen[wi];
    for (uint32_t j = 0; j < len; j++) {
      const int32_t id = code_to_id(ids[base + j], P);
      if (id >= 0 && static_cast<uint64_t>(id) < T) atomicAdd(&freq[id], c);  // bpe.cpp:413
    }
  }
}

// ============================================================================================================ engine

inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline uint64_t next_pow2(uint64_t x) { uint64_t p = 1; while (p < x) p <<= 1; return p; }

class CudaEngine : public Engine {
 public:
  CudaEngine(int dev, const cudaDeviceProp& prop) : dev_(dev), n_sm_(prop.multiProcessorCount > 0 ? prop.multiProcessorCount : N_SM_FALLBACK) {
    std::snprintf(name_, sizeof name_, "%s sm_%d%d %d SMs", prop.name, prop.major, prop.minor, n_sm_);
  }
  ~CudaEngine() override { release_all(); }

  int init() {
    CK(cudaSetDevice(dev_));
    CK(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
    void* cp = nullptr;
    CK(cudaHostAlloc(&cp, sizeof(Ctrl), cudaHostAllocMapped));
    std::memset(cp, 0, sizeof(Ctrl));
    ctrl_ = static_cast<volatile Ctrl*>(cp);
    CK(cudaMallocAsync(reinterpret_cast<void**>(&ctr_), sizeof(DevCounters), st_));
    CK(cudaMemset(ctr_, 0, sizeof(DevCounters)));
    CK(cudaEventCreate(&ev0_));
    CK(cudaEventCreate(&ev1_));
    CK(cudaEventCreate(&evm0_));
    CK(cudaEventCreate(&evm1_));
    cudaMemPool_t pool;  // stream-ordered allocations; keep freed blocks cached so repeated loads do not pay cudaMalloc
    if (cudaDeviceGetDefaultMemPool(&pool, dev_) == cudaSuccess) { uint64_t thr = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr); }
    {  // in-kernel phase timestamps (%globaltimer) of the timed launches
      void* dp = nullptr;
      CK(cudaHostAlloc(&dp, 8 * sizeof(ull), cudaHostAllocMapped));
      dbg_ = static_cast<ull*>(dp);
      std::memset(dp, 0, 8 * sizeof(ull));
      const char* d = std::getenv("SHRED_DEBUG_TIMING");
      dbg_print_ = d && *d && *d != '0';
    }
    CK(cudaFuncSetAttribute(k_sort_buckets, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(SORT_CAP * sizeof(ull))));
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_merge<4, false>, 256, 0) == cudaSuccess && nb > 0) scan_ctas_per_sm_ = nb;
    if (const char* w = std::getenv("SHRED_WORLD")) world_ = std::atoi(w);
    if (world_ > 1) {
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_merge<4, true>, 256, 0) == cudaSuccess && nb > 0 && nb < scan_ctas_per_sm_) scan_ctas_per_sm_ = nb;
      const char* r = std::getenv("SHRED_RANK");
      rank_ = r ? std::atoi(r) : 0;
      if (world_ > MAX_RANKS || rank_ < 0 || rank_ >= world_) { std::fprintf(stderr, "[ERROR]\t bad SHRED_RANK/SHRED_WORLD (%d/%d, at most %d ranks)\n", rank_, world_, MAX_RANKS); return -1; }
      RC(dist_setup());
    } else { world_ = 1; rank_ = 0; }
    if (const char* pl = std::getenv("SHRED_PLAIN_LAUNCH")) plain_launch_ = *pl && *pl != '0';
    const char* e = std::getenv("SHRED_TIMING");
    timing_every_ = e && *e ? std::atoi(e) : 0;
    return 0;
  }

  // ------------------------------------------------------------------------------------------------- multi-GPU setup
  // One process per GPU.  Every rank allocates its inbox with cudaMalloc, exports it with CUDA IPC through a small file
  // in the rendezvous directory SHRED_RDV (shared by the ranks of one job) and maps every peer's inbox.
  int dist_setup() {
    static int instance = 0;  // ranks create their trainers in the same order, so instance numbers agree
    const int inst = instance++;
    const char* rdv = std::getenv("SHRED_RDV");
    if (!rdv || !*rdv) { std::fprintf(stderr, "[ERROR]\t SHRED_WORLD > 1 needs SHRED_RDV (a directory shared by the ranks)\n"); return -1; }
    const size_t bytes = 2ull * world_ * INBOX_BYTES;
    CK(cudaMalloc(reinterpret_cast<void**>(&inbox_), bytes));
    CK(cudaMemset(inbox_, 0, bytes));
    CK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t mine;
    CK(cudaIpcGetMemHandle(&mine, inbox_));
    auto path = [&](int r, const char* ext) { return std::string(rdv) + "/inst" + std::to_string(inst) + "_rank" + std::to_string(r) + ext; };
    {
      const std::string tmp = path(rank_, ".tmp"), fin = path(rank_, ".ipc");
      FILE* f = std::fopen(tmp.c_str(), "wb");
      if (!f) { std::fprintf(stderr, "[ERROR]\t cannot write %s\n", tmp.c_str()); return -1; }
      std::fwrite(&mine, sizeof mine, 1, f);
      std::fclose(f);
      if (std::rename(tmp.c_str(), fin.c_str()) != 0) return -1;
    }
    for (int r = 0; r < MAX_RANKS; r++) dist_.peer[r] = nullptr;
    dist_.rank = rank_; dist_.world = world_; dist_.xseq = 0;
    dist_.peer[rank_] = inbox_;
    const double t0 = now_ms();
    for (int r = 0; r < world_; r++) if (r != rank_) {
      cudaIpcMemHandle_t h;
      for (;;) {
        FILE* f = std::fopen(path(r, ".ipc").c_str(), "rb");
        if (f) { size_t got = std::fread(&h, sizeof h, 1, f); std::fclose(f); if (got == 1) break; }
        if (now_ms() - t0 > 120000.0) { std::fprintf(stderr, "[ERROR]\t rank %d: timed out waiting for rank %d in %s\n", rank_, r, rdv); return -1; }
        usleep(1000);
      }
      void* p = nullptr;
      CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      dist_.peer[r] = static_cast<uint8_t*>(p);
    }
    // nobody may unlink its file before every rank has opened every handle
    { FILE* f = std::fopen(path(rank_, ".ok").c_str(), "wb"); if (f) std::fclose(f); }
    for (int r = 0; r < world_; r++) {
      struct stat st;
      while (stat(path(r, ".ok").c_str(), &st) != 0) {
        if (now_ms() - t0 > 120000.0) { std::fprintf(stderr, "[ERROR]\t rank %d: rank %d never finished its setup\n", rank_, r); return -1; }
        usleep(1000);
      }
    }
    rdv_prefix_ = std::string(rdv) + "/inst" + std::to_string(inst) + "_rank";
    return 0;
  }
  void dist_teardown() {
    if (world_ <= 1 || !inbox_) return;
    for (int r = 0; r < world_; r++) if (r != rank_ && dist_.peer[r]) cudaIpcCloseMemHandle(dist_.peer[r]);
    // exported memory must outlive every peer's mapping: free only after all ranks have closed their handles
    if (!rdv_prefix_.empty()) {
      { FILE* f = std::fopen((rdv_prefix_ + std::to_string(rank_) + ".closed").c_str(), "wb"); if (f) std::fclose(f); }
      const double t0 = now_ms();
      for (int r = 0; r < world_; r++) {
        struct stat st;
        while (stat((rdv_prefix_ + std::to_string(r) + ".closed").c_str(), &st) != 0 && now_ms() - t0 < 20000.0) usleep(500);
      }
    }
    cudaFree(inbox_);
    inbox_ = nullptr;
    // the small rendezvous files stay: the job that created the directory removes it
  }
  DistArgs next_exchange() { DistArgs d = dist_; d.xseq = ++dist_.xseq; return d; }

  // ---------------------------------------------------------------------------------------------------------- load
  int load(const uint8_t* text, size_t n, const EngineConfig& cfg, LoadInfo* info) override { return load_impl(text, -1, n, cfg, info); }

  // bpe_load_corpus(path): the file goes to HBM through a ring of pinned staging buffers, one reader thread per buffer
  // (pread from the page cache) while the previous chunks are already in flight over PCIe.
  int load_file(int fd, size_t n, const EngineConfig& cfg, LoadInfo* info) override { return load_impl(nullptr, fd, n, cfg, info); }

  int load_impl(const uint8_t* text, int fd, size_t n, const EngineConfig& cfg, LoadInfo* info) {
    CK(cudaSetDevice(dev_));
    cfg_ = cfg;
    vocab_hint_ = cfg.vocab_size < (1ull << 22) ? cfg.vocab_size : (1ull << 22);
    P_.unk_id = cfg.unk_id;
    P_.unk_code = cfg.unk_id >= 0 ? cfg.unk_id : UNK_CODE_NEG;
    P_.min_freq = cfg.min_freq;
    release_corpus();
    std::memset(info, 0, sizeof *info);
    std::memset(&es_, 0, sizeof es_);
    // --- corpus bytes to HBM, padded with spaces so token walks and 16-byte loads stay in bounds
    const uint64_t padded = ((n + 15) & ~15ull) + 64;
    uint8_t* d_text = nullptr;
    CK(cudaMallocAsync(reinterpret_cast<void**>(&d_text), padded, st_));
    double t0 = now_ms();
    if (n && text) CK(cudaMemcpyAsync(d_text, text, n, cudaMemcpyHostToDevice, st_));
    if (n && !text && stream_file(fd, n, d_text) != 0) { cudaFreeAsync(d_text, st_); return -1; }
    CK(cudaMemsetAsync(d_text + n, ' ', padded - n, st_));
    CK(cudaStreamSynchronize(st_));
    es_.h2d_ms += now_ms() - t0; es_.h2d_bytes += n;

    CK(cudaEventRecord(ev0_, st_));
    int rc = ingest(d_text, n, info);
    cudaFreeAsync(d_text, st_);
    if (rc != 0) return rc;
    CK(cudaEventRecord(ev1_, st_));
    CK(cudaStreamSynchronize(st_));
    float ms = 0; cudaEventElapsedTime(&ms, ev0_, ev1_);
    es_.ingest_device_ms = ms;
    es_.ingest_bytes = static_cast<double>(n) + 4.0 * info->n_symbols + 12.0 * info->n_words;
    return 0;
  }

  // pinned staging ring shared by all trainers of the process (cudaHostAlloc is slow, so it is done once)
  static constexpr int STAGE_BUFS = 8;
  static constexpr size_t STAGE_BYTES = 16u << 20;
  struct StageRing { uint8_t* buf[STAGE_BUFS] = {}; std::mutex mu; };
  static StageRing& stage_ring() { static StageRing r; return r; }

  int stream_file(int fd, size_t n, uint8_t* d_text) {
    StageRing& ring = stage_ring();
    std::lock_guard<std::mutex> hold(ring.mu);  // one load at a time uses the ring
    for (int b = 0; b < STAGE_BUFS; b++) if (!ring.buf[b]) CK(cudaHostAlloc(reinterpret_cast<void**>(&ring.buf[b]), STAGE_BYTES, cudaHostAllocDefault));
    const size_t n_chunks = (n + STAGE_BYTES - 1) / STAGE_BYTES;
    cudaEvent_t done[STAGE_BUFS];
    for (int b = 0; b < STAGE_BUFS; b++) CK(cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming));
    // state[b]: 0 = reader owns the buffer, 1 = filled (main may copy), 2 = copy issued (reader waits for the event)
    std::atomic<int> state[STAGE_BUFS];
    std::atomic<bool> failed{false};
    for (int b = 0; b < STAGE_BUFS; b++) state[b].store(0);
    std::vector<std::thread> readers;
    for (int b = 0; b < STAGE_BUFS; b++) {
      readers.emplace_back([&, b]() {
        cudaSetDevice(dev_);
        for (size_t c = b; c < n_chunks && !failed.load(); c += STAGE_BUFS) {
          const size_t off = c * STAGE_BYTES, len = std::min(STAGE_BYTES, n - off);
          size_t got = 0;
          while (got < len) {
            const ssize_t r = pread(fd, ring.buf[b] + got, len - got, static_cast<off_t>(off + got));
            if (r <= 0) { failed.store(true); break; }
            got += static_cast<size_t>(r);
          }
          state[b].store(1, std::memory_order_release);
          while (state[b].load(std::memory_order_acquire) != 2 && !failed.load()) std::this_thread::yield();
          if (failed.load()) break;
          if (cudaEventSynchronize(done[b]) != cudaSuccess) { failed.store(true); break; }  // the buffer may be overwritten again
          state[b].store(0, std::memory_order_release);
        }
      });
    }
    for (size_t c = 0; c < n_chunks && !failed.load(); c++) {  // issue the copies in file order
      const int b = static_cast<int>(c % STAGE_BUFS);
      const size_t off = c * STAGE_BYTES, len = std::min(STAGE_BYTES, n - off);
      while (state[b].load(std::memory_order_acquire) != 1 && !failed.load()) std::this_thread::yield();
      if (failed.load()) break;
      if (cudaMemcpyAsync(d_text + off, ring.buf[b], len, cudaMemcpyHostToDevice, st_) != cudaSuccess || cudaEventRecord(done[b], st_) != cudaSuccess) { failed.store(true); break; }
      state[b].store(2, std::memory_order_release);
    }
    for (auto& t : readers) t.join();
    for (int b = 0; b < STAGE_BUFS; b++) cudaEventDestroy(done[b]);
    if (failed.load()) { std::fprintf(stderr, "[ERROR]\t reading the corpus file failed\n"); return -1; }
    return 0;
  }

  int ingest(const uint8_t* d_text, uint64_t n, LoadInfo* info) {
    DevCounters zero; std::memset(&zero, 0, sizeof zero);
    WordTable wt; std::memset(&wt, 0, sizeof wt);
    uint32_t N = 0; ull n_tokens = 0;
    uint64_t cap = next_pow2(n / 64 + 1); if (cap < (1u << 16)) cap = 1u << 16;  // grown 4x and redone if more than half fills up
    uint32_t seed = 0x5bd1e995u;
    for (int attempt = 0;; ++attempt) {
      if (attempt > 8) { std::fprintf(stderr, "[ERROR]\t unique-word table did not converge\n"); return -1; }
      if (cap > (1ull << 32)) { std::fprintf(stderr, "[ERROR]\t unique-word table too large\n"); return -1; }
      CK(cudaMallocAsync(reinterpret_cast<void**>(&wt.tag), cap * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&wt.first), cap * 8, st_));
      CK(cudaMallocAsync(reinterpret_cast<void**>(&wt.count), cap * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&wt.len), cap * 4, st_));
      CK(cudaMallocAsync(reinterpret_cast<void**>(&wt.bucket), cap * 4, st_));
      wt.cap = cap; wt.mask = cap - 1;
      CK(cudaMemsetAsync(wt.tag, 0, cap * 8, st_)); CK(cudaMemsetAsync(wt.first, 0xFF, cap * 8, st_)); CK(cudaMemsetAsync(wt.count, 0, cap * 8, st_));
      CK(cudaMemcpyAsync(ctr_, &zero, sizeof zero, cudaMemcpyHostToDevice, st_));
      bar_count_ = 0;
      if (n) { k_tokenize<<<grid_for((n + 15) / 16, 256), 256, 0, st_>>>(d_text, n, wt, ctr_, seed); launches_++; es_.ingest_launches++; }
      DevCounters c;
      CK(cudaMemcpyAsync(&c, ctr_, sizeof c, cudaMemcpyDeviceToHost, st_));
      CK(cudaStreamSynchronize(st_));
      CK(cudaGetLastError());
      if (c.err & ERR_HAS_NUL) {
        cudaFreeAsync(wt.tag, st_); cudaFreeAsync(wt.first, st_); cudaFreeAsync(wt.count, st_); cudaFreeAsync(wt.len, st_); cudaFreeAsync(wt.bucket, st_);
        return 1;
      }
      const bool too_full = static_cast<uint64_t>(c.n_unique) * 2 > cap;
      if ((c.err & (ERR_WT_FULL | ERR_WT_COLLISION)) || too_full) {
        cudaFreeAsync(wt.tag, st_); cudaFreeAsync(wt.first, st_); cudaFreeAsync(wt.count, st_); cudaFreeAsync(wt.len, st_); cudaFreeAsync(wt.bucket, st_);
        if ((c.err & ERR_WT_FULL) || too_full) cap *= 4;
        if (c.err & ERR_WT_COLLISION) seed = seed * 2654435761u + 12345u;
        continue;
      }
      N = c.n_unique; n_tokens = c.n_tokens;
      break;
    }
    auto free_wt = [&]() { cudaFreeAsync(wt.tag, st_); cudaFreeAsync(wt.first, st_); cudaFreeAsync(wt.count, st_); cudaFreeAsync(wt.len, st_); cudaFreeAsync(wt.bucket, st_); };
    if (N >= 0x7FFFFFF0u) { free_wt(); std::fprintf(stderr, "[ERROR]\t too many unique words\n"); return -1; }
    n_words_ = N;
    info->n_words = N; info->n_tokens = n_tokens;
    // --- reference word order: bucket = djb2 & 4095 ascending, first occurrence ascending inside a bucket
    uint32_t *u_slot = nullptr, *u_n = nullptr, *bcnt = nullptr, *bstart = nullptr, *cursor = nullptr, *tmp_slot = nullptr, *order_slot = nullptr;
    ull *tmp_first = nullptr, *d_hist = nullptr, *len1 = nullptr, *sums = nullptr;
    uint8_t* d_keep = nullptr;
    const uint64_t Na = N ? N : 1;
    CK(cudaMallocAsync(reinterpret_cast<void**>(&u_slot), Na * 4, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&tmp_slot), Na * 4, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&order_slot), Na * 4, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&tmp_first), Na * 8, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&len1), Na * 8, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&u_n), 4, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&bcnt), 4096 * 4, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&bstart), 4097 * 4, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&cursor), 4096 * 4, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&d_hist), 256 * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&d_keep), 256, st_));
    const uint32_t nb_scan = static_cast<uint32_t>((Na + SCAN_TILE - 1) / SCAN_TILE);
    CK(cudaMallocAsync(reinterpret_cast<void**>(&sums), (static_cast<uint64_t>(nb_scan) + 1) * 8, st_));
    auto free_tmp = [&]() {
      cudaFreeAsync(u_slot, st_); cudaFreeAsync(tmp_slot, st_); cudaFreeAsync(order_slot, st_); cudaFreeAsync(tmp_first, st_); cudaFreeAsync(len1, st_); cudaFreeAsync(u_n, st_); cudaFreeAsync(bcnt, st_);
      cudaFreeAsync(bstart, st_); cudaFreeAsync(cursor, st_); cudaFreeAsync(d_hist, st_); cudaFreeAsync(d_keep, st_); cudaFreeAsync(sums, st_);
    };
    CK(cudaMemsetAsync(u_n, 0, 4, st_)); CK(cudaMemsetAsync(bcnt, 0, 4096 * 4, st_)); CK(cudaMemsetAsync(cursor, 0, 4096 * 4, st_));
    CK(cudaMemsetAsync(d_hist, 0, 256 * 8, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&wcnt_), Na * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&wlen_), Na * 4, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&woff_[0]), (Na + 1) * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&woff_[1]), (Na + 1) * 8, st_));
    ull S1 = 0;  // total slots = symbols + headers
    if (N) {
      k_collect<<<grid_for(cap, 256), 256, 0, st_>>>(wt, u_slot, u_n, bcnt);
      k_scan4096<<<1, 1024, 0, st_>>>(bcnt, bstart);
      k_scatter<<<grid_for(N, 256), 256, 0, st_>>>(wt, u_slot, N, bstart, cursor, tmp_slot, tmp_first);
      k_sort_buckets<<<4096, SORT_THREADS, SORT_CAP * sizeof(ull), st_>>>(tmp_slot, tmp_first, bstart, order_slot);
      k_rank_big<<<grid_for(N, 128), 128, 0, st_>>>(wt, tmp_slot, tmp_first, N, bstart, order_slot); launches_++;  // only buckets above SORT_CAP do work
      k_hist_words<<<grid_for(N, 256), 256, 0, st_>>>(d_text, wt, order_slot, N, d_hist, wcnt_, wlen_, len1);
      k_scan_sums<<<nb_scan, SCAN_THREADS, 0, st_>>>(len1, N, sums);
      k_scan_top<<<1, SCAN_THREADS, 0, st_>>>(sums, nb_scan, sums + nb_scan);
      k_scan_apply<<<nb_scan, SCAN_THREADS, 0, st_>>>(len1, N, sums, woff_[0]);
      launches_ += 8; es_.ingest_launches += 8;
      CK(cudaMemcpyAsync(&S1, sums + nb_scan, 8, cudaMemcpyDeviceToHost, st_));
    }
    CK(cudaMemcpyAsync(info->hist, d_hist, 256 * 8, cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    CK(cudaGetLastError());
    es_.d2h_bytes += 256 * 8 + 8 + sizeof(DevCounters);
    charset_keep(info->hist, cfg_.coverage, info->keep, &info->n_distinct, &info->n_keep);
    info->n_symbols = S1 - N;
    uint32_t lo = 0, n_local = N;
    host_counts_.clear();
    if (world_ > 1 && N) {  // keep only this rank's contiguous range of words (shard.hpp); ingest itself is replicated
      std::vector<ull> hoff(static_cast<size_t>(N) + 1);
      CK(cudaMemcpyAsync(hoff.data(), woff_[0], static_cast<uint64_t>(N) * 8, cudaMemcpyDeviceToHost, st_));
      host_counts_.resize(N);
      CK(cudaMemcpyAsync(host_counts_.data(), wcnt_, static_cast<uint64_t>(N) * 8, cudaMemcpyDeviceToHost, st_));
      CK(cudaStreamSynchronize(st_));
      hoff[N] = S1;
      lo = static_cast<uint32_t>(shard_begin(hoff.data(), N, rank_, world_));
      const uint32_t hi = static_cast<uint32_t>(shard_begin(hoff.data(), N, rank_ + 1, world_));
      n_local = hi - lo;
      const ull base_off = hoff[lo];
      S1 = hoff[hi] - base_off;
      ull *wc = nullptr, *wo0 = nullptr, *wo1 = nullptr; uint32_t* wl = nullptr;
      const uint64_t nl = n_local ? n_local : 1;
      CK(cudaMallocAsync(reinterpret_cast<void**>(&wc), nl * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&wl), nl * 4, st_));
      CK(cudaMallocAsync(reinterpret_cast<void**>(&wo0), (nl + 1) * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&wo1), (nl + 1) * 8, st_));
      if (n_local) {
        CK(cudaMemcpyAsync(wc, wcnt_ + lo, static_cast<uint64_t>(n_local) * 8, cudaMemcpyDeviceToDevice, st_));
        CK(cudaMemcpyAsync(wl, wlen_ + lo, static_cast<uint64_t>(n_local) * 4, cudaMemcpyDeviceToDevice, st_));
        k_rebase<<<grid_for(n_local, 256), 256, 0, st_>>>(woff_[0] + lo, n_local, base_off, wo0); launches_++;
      }
      cudaFreeAsync(wcnt_, st_); cudaFreeAsync(wlen_, st_); cudaFreeAsync(woff_[0], st_); cudaFreeAsync(woff_[1], st_);
      wcnt_ = wc; wlen_ = wl; woff_[0] = wo0; woff_[1] = wo1;
      es_.d2h_bytes += static_cast<uint64_t>(N) * 16;
    }
    n_words_ = n_local;
    n_slots_ = S1; n_live_ = S1;
    if (S1 + 64 >= (1ull << 32)) { free_tmp(); free_wt(); std::fprintf(stderr, "[ERROR]\t corpus needs more than 2^32 symbol slots on one GPU\n"); return -1; }
    ids_cap_ = ((S1 + 8 + 1023) / 1024) * 1024;
    CK(cudaMallocAsync(reinterpret_cast<void**>(&ids_[0]), ids_cap_ * 4, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&ids_[1]), ids_cap_ * 4, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&wl_), ids_cap_ * 4, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&wid_[0]), ids_cap_ * 4, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&wid_[1]), ids_cap_ * 4, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&claimed_), Na * 4, st_));
    CK(cudaMemsetAsync(claimed_, 0, Na * 4, st_));
    CK(cudaMemsetAsync(wid_[0], 0, ids_cap_ * 4, st_)); CK(cudaMemsetAsync(wid_[1], 0, ids_cap_ * 4, st_));
    merge_no_ = 0;
    cur_ = 0;
    CK(cudaMemcpyAsync(d_keep, info->keep, 256, cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(woff_[0] + n_local, &S1, 8, cudaMemcpyHostToDevice, st_));
    if (n_local) { k_symbolize<<<grid_for(n_local, 256), 256, 0, st_>>>(d_text, wt, order_slot + lo, n_local, woff_[0], d_keep, P_.unk_code, ids_[0], wid_[0]); launches_++; es_.ingest_launches++; }
    k_fill_i32<<<grid_for(ids_cap_ - S1, 256), 256, 0, st_>>>(ids_[0], S1, ids_cap_, DEAD); launches_++; es_.ingest_launches++;
    CK(cudaStreamSynchronize(st_));
    CK(cudaGetLastError());
    free_tmp(); free_wt();
    RC(build_planes());
    // --- pair/delta tables sized for this trainer
    RC(alloc_tables());
    loaded_ = true;
    return 0;
  }

  // (re)build the tile occurrence index for the current ids buffer
  int build_planes() {
    if (std::getenv("SHRED_NO_TILE_INDEX")) return 0;
    const uint32_t n_tiles_cap = static_cast<uint32_t>((ids_cap_ + TILE_SLOTS - 1) >> TILE_SHIFT);
    const uint32_t W = (n_tiles_cap + 31) / 32;
    const uint32_t id_cap = static_cast<uint32_t>(256 + vocab_hint_ + 64);
    const uint64_t bytes = static_cast<uint64_t>(id_cap) * W * 4;
    if (!planes_ || W != plane_words_ || id_cap != id_cap_) {
      if (planes_) { cudaFreeAsync(planes_, st_); planes_ = nullptr; }
      size_t free_b = 0, total_b = 0;
      if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || bytes > free_b / 3) { plane_words_ = 0; id_cap_ = 0; return 0; }  // too big: scan every tile
      CK(cudaMallocAsync(reinterpret_cast<void**>(&planes_), bytes, st_));
      plane_words_ = W; id_cap_ = id_cap;
    }
    CK(cudaMemsetAsync(planes_, 0, bytes, st_));
    if (n_slots_) { k_build_planes<<<grid_for(n_slots_, 256), 256, 0, st_>>>(ids_[cur_], n_slots_, planes_, plane_words_, id_cap_); launches_++; }
    return 0;
  }

  int alloc_tables() {
    if (!dt_.keys) {
      uint64_t cap = next_pow2(8ull * (256 + vocab_hint_) + 1024); if (cap < (1u << 16)) cap = 1u << 16;
      RC(alloc_dt(cap));
    }
    if (!pt_.ent) RC(alloc_pt(&pt_, 1ull << 20));
    if (!recs_) {
      rec_cap_ = dt_.cap;
      CK(cudaHostAlloc(reinterpret_cast<void**>(&recs_), static_cast<size_t>(rec_cap_) * sizeof(Rec), cudaHostAllocMapped));
    }
    return 0;
  }
  int alloc_dt(uint64_t cap) {
    if (dt_.keys) { cudaFreeAsync(dt_.keys, st_); cudaFreeAsync(dt_.delta, st_); cudaFreeAsync(dt_.seq, st_); cudaFreeAsync(dt_.list, st_); cudaFreeAsync(dt_.klist, st_); dt_.keys = nullptr; }
    CK(cudaMallocAsync(reinterpret_cast<void**>(&dt_.keys), cap * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&dt_.delta), cap * 8, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&dt_.seq), cap * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&dt_.list), cap * 4, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&dt_.klist), cap * 8, st_));
    dt_.cap = static_cast<uint32_t>(cap); dt_.mask = cap - 1;
    // a key value no pair can produce: high word >= 2^31 that is neither all-ones nor unk_id
    uint32_t hi = 0x80000000u; if (static_cast<uint32_t>(cfg_.unk_id) == hi) hi = 0x80000001u;
    dt_.empty = static_cast<uint64_t>(hi) << 32;
    k_fill_u64<<<grid_for(cap, 256), 256, 0, st_>>>(reinterpret_cast<ull*>(dt_.keys), cap, dt_.empty);
    CK(cudaMemsetAsync(dt_.delta, 0, cap * 8, st_)); CK(cudaMemsetAsync(dt_.seq, 0xFF, cap * 8, st_));
    launches_++;
    if (recs_ && rec_cap_ < dt_.cap) { cudaFreeHost(recs_); recs_ = nullptr; rec_cap_ = dt_.cap; CK(cudaHostAlloc(reinterpret_cast<void**>(&recs_), static_cast<size_t>(rec_cap_) * sizeof(Rec), cudaHostAllocMapped)); }
    return 0;
  }
  int alloc_pt(PairTable* pt, uint64_t cap) {
    CK(cudaMallocAsync(reinterpret_cast<void**>(&pt->ent), cap * sizeof(PairEnt), st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&pt->serial), cap * 4, st_));
    pt->cap = cap; pt->mask = cap - 1;
    CK(cudaMemsetAsync(pt->ent, 0xFF, cap * sizeof(PairEnt), st_));  // key = EMPTY; freq is written when the entry is claimed
    return 0;
  }
  int grow_pt(uint64_t need_entries) {
    uint64_t cap = pt_.cap; while (need_entries * 2 > cap) cap *= 2;
    if (cap == pt_.cap) return 0;
    PairTable nt; std::memset(&nt, 0, sizeof nt);
    RC(alloc_pt(&nt, cap));
    k_rehash<<<grid_for(pt_.cap, 256), 256, 0, st_>>>(pt_, nt, ctr_); launches_++;
    CK(cudaStreamSynchronize(st_));
    cudaFreeAsync(pt_.ent, st_); cudaFreeAsync(pt_.serial, st_);
    pt_ = nt;
    return 0;
  }

  // --------------------------------------------------------------------------------------------------------- count
  int count_pairs(const Rec** recs, size_t* n) override {
    CK(cudaSetDevice(dev_));
    *recs = recs_; *n = 0;
    if (!loaded_) return 0;
    for (int attempt = 0; attempt < 12; ++attempt) {
      CK(cudaMemsetAsync(pt_.ent, 0xFF, pt_.cap * sizeof(PairEnt), st_));
      CK(cudaMemsetAsync(ctr_, 0, sizeof(DevCounters), st_));
      bar_count_ = 0;
      pt_n_ = 0;
      ++flag_;
      CK(cudaEventRecord(ev0_, st_));
      if (n_words_) {
        const uint32_t n4c = static_cast<uint32_t>((n_slots_ + 3) / 4);
        k_count<<<n_sm_ * 4, 256, 0, st_>>>(reinterpret_cast<const int4*>(ids_[cur_]), reinterpret_cast<const uint4*>(wid_[cur_]), n4c, wcnt_, P_, dt_, ctr_,
                                            world_ > 1 ? seq_base(rank_) : 0ull);
        launches_++;
      }
      CK(cudaEventRecord(ev1_, st_));
      // the finalize pass needs dt_n <= cap/2 and room in the pair table: check before consuming the delta table
      DevCounters c;
      CK(cudaMemcpyAsync(&c, ctr_, sizeof c, cudaMemcpyDeviceToHost, st_));
      CK(cudaStreamSynchronize(st_));
      CK(cudaGetLastError());
      float ms = 0; cudaEventElapsedTime(&ms, ev0_, ev1_);
      if ((c.err & ERR_DT_FULL) || static_cast<uint64_t>(c.dt_n) * 2 > dt_.cap) {  // enlarge the scratch table, redo (ids untouched)
        RC(alloc_dt(static_cast<uint64_t>(dt_.cap) * 4));
        continue;
      }
      RC(grow_pt((world_ > 1 ? static_cast<uint64_t>(dt_.cap) / 2 : static_cast<uint64_t>(c.dt_n)) + 4ull * (256 + vocab_hint_) + 1024));  // replicas must size identically
      if (world_ > 1) {
        DistArgs a_D = next_exchange();
        const int grid = n_sm_ * 2;
        uint32_t a_reccap = rec_cap_, a_bar = bar_count_;
        Ctrl* a_ctrl = const_cast<Ctrl*>(ctrl_);
        uint64_t a_flag = flag_;
        bar_count_ += 1u * static_cast<uint32_t>(grid);
        void* args[] = {&dt_, &pt_, &ctr_, &recs_, &a_reccap, &a_ctrl, &P_, &a_flag, &a_D, &a_bar};
        CK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_dist_count_finalize), dim3(grid), dim3(256), args, 0, st_));
      } else {
        k_finalize_count<<<1, 256, 0, st_>>>(dt_, pt_, ctr_, recs_, rec_cap_, const_cast<Ctrl*>(ctrl_), P_, flag_);
      }
      launches_++;
      RC(wait_flag());
      if (ctrl_->err) { std::fprintf(stderr, "[ERROR]\t device count pass failed (err=%u)\n", ctrl_->err); return -1; }
      es_.count_launches++; es_.count_device_ms += ms; es_.count_bytes += 4.0 * static_cast<double>(n_slots_) + 12.0 * n_words_;  // 4S + 12N; the kernel reads 8 B per slot + counts
      pt_n_ = ctrl_->pt_n;
      *n = ctrl_->n_recs;
      es_.d2h_bytes += *n * sizeof(Rec) + sizeof(Ctrl);
      return 0;
    }
    return -1;
  }

  // --------------------------------------------------------------------------------------------------------- merge
  int merge(int32_t a, int32_t b, int32_t new_id, const Rec** recs, size_t* n, uint64_t* occurrences) override {
    *recs = recs_; *n = 0; *occurrences = 0;
    const double tm0 = now_ms();
    CK(cudaSetDevice(dev_));  // the caller's thread may have another current device
    // keep the pair table at most half full even if this merge creates every key it can (4 per distinct id)
    const uint64_t worst_new = 4ull * (static_cast<uint64_t>(new_id) + 2);
    if ((pt_n_ + worst_new) * 2 > pt_.cap) RC(grow_pt(pt_n_ + worst_new));
    if (worst_new * 2 > dt_.cap) RC(alloc_dt(next_pow2(worst_new * 2)));
    // reclaim dead slots once a quarter of the scanned array is dead
    if (n_slots_ > 4096 && (n_slots_ - n_live_) * 4 > n_slots_) RC(compact());
    const uint32_t n4 = static_cast<uint32_t>((n_slots_ + 3) / 4);
    const double tl0 = now_ms();
    ++flag_;
    const bool timed = timing_every_ > 0 && (merge_seq_++ % timing_every_) == 0;
    if (timed) CK(cudaEventRecord(ev0_, st_));
    ++merge_no_;
    ull* a_dbg = timed ? dbg_ : nullptr;
    const uint32_t n_tiles = static_cast<uint32_t>((n_slots_ + TILE_SLOTS - 1) >> TILE_SHIFT);
    int grid = detect_grid(n4);
    const uint32_t tiles_per_cta = (n_tiles + grid - 1) / grid;
    const bool indexed = planes_ && static_cast<uint32_t>(a) < id_cap_ && static_cast<uint32_t>(b) < id_cap_;
    const uint32_t* pa = indexed ? planes_ + static_cast<uint64_t>(a) * plane_words_ : nullptr;
    const uint32_t* pb = indexed ? planes_ + static_cast<uint64_t>(b) * plane_words_ : nullptr;
    {
      int4* a_ids = reinterpret_cast<int4*>(ids_[cur_]);
      uint32_t a_n4 = n4, a_nt = n_tiles, a_tpc = tiles_per_cta, a_W = plane_words_, a_idcap = id_cap_, a_mno = merge_no_, a_reccap = rec_cap_;
      const uint32_t* a_wid = wid_[cur_]; const ull* a_wcnt = wcnt_; const ull* a_woff = woff_[cur_];
      int32_t a_A = a, a_B = b, a_N = new_id;
      Ctrl* a_ctrl = const_cast<Ctrl*>(ctrl_);
      uint64_t a_flag = flag_;
      uint32_t a_bar = bar_count_;  // barrier counter before this launch; it only grows (wraps mod 2^32)
      DistArgs a_D = dist_;
      if (world_ > 1) a_D = next_exchange();
      bar_count_ += (world_ > 1 ? 2u : 1u) * static_cast<uint32_t>(grid);
      void* args[] = {&a_ids, &a_n4, &a_nt, &a_tpc, &pa, &pb, &planes_, &a_W, &a_idcap, &a_wid, &a_wcnt, &a_woff, &wlen_, &claimed_, &a_mno, &a_A, &a_B, &a_N,
                      &P_, &dt_, &pt_, &ctr_, &wl_, &recs_, &a_reccap, &a_ctrl, &a_flag, &a_bar, &a_dbg, &a_D};
      const void* kfn = world_ > 1 ? reinterpret_cast<const void*>(k_merge<4, true>) : reinterpret_cast<const void*>(k_merge<4, false>);
      if (plain_launch_) CK(cudaLaunchKernel(kfn, dim3(grid), dim3(256), args, 0, st_));  // experiment: same grid, no co-residency check by the driver
      else CK(cudaLaunchCooperativeKernel(kfn, dim3(grid), dim3(256), args, 0, st_));
    }
    if (timed) CK(cudaEventRecord(ev1_, st_));
    launches_ += 1;
    launch_ms_ += now_ms() - tl0;
    RC(wait_flag());
    if (ctrl_->err) { std::fprintf(stderr, "[ERROR]\t device merge pass failed (err=%u)\n", ctrl_->err); return -1; }
    if (timed) {
      float ms = 0;
      CK(cudaEventSynchronize(ev1_));
      cudaEventElapsedTime(&ms, ev0_, ev1_);
      const double algo = 4.0 * static_cast<double>(n4) * 4.0, touched = 4.0 * TILE_SLOTS * static_cast<double>(ctrl_->cand_tiles);
      es_.scan_launches++; es_.scan_device_ms += ms; es_.scan_bytes += algo; es_.scan_bytes_touched += touched;
      const double p1 = (dbg_[1] - dbg_[0]) * 1e-6, p2 = (dbg_[2] - dbg_[1]) * 1e-6, p3 = (dbg_[3] - dbg_[2]) * 1e-6;  // ms: scan+emit+barrier | fold+publish | rewrite
      es_.scan_phase_ms += p1;
      if (ctrl_->cand_tiles * 10 >= static_cast<uint64_t>(n_tiles) * 9) { es_.dense_launches++; es_.dense_device_ms += ms; es_.dense_bytes += algo; es_.dense_phase_ms += p1; }  // streams >= 90 % of the array
      dbg_acc_[0] += p1; dbg_acc_[1] += p2; dbg_acc_[2] += p3; dbg_acc_[3] += ms; dbg_n_++;
      if (dbg_print_ && (dbg_n_ % 500) == 0)
        std::fprintf(stderr, "[KTIME]\t %llu timed merges: scan+emit+barrier %.1f us, fold+publish %.1f us, rewrite+rearm %.1f us | kernel (events) %.1f us (averages)\n",
                     (unsigned long long)dbg_n_, 1e3 * dbg_acc_[0] / dbg_n_, 1e3 * dbg_acc_[1] / dbg_n_, 1e3 * dbg_acc_[2] / dbg_n_, 1e3 * dbg_acc_[3] / dbg_n_);
    }
    cand_tiles_total_ += ctrl_->cand_tiles; tiles_total_ += n_tiles;
    *n = ctrl_->n_recs; *occurrences = ctrl_->occ;
    pt_n_ = ctrl_->pt_n;
    n_live_ -= ctrl_->occ_local;
    es_.d2h_bytes += *n * sizeof(Rec) + sizeof(Ctrl);
    merge_ms_ += now_ms() - tm0;
    return 0;
  }

  int compact() {
    const uint32_t N = n_words_;
    if (!N) return 0;
    ull *len1 = nullptr, *sums = nullptr;
    const uint32_t nb_scan = (N + SCAN_TILE - 1) / SCAN_TILE;
    CK(cudaMallocAsync(reinterpret_cast<void**>(&len1), static_cast<uint64_t>(N) * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&sums), (static_cast<uint64_t>(nb_scan) + 1) * 8, st_));
    const int nxt = cur_ ^ 1;
    k_len1<<<grid_for(N, 256), 256, 0, st_>>>(wlen_, N, len1);
    k_scan_sums<<<nb_scan, SCAN_THREADS, 0, st_>>>(len1, N, sums);
    k_scan_top<<<1, SCAN_THREADS, 0, st_>>>(sums, nb_scan, sums + nb_scan);
    k_scan_apply<<<nb_scan, SCAN_THREADS, 0, st_>>>(len1, N, sums, woff_[nxt]);
    ull S1 = 0;
    CK(cudaMemcpyAsync(&S1, sums + nb_scan, 8, cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    CK(cudaMemcpyAsync(woff_[nxt] + N, &S1, 8, cudaMemcpyHostToDevice, st_));
    k_compact<<<grid_for(N, 256), 256, 0, st_>>>(ids_[cur_], woff_[cur_], woff_[nxt], wlen_, N, ids_[nxt], wid_[nxt]);
    const uint64_t pad_to = ((S1 + 8 + 1023) / 1024) * 1024;
    k_fill_i32<<<grid_for(pad_to - S1, 256), 256, 0, st_>>>(ids_[nxt], S1, pad_to < ids_cap_ ? pad_to : ids_cap_, DEAD);
    launches_ += 6;
    CK(cudaStreamSynchronize(st_));
    CK(cudaGetLastError());
    cudaFreeAsync(len1, st_); cudaFreeAsync(sums, st_);
    cur_ = nxt; n_slots_ = S1; n_live_ = S1;
    es_.compactions++;
    RC(build_planes());
    return 0;
  }

  // ---------------------------------------------------------------------------------------------------------- save
  int token_freqs(uint64_t* freq, size_t T) override {
    CK(cudaSetDevice(dev_));
    if (!loaded_ || (!n_words_ && world_ == 1) || !T) return 0;
    ull* d = nullptr;
    CK(cudaMallocAsync(reinterpret_cast<void**>(&d), T * 8, st_));
    CK(cudaMemsetAsync(d, 0, T * 8, st_));
    if (n_words_) { k_token_freq<<<grid_for(n_words_, 256), 256, 0, st_>>>(ids_[cur_], woff_[cur_], wlen_, wcnt_, n_words_, P_, d, T); launches_++; }
    if (world_ > 1) {
      if (T > INBOX_ENTRIES * 3) { cudaFreeAsync(d, st_); std::fprintf(stderr, "[ERROR]\t vocabulary too large for the exchange buffer\n"); return -1; }
      DistArgs a_D = next_exchange();
      const int grid = n_sm_ * 2;
      uint64_t a_T = T;
      uint32_t a_bar = bar_count_;
      bar_count_ += 2u * static_cast<uint32_t>(grid);
      void* args[] = {&d, &a_T, &ctr_, &a_D, &a_bar};
      CK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_dist_sum_u64), dim3(grid), dim3(256), args, 0, st_));
      launches_++;
    }
    CK(cudaMemcpyAsync(freq, d, T * 8, cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    cudaFreeAsync(d, st_);
    es_.d2h_bytes += T * 8;
    return 0;
  }
  int word_counts(uint64_t* out) override {
    if (world_ > 1) { std::memcpy(out, host_counts_.data(), host_counts_.size() * 8); return 0; }  // global counts, kept from ingest
    if (!n_words_) return 0;
    CK(cudaMemcpyAsync(out, wcnt_, static_cast<uint64_t>(n_words_) * 8, cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    es_.d2h_bytes += static_cast<uint64_t>(n_words_) * 8;
    return 0;
  }
  int get_words(uint64_t* counts, uint64_t* off, int32_t* ids, uint64_t ids_cap) override {
    if (!loaded_) return -1;
    CK(cudaStreamSynchronize(st_));
    const uint32_t N = n_words_;
    std::vector<ull> ho(N + 1); std::vector<uint32_t> hl(N ? N : 1); std::vector<int32_t> hi(n_slots_ ? n_slots_ : 1);
    if (N) {
      CK(cudaMemcpy(ho.data(), woff_[cur_], (static_cast<uint64_t>(N) + 1) * 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hl.data(), wlen_, static_cast<uint64_t>(N) * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hi.data(), ids_[cur_], n_slots_ * 4, cudaMemcpyDeviceToHost));
      if (counts) CK(cudaMemcpy(counts, wcnt_, static_cast<uint64_t>(N) * 8, cudaMemcpyDeviceToHost));
    }
    uint64_t at = 0;
    for (uint32_t wi = 0; wi < N; wi++) {
      if (off) off[wi] = at;
      if (static_cast<uint32_t>(hi[ho[wi]]) != (HDR_BIT | wi)) return -2;  // layout invariant
      for (uint32_t j = 0; j < hl[wi]; j++) {
        int32_t code = hi[ho[wi] + 1 + j];
        if (ids && at < ids_cap) ids[at] = (cfg_.unk_id < 0 && code == UNK_CODE_NEG) ? cfg_.unk_id : code;
        at++;
      }
    }
    if (off) off[N] = at;
    return 0;
  }
  uint64_t get_pairs(int32_t* ab, uint64_t* freq, uint64_t cap) override {
    if (!pt_.ent) return 0;
    cudaStreamSynchronize(st_);
    std::vector<PairEnt> e(pt_.cap);
    if (cudaMemcpy(e.data(), pt_.ent, pt_.cap * sizeof(PairEnt), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    uint64_t n = 0;
    for (uint64_t s = 0; s < pt_.cap; s++) if (e[s].key != PT_EMPTY) {
      if (n < cap) { ab[2 * n] = static_cast<int32_t>(e[s].key >> 32); ab[2 * n + 1] = static_cast<int32_t>(e[s].key & 0xFFFFFFFFu); freq[n] = e[s].freq; }
      n++;
    }
    return n;
  }
  int mark_begin() override { CK(cudaEventRecord(evm0_, st_)); return 0; }
  double mark_end() override {
    if (cudaEventRecord(evm1_, st_) != cudaSuccess || cudaEventSynchronize(evm1_) != cudaSuccess) return -1.0;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, evm0_, evm1_) != cudaSuccess) return -1.0;
    return ms;
  }
  void stats(EngineStats* out) override {
    *out = es_;
    out->n_slots = n_slots_; out->n_symbols_live = n_live_ >= n_words_ ? n_live_ - n_words_ : 0; out->pair_entries = pt_n_;
    out->kernel_launches = launches_; out->wait_ms = wait_ms_; out->launch_ms = launch_ms_; out->merge_ms = merge_ms_;
    out->cand_tiles = cand_tiles_total_; out->tiles_total = tiles_total_;
  }
  const char* name() override { return name_; }

 private:
  int grid_for(uint64_t n, int block) const {
    uint64_t g = (n + block - 1) / block, maxg = static_cast<uint64_t>(n_sm_) * 16;
    if (g < 1) g = 1;
    return static_cast<int>(g < maxg ? g : maxg);
  }
  int detect_grid(uint32_t n4) const {  // persistent grid: SM count x resident CTAs per SM (occupancy query), never more than the work
    uint64_t warps_needed = (static_cast<uint64_t>(n4) + 127) / 128, ctas = (warps_needed + 7) / 8;
    uint64_t maxg = static_cast<uint64_t>(n_sm_) * scan_ctas_per_sm_;
    if (ctas < 1) ctas = 1;
    return static_cast<int>(ctas < maxg ? ctas : maxg);
  }

  int wait_flag() {
    double t0 = now_ms();
    uint64_t spins = 0;
    while (__atomic_load_n(&ctrl_->flag, __ATOMIC_ACQUIRE) != flag_) {
      if ((++spins & 0x3FFF) == 0) {
        cudaError_t q = cudaStreamQuery(st_);
        if (q == cudaSuccess) { if (__atomic_load_n(&ctrl_->flag, __ATOMIC_ACQUIRE) == flag_) break; std::fprintf(stderr, "[ERROR]\t device pass finished without publishing its result\n"); return -1; }
        if (q != cudaErrorNotReady) { std::fprintf(stderr, "[ERROR]\t CUDA: %s\n", cudaGetErrorString(q)); return -1; }
        if (now_ms() - t0 > 120000.0) { std::fprintf(stderr, "[ERROR]\t device pass timed out\n"); return -1; }
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    wait_ms_ += now_ms() - t0;
    return 0;
  }

  void release_corpus() {
    for (int i = 0; i < 2; i++) { if (ids_[i]) cudaFreeAsync(ids_[i], st_); ids_[i] = nullptr; if (woff_[i]) cudaFreeAsync(woff_[i], st_); woff_[i] = nullptr; }
    if (wcnt_) cudaFreeAsync(wcnt_, st_); wcnt_ = nullptr;
    if (wlen_) cudaFreeAsync(wlen_, st_); wlen_ = nullptr;
    if (wl_) cudaFreeAsync(wl_, st_); wl_ = nullptr;
    for (int i = 0; i < 2; i++) { if (wid_[i]) cudaFreeAsync(wid_[i], st_); wid_[i] = nullptr; }
    if (claimed_) cudaFreeAsync(claimed_, st_); claimed_ = nullptr;
    if (planes_) cudaFreeAsync(planes_, st_); planes_ = nullptr; plane_words_ = 0; id_cap_ = 0;
    n_words_ = 0; n_slots_ = n_live_ = 0; loaded_ = false; pt_n_ = 0;
  }
  void release_all() {
    cudaSetDevice(dev_);
    release_corpus();
    if (dt_.keys) { cudaFreeAsync(dt_.keys, st_); cudaFreeAsync(dt_.delta, st_); cudaFreeAsync(dt_.seq, st_); cudaFreeAsync(dt_.list, st_); cudaFreeAsync(dt_.klist, st_); }
    if (pt_.ent) { cudaFreeAsync(pt_.ent, st_); cudaFreeAsync(pt_.serial, st_); }
    if (recs_) cudaFreeHost(recs_);
    if (ctrl_) cudaFreeHost(const_cast<Ctrl*>(ctrl_));
    if (ctr_) cudaFreeAsync(ctr_, st_);
    if (st_) cudaStreamSynchronize(st_);
    dist_teardown();
    if (ev0_) cudaEventDestroy(ev0_);
    if (ev1_) cudaEventDestroy(ev1_);
    if (evm0_) cudaEventDestroy(evm0_);
    if (evm1_) cudaEventDestroy(evm1_);
    if (st_) cudaStreamDestroy(st_);
  }

  int dev_, n_sm_;
  char name_[320];
  cudaStream_t st_ = nullptr;
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr, evm0_ = nullptr, evm1_ = nullptr;
  EngineConfig cfg_{};
  Params P_{};
  bool loaded_ = false;
  uint32_t n_words_ = 0;
  uint64_t n_slots_ = 0, n_live_ = 0, ids_cap_ = 0;
  int32_t* ids_[2] = {nullptr, nullptr};
  ull* woff_[2] = {nullptr, nullptr};
  int cur_ = 0;
  ull* wcnt_ = nullptr;
  uint32_t* wlen_ = nullptr;
  uint32_t* wl_ = nullptr;
  uint32_t* wid_[2] = {nullptr, nullptr};
  uint32_t* claimed_ = nullptr;
  uint32_t merge_no_ = 0, bar_count_ = 0;
  int rank_ = 0, world_ = 1;
  DistArgs dist_{};
  uint8_t* inbox_ = nullptr;
  std::vector<uint64_t> host_counts_;
  std::string rdv_prefix_;
  uint32_t* planes_ = nullptr;
  uint32_t plane_words_ = 0, id_cap_ = 0;
  uint64_t cand_tiles_total_ = 0, tiles_total_ = 0;
  DeltaTable dt_{};
  PairTable pt_{};
  uint64_t pt_n_ = 0;
  Rec* recs_ = nullptr;
  uint32_t rec_cap_ = 0;
  volatile Ctrl* ctrl_ = nullptr;
  DevCounters* ctr_ = nullptr;
  uint64_t flag_ = 0;
  uint64_t vocab_hint_ = 32768;
  EngineStats es_{};
  uint64_t launches_ = 0, merge_seq_ = 0;
  double wait_ms_ = 0, launch_ms_ = 0, merge_ms_ = 0;
  int timing_every_ = 0;
  ull* dbg_ = nullptr;
  bool dbg_print_ = false, plain_launch_ = false;
  double dbg_acc_[5] = {0, 0, 0, 0, 0};
  uint64_t dbg_n_ = 0;
  int scan_ctas_per_sm_ = 4;
};

char g_devname[320] = "no CUDA device";

}  // namespace

Engine* make_device_engine() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    std::fprintf(stderr, "[ERROR]\t CUDA: no usable device (%s)\n", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return nullptr;
  }
  int dev = 0;
  if (const char* s = std::getenv("SHRED_DEVICE")) dev = std::atoi(s);
  else if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  if (dev < 0 || dev >= n) dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return nullptr;
  if (prop.major != 10) {
    std::fprintf(stderr, "[ERROR]\t CUDA: device %d (%s, sm_%d%d) is not a Blackwell sm_100 part; this library carries sm_100a code only\n", dev, prop.name,
                 prop.major, prop.minor);
    return nullptr;
  }
  CudaEngine* eng = new CudaEngine(dev, prop);
  if (eng->init() != 0) { delete eng; return nullptr; }
  std::snprintf(g_devname, sizeof g_devname, "%s", eng->name());
  return eng;
}

}  // namespace shred

extern "C" const char* bpe_b200_device_name(void) {
  static char buf[320];
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return "no CUDA device";
  int dev = 0; cudaGetDevice(&dev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return "no CUDA device";
  std::snprintf(buf, sizeof buf, "%s sm_%d%d %d SMs", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  return buf;
}
