// kernels_ingest.cuh -- corpus ingest: tokenise + unique-word table, reference word order, histogram, symbolise; fills; device-wide scan.
// Fragment of engine_cuda.cu: included inside `namespace shred { namespace {`, in the order listed there.
#pragma once

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
