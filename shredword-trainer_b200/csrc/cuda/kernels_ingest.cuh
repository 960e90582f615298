// kernels_ingest.cuh -- corpus ingest after tokenisation (kernels_tokenize.cuh): reference word order, histogram, symbolise; fills.
// Fragment of engine_cuda.cu: included inside `namespace shred { namespace {`, in the order listed there.
#pragma once

// ------------------------------------------------------------------------------------------------------------ ingest

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
