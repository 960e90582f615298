// kernels_count.cuh -- bigram count pass.
// Fragment of engine_cuda.cu: included inside `namespace shred { namespace {`, in the order listed there.
#pragma once

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
