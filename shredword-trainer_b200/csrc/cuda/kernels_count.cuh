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
// Shared-memory atomics are native for 32-bit words only (a 64-bit add compiles to a load + CAS spin loop, which is what hot
// pairs used to serialise on): the sum is kept as two 32-bit halves, the low one takes the add and reports the carry.
constexpr uint32_t CNT_SLOTS = 2048, CNT_PROBES = 12, CNT_UNROLL = 2;
__device__ __forceinline__ void cnt_add(ull* s_key, uint32_t* s_lo, uint32_t* s_hi, ull* s_seq, const DeltaTable& dt, DevCounters* ctr, uint64_t key, uint64_t c,
                                        uint64_t seq) {
  uint32_t slot = static_cast<uint32_t>((key * 0x9E3779B97F4A7C15ull) >> 53) & (CNT_SLOTS - 1);
  for (uint32_t probe = 0; probe < CNT_PROBES; ++probe) {
    ull cur = s_key[slot];
    if (cur == ~0ull) { const ull prev = atomicCAS(&s_key[slot], ~0ull, static_cast<ull>(key)); cur = prev == ~0ull ? key : prev; }
    if (cur == key) {
      const uint32_t c_lo = static_cast<uint32_t>(c), c_hi = static_cast<uint32_t>(c >> 32);
      const uint32_t old = atomicAdd(&s_lo[slot], c_lo);
      const uint32_t up = c_hi + (old + c_lo < old ? 1u : 0u);
      if (up) atomicAdd(&s_hi[slot], up);
      if (seq < *reinterpret_cast<volatile ull*>(&s_seq[slot])) atomicMin(&s_seq[slot], static_cast<ull>(seq));  // positions grow along the grid-stride loop: rarely taken
      return;
    }
    slot = (slot + 1) & (CNT_SLOTS - 1);
  }
  dt_add(dt, ctr, key, static_cast<int64_t>(c), seq);  // shared table crowded: straight to the global one
}
__global__ void __launch_bounds__(256, 4) k_count(const int4* __restrict__ ids4, const uint4* __restrict__ wid4, uint32_t n4, const ull* __restrict__ wcnt, Params P,
                                               DeltaTable dt, DevCounters* ctr, uint64_t seq_base) {
  __shared__ ull s_key[CNT_SLOTS], s_seq[CNT_SLOTS];
  __shared__ uint32_t s_lo[CNT_SLOTS], s_hi[CNT_SLOTS];
  for (uint32_t i = threadIdx.x; i < CNT_SLOTS; i += blockDim.x) { s_key[i] = ~0ull; s_lo[i] = 0u; s_hi[i] = 0u; s_seq[i] = SEQ_MAX; }
  __syncthreads();
  const int32_t* ids = reinterpret_cast<const int32_t*>(ids4);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t n4_ceil = (n4 + 31u) & ~31u;  // whole warps stay in the loop so the shuffle below is full-width
  // CNT_UNROLL independent 16-byte loads of ids and of word indices per thread are issued before any of them is consumed: the
  // pass is a stream, and with the shared table capping residency at 4 CTAs per SM one load per thread left HBM idle (ncu:
  // long-scoreboard stalls, 1.8 TB/s)
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n4_ceil; i0 += stride * CNT_UNROLL) {
    int4 v[CNT_UNROLL];
    uint4 w[CNT_UNROLL];
    int32_t after[CNT_UNROLL];
#pragma unroll
    for (int u = 0; u < CNT_UNROLL; u++) {
      const uint64_t i = static_cast<uint64_t>(i0) + static_cast<uint64_t>(u) * stride;
      const bool in = i < n4;
      v[u] = in ? __ldg(ids4 + i) : make_int4(DEAD, DEAD, DEAD, DEAD);
      w[u] = in ? __ldg(wid4 + i) : make_uint4(0, 0, 0, 0);
      after[u] = (lane == 31 && i + 1 < n4) ? __ldg(ids + 4 * (i + 1)) : DEAD;
    }
#pragma unroll
    for (int u = 0; u < CNT_UNROLL; u++) {
      const uint64_t i = static_cast<uint64_t>(i0) + static_cast<uint64_t>(u) * stride;
      if (i >= n4_ceil) break;  // uniform per warp: stride is a multiple of 32
      int32_t nxt = __shfl_down_sync(0xFFFFFFFFu, v[u].x, 1);
      if (lane == 31) nxt = after[u];
      const int32_t s[5] = {v[u].x, v[u].y, v[u].z, v[u].w, nxt};
      uint32_t m = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) m |= (s[k] >= 0 && s[k + 1] >= 0 && s[k] != P.unk_code && s[k + 1] != P.unk_code) ? (1u << k) : 0u;
      if (m) {
        const uint32_t ws[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
        for (int k = 0; k < 4; k++) if (m & (1u << k))
          cnt_add(s_key, s_lo, s_hi, s_seq, dt, ctr, fc_key(s[k], s[k + 1]), wcnt[ws[k]], seq_base | (4ull * i + k));
      }
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < CNT_SLOTS; i += blockDim.x)
    if (s_key[i] != ~0ull) dt_add(dt, ctr, s_key[i], static_cast<int64_t>((static_cast<ull>(s_hi[i]) << 32) | s_lo[i]), s_seq[i]);
}
