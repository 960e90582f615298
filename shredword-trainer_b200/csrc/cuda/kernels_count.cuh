// kernels_count.cuh -- bigram count pass (pair counts, first sightings, occurrence-list sizes) and the occurrence-list fill.
// Fragment of engine_cuda.cu: included inside `namespace shred { namespace {`, in the order listed there.
#pragma once

// -------------------------------------------------------------------------------------------------------------- count

// bpe.cpp:197-214: every adjacent pair without unk adds the word's count; first sighting = flat position.
// Flat, coalesced pass over the symbol array (int4 of ids + int4 of word indices per thread); the handful of distinct
// pairs of a fresh corpus would serialise on global atomics, so each CTA first aggregates into a shared-memory hash
// table (sum of counts, min position, number of occurrences) and flushes one delta-table update per distinct pair at the end.
// Shared-memory atomics are native for 32-bit words only (a 64-bit add compiles to a load + CAS spin loop, which is what hot
// pairs used to serialise on): the sum is kept as two 32-bit halves, the low one takes the add and reports the carry.
// On a fresh corpus every token is one byte = one slot; on a corpus that already holds merged tokens (bpe_train called
// twice) a token's right neighbour is found through its SKIP mark (layout.hpp), one extra gather.
constexpr uint32_t CNT_SLOTS = 2048, CNT_PROBES = 12, CNT_UNROLL = 2;
struct CountStage {
  ull key[CNT_SLOTS], seq[CNT_SLOTS];
  uint32_t lo[CNT_SLOTS], hi[CNT_SLOTS], n[CNT_SLOTS];
};
__device__ __forceinline__ void cnt_add(CountStage& s, const DeltaTable& dt, DevCounters* ctr, uint64_t key, uint64_t c, uint64_t seq) {
  uint32_t slot = static_cast<uint32_t>((key * 0x9E3779B97F4A7C15ull) >> 53) & (CNT_SLOTS - 1);
  for (uint32_t probe = 0; probe < CNT_PROBES; ++probe) {
    ull cur = s.key[slot];
    if (cur == ~0ull) { const ull prev = atomicCAS(&s.key[slot], ~0ull, static_cast<ull>(key)); cur = prev == ~0ull ? key : prev; }
    if (cur == key) {
      const uint32_t c_lo = static_cast<uint32_t>(c), c_hi = static_cast<uint32_t>(c >> 32);
      const uint32_t old = atomicAdd(&s.lo[slot], c_lo);
      const uint32_t up = c_hi + (old + c_lo < old ? 1u : 0u);
      if (up) atomicAdd(&s.hi[slot], up);
      atomicAdd(&s.n[slot], 1u);
      if (seq < *reinterpret_cast<volatile ull*>(&s.seq[slot])) atomicMin(&s.seq[slot], static_cast<ull>(seq));  // positions grow along the grid-stride loop: rarely taken
      return;
    }
    slot = (slot + 1) & (CNT_SLOTS - 1);
  }
  const uint32_t ds = dt_add(dt, ctr, key, static_cast<int64_t>(c), seq);  // shared table crowded: straight to the global one
  if (ds != NONE32) atomicAdd(&dt.nocc[ds], 1u);
}

// the pair that starts at slot p (token s0 at p, s1 = content of slot p + 1): right neighbour through the SKIP mark if s0 is a merged token
__device__ __forceinline__ bool pair_at(const int32_t* __restrict__ ids, uint64_t p, int32_t s0, int32_t s1, const Params& P, uint64_t* key) {
  if (s0 < 0) return false;
  const int32_t y = lay::is_skip(s1) ? __ldg(ids + p + lay::skip_len(s1)) : s1;
  if (y < 0 || s0 == P.unk_code || y == P.unk_code) return false;  // bpe.cpp:201
  *key = fc_key(s0, y);
  return true;
}

__global__ void __launch_bounds__(256, 4) k_count(const int4* __restrict__ ids4, const uint4* __restrict__ wid4, uint32_t n4, const ull* __restrict__ wcnt, Params P,
                                               DeltaTable dt, DevCounters* ctr, uint64_t seq_base) {
  extern __shared__ __align__(16) unsigned char count_smem[];  // sizeof(CountStage) = 56 KB: opt-in dynamic shared memory, 4 CTAs per SM
  CountStage& st = *reinterpret_cast<CountStage*>(count_smem);
  for (uint32_t i = threadIdx.x; i < CNT_SLOTS; i += blockDim.x) { st.key[i] = ~0ull; st.lo[i] = 0u; st.hi[i] = 0u; st.n[i] = 0u; st.seq[i] = SEQ_MAX; }
  __syncthreads();
  const int32_t* ids = reinterpret_cast<const int32_t*>(ids4);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t n4_ceil = (n4 + 31u) & ~31u;  // whole warps stay in the loop so the shuffle below is full-width
  // CNT_UNROLL independent 16-byte loads of ids and of word indices per thread are issued before any of them is consumed
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
      const uint32_t ws[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        uint64_t key;
        if (pair_at(ids, 4ull * i + k, s[k], s[k + 1], P, &key)) cnt_add(st, dt, ctr, key, wcnt[ws[k]], seq_base | (4ull * i + k));
      }
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < CNT_SLOTS; i += blockDim.x)
    if (st.key[i] != ~0ull) {
      const uint32_t ds = dt_add(dt, ctr, st.key[i], static_cast<int64_t>((static_cast<ull>(st.hi[i]) << 32) | st.lo[i]), st.seq[i]);
      if (ds != NONE32) atomicAdd(&dt.nocc[ds], st.n[i]);
    }
}

// Count pass of a FRESH corpus (every token is one byte): the pair space is at most 256 x 256, so each CTA aggregates into a
// DIRECT-INDEXED shared-memory table over the dense codes of the bytes in use (K <= 111 distinct bytes: K^2 x 16 B <= 197 KB) --
// no hashing, no probing, no CAS: per adjacent pair one LUT lookup per byte and four native 32-bit shared atomics (sum low half,
// carry, occurrences, minimum position).  ncu on k_count showed the hashed table's instructions (84 per slot), not HBM, as the
// limit.  One flush per CTA and used pair into the delta table at the end.
struct ByteLut { uint8_t code[256]; };  // dense code of a byte value, 0xFF = not a countable symbol (unused, or the value unk symbols carry)
__global__ void __launch_bounds__(256) k_count_dense(const int4* __restrict__ ids4, const uint4* __restrict__ wid4, uint32_t n4, const ull* __restrict__ wcnt, const ByteLut lut,
                                                     uint32_t K, DeltaTable dt, DevCounters* ctr, uint64_t seq_base) {
  extern __shared__ __align__(16) unsigned char dense_smem[];
  uint32_t* lo = reinterpret_cast<uint32_t*>(dense_smem);
  const uint32_t KK = K * K;
  uint32_t *hi = lo + KK, *cnt = hi + KK, *mp = cnt + KK;
  __shared__ uint8_t s_code[256];
  __shared__ uint8_t s_byte[256];  // dense code -> byte value
  s_code[threadIdx.x] = lut.code[threadIdx.x];
  if (lut.code[threadIdx.x] != 0xFF) s_byte[lut.code[threadIdx.x]] = static_cast<uint8_t>(threadIdx.x);
  for (uint32_t i = threadIdx.x; i < KK; i += blockDim.x) { lo[i] = 0u; hi[i] = 0u; cnt[i] = 0u; mp[i] = 0xFFFFFFFFu; }
  __syncthreads();
  const int32_t* ids = reinterpret_cast<const int32_t*>(ids4);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t n4_ceil = (n4 + 31u) & ~31u, stride = gridDim.x * blockDim.x;
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
      if (i >= n4_ceil) break;  // uniform per warp
      int32_t nxt = __shfl_down_sync(0xFFFFFFFFu, v[u].x, 1);
      if (lane == 31) nxt = after[u];
      const int32_t sv[5] = {v[u].x, v[u].y, v[u].z, v[u].w, nxt};
      const uint32_t ws[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
      uint32_t cd[5];
#pragma unroll
      for (int k = 0; k < 5; k++) cd[k] = (static_cast<uint32_t>(sv[k]) < 256u) ? s_code[sv[k]] : 0xFFu;  // headers, padding and out-of-range unk codes are not symbols
#pragma unroll
      for (int k = 0; k < 4; k++) if (cd[k] != 0xFFu && cd[k + 1] != 0xFFu) {  // bpe.cpp:201
        const uint32_t e = cd[k] * K + cd[k + 1];
        const ull c = wcnt[ws[k]];
        const uint32_t c_lo = static_cast<uint32_t>(c), c_hi = static_cast<uint32_t>(c >> 32);
        const uint32_t old = atomicAdd(&lo[e], c_lo);
        const uint32_t up = c_hi + (old + c_lo < old ? 1u : 0u);
        if (up) atomicAdd(&hi[e], up);
        atomicAdd(&cnt[e], 1u);
        const uint32_t pos = static_cast<uint32_t>(4ull * i + k);
        if (pos < *reinterpret_cast<volatile uint32_t*>(&mp[e])) atomicMin(&mp[e], pos);  // positions grow along the grid-stride loop: rarely taken
      }
    }
  }
  __syncthreads();
  for (uint32_t e = threadIdx.x; e < KK; e += blockDim.x)
    if (cnt[e]) {
      const uint64_t key = fc_key(static_cast<int32_t>(s_byte[e / K]), static_cast<int32_t>(s_byte[e % K]));
      const uint32_t ds = dt_add(dt, ctr, key, static_cast<int64_t>((static_cast<ull>(hi[e]) << 32) | lo[e]), seq_base | mp[e]);
      if (ds != NONE32) atomicAdd(&dt.nocc[ds], cnt[e]);
    }
}

// read-only probe of the pair table: the slot of `key`, or ~0 if it is absent
__device__ __forceinline__ uint64_t pt_lookup(const PairTable& pt, uint64_t key) {
  uint64_t slot = mix64(key) & pt.mask;
  for (uint64_t probe = 0; probe < pt.cap; ++probe) {
    const uint64_t k = pt.ent[slot].key;
    if (k == key) return slot;
    if (k == PT_EMPTY) return ~0ull;
    slot = (slot + 1) & pt.mask;
  }
  return ~0ull;
}

// Second pass of the count: every position whose pair received a list (pairs that reached min_pair_freq) is stored in it.
// Ranks come from the list's fill cursor; the lanes of a warp that hold the same pair reserve together (one atomic per
// distinct pair per warp and quarter), so that the few pairs most of a fresh corpus consists of do not serialise on one address.
__global__ void __launch_bounds__(256) k_fill_lists(const int4* __restrict__ ids4, const uint4* __restrict__ wid4, const ull* __restrict__ wcnt, uint32_t n4, Params P, PairTable pt,
                                                    PoolEnt* __restrict__ pool, DevCounters* ctr) {
  const int32_t* ids = reinterpret_cast<const int32_t*>(ids4);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t n4_ceil = (n4 + 31u) & ~31u;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4_ceil; i += stride) {
    const bool in = i < n4;
    const int4 v = in ? __ldg(ids4 + i) : make_int4(DEAD, DEAD, DEAD, DEAD);
    const uint4 w = in ? __ldg(wid4 + i) : make_uint4(0, 0, 0, 0);
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
    const int32_t after = (lane == 31 && i + 1 < n4) ? __ldg(ids + 4 * (static_cast<uint64_t>(i) + 1)) : DEAD;
    int32_t nxt = __shfl_down_sync(0xFFFFFFFFu, v.x, 1);
    if (lane == 31) nxt = after;
    const int32_t s[5] = {v.x, v.y, v.z, v.w, nxt};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint64_t key;
      uint32_t serial = NONE32;
      if (pair_at(ids, 4ull * i + k, s[k], s[k + 1], P, &key)) {
        const uint64_t slot = pt_lookup(pt, key);
        if (slot != ~0ull) { serial = pt.ent[slot].serial; if (pt.lists[serial].len == 0) serial = NONE32; }
      }
      const uint32_t peers = __match_any_sync(0xFFFFFFFFu, serial);
      if (serial != NONE32) {
        const uint32_t leader = __ffs(peers) - 1u, rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&pt.lists[serial].fill, static_cast<uint32_t>(__popc(peers)));
        base = __shfl_sync(peers, base, leader);
        const ListRef lr = pt.lists[serial];
        const ull c = wcnt[ws[k]];
        PoolEnt e; e.pos = static_cast<uint32_t>(4ull * i + k); e.cnt = c < CNT_SAT ? static_cast<uint32_t>(c) : CNT_SAT;
        if (base + rank < lr.len) pool[lr.off + base + rank] = e;
        else atomicOr(&ctr->err, ERR_BAD_LIST);
      }
    }
  }
}
