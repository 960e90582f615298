// kernels_merge.cuh -- the per-merge kernel (occurrence-list driven), pair-table rehash and the word walkers (token frequencies).
// Fragment of engine_cuda.cu: included inside `namespace shred { namespace {`, in the order listed there.
#pragma once

// Scratch of one merge, one entry per occurrence found in phase 1 (dense index from a warp-aggregated counter).
struct OccScratch {
  uint4* a;   // {slot of the occurrence, slot where its left neighbour starts after the merge, delta-table slot of (L,N), rank in that key's list}
  uint4* b;   // {delta-table slot of (N,R), rank in that key's list, the word's count as stored in list entries, -}
  uint32_t cap;
};

struct MergeArgs {
  int32_t* ids; uint64_t ids_cap; const uint32_t* wid; const ull* wcnt;
  PoolEnt* pool; uint64_t pool_cap;
  OccScratch sc;
  int32_t A, B, N; uint32_t lenA, lenB;
  uint32_t serial;     // of the pair (A,B): pt.lists[serial] is its occurrence list
  uint32_t par;        // parity of this pass: which half of the double-buffered counters it uses
  Params P; DeltaTable dt; PairTable pt; DevCounters* ctr;
  WireRec* recs; uint32_t rec_cap; Ctrl* ctrl; uint32_t tag;
  uint32_t bar_base;   // value of the grid-barrier counter before this launch
  uint64_t seq_base;   // multi-GPU: rank << kSeqRankShift, so that sequence numbers compare globally
  ull* dbg;
  DistArgs D;
};

// One list entry: load everything the probe certainly or probably reads in one round trip (the token, its right neighbour, the
// neighbour after that, the slot on its left), then probe.  The word's count travels in the list entry.
struct EntrySyms { int32_t v_p, v_b, v_r, v_m; };
__device__ __forceinline__ EntrySyms load_entry(const MergeArgs& a, const PoolEnt e) {
  const int32_t* ids = a.ids;
  const uint32_t p = e.pos;
  const uint64_t pB = static_cast<uint64_t>(p) + a.lenA, pR = min(pB + a.lenB, a.ids_cap - 1);
  EntrySyms y;
  y.v_p = ids[p]; y.v_b = ids[pB < a.ids_cap ? pB : a.ids_cap - 1]; y.v_r = ids[pR]; y.v_m = ids[p - 1];
  return y;
}
__device__ __forceinline__ bool probe_loaded(const MergeArgs& a, const PoolEnt e, const EntrySyms& y, lay::Occ* o, ull* c) {
  if (y.v_p != a.A || y.v_b != a.B) return false;
  const int32_t* ids = a.ids;
  const uint32_t p = e.pos;
  const uint64_t pB = static_cast<uint64_t>(p) + a.lenA, pR = min(pB + a.lenB, a.ids_cap - 1);
  *c = e.cnt != CNT_SAT ? static_cast<ull>(e.cnt) : a.wcnt[a.wid[p]];
  auto ld = [&](uint64_t q) { return q == p ? y.v_p : q == pB ? y.v_b : q == pR ? y.v_r : q + 1 == p ? y.v_m : ids[q]; };
  return lay::probe_occurrence(ld, p, a.A, a.B, a.lenA, a.lenB, a.N, a.P, o);
}
__device__ __forceinline__ bool probe_entry(const MergeArgs& a, const PoolEnt e, lay::Occ* o, ull* c) { return probe_loaded(a, e, load_entry(a, e), o, c); }

// The per-merge kernel: one cooperative launch per merge, grid sized by the host from the length of the pair's occurrence list.
// The reference scans every word for the pair (bpe.cpp:265-296); here the pair's occurrence list names the only slots that can
// hold it, and every entry is re-validated against the symbol array (lists are never updated when occurrences disappear).
//   phase 1  probe the list entries in parallel against the PRE-merge symbols (layout.hpp probe_occurrence); every occurrence
//            adds its four count deltas to the delta table, draws its ranks in the lists of the two pairs it creates -- (L,N)
//            and (N,R) -- and notes itself in the scratch
//   barrier  (multi-GPU: the aggregated deltas are exchanged over NVLink peer memory here, kernels_dist.cuh)
//   phase 2  the touched keys are folded into the pair table, the new pairs' lists reserved in the pool, the host records written
//            (bpe.cpp:297-318); the last CTA to arrive at the second barrier publishes the counters (self-validating 16-byte
//            blocks: no fence towards the host anywhere, common.cuh)
//   barrier
//   phase 3  (host already replaying its heap) rewrite the occurrences in place -- 4 stores each, nothing moves -- and store
//            their slots in the new pairs' lists
// Lists of up to SMALL_MAX entries (most merges of a run) take k_merge_small below instead: one CTA, shared-memory tables.
// The merge is a chain of dependent memory round trips (list entry -> symbols -> table slot -> ...), not a stream: independent
// loads and atomics are issued together, and nothing on the path waits for the host.
constexpr uint32_t CAS_FIRST_MAX = 16384;  // launches with at most this many list entries claim delta-table slots without looking first
template <bool DIST>
__device__ __forceinline__ void merge_body(const MergeArgs& a, const VGrid vg) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t gtid = vg.b * blockDim.x + threadIdx.x, gthreads = vg.g * blockDim.x;
  DevCounters* const ctr = a.ctr;
  const uint32_t par = a.par;
  if (gtid == 0) {
    if (a.dbg) a.dbg[0] = gtime();
    ctr->n_occ[par ^ 1u] = 0u; ctr->dt_n[par ^ 1u] = 0u; ctr->rec_n[par ^ 1u] = 0u;  // for the next pass (a later launch)
  }
  // ---- phase 1
  const ListRef lr = a.pt.lists[a.serial];
  const uint32_t len_ceil = (lr.len + 31u) & ~31u;
  for (uint32_t i = gtid; i < len_ceil; i += gthreads) {
    bool ok = false;
    PoolEnt e; e.pos = 0; e.cnt = 0;
    ull c = 0;
    lay::Occ o;
    o.has_l = o.has_r = false; o.lid = o.rid = 0; o.pl = 0;
    if (i < lr.len) { e = a.pool[lr.off + i]; ok = probe_entry(a, e, &o, &c); }
    const uint32_t found = __ballot_sync(0xFFFFFFFFu, ok);
    if (!found) continue;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&ctr->n_occ[par], static_cast<uint32_t>(__popc(found)));  // consumed after the table updates: its round trip overlaps theirs
    const uint32_t p = e.pos;
    const uint64_t seq = a.seq_base | (static_cast<uint64_t>(p) * 4ull);
    const uint64_t key[4] = {fc_key(o.lid, a.A), fc_key(o.lid, a.N), fc_key(a.B, o.rid), fc_key(a.N, o.rid)};  // bpe.cpp:274-290
    uint32_t s1 = NONE32, r1 = 0, s2 = NONE32, r2 = 0;
    if (ok) {
      const int64_t cc = static_cast<int64_t>(c);
      const int64_t delta[4] = {-cc, cc, -cc, cc};
      const uint64_t sq[4] = {seq + 0, seq + 1, seq + 2, seq + 3};
      const uint32_t valid = (o.has_l ? 3u : 0u) | (o.has_r ? 12u : 0u);
      const uint32_t want = ((valid & 2u) && !lay::key_has_unk(key[1], a.P) ? 2u : 0u) | ((valid & 8u) && !lay::key_has_unk(key[3], a.P) ? 8u : 0u);
      uint32_t slot[4], rank[4];
      dt_emit4(a.dt, ctr, key, delta, sq, valid, want, lr.len <= CAS_FIRST_MAX, slot, rank);
      if (want & 2u) { s1 = slot[1]; r1 = rank[1]; }
      if (want & 8u) { s2 = slot[3]; r2 = rank[3]; }
    }
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (ok) {
      const uint32_t idx = base + __popc(found & ((1u << lane) - 1u));
      if (idx < a.sc.cap) { a.sc.a[idx] = make_uint4(p, o.pl, s1, r1); a.sc.b[idx] = make_uint4(s2, r2, c < CNT_SAT ? static_cast<uint32_t>(c) : CNT_SAT, 0u); }
      else atomicOr(&ctr->err, ERR_SCRATCH_FULL);
    }
  }
  grid_barrier(&ctr->bar, a.bar_base + vg.g, &ctr->err);
  if (a.dbg && gtid == 0) a.dbg[1] = gtime();
  const ull occ_local = *reinterpret_cast<volatile uint32_t*>(&ctr->n_occ[par]);
  ull occ_global = occ_local;
  if (DIST) exchange_deltas(a.dt, ctr, a.D, vg, a.bar_base, 2, occ_local, &occ_global);

  // ---- phase 2: fold the aggregated deltas into the pair table
  const uint32_t n_keys = min(*reinterpret_cast<volatile uint32_t*>(a.dt.n), a.dt.cap);
  if (gtid == gthreads - 1) {  // bpe.cpp:315: the merged pair's frequency becomes 0
    const uint64_t k = fc_key(a.A, a.B);
    uint64_t old;
    const uint64_t sl = pt_find_or_insert(a.pt, ctr, k, ld_ent(&a.pt.ent[mix64(k) & a.pt.mask]), &old);
    a.pt.ent[sl].freq = 0ull;
  }
  fold_keys<false>(a.dt, a.pt, ctr, n_keys, gtid >> 5, gthreads >> 5, a.A, a.B, a.P, a.pool_cap, a.recs, a.rec_cap, &ctr->rec_n[par], a.tag);
  // second barrier; its last arrival publishes (the others are already released and rewriting)
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t target2 = a.bar_base + (DIST ? 3u : 2u) * vg.g;
    __threadfence();
    const bool last = atomicAdd(&ctr->bar, 1u) + 1u == target2;
    if (last) {
      __threadfence();
      const uint32_t nr = *reinterpret_cast<volatile uint32_t*>(&ctr->rec_n[par]);
      wire_ctrl(a.ctrl, a.tag, nr < a.rec_cap ? nr : a.rec_cap, *reinterpret_cast<volatile uint32_t*>(&ctr->err), lr.len, static_cast<uint32_t>(occ_local), occ_global, n_keys,
                *reinterpret_cast<volatile ull*>(&ctr->pt_n), *reinterpret_cast<volatile ull*>(&ctr->pool_top));
      if (a.dbg) a.dbg[2] = gtime();
    } else {
      const long long t0 = clock64();
      while (static_cast<int32_t>(*reinterpret_cast<volatile uint32_t*>(&ctr->bar) - target2) < 0) {
        if (clock64() - t0 > 4000000000ll) { atomicOr(&ctr->err, ERR_BARRIER); break; }
      }
    }
    __threadfence();
  }
  __syncthreads();

  // ---- phase 3: rewrite the occurrences, fill the new pairs' lists
  int32_t* idsw = a.ids;
  auto st = [idsw](uint64_t q, int32_t v) { idsw[q] = v; };
  const uint32_t n_occ = min(static_cast<uint32_t>(occ_local), a.sc.cap);
  for (uint32_t i = gtid; i < n_occ; i += gthreads) {
    const uint4 x = a.sc.a[i];
    const uint4 y = a.sc.b[i];
    const ull b1 = x.z != NONE32 ? a.dt.base[x.z] : NO_LIST, b2 = y.x != NONE32 ? a.dt.base[y.x] : NO_LIST;
    lay::rewrite_occurrence(st, x.x, a.lenA, a.lenB, a.N);
    PoolEnt e; e.cnt = y.z;
    if (b1 != NO_LIST) { e.pos = x.y; a.pool[b1 + x.w] = e; }
    if (b2 != NO_LIST) { e.pos = x.x; a.pool[b2 + y.y] = e; }
  }
  if (a.dbg && gtid == 0) a.dbg[3] = gtime();
}
template <bool DIST>
__global__ void __launch_bounds__(256) k_merge(const MergeArgs a) { merge_body<DIST>(a, VGrid{blockIdx.x, gridDim.x}); }
// virtual ranks (tests on fewer GPUs than ranks): every rank's CTA group in one cooperative launch
__global__ void __launch_bounds__(256) k_merge_virtual(const MergeArgs* argv, uint32_t g) { merge_body<true>(argv[blockIdx.x / g], VGrid{blockIdx.x % g, g}); }

// ---------------------------------------------------------------------------------------------------- one-CTA merge
// Most merges of a run have occurrence lists of a few dozen to a few hundred entries, and for them a launch is pure latency:
// ncu on such launches shows ~340 instructions per warp spread over ~10 000 cycles -- barriers waiting for the slowest warp,
// instruction fetch, and a chain of dependent round trips (profiles/r02).  k_merge_small therefore runs the three phases in ONE
// CTA sized to the list (up to four entries per thread), with the delta table and the touched-key list in SHARED memory and as few block barriers and dependent round trips as the data flow allows:
//   phase 1  list entry -> symbols (one batch of loads) -> the occurrence's four keys go into the shared table together (loads,
//            claims and adds of all four issued before the first result is used); the first thread to see a key prefetches the
//            key's pair-table sector into L2
//   phase 2  per warp, 32 keys at a time: keys that contain the new token are new by construction, so their claim in the pair
//            table (a CAS on the prefetched sector) is issued together with the loads of the existing keys and with the warp's
//            reservations of serials and pool space (one global atomic each per warp); records leave as self-validating blocks
//   phase 3  rewrite + new lists, after the control block has been published from shared counters
// Sequence numbers are 32 bits here (slot * 4 + delta slot), so the host uses it only below 2^30 slots, on one GPU, for lists
// of at most SMALL_MAX entries.  If the shared table cannot place a key the kernel publishes ERR_RETRY before it has changed
// anything in global memory, and the host runs the general kernel instead.
constexpr uint32_t SMALL_MAX = 4096, SMALL_PER_THREAD = 4, SMALL_SLOTS_MAX = 4096, SMALL_PROBES = 32;  // up to 1024 threads x 4 list entries
constexpr uint32_t ERR_RETRY = 0x40000000u;
struct SmallStage {
  ull key[SMALL_SLOTS_MAX];                                   // dt.empty = free; the launch uses the first `slots` (a power of two >= 4 x list length)
  uint32_t lo[SMALL_SLOTS_MAX], hi[SMALL_SLOTS_MAX];          // net delta as two 32-bit halves (shared adds are native for 32 bits only); after the fold: list base
  uint32_t seq[SMALL_SLOTS_MAX], nocc[SMALL_SLOTS_MAX];
  uint32_t used[SMALL_SLOTS_MAX];                             // dense list of claimed slots
  ull pt_after, pool_after;                                   // counters as this pass leaves them (0: untouched)
  uint32_t n_keys, n_occ, n_recs, overflow;
};
__host__ __device__ inline uint32_t small_slots_for(uint32_t list_len) {
  uint32_t n = 128;
  while (n < 4u * list_len && n < SMALL_SLOTS_MAX) n <<= 1;  // (longer lists: the keys of a merge are far fewer than 4 per entry; ERR_RETRY otherwise)
  return n;
}
__device__ __forceinline__ void small_apply(SmallStage& s, uint32_t slot, int64_t delta, uint32_t seq, bool want_rank, uint32_t* rank_out) {
  const uint32_t d_lo = static_cast<uint32_t>(static_cast<uint64_t>(delta)), d_hi = static_cast<uint32_t>(static_cast<uint64_t>(delta) >> 32);
  const uint32_t old = atomicAdd(&s.lo[slot], d_lo);
  if (want_rank) *rank_out = atomicAdd(&s.nocc[slot], 1u);
  atomicMin(&s.seq[slot], seq);
  const uint32_t up = d_hi + (old + d_lo < old ? 1u : 0u);  // two's complement: exact modulo 2^64
  if (up) atomicAdd(&s.hi[slot], up);
}
__device__ __forceinline__ void small_claimed(SmallStage& s, const MergeArgs& a, uint32_t slot, uint64_t key) {
  s.used[atomicAdd(&s.n_keys, 1u)] = slot;
  asm volatile("prefetch.global.L2 [%0];" ::"l"(&a.pt.ent[mix64(key) & a.pt.mask]));  // the fold reads or claims this sector
}
// probing insert for a key whose home slot is taken by another key
__device__ __forceinline__ bool small_add_probe(SmallStage& s, const MergeArgs& a, uint32_t mask, uint64_t key, int64_t delta, uint32_t seq, bool want_rank, uint32_t* slot_out,
                                                uint32_t* rank_out) {
  const ull empty = a.dt.empty;
  uint32_t slot = (static_cast<uint32_t>(mix64(key)) + 1u) & mask;
  for (uint32_t probe = 0; probe < SMALL_PROBES; ++probe) {
    ull cur = s.key[slot];
    if (cur == empty) {
      const ull prev = atomicCAS(&s.key[slot], empty, static_cast<ull>(key));
      if (prev == empty) { small_claimed(s, a, slot, key); cur = key; } else cur = prev;
    }
    if (cur == key) { small_apply(s, slot, delta, seq, want_rank, rank_out); *slot_out = slot; return true; }
    slot = (slot + 1) & mask;
  }
  return false;
}

__device__ __forceinline__ void small_clear(SmallStage& s, uint32_t slots, ull empty) {
  for (uint32_t i = threadIdx.x; i < slots; i += blockDim.x) { s.key[i] = empty; s.lo[i] = 0u; s.hi[i] = 0u; s.seq[i] = 0xFFFFFFFFu; s.nocc[i] = 0u; }
  if (threadIdx.x == 0) { s.n_keys = 0; s.n_occ = 0; s.n_recs = 0; s.overflow = 0; s.pt_after = 0ull; s.pool_after = 0ull; }
}
// the whole merge in one CTA (all threads of the block call it); returns false if it gave the merge up (ERR_RETRY published).
// PRECLEARED: the shared tables are already clean and a block barrier has passed since (the resident server cleans up after every
// merge, off the critical path) -- ncu's stall samples put 15 % of a launch into that first barrier.
template <bool PRECLEARED>
__device__ __forceinline__ bool small_merge_body(SmallStage& s, const MergeArgs& a, const uint32_t slots) {
  const uint32_t t = threadIdx.x, lane = t & 31u, nt = blockDim.x, mask = slots - 1u;
  DevCounters* const ctr = a.ctr;
  const ull empty = a.dt.empty;
  if (t == 0 && a.dbg) a.dbg[0] = gtime();
  const ListRef lr = a.pt.lists[a.serial];  // in flight while the tables are cleared
  if (t == nt - 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(&a.pt.ent[mix64(fc_key(a.A, a.B)) & a.pt.mask]));  // phase 2 zeroes the merged pair's frequency there
  if (!PRECLEARED) { small_clear(s, slots, empty); __syncthreads(); }
  // ---- phase 1: up to SMALL_PER_THREAD list entries per thread; entries first, then all their symbols, then the probes
  if (lr.len <= nt * SMALL_PER_THREAD) {
    PoolEnt e[SMALL_PER_THREAD];
    EntrySyms y[SMALL_PER_THREAD];
#pragma unroll
    for (uint32_t j = 0; j < SMALL_PER_THREAD; j++) { const uint32_t i = t + j * nt; if (i < lr.len) e[j] = a.pool[lr.off + i]; }
#pragma unroll
    for (uint32_t j = 0; j < SMALL_PER_THREAD; j++) { const uint32_t i = t + j * nt; if (i < lr.len) y[j] = load_entry(a, e[j]); }
#pragma unroll
    for (uint32_t j = 0; j < SMALL_PER_THREAD; j++) {
      const uint32_t i = t + j * nt;
      if (i >= lr.len) break;
      ull c = 0;
      lay::Occ o;
      if (!probe_loaded(a, e[j], y[j], &o, &c)) continue;
      const uint32_t p = e[j].pos, seq = p * 4u;
      const int64_t cc = static_cast<int64_t>(c);
      const uint64_t key[4] = {fc_key(o.lid, a.A), fc_key(o.lid, a.N), fc_key(a.B, o.rid), fc_key(a.N, o.rid)};  // bpe.cpp:274-290
      const bool has[4] = {o.has_l, o.has_l, o.has_r, o.has_r};
      const bool list[4] = {false, o.has_l && !lay::key_has_unk(key[1], a.P), false, o.has_r && !lay::key_has_unk(key[3], a.P)};
      uint32_t slot[4], rank[4] = {0, 0, 0, 0};
      ull cur[4];
      // the four keys together: home slots, claims, then every add before any result is used
#pragma unroll
      for (int q = 0; q < 4; q++) { slot[q] = static_cast<uint32_t>(mix64(key[q])) & mask; cur[q] = has[q] ? s.key[slot[q]] : 0ull; }
#pragma unroll
      for (int q = 0; q < 4; q++) if (has[q] && cur[q] == empty) {
        const ull prev = atomicCAS(&s.key[slot[q]], empty, static_cast<ull>(key[q]));
        if (prev == empty) { small_claimed(s, a, slot[q], key[q]); cur[q] = key[q]; } else cur[q] = prev;
      }
#pragma unroll
      for (int q = 0; q < 4; q++) if (has[q] && cur[q] == key[q]) small_apply(s, slot[q], (q & 1) ? cc : -cc, seq + static_cast<uint32_t>(q), list[q], &rank[q]);
      bool fit = true;
#pragma unroll
      for (int q = 0; q < 4; q++) if (has[q] && cur[q] != key[q]) fit &= small_add_probe(s, a, mask, key[q], (q & 1) ? cc : -cc, seq + static_cast<uint32_t>(q), list[q], &slot[q], &rank[q]);
      if (!fit) s.overflow = 1u;
      const uint32_t idx = atomicAdd(&s.n_occ, 1u);
      if (idx < a.sc.cap) {
        a.sc.a[idx] = make_uint4(p, o.pl, list[1] ? slot[1] : NONE32, rank[1]);
        a.sc.b[idx] = make_uint4(list[3] ? slot[3] : NONE32, rank[3], c < CNT_SAT ? static_cast<uint32_t>(c) : CNT_SAT, 0u);
      } else s.overflow = 1u;
    }
  }
  __syncthreads();
  if (t == 0 && a.dbg) a.dbg[1] = gtime();
  if (s.overflow || lr.len > nt * SMALL_PER_THREAD || s.n_keys > SMALL_SLOTS_MAX) {  // nothing but scratch has been written: hand the merge to the general kernel
    if (t == 0) wire_ctrl(a.ctrl, a.tag, 0u, ERR_RETRY, lr.len, 0u, 0ull, 0u, 0ull, 0ull);
    return false;
  }
  // ---- phase 2: fold, 32 keys per warp at a time
  const uint32_t n_keys = s.n_keys, n_occ = s.n_occ;
  const uint32_t err_before = *reinterpret_cast<volatile uint32_t*>(&ctr->err);  // consumed at the very end
  const uint32_t lt = (1u << lane) - 1u;
  for (uint32_t i0 = (t >> 5) * 32u; i0 < n_keys; i0 += (nt >> 5) * 32u) {
    const uint32_t i = i0 + lane;
    bool normal = false, phantom = false, is_new = false, emit = false;
    uint64_t key = 0, sl = 0, nf = 0, old = 0;
    int64_t d = 0;
    uint32_t ds = 0, no = 0, sq = 0, serial = REC_NO_SERIAL, kind = REC_PUSH;
    if (i < n_keys) {
      ds = s.used[i];
      key = s.key[ds];
      d = static_cast<int64_t>((static_cast<ull>(s.hi[ds]) << 32) | s.lo[ds]);
      sq = s.seq[ds]; no = s.nocc[ds];
      s.lo[ds] = 0xFFFFFFFFu; s.hi[ds] = 0xFFFFFFFFu;  // from here on: the key's list base (NO_LIST)
      const int32_t pa = static_cast<int32_t>(key >> 32), pb = static_cast<int32_t>(key & 0xFFFFFFFFu);  // bpe.cpp:301
      if (pa == a.A && pb == a.B) {  // bpe.cpp:302
      } else if (pa == a.P.unk_id || pb == a.P.unk_id) {  // phantom pair: tracked by the host (Appendix A12)
        phantom = true; emit = true; kind = REC_PHANTOM;
      } else {
        normal = true;
        is_new = pa == a.N || pb == a.N;  // a pair with the token this merge creates cannot exist yet
      }
    }
    // existing keys: load the home sector; new keys: claim it.  Both are issued before either result is used.
    uint64_t hs = mix64(key) & a.pt.mask;
    ulonglong2 home = make_ulonglong2(0ull, 0ull);
    uint32_t home_serial = 0;
    uint64_t prev = 0;
    if (normal && !is_new) { home = ld_ent(&a.pt.ent[hs]); home_serial = a.pt.ent[hs].serial; }
    if (normal && is_new) prev = atomicCAS(reinterpret_cast<ull*>(&a.pt.ent[hs].key), static_cast<ull>(PT_EMPTY), static_cast<ull>(key));
    // the warp's reservations do not depend on those results
    const bool may_list = normal && is_new && no > 0;  // whether it gets a list is decided by its frequency = its delta (old = 0)
    const bool wants_list = may_list && d > 0 && static_cast<uint64_t>(d) >= a.P.min_freq;
    const uint32_t m_new = __ballot_sync(0xFFFFFFFFu, is_new);
    uint32_t pre = wants_list ? no : 0u;  // inclusive warp scan of the list lengths
#pragma unroll
    for (int o2 = 1; o2 < 32; o2 <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, pre, o2); if (lane >= static_cast<uint32_t>(o2)) pre += y; }
    const uint32_t pool_need = __shfl_sync(0xFFFFFFFFu, pre, 31);
    ull b_pt = 0, b_pool = 0;
    if (lane == 0) {  // both reservations in flight before either result is used
      if (m_new) b_pt = atomicAdd(&ctr->pt_n, static_cast<ull>(__popc(m_new)));
      if (pool_need) b_pool = atomicAdd(&ctr->pool_top, static_cast<ull>(pool_need));
    }
    // now the pair table's answers
    if (normal && !is_new) {
      bool fresh = false;
      sl = pt_find_or_claim(a.pt, ctr, key, home, &old, &fresh);
      if (fresh) { serial = static_cast<uint32_t>(atomicAdd(&ctr->pt_n, 1ull)); a.pt.ent[sl].serial = serial; atomicMax(&s.pt_after, static_cast<ull>(serial) + 1ull); }  // (not expected)
      else serial = (sl == hs) ? home_serial : a.pt.ent[sl].serial;
    } else if (normal) {
      sl = hs;
      if (prev != PT_EMPTY) {  // home slot taken by another pair: probe on
        bool fresh = false;
        sl = pt_find_or_claim(a.pt, ctr, key, make_ulonglong2(prev, 0ull), &old, &fresh);
      }
    }
    if (normal) {
      if (d < 0) { const uint64_t ad = static_cast<uint64_t>(-d); nf = old >= ad ? old - ad : 0; } else nf = old + static_cast<uint64_t>(d);  // bpe.cpp:303-307
      a.pt.ent[sl].freq = nf;
      if (nf >= a.P.min_freq) { emit = true; kind = REC_PUSH; }                        // bpe.cpp:308-311
      else if (old >= a.P.min_freq) { emit = true; kind = REC_DEMOTE; }
    }
    const uint32_t m_emit = __ballot_sync(0xFFFFFFFFu, emit);
    uint32_t b_rec = 0;
    if (lane == 0 && m_emit) b_rec = atomicAdd(&s.n_recs, static_cast<uint32_t>(__popc(m_emit)));
    if (lane == 0) {
      if (m_new) atomicMax(&s.pt_after, b_pt + __popc(m_new));
      if (pool_need) atomicMax(&s.pool_after, b_pool + pool_need);
    }
    b_rec = __shfl_sync(0xFFFFFFFFu, b_rec, 0);
    b_pt = __shfl_sync(0xFFFFFFFFu, b_pt, 0); b_pool = __shfl_sync(0xFFFFFFFFu, b_pool, 0);
    if (is_new) { serial = static_cast<uint32_t>(b_pt) + __popc(m_new & lt); a.pt.ent[sl].serial = serial; }
    uint32_t list_len = 0;
    if (wants_list) {
      const ull off = b_pool + pre - no;
      if (off + no <= a.pool_cap && serial < a.pt.lists_cap) {
        ListRef nl; nl.off = off; nl.len = no; nl.fill = 0;
        a.pt.lists[serial] = nl;
        s.lo[ds] = static_cast<uint32_t>(off); s.hi[ds] = static_cast<uint32_t>(off >> 32);
        list_len = no;
      } else atomicOr(&ctr->err, ERR_POOL_FULL);
    }
    if (emit) {
      const uint32_t idx = b_rec + __popc(m_emit & lt);
      const uint64_t val = phantom ? static_cast<uint64_t>(d) : nf;
      if (idx < a.rec_cap) wire_rec(a.recs + idx, a.tag, key, val, sq, kind == REC_PUSH ? rec_pack(REC_PUSH, list_len) : kind, phantom ? REC_NO_SERIAL : serial);
      else atomicOr(&ctr->err, ERR_REC_FULL);
    }
  }
  __syncthreads();
  if (t == 0) {  // publish from shared counters; device counters this pass did not move are read now
    const ull pt_n_after = s.pt_after ? s.pt_after : *reinterpret_cast<volatile ull*>(&ctr->pt_n);
    const ull pool_after = s.pool_after ? s.pool_after : *reinterpret_cast<volatile ull*>(&ctr->pool_top);
    const uint32_t nr = s.n_recs;
    wire_ctrl(a.ctrl, a.tag, nr < a.rec_cap ? nr : a.rec_cap, err_before | (nr > a.rec_cap ? ERR_REC_FULL : 0u), lr.len, n_occ, n_occ, n_keys, pt_n_after, pool_after);
    if (a.dbg) a.dbg[2] = gtime();
  }
  // ---- phase 3: rewrite the occurrences, fill the new pairs' lists (the host is already replaying its heap)
  if (t == nt - 1) {  // bpe.cpp:315: the merged pair's frequency becomes 0.  Nothing reads that entry before the next count pass
    const uint64_t sl = pt_lookup(a.pt, fc_key(a.A, a.B));  // (a new adjacency always contains a new token), so it is kept off the path to the publish
    if (sl != ~0ull) a.pt.ent[sl].freq = 0ull;              // (the host only merges pairs the table holds: its entry exists)
  }
  int32_t* idsw = a.ids;
  auto st = [idsw](uint64_t q, int32_t v2) { idsw[q] = v2; };
  for (uint32_t i = t; i < n_occ; i += nt) {
    const uint4 x = a.sc.a[i];
    const uint4 y = a.sc.b[i];
    const ull b1 = x.z != NONE32 ? ((static_cast<ull>(s.hi[x.z]) << 32) | s.lo[x.z]) : NO_LIST, b2 = y.x != NONE32 ? ((static_cast<ull>(s.hi[y.x]) << 32) | s.lo[y.x]) : NO_LIST;
    lay::rewrite_occurrence(st, x.x, a.lenA, a.lenB, a.N);
    PoolEnt ne; ne.cnt = y.z;
    if (b1 != NO_LIST) { ne.pos = x.y; a.pool[b1 + x.w] = ne; }
    if (b2 != NO_LIST) { ne.pos = x.x; a.pool[b2 + y.y] = ne; }
  }
  if (t == 0 && a.dbg) a.dbg[3] = gtime();
  return true;
}
__global__ void __launch_bounds__(1024) k_merge_small(const MergeArgs a, const uint32_t slots) {
  extern __shared__ __align__(16) unsigned char small_smem[];
  small_merge_body<false>(*reinterpret_cast<SmallStage*>(small_smem), a, slots);
}

// ------------------------------------------------------------------------------------------------ resident merge server
// A launch costs the host ~3 us of API time and the device ~3-5 us before the first instruction runs, with a cold instruction
// cache -- as much as the work of a short merge itself.  While the host replays its heap, ONE CTA therefore stays resident
// (k_merge_server, 1024 threads, its own stream) and takes short merges as 64-byte commands from mapped host memory: the host
// writes (A, B, N, lengths, serial, list length, tag), warp 0 of the CTA polls the block -- four lanes read one 16-byte quarter
// each with ONE load instruction, i.e. one 64-byte PCIe read per poll: polling with four separate reads cost 2.8 us more per
// merge, polling from eight threads 15 us more (non-posted reads queue up) -- and the CTA
// runs small_merge_body and reports the end of phase 3 in a second self-validating block, after which the host may run other
// kernels on the data.  The server leaves when told to (end of bpe_merge_batch / bpe_train, any other operation) or after
// SERVER_IDLE_NS without a command -- it can never outlive a dead host by more than that.
// Command block: four 16-byte quarters {data (64 bits), data (32 bits) | sequence number << 32}; the host writes a quarter's first
// word before its second, the device reads a quarter with one 16-byte load and accepts a command only when all four carry the
// sequence number it expects (so neither a half-written quarter nor quarters of two different commands can be taken for one).
//   q0: A | B << 32, op      q1: N | serial << 32, lenA      q2: list_len | tag << 32, lenB      q3: timed, -
struct ServerCmd { ull w[8]; };
struct ServerDone { ull w[2]; }; // w0 = tag | state << 32 (1 = phase 3 of that merge finished, 2 = server has left), w1 = sequence number it waits for
constexpr ull SERVER_IDLE_NS = 300000000ull;  // 0.3 s
enum : uint32_t { SRV_OP_MERGE = 1, SRV_OP_QUIT = 2 };
__global__ void __launch_bounds__(1024) k_merge_server(MergeArgs a, const ServerCmd* cmd, ServerDone* done, ull first_seq) {
  extern __shared__ __align__(16) unsigned char small_smem[];
  SmallStage& s = *reinterpret_cast<SmallStage*>(small_smem);
  __shared__ ull c[8];
  __shared__ uint32_t op;
  ull* const dbg = a.dbg;
  ull next = first_seq;
  small_clear(s, SMALL_SLOTS_MAX, a.dt.empty);  // the poll loop's barrier below separates this from the first merge
  for (;;) {
    if (threadIdx.x < 32u) {  // warp 0 polls together: lanes 0-3 read one 16-byte quarter each with ONE instruction (one 64-byte PCIe read per poll)
      const uint32_t ln = threadIdx.x;
      const ull t0 = gtime();
      if (ln == 0) op = SRV_OP_QUIT;
      for (;;) {
        ull x = 0, y = 0;
        if (ln < 4u) asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "l"(&cmd->w[2u * ln]));
        const ull want = next & 0xFFFFFFFFull;
        if (__all_sync(0xFFFFFFFFu, ln >= 4u || (y >> 32) == want)) {
          if (ln < 4u) { c[2u * ln] = x; c[2u * ln + 1u] = y; }
          if (ln == 0) op = static_cast<uint32_t>(y & 0xFFFFFFFFull);
          break;
        }
        if (__any_sync(0xFFFFFFFFu, gtime() - t0 > SERVER_IDLE_NS)) break;  // nobody is talking to us any more
      }
    }
    __syncthreads();
    if (op != SRV_OP_MERGE) break;
    __threadfence();  // drop this SM's L1: other kernels may have rewritten symbols, lists and tables since the last command
    a.A = static_cast<int32_t>(c[0] & 0xFFFFFFFFull); a.B = static_cast<int32_t>(c[0] >> 32);
    a.N = static_cast<int32_t>(c[2] & 0xFFFFFFFFull); a.serial = static_cast<uint32_t>(c[2] >> 32);
    a.lenA = static_cast<uint32_t>(c[3] & 0xFFFFFFFFull); a.lenB = static_cast<uint32_t>(c[5] & 0xFFFFFFFFull);
    const uint32_t list_len = static_cast<uint32_t>(c[4] & 0xFFFFFFFFull);
    a.tag = static_cast<uint32_t>(c[4] >> 32);
    a.dbg = c[6] ? dbg : nullptr;
    const uint32_t slots = small_slots_for(list_len);
    small_merge_body<true>(s, a, slots);
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); st_wire(&done->w[0], a.tag | (1ull << 32), next); }  // phase 3 is complete and visible to later kernels
    ++next;
    small_clear(s, slots, a.dt.empty);  // for the next merge, while the host is busy with this one's records
    __syncthreads();
  }
  if (threadIdx.x == 0) { __threadfence(); st_wire(&done->w[0], 0xFFFFFFFFull | (2ull << 32), next); }
}

__global__ void k_rehash(PairTable oldt, PairTable newt, DevCounters* ctr) {
  for (uint64_t s = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; s < oldt.cap; s += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const ulonglong2 e = ld_ent(&oldt.ent[s]);
    if (e.x == PT_EMPTY) continue;
    uint64_t slot = mix64(e.x) & newt.mask;
    for (;;) {
      uint64_t prev = atomicCAS(reinterpret_cast<ull*>(&newt.ent[slot].key), static_cast<ull>(PT_EMPTY), static_cast<ull>(e.x));
      if (prev == PT_EMPTY) { newt.ent[slot].freq = e.y; newt.ent[slot].serial = oldt.ent[s].serial; break; }
      slot = (slot + 1) & newt.mask;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------- save

__global__ void k_rebase(const ull* in, uint32_t n, ull base, ull* out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = in[i] - base;
}
// freq[id] += word count over the live tokens of every word (bpe.cpp:409-415): one thread walks one word through its SKIP marks
__global__ void k_token_freq(const int32_t* __restrict__ ids, const ull* __restrict__ woff, const ull* __restrict__ wcnt, uint32_t n, Params P, ull* freq, uint64_t T) {
  auto ld = [ids](uint64_t q) { return ids[q]; };
  for (uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x; wi < n; wi += gridDim.x * blockDim.x) {
    const ull end = woff[wi + 1], c = wcnt[wi];
    for (ull p = woff[wi] + 1; p < end; p = lay::next_start(ld, p)) {
      const int32_t id = code_to_id(ids[p], P);
      if (id >= 0 && static_cast<uint64_t>(id) < T) atomicAdd(&freq[id], c);  // bpe.cpp:413
    }
  }
}
