// kernels_fold.cuh -- tile occurrence index, delta-table -> pair-table fold (finalize_block), per-occurrence delta emission.
// Fragment of engine_cuda.cu: included inside `namespace shred { namespace {`, in the order listed there.
#pragma once

// -------------------------------------------------------------------------------------------------------------- merge

// ---- tile occurrence index ------------------------------------------------------------------------------------
// planes[id * W + (t >> 5)] bit (t & 31) is set if token `id` occurs in slots [T t, T t + T] with T = 2^tile_shift (the first slot of
// the next tile included, so a pair that straddles the boundary is found from tile t).  Bits are only ever added between
// two compactions, so the index is a superset of the truth: a merge scans exactly the tiles whose bit is set for both A
// and B and provably misses nothing.  Late merges touch a few hundred of tens of thousands of tiles.
// The tile size is chosen per corpus (2^tile_shift slots, 128 ... 4096): the finest tiling whose planes stay within a
// few GB, because finer tiles mean fewer bytes scanned per merge.
constexpr uint32_t MIN_TILE_SHIFT = 7, MAX_TILE_SHIFT = 12, MAX_TILES_PER_CTA = 1024;

__device__ __forceinline__ void plane_set(uint32_t* planes, uint32_t W, uint32_t id_cap, uint32_t tile_shift, int32_t id, uint64_t slot) {
  if (id < 0 || static_cast<uint32_t>(id) >= id_cap) return;
  uint32_t t = static_cast<uint32_t>(slot >> tile_shift);
  uint32_t* wp = planes + static_cast<uint64_t>(id) * W + (t >> 5);
  uint32_t bit = 1u << (t & 31);
  if (!(*wp & bit)) atomicOr(wp, bit);
  if ((slot & ((1ull << tile_shift) - 1)) == 0 && t > 0) {
    --t;
    wp = planes + static_cast<uint64_t>(id) * W + (t >> 5);
    bit = 1u << (t & 31);
    if (!(*wp & bit)) atomicOr(wp, bit);
  }
}

__global__ void __launch_bounds__(256) k_build_planes(const int32_t* __restrict__ ids, uint64_t n_slots, uint32_t* planes, uint32_t W, uint32_t id_cap,
                                                      uint32_t tile_shift) {
  for (uint64_t p = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; p < n_slots; p += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    plane_set(planes, W, id_cap, tile_shift, ids[p], p);
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

// Hot CTAs.  The first merges of a corpus have 10^4-10^6 occurrences: nearly every 128-slot row holds one, so the scan
// loop would run the long, latency-bound emission (five dependent loads, eight table atomics) once per row with one or two
// lanes active, on keys that all rows share (merge 0 of the 1 GB corpus: 545 us for a 150 MB scan).  From its HOT_AFTER-th
// match of a launch on, a CTA therefore only notes the position in shared memory and keeps streaming; after its scan all
// 256 threads emit the noted occurrences in parallel, aggregating the deltas (sum, minimum sequence number: both
// associative, so the result is unchanged) in a shared-memory table that is flushed once, before the grid barrier.
// The host switches this on per launch (hot_on) when the previous merge had many occurrences -- their number falls quickly
// over the first few hundred merges -- so the sparse launches (nearly all of them) pay nothing for it and keep the immediate
// path, whose round trips overlap the scan.
constexpr uint32_t HOT_AFTER = 16, HOT_ML = 3072, HOT_SLOTS = 512, HOT_PROBES = 8;
struct HotStage {
  uint32_t seen;              // matches this CTA has met in this launch
  uint32_t ml[HOT_ML];        // positions noted for the deferred emission (they double as its match-list entries)
  ull key[HOT_SLOTS];         // ~0 = empty
  uint32_t lo[HOT_SLOTS], hi[HOT_SLOTS];  // delta sum as two 32-bit halves (shared-memory adds are native for 32 bits only)
  ull seq[HOT_SLOTS];
};
__device__ __forceinline__ void hot_init(HotStage& h) {
  for (uint32_t i = threadIdx.x; i < HOT_SLOTS; i += blockDim.x) { h.key[i] = ~0ull; h.lo[i] = 0u; h.hi[i] = 0u; h.seq[i] = SEQ_MAX; }
  if (threadIdx.x == 0) h.seen = 0;
}
__device__ __forceinline__ bool hot_add(HotStage& h, uint64_t key, int64_t delta, uint64_t seq) {
  if (key == ~0ull) return false;  // every pair (x, -1) has this key (sign extension, bpe.cpp:277-278) and it is the table's empty marker
  uint32_t slot = static_cast<uint32_t>((key * 0x9E3779B97F4A7C15ull) >> 55) & (HOT_SLOTS - 1);
  for (uint32_t probe = 0; probe < HOT_PROBES; ++probe) {
    ull cur = h.key[slot];
    if (cur == ~0ull) { const ull prev = atomicCAS(&h.key[slot], ~0ull, static_cast<ull>(key)); cur = prev == ~0ull ? key : prev; }
    if (cur == key) {
      const uint32_t d_lo = static_cast<uint32_t>(static_cast<uint64_t>(delta)), d_hi = static_cast<uint32_t>(static_cast<uint64_t>(delta) >> 32);
      const uint32_t old = atomicAdd(&h.lo[slot], d_lo);
      const uint32_t up = d_hi + (old + d_lo < old ? 1u : 0u);  // two's complement: exact modulo 2^64
      if (up) atomicAdd(&h.hi[slot], up);
      if (seq < *reinterpret_cast<volatile ull*>(&h.seq[slot])) atomicMin(&h.seq[slot], static_cast<ull>(seq));
      return true;
    }
    slot = (slot + 1) & (HOT_SLOTS - 1);
  }
  return false;  // crowded: the caller goes to the global table
}

// One occurrence of (A,B) at flat position p: the four count deltas of bpe.cpp:274-290, computed independently per
// occurrence.  Left neighbour = the id that stands there when the reference's left-to-right pass reaches p (N if the two
// symbols before p were themselves merged in this pass), right neighbour = the raw id two slots on.
// HOT: deferred emission of a hot CTA -- deltas go to its shared-memory table, the match-list entry is already staged.
template <bool HOT>
__device__ __forceinline__ void emit_core(const int32_t* ids, uint64_t p, const uint32_t* __restrict__ wid, const ull* __restrict__ wcnt, int32_t A, int32_t B,
                                          int32_t N, const Params& P, const DeltaTable& dt, DevCounters* ctr, uint32_t* ml, uint32_t& my_occ, uint64_t seq_base,
                                          HotStage& hot) {
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
  const int32_t lid = left_merged ? N : code_to_id(l1, P), rid = code_to_id(r2, P);
  const uint64_t key[4] = {fc_key(lid, A), fc_key(lid, N), fc_key(B, rid), fc_key(N, rid)};
  const int64_t delta[4] = {-c, c, -c, c};
  const uint64_t sq[4] = {seq + 0, seq + 1, seq + 2, seq + 3};
  const uint32_t valid = (l1 >= 0 ? 3u : 0u) | (r2 >= 0 ? 12u : 0u);
  ++my_occ;
  if (!HOT) {
    const uint32_t slot_ml = atomicAdd(&ctr->wl_n, 1u);  // issued before the table updates so its round trip overlaps theirs
    dt_add4(dt, ctr, key, delta, sq, valid);
    ml[slot_ml] = static_cast<uint32_t>(p);
  } else {
#pragma unroll
    for (int j = 0; j < 4; j++)
      if ((valid >> j) & 1u) { if (!hot_add(hot, key[j], delta[j], sq[j])) dt_add(dt, ctr, key[j], delta[j], sq[j]); }
  }
}

// called from the scan loop for every match: immediate emission, or (hot CTA) just a note
__device__ __forceinline__ void emit_occurrence(const int32_t* ids, uint64_t p, const uint32_t* __restrict__ wid, const ull* __restrict__ wcnt,
                                                int32_t A, int32_t B, int32_t N, const Params& P, const DeltaTable& dt, DevCounters* ctr, uint32_t* ml,
                                                uint32_t& my_occ, uint64_t seq_base, HotStage& hot, bool hot_on) {
  if (hot_on) {
    const uint32_t nth = atomicAdd(&hot.seen, 1u);  // shared memory
    if (nth >= HOT_AFTER && nth - HOT_AFTER < HOT_ML) { hot.ml[nth - HOT_AFTER] = static_cast<uint32_t>(p); return; }
  }
  emit_core<false>(ids, p, wid, wcnt, A, B, N, P, dt, ctr, ml, my_occ, seq_base, hot);
}

// all threads of the CTA, after a __syncthreads() that follows its scan: emit what was noted, flush table and match list
__device__ __forceinline__ void hot_finish(HotStage& h, const int32_t* ids, const uint32_t* __restrict__ wid, const ull* __restrict__ wcnt, int32_t A, int32_t B,
                                           int32_t N, const Params& P, const DeltaTable& dt, DevCounters* ctr, uint32_t* ml, uint32_t& my_occ, uint64_t seq_base,
                                           uint32_t* s_base) {
  if (h.seen <= HOT_AFTER) return;  // uniform
  const uint32_t staged = min(h.seen - HOT_AFTER, HOT_ML);
  if (threadIdx.x == 0) *s_base = atomicAdd(&ctr->wl_n, staged);
  for (uint32_t i = threadIdx.x; i < staged; i += blockDim.x) emit_core<true>(ids, h.ml[i], wid, wcnt, A, B, N, P, dt, ctr, ml, my_occ, seq_base, h);
  __syncthreads();
  const uint32_t base = *s_base;
  for (uint32_t i = threadIdx.x; i < staged; i += blockDim.x) ml[base + i] = h.ml[i];  // (A,A) runs: second halves ride along; phase 3 only needs the words
  for (uint32_t i = threadIdx.x; i < HOT_SLOTS; i += blockDim.x)
    if (h.key[i] != ~0ull) dt_add(dt, ctr, h.key[i], static_cast<int64_t>((static_cast<uint64_t>(h.hi[i]) << 32) | h.lo[i]), h.seq[i]);
}
