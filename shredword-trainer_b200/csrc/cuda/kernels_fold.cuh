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
  const int32_t lid = left_merged ? N : code_to_id(l1, P), rid = code_to_id(r2, P);
  const uint64_t key[4] = {fc_key(lid, A), fc_key(lid, N), fc_key(B, rid), fc_key(N, rid)};
  const int64_t delta[4] = {-c, c, -c, c};
  const uint64_t sq[4] = {seq + 0, seq + 1, seq + 2, seq + 3};
  const uint32_t slot_ml = atomicAdd(&ctr->wl_n, 1u);  // issued before the table updates so its round trip overlaps theirs
  dt_add4(dt, ctr, key, delta, sq, (l1 >= 0 ? 3u : 0u) | (r2 >= 0 ? 12u : 0u));
  ml[slot_ml] = static_cast<uint32_t>(p);
  ++my_occ;
}
