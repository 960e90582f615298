// kernels_merge.cuh -- the cooperative per-merge kernel, pair-table rehash, compaction and token-frequency kernels.
// Fragment of engine_cuda.cu: included inside `namespace shred { namespace {`, in the order listed there.
#pragma once

// The per-merge kernel (cooperative launch, persistent grid = SM count x resident CTAs).
//   phase 1  HBM-bound scan of the candidate tiles: every thread streams int4 (4 symbols) and tests the 4 adjacent pairs
//            that start in it; an occurrence emits its count deltas straight into the delta table and is remembered
//   barrier
//   phase 2  every thread folds a share of the touched keys into the pair table and writes the records (bpe.cpp:297-318);
//            the last CTA to finish publishes the counters and the flag the host spins on
//   phase 3  in-place left-packed rewrite of the touched words (bpe.cpp:291-296), off the host's critical path: the
//            first occurrence to claim a word (claimed[wi] = merge number) rewrites it; the last CTA re-arms the counters
template <int UNROLL, bool DIST>
__global__ void __launch_bounds__(256, 4) k_merge(int4* ids4, uint32_t n4, uint32_t n_tiles, uint32_t tiles_per_cta, uint32_t tile_shift,
                                               const uint32_t* __restrict__ planeA, const uint32_t* __restrict__ planeB, uint32_t* planes, uint32_t W, uint32_t id_cap,
                                               const uint32_t* __restrict__ wid, const ull* __restrict__ wcnt, const ull* __restrict__ woff, uint32_t* wlen,
                                               uint32_t* claimed, uint32_t merge_no, int32_t A, int32_t B, int32_t N, Params P, DeltaTable dt, PairTable pt,
                                               DevCounters* ctr, uint32_t* __restrict__ ml, Rec* recs, uint32_t rec_cap, Ctrl* ctrl, uint64_t flag_value,
                                               uint32_t bar_base, ull* dbg, DistArgs D, uint32_t hot_on) {
  __shared__ uint32_t cand[MAX_TILES_PER_CTA];
  __shared__ uint32_t n_cand;
  __shared__ bool last;
  __shared__ HotStage hot;
  __shared__ uint32_t hot_base;
  if (hot_on) hot_init(hot);  // ordered before the first emit_occurrence by the __syncthreads() of the candidate search below
  int32_t* ids = reinterpret_cast<int32_t*>(ids4);
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
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
  // the candidate tiles as one flat list of rows (a row = 32 lanes x int4 = 128 slots = 512 B); every warp keeps UNROLL rows,
  // possibly of different tiles, in flight
  const uint32_t rpt_shift = tile_shift - 7u, n_rows = nc << rpt_shift;
  for (uint32_t r0 = warp * UNROLL; r0 < n_rows; r0 += warps * UNROLL) {
    int4 v[UNROLL];
    uint64_t rowi[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const uint32_t r = r0 + u;
      rowi[u] = r < n_rows ? ((static_cast<uint64_t>(cand[r >> rpt_shift]) << rpt_shift) + (r & ((1u << rpt_shift) - 1u))) * 32u : ~0ull;
      const uint64_t i = rowi[u] + lane;
      v[u] = (rowi[u] != ~0ull && i < n4) ? __ldcv(ids4 + i) : make_int4(DEAD, DEAD, DEAD, DEAD);
    }
    int32_t after[UNROLL];  // first symbol after each row (lane 31's right neighbour)
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      after[u] = DEAD;
      if (lane == 31 && rowi[u] != ~0ull && rowi[u] + 32u < n4) after[u] = __ldcv(ids + 4 * (rowi[u] + 32u));
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      int32_t nxt = __shfl_down_sync(0xFFFFFFFFu, v[u].x, 1);
      if (lane == 31) nxt = after[u];
      uint32_t m = 0;
      m |= (v[u].x == A && v[u].y == B) ? 1u : 0u;
      m |= (v[u].y == A && v[u].z == B) ? 2u : 0u;
      m |= (v[u].z == A && v[u].w == B) ? 4u : 0u;
      m |= (v[u].w == A && nxt == B) ? 8u : 0u;
      if (__any_sync(0xFFFFFFFFu, m != 0)) {
        const uint64_t p0 = (rowi[u] + lane) * 4u;
        while (m) {
          const int k = __ffs(m) - 1;
          m &= m - 1;
          emit_occurrence(ids, p0 + k, wid, wcnt, A, B, N, P, dt, ctr, ml, my_occ, DIST ? (static_cast<uint64_t>(D.rank) << kSeqRankShift) : 0ull, hot, hot_on != 0);
        }
      }
    }
  }
  }
  if (hot_on) {  // uniform
    __syncthreads();
    hot_finish(hot, ids, wid, wcnt, A, B, N, P, dt, ctr, ml, my_occ, DIST ? (static_cast<uint64_t>(D.rank) << kSeqRankShift) : 0ull, &hot_base);
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
    const uint64_t home_slot = mix64(key) & pt.mask;
    const ulonglong2 home = ld_ent(&pt.ent[home_slot]);
    const uint32_t home_serial = pt.serial[home_slot];  // issued with the entry: a record needs it, and most keys sit in their home slot
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
      if (emit) out.serial = (sl == home_slot && home.x == key) ? home_serial : pt.serial[sl];
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
        if (planes) plane_set(planes, W, id_cap, tile_shift, N, w);
        ++w; r += 2;
        cur = nn;
      } else {
        if (w != r) { ids[w] = cur; if (planes) plane_set(planes, W, id_cap, tile_shift, cur, w); }  // a moved symbol may enter another tile
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
