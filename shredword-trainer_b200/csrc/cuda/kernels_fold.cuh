// kernels_fold.cuh -- delta-table -> pair-table fold: one touched key in, pair-table update + host record + occurrence-list
// reservation out.  Shared by the count pass (k_finalize_count, bpe.cpp:219-227) and the merge pass (k_merge phase 2,
// bpe.cpp:297-318).  Fragment of engine_cuda.cu: included inside `namespace shred { namespace {`, in the order listed there.
#pragma once

// One key of the aggregated delta list (dense index i of the delta table).
//   * re-arms its scratch slot for the next pass,
//   * applies the net delta to the pair table with the reference's clamp (bpe.cpp:303-307) -- COUNT: the sum is the frequency,
//   * reserves the occurrence list of a pair this pass created, if the pair can ever be merged (freq >= min_pair_freq at
//     creation: a pair's occurrences only disappear afterwards, so a pair born below the threshold never reaches the heap),
//   * writes the host record (PUSH / DEMOTE / PHANTOM, engine.hpp) when the host has to act on the key.
template <bool COUNT>
__device__ __forceinline__ void fold_key(const DeltaTable& dt, const PairTable& pt, DevCounters* ctr, uint32_t i, ulonglong2 home, int32_t A, int32_t B, const Params& P,
                                         uint64_t pool_cap, WireRec* recs, uint32_t rec_cap, uint32_t* rec_n, uint32_t tag) {
  const uint64_t key = dt.klist[i];
  const uint32_t ds = dt.list[i];
  const int64_t d = static_cast<int64_t>(dt.delta[ds]);
  const uint64_t sq = dt.seq[ds];
  const uint32_t no = dt.nocc[ds];
  dt.keys[ds] = dt.empty; dt.delta[ds] = 0ull; dt.seq[ds] = SEQ_MAX; dt.nocc[ds] = 0u;  // re-arm the scratch slot
  dt.base[ds] = NO_LIST;
  const int32_t pa = static_cast<int32_t>(key >> 32), pb = static_cast<int32_t>(key & 0xFFFFFFFFu);  // bpe.cpp:301
  if (!COUNT && pa == A && pb == B) return;  // bpe.cpp:302
  Rec out; out.key = key; out.seq = sq; out.serial = REC_NO_SERIAL; out.kind = REC_PUSH; out.val = 0;
  bool emit = false;
  if (!COUNT && (pa == P.unk_id || pb == P.unk_id)) {  // phantom pair: tracked by the host (Appendix A12)
    out.kind = REC_PHANTOM; out.val = static_cast<uint64_t>(d); emit = true;
  } else {
    uint64_t old;
    const uint64_t sl = pt_find_or_insert(pt, ctr, key, home, &old);
    uint64_t nf;
    if (d < 0) { const uint64_t ad = static_cast<uint64_t>(-d); nf = old >= ad ? old - ad : 0; } else nf = old + static_cast<uint64_t>(d);  // bpe.cpp:303-307
    pt.ent[sl].freq = nf;
    const uint32_t serial = pt.serial[sl];
    uint32_t list_len = 0;
    if (no > 0 && nf >= P.min_freq) {  // the pair was created by this pass and can reach the heap: it gets its occurrence list
      const ull off = atomicAdd(&ctr->pool_top, static_cast<ull>(no));
      if (off + no <= pool_cap && serial < pt.lists_cap) {
        ListRef lr; lr.off = off; lr.len = no; lr.fill = 0;
        pt.lists[serial] = lr;
        dt.base[ds] = off;
        list_len = no;
      } else atomicOr(&ctr->err, ERR_POOL_FULL);
    }
    if (nf >= P.min_freq) { out.kind = rec_pack(REC_PUSH, list_len); out.val = nf; emit = true; }            // bpe.cpp:308-311
    else if (!COUNT && old >= P.min_freq) { out.kind = REC_DEMOTE; out.val = nf; emit = true; }
    out.serial = serial;
  }
  if (emit) {
    const uint32_t idx = atomicAdd(rec_n, 1u);
    if (idx < rec_cap) wire_rec(recs + idx, tag, out.key, out.val, out.seq, out.kind, out.serial); else atomicOr(&ctr->err, ERR_REC_FULL);
  }
}

// Count pass: ONE block folds the aggregated counts and publishes the counters to the host.
__device__ __forceinline__ void finalize_count_block(const DeltaTable& dt, const PairTable& pt, DevCounters* ctr, uint32_t par, uint64_t pool_cap, WireRec* recs, uint32_t rec_cap,
                                                     Ctrl* ctrl, const Params& P, uint32_t tag) {
  const uint32_t n = min(*reinterpret_cast<volatile uint32_t*>(dt.n), dt.cap);
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
    fold_key<true>(dt, pt, ctr, i, ld_ent(&pt.ent[mix64(dt.klist[i]) & pt.mask]), 0, 0, P, pool_cap, recs, rec_cap, &ctr->rec_n[par], tag);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t nr = *reinterpret_cast<volatile uint32_t*>(&ctr->rec_n[par]);
    wire_ctrl(ctrl, tag, nr < rec_cap ? nr : rec_cap, *reinterpret_cast<volatile uint32_t*>(&ctr->err), 0u, 0u, 0ull, n, *reinterpret_cast<volatile ull*>(&ctr->pt_n),
              *reinterpret_cast<volatile ull*>(&ctr->pool_top));
  }
}
__global__ void __launch_bounds__(1024) k_finalize_count(DeltaTable dt, PairTable pt, DevCounters* ctr, uint32_t par, uint64_t pool_cap, WireRec* recs, uint32_t rec_cap, Ctrl* ctrl,
                                                         Params P, uint32_t tag) {
  finalize_count_block(dt, pt, ctr, par, pool_cap, recs, rec_cap, ctrl, P, tag);
}
