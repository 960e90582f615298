// kernels_fold.cuh -- delta-table -> pair-table fold: touched keys in, pair-table updates + host records + occurrence-list
// reservations out.  Shared by the count pass (k_finalize_count, bpe.cpp:219-227) and the merge pass (k_merge phase 2,
// bpe.cpp:297-318).  Fragment of engine_cuda.cu: included inside `namespace shred { namespace {`, in the order listed there.
#pragma once

// The fold of one pass: the n touched keys of the delta table, 32 consecutive keys per warp.  For every key it
//   * re-arms the scratch slot for the next pass,
//   * applies the net delta to the pair table with the reference's clamp (bpe.cpp:303-307) -- COUNT: the sum is the frequency,
//   * reserves the occurrence list of a pair this pass created, if the pair can ever be merged (freq >= min_pair_freq at
//     creation: a pair's occurrences only disappear afterwards, so a pair born below the threshold never reaches the heap),
//   * writes the host record (PUSH / DEMOTE / PHANTOM, engine.hpp) when the host has to act on the key.
// The three dense counters a key may need -- pair serial, pool space, record slot -- are reserved once per warp, the three
// atomics issued together: the pass is a chain of dependent round trips, and each per-key atomic used to be a link of its own.
// Must be called by whole warps (warp_id / n_warps: this warp's index and the number of warps sharing the fold).
template <bool COUNT>
__device__ __forceinline__ void fold_keys(const DeltaTable& dt, const PairTable& pt, DevCounters* ctr, uint32_t n, uint32_t warp_id, uint32_t n_warps, int32_t A, int32_t B,
                                          const Params& P, uint64_t pool_cap, WireRec* recs, uint32_t rec_cap, uint32_t* rec_n, uint32_t tag) {
  const uint32_t lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
  for (uint32_t i0 = warp_id * 32u; i0 < n; i0 += n_warps * 32u) {
    const uint32_t i = i0 + lane;
    bool normal = false, phantom = false, is_new = false, emit = false;
    uint64_t key = 0, sq = 0, sl = 0, nf = 0, old = 0;
    int64_t d = 0;
    uint32_t ds = 0, no = 0, serial = REC_NO_SERIAL, kind = REC_PUSH;
    if (i < n) {
      key = dt.klist[i];
      ds = dt.list[i];
      const ulonglong2 home = ld_ent(&pt.ent[mix64(key) & pt.mask]);
      d = static_cast<int64_t>(dt.delta[ds]);
      sq = dt.seq[ds];
      no = dt.nocc[ds];
      dt.keys[ds] = dt.empty; dt.delta[ds] = 0ull; dt.seq[ds] = SEQ_MAX; dt.nocc[ds] = 0u; dt.base[ds] = NO_LIST;  // re-arm the scratch slot
      const int32_t pa = static_cast<int32_t>(key >> 32), pb = static_cast<int32_t>(key & 0xFFFFFFFFu);  // bpe.cpp:301
      if (!COUNT && pa == A && pb == B) {  // bpe.cpp:302
      } else if (!COUNT && (pa == P.unk_id || pb == P.unk_id)) {  // phantom pair: tracked by the host (Appendix A12)
        phantom = true; emit = true; kind = REC_PHANTOM;
      } else {
        normal = true;
        sl = pt_find_or_claim(pt, ctr, key, home, &old, &is_new);
        if (!is_new) serial = pt.ent[sl].serial;  // same sector as the entry: in flight while the counters are reserved
        if (d < 0) { const uint64_t ad = static_cast<uint64_t>(-d); nf = old >= ad ? old - ad : 0; } else nf = old + static_cast<uint64_t>(d);  // bpe.cpp:303-307
        pt.ent[sl].freq = nf;
        if (nf >= P.min_freq) { emit = true; kind = REC_PUSH; }                        // bpe.cpp:308-311
        else if (!COUNT && old >= P.min_freq) { emit = true; kind = REC_DEMOTE; }
      }
    }
    __syncwarp();
    const bool wants_list = normal && no > 0 && nf >= P.min_freq;  // the pair was created by this pass and can reach the heap
    const uint32_t m_new = __ballot_sync(0xFFFFFFFFu, is_new), m_emit = __ballot_sync(0xFFFFFFFFu, emit);
    uint32_t pre = wants_list ? no : 0u;  // inclusive warp scan of the list lengths
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, pre, o); if (lane >= static_cast<uint32_t>(o)) pre += y; }
    const uint32_t pool_need = __shfl_sync(0xFFFFFFFFu, pre, 31);
    ull b_pt = 0, b_pool = 0;
    uint32_t b_rec = 0;
    if (lane == 0) {  // the three reservations of this warp, issued together
      if (m_new) b_pt = atomicAdd(&ctr->pt_n, static_cast<ull>(__popc(m_new)));
      if (pool_need) b_pool = atomicAdd(&ctr->pool_top, static_cast<ull>(pool_need));
      if (m_emit) b_rec = atomicAdd(rec_n, static_cast<uint32_t>(__popc(m_emit)));
    }
    b_pt = __shfl_sync(0xFFFFFFFFu, b_pt, 0); b_pool = __shfl_sync(0xFFFFFFFFu, b_pool, 0); b_rec = __shfl_sync(0xFFFFFFFFu, b_rec, 0);
    if (is_new) { serial = static_cast<uint32_t>(b_pt) + __popc(m_new & lt); pt.ent[sl].serial = serial; }
    uint32_t list_len = 0;
    if (wants_list) {
      const ull off = b_pool + pre - no;
      if (off + no <= pool_cap && serial < pt.lists_cap) {
        ListRef lr; lr.off = off; lr.len = no; lr.fill = 0;
        pt.lists[serial] = lr;
        dt.base[ds] = off;
        list_len = no;
      } else atomicOr(&ctr->err, ERR_POOL_FULL);
    }
    if (emit) {
      const uint32_t idx = b_rec + __popc(m_emit & lt);
      const uint64_t val = phantom ? static_cast<uint64_t>(d) : nf;
      if (idx < rec_cap) wire_rec(recs + idx, tag, key, val, sq, kind == REC_PUSH ? rec_pack(REC_PUSH, list_len) : kind, phantom ? REC_NO_SERIAL : serial);
      else atomicOr(&ctr->err, ERR_REC_FULL);
    }
  }
}

// Count pass: ONE block folds the aggregated counts and publishes the counters to the host.
__device__ __forceinline__ void finalize_count_block(const DeltaTable& dt, const PairTable& pt, DevCounters* ctr, uint32_t par, uint64_t pool_cap, WireRec* recs, uint32_t rec_cap,
                                                     Ctrl* ctrl, const Params& P, uint32_t tag) {
  const uint32_t n = min(*reinterpret_cast<volatile uint32_t*>(dt.n), dt.cap);
  fold_keys<true>(dt, pt, ctr, n, threadIdx.x >> 5, blockDim.x >> 5, 0, 0, P, pool_cap, recs, rec_cap, &ctr->rec_n[par], tag);
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
