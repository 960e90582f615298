// kernels_merge.cuh -- the per-merge kernel (occurrence-list driven), pair-table rehash and the word walkers (token frequencies).
// Fragment of engine_cuda.cu: included inside `namespace shred { namespace {`, in the order listed there.
#pragma once

// Scratch of one merge, one entry per occurrence found in phase 1 (dense index from a warp-aggregated counter).
struct OccScratch {
  uint4* a;   // {slot of the occurrence, slot where its left neighbour starts after the merge, delta-table slot of (L,N), rank in that key's list}
  uint2* b;   // {delta-table slot of (N,R), rank in that key's list}
  uint32_t cap;
};

struct MergeArgs {
  int32_t* ids; uint64_t ids_cap; const uint32_t* wid; const ull* wcnt;
  uint32_t* pool; uint64_t pool_cap;
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

// The per-merge kernel: one launch per merge, grid sized by the host from the length of the pair's occurrence list.
// The reference scans every word for the pair (bpe.cpp:265-296); here the pair's occurrence list names the only slots that can
// hold it, and every entry is re-validated against the symbol array (lists are never updated when occurrences disappear).
//   phase 1  probe the list entries in parallel against the PRE-merge symbols (layout.hpp probe_occurrence); every occurrence
//            adds its four count deltas to the delta table, draws its ranks in the lists of the two pairs it creates -- (L,N)
//            and (N,R) -- and notes itself in the scratch
//   barrier  (multi-GPU: the aggregated deltas are exchanged over NVLink peer memory here, kernels_dist.cuh)
//   phase 2  every thread folds a share of the touched keys into the pair table, reserves the new pairs' lists in the pool and
//            writes the host records (bpe.cpp:297-318); the last CTA to arrive at the second barrier publishes the counters
//            (self-validating 16-byte blocks: no fence towards the host anywhere, common.cuh)
//   barrier
//   phase 3  (host already replaying its heap) rewrite the occurrences in place -- 4 stores each, nothing moves -- and store
//            their slots in the new pairs' lists
// SINGLE: the whole merge in ONE CTA of 1024 threads (lists up to a few thousand entries: nearly all merges of a run); the
// two barriers are __syncthreads() and the launch is an ordinary one.  Otherwise a cooperative launch of 256-thread CTAs.
// The merge is a chain of dependent memory round trips (list entry -> symbols -> word count -> table slot -> ...), not a stream:
// independent loads are issued together, and nothing on the path waits for the host.
template <bool DIST, bool SINGLE>
__global__ void __launch_bounds__(SINGLE ? 1024 : 256) k_merge(const MergeArgs a) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  DevCounters* const ctr = a.ctr;
  const uint32_t par = a.par;
  if (gtid == 0) {
    if (a.dbg) a.dbg[0] = gtime();
    ctr->n_occ[par ^ 1u] = 0u; ctr->dt_n[par ^ 1u] = 0u; ctr->rec_n[par ^ 1u] = 0u;  // for the next pass (a later launch)
  }
  const int32_t* ids = a.ids;
  // ---- phase 1
  const ListRef lr = a.pt.lists[a.serial];
  const uint32_t len_ceil = (lr.len + 31u) & ~31u;
  for (uint32_t i = gtid; i < len_ceil; i += gthreads) {
    bool ok = false;
    uint32_t p = 0;
    ull c = 0;
    lay::Occ o;
    if (i < lr.len) {
      p = a.pool[lr.off + i];
      // everything the probe certainly or probably reads, in one round trip: the token, its right neighbour, the neighbour after
      // that, the slot on its left, the word index
      const uint64_t pB = static_cast<uint64_t>(p) + a.lenA, pR = min(pB + a.lenB, a.ids_cap - 1);
      const int32_t v_p = ids[p], v_b = ids[pB < a.ids_cap ? pB : a.ids_cap - 1], v_r = ids[pR], v_m = ids[p - 1];
      const uint32_t wi = a.wid[p];
      if (v_p == a.A && v_b == a.B) {
        c = a.wcnt[wi];  // in flight while the probe walks to the left
        auto ld = [&](uint64_t q) { return q == p ? v_p : q == pB ? v_b : q == pR ? v_r : q + 1 == p ? v_m : ids[q]; };
        ok = lay::probe_occurrence(ld, p, a.A, a.B, a.lenA, a.lenB, a.N, a.P, &o);
      }
    }
    const uint32_t found = __ballot_sync(0xFFFFFFFFu, ok);
    if (!found) continue;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&ctr->n_occ[par], static_cast<uint32_t>(__popc(found)));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (!ok) continue;
    const uint32_t idx = base + __popc(found & ((1u << lane) - 1u));
    const int64_t cc = static_cast<int64_t>(c);
    const uint64_t seq = a.seq_base | (static_cast<uint64_t>(p) * 4ull);
    const uint64_t key[4] = {fc_key(o.lid, a.A), fc_key(o.lid, a.N), fc_key(a.B, o.rid), fc_key(a.N, o.rid)};  // bpe.cpp:274-290
    const int64_t delta[4] = {-cc, cc, -cc, cc};
    const uint64_t sq[4] = {seq + 0, seq + 1, seq + 2, seq + 3};
    const uint32_t valid = (o.has_l ? 3u : 0u) | (o.has_r ? 12u : 0u);
    uint32_t slot[4];
    dt_add4(a.dt, ctr, key, delta, sq, valid, slot);
    uint32_t s1 = NONE32, r1 = 0, s2 = NONE32, r2 = 0;
    if (slot[1] != NONE32 && !lay::key_has_unk(key[1], a.P)) { s1 = slot[1]; r1 = atomicAdd(&a.dt.nocc[s1], 1u); }
    if (slot[3] != NONE32 && !lay::key_has_unk(key[3], a.P)) { s2 = slot[3]; r2 = atomicAdd(&a.dt.nocc[s2], 1u); }
    if (idx < a.sc.cap) { a.sc.a[idx] = make_uint4(p, o.pl, s1, r1); a.sc.b[idx] = make_uint2(s2, r2); }
    else atomicOr(&ctr->err, ERR_SCRATCH_FULL);
  }
  if (SINGLE) __syncthreads(); else grid_barrier(&ctr->bar, a.bar_base + gridDim.x, &ctr->err);
  if (a.dbg && gtid == 0) a.dbg[1] = gtime();
  const ull occ_local = *reinterpret_cast<volatile uint32_t*>(&ctr->n_occ[par]);
  ull occ_global = occ_local;
  if (DIST) exchange_deltas(a.dt, ctr, a.D, a.bar_base, 2, occ_local, &occ_global);

  // ---- phase 2: fold the aggregated deltas into the pair table, one key per thread
  const uint32_t n_keys = min(*reinterpret_cast<volatile uint32_t*>(a.dt.n), a.dt.cap);
  if (gtid == gthreads - 1) {  // bpe.cpp:315: the merged pair's frequency becomes 0
    const uint64_t k = fc_key(a.A, a.B);
    uint64_t old;
    const uint64_t sl = pt_find_or_insert(a.pt, ctr, k, ld_ent(&a.pt.ent[mix64(k) & a.pt.mask]), &old);
    a.pt.ent[sl].freq = 0ull;
  }
  for (uint32_t i = gtid; i < n_keys; i += gthreads)
    fold_key<false>(a.dt, a.pt, ctr, i, ld_ent(&a.pt.ent[mix64(a.dt.klist[i]) & a.pt.mask]), a.A, a.B, a.P, a.pool_cap, a.recs, a.rec_cap, &ctr->rec_n[par], a.tag);
  // second barrier; its last arrival publishes (the others are already released and rewriting)
  __syncthreads();
  if (threadIdx.x == 0) {
    bool last = true;
    const uint32_t target2 = a.bar_base + (DIST ? 3u : 2u) * gridDim.x;
    if (!SINGLE) { __threadfence(); last = atomicAdd(&ctr->bar, 1u) + 1u == target2; }
    if (last) {
      if (!SINGLE) __threadfence();
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
    if (!SINGLE) __threadfence();
  }
  __syncthreads();

  // ---- phase 3: rewrite the occurrences, fill the new pairs' lists
  int32_t* idsw = a.ids;
  auto st = [idsw](uint64_t q, int32_t v) { idsw[q] = v; };
  const uint32_t n_occ = min(static_cast<uint32_t>(occ_local), a.sc.cap);
  for (uint32_t i = gtid; i < n_occ; i += gthreads) {
    const uint4 x = a.sc.a[i];
    const uint2 y = a.sc.b[i];
    const ull b1 = x.z != NONE32 ? a.dt.base[x.z] : NO_LIST, b2 = y.x != NONE32 ? a.dt.base[y.x] : NO_LIST;
    lay::rewrite_occurrence(st, x.x, a.lenA, a.lenB, a.N);
    if (b1 != NO_LIST) a.pool[b1 + x.w] = x.y;
    if (b2 != NO_LIST) a.pool[b2 + y.y] = x.x;
  }
  if (a.dbg && gtid == 0) a.dbg[3] = gtime();
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
