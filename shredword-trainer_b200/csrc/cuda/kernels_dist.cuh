// kernels_dist.cuh -- multi-GPU: inbox protocol over NVLink peer memory, grid barrier, delta exchange, sharded count/sum kernels.
// Fragment of engine_cuda.cu: included inside `namespace shred { namespace {`, in the order listed there.
#pragma once

// ---- multi-GPU exchange over NVLink peer memory ---------------------------------------------------------------
// Every rank owns a contiguous range of the unique words and a full replica of the pair table and of the host heap.
// Per pass (count, merge, token frequencies) each rank's aggregated (key, delta, sequence) list is the only thing that
// crosses GPUs: the kernel STORES it straight into every peer's inbox (memory mapped with CUDA IPC, NVLink/NVSwitch),
// raises a sequence flag there, waits for the peers' flags in its own inbox, and folds their entries into its own delta
// table.  No host round trip and no NCCL call sits between the scan and the pair-table update.
constexpr int MAX_RANKS = 8;
constexpr uint64_t INBOX_ENTRIES = 1ull << 20, INBOX_HDR = 64, INBOX_BYTES = INBOX_HDR + INBOX_ENTRIES * 24;
struct InboxHdr { ull seq; ull n; ull aux; };
// Grid coordinates of a CTA inside ITS rank's grid.  One process per GPU: (blockIdx.x, gridDim.x).  Virtual ranks (test mode, all
// ranks of a job as CTA groups of ONE cooperative launch on one GPU, see engine_cuda.cu VirtualCluster): the CTA's index inside its
// rank's group and the group size -- the guide's rule for machines with fewer GPUs than ranks: kernels that wait on one another
// must be one launch.
struct VGrid { uint32_t b, g; };
struct DistArgs {
  int rank, world;
  uint8_t* peer[MAX_RANKS];  // inbox base of every rank (peer[rank] is local memory)
  ull xseq;                  // exchange number (>= 1); its parity selects the inbox half
};
__device__ __forceinline__ uint8_t* inbox_region(uint8_t* base, int world, ull xseq, int src) {
  return base + ((xseq & 1ull) * static_cast<ull>(world) + static_cast<ull>(src)) * INBOX_BYTES;
}

// Software grid barrier for the cooperative per-merge kernel (all CTAs are co-resident: cudaLaunchCooperativeKernel).
// `bar` only ever grows; `target` = value it reaches when every CTA of this launch has arrived at this barrier.
__device__ __forceinline__ void grid_barrier(uint32_t* bar, uint32_t target, uint32_t* err) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    const long long t0 = clock64();
    while (static_cast<int32_t>(*reinterpret_cast<volatile uint32_t*>(bar) - target) < 0) {
      if (clock64() - t0 > 4000000000ll) { atomicOr(err, ERR_BARRIER); break; }  // ~2 s: never hang the GPU on a host-side accounting bug
    }
    __threadfence();
  }
  __syncthreads();
}

// Cooperative exchange of the delta table's dense list (klist/list/delta/seq) between ranks.  Must be entered after a
// grid barrier (the local list is complete); ends with one grid barrier (number `barrier_no` of this launch).
// On return the local delta table holds the GLOBAL aggregate and *occ_global the global occurrence count.
__device__ __forceinline__ void exchange_deltas(const DeltaTable& dt, DevCounters* ctr, const DistArgs& D, const VGrid vg, uint32_t bar_base, int barrier_no, ull occ_local,
                                                ull* occ_global) {
  __shared__ bool last_sender;
  const uint32_t gtid = vg.b * blockDim.x + threadIdx.x, gthreads = vg.g * blockDim.x;
  // Length of this rank's own list.  Folding the peers' entries (below) appends to the same list, so no CTA may start folding
  // before every CTA has read the length: the fold waits for sent_epoch, which the last CTA to finish sending raises.
  const uint32_t n_raw = *reinterpret_cast<volatile uint32_t*>(dt.n);
  const uint32_t n_local = n_raw < dt.cap ? n_raw : dt.cap;
  if (n_local > INBOX_ENTRIES && gtid == 0) atomicOr(&ctr->err, ERR_INBOX_FULL);
  const uint32_t n_send = n_local < INBOX_ENTRIES ? n_local : static_cast<uint32_t>(INBOX_ENTRIES);
  bool sent = false;
  if (gtid == 0) {  // list length and occurrence count travel with the data, under the same fence
    for (int dst = 0; dst < D.world; dst++) if (dst != D.rank) {
      InboxHdr* h = reinterpret_cast<InboxHdr*>(inbox_region(D.peer[dst], D.world, D.xseq, D.rank));
      h->n = n_send; h->aux = occ_local;
    }
    sent = true;
  }
  for (uint32_t i = gtid; i < n_send; i += gthreads) {  // P2P stores into every peer's inbox
    const uint32_t ds = dt.list[i];
    const ull k = dt.klist[i], d = dt.delta[ds], sq = dt.seq[ds];
    for (int dst = 0; dst < D.world; dst++) if (dst != D.rank) {
      ull* e = reinterpret_cast<ull*>(inbox_region(D.peer[dst], D.world, D.xseq, D.rank) + INBOX_HDR) + 3ull * i;
      e[0] = k; e[1] = d; e[2] = sq;
    }
    sent = true;
  }
  if (sent) __threadfence_system();  // only threads with stores in flight pay for the system-scope fence
  __syncthreads();
  if (threadIdx.x == 0) last_sender = atomicAdd(&ctr->sent_ctas, 1u) == vg.g - 1;
  __syncthreads();
  if (last_sender && threadIdx.x == 0) {  // every CTA's stores are out and fenced: raise the flag at the peers
    __threadfence();
    ctr->sent_ctas = 0;
    *reinterpret_cast<volatile ull*>(&ctr->sent_epoch) = D.xseq;  // every local CTA has read its list length: folding may start
    for (int dst = 0; dst < D.world; dst++) if (dst != D.rank)
      *reinterpret_cast<volatile ull*>(&reinterpret_cast<InboxHdr*>(inbox_region(D.peer[dst], D.world, D.xseq, D.rank))->seq) = D.xseq;
  }
  if (threadIdx.x == 0) {  // every CTA waits for the peers' lists to land in MY inbox (local memory)
    const long long t0 = clock64();
    while (*reinterpret_cast<volatile ull*>(&ctr->sent_epoch) != D.xseq) {
      if (clock64() - t0 > 8000000000ll) { atomicOr(&ctr->err, ERR_BARRIER); break; }
    }
    for (int src = 0; src < D.world; src++) if (src != D.rank) {
      volatile ull* f = &reinterpret_cast<InboxHdr*>(inbox_region(D.peer[D.rank], D.world, D.xseq, src))->seq;
      while (*f != D.xseq) {
        __nanosleep(64);  // hundreds of CTAs poll this line while the peer's NVLink write has to get in
        if (clock64() - t0 > 8000000000ll) { atomicOr(&ctr->err, ERR_PEER_TIMEOUT); break; }  // ~4 s: never hang the GPU
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  ull occ = occ_local;
  for (int src = 0; src < D.world; src++) if (src != D.rank) {  // fold the peers' entries into my delta table
    const uint8_t* reg = inbox_region(D.peer[D.rank], D.world, D.xseq, src);
    const InboxHdr* h = reinterpret_cast<const InboxHdr*>(reg);
    const ull n_src = __ldcv(&h->n);
    occ += __ldcv(&h->aux);
    const ull* e = reinterpret_cast<const ull*>(reg + INBOX_HDR);
    for (ull i = gtid; i < n_src && i < INBOX_ENTRIES; i += gthreads)
      dt_add(dt, ctr, __ldcv(e + 3 * i), static_cast<int64_t>(__ldcv(e + 3 * i + 1)), __ldcv(e + 3 * i + 2));
  }
  *occ_global = occ;
  grid_barrier(&ctr->bar, bar_base + barrier_no * vg.g, &ctr->err);
}

// count pass, sharded: exchange the local pair counts, then block 0 folds the global aggregate and publishes
struct CountFinArgs { DeltaTable dt; PairTable pt; DevCounters* ctr; uint32_t par; uint64_t pool_cap; WireRec* recs; uint32_t rec_cap; Ctrl* ctrl; Params P; uint32_t tag; DistArgs D; uint32_t bar_base; };
__device__ __forceinline__ void dist_count_finalize_body(const CountFinArgs& a, const VGrid vg) {
  ull occ;
  exchange_deltas(a.dt, a.ctr, a.D, vg, a.bar_base, 1, 0ull, &occ);  // entered at kernel start: k_count has completed
  if (vg.b == 0) finalize_count_block(a.dt, a.pt, a.ctr, a.par, a.pool_cap, a.recs, a.rec_cap, a.ctrl, a.P, a.tag);
}
__global__ void __launch_bounds__(256) k_dist_count_finalize(const CountFinArgs a) { dist_count_finalize_body(a, VGrid{blockIdx.x, gridDim.x}); }
__global__ void __launch_bounds__(256) k_dist_count_finalize_virtual(const CountFinArgs* argv, uint32_t g) { dist_count_finalize_body(argv[blockIdx.x / g], VGrid{blockIdx.x % g, g}); }

// token frequencies, sharded: sum of the ranks' partial arrays (T x uint64), same inbox protocol
struct SumArgs { ull* vals; uint64_t T; DevCounters* ctr; DistArgs D; uint32_t bar_base; };
__device__ __forceinline__ void dist_sum_body(const SumArgs& a, const VGrid vg) {
  ull* vals = a.vals; const uint64_t T = a.T; DevCounters* ctr = a.ctr; const DistArgs& D = a.D; const uint32_t bar_base = a.bar_base;
  const uint32_t gtid = vg.b * blockDim.x + threadIdx.x, gthreads = vg.g * blockDim.x;
  for (uint64_t i = gtid; i < T; i += gthreads) {
    const ull v = vals[i];
    for (int dst = 0; dst < D.world; dst++) if (dst != D.rank)
      reinterpret_cast<ull*>(inbox_region(D.peer[dst], D.world, D.xseq, D.rank) + INBOX_HDR)[i] = v;
  }
  __threadfence_system();
  grid_barrier(&ctr->bar, bar_base + 1 * vg.g, &ctr->err);
  if (gtid == 0) {
    for (int dst = 0; dst < D.world; dst++) if (dst != D.rank)
      *reinterpret_cast<volatile ull*>(&reinterpret_cast<InboxHdr*>(inbox_region(D.peer[dst], D.world, D.xseq, D.rank))->seq) = D.xseq;
    const long long t0 = clock64();
    for (int src = 0; src < D.world; src++) if (src != D.rank) {
      volatile ull* f = &reinterpret_cast<InboxHdr*>(inbox_region(D.peer[D.rank], D.world, D.xseq, src))->seq;
      while (*f != D.xseq) if (clock64() - t0 > 8000000000ll) { atomicOr(&ctr->err, ERR_PEER_TIMEOUT); break; }
    }
    __threadfence_system();
  }
  grid_barrier(&ctr->bar, bar_base + 2 * vg.g, &ctr->err);
  for (uint64_t i = gtid; i < T; i += gthreads) {
    ull v = vals[i];
    for (int src = 0; src < D.world; src++) if (src != D.rank)
      v += __ldcv(reinterpret_cast<const ull*>(inbox_region(D.peer[D.rank], D.world, D.xseq, src) + INBOX_HDR) + i);
    vals[i] = v;
  }
}
__global__ void __launch_bounds__(256) k_dist_sum_u64(const SumArgs a) { dist_sum_body(a, VGrid{blockIdx.x, gridDim.x}); }
__global__ void __launch_bounds__(256) k_dist_sum_u64_virtual(const SumArgs* argv, uint32_t g) { dist_sum_body(argv[blockIdx.x / g], VGrid{blockIdx.x % g, g}); }
