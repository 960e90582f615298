// engine_cuda.cu -- the B200 (sm_100a) device engine of the BPE trainer: corpus ingest, pair counting, merge
// application and pair-table maintenance as hand-written CUDA kernels.  Implements shred::Engine (../engine.hpp).
// This file holds the host side (buffers, launches, multi-GPU rendezvous); the kernels live in the .cuh fragments
// included below.
//
// Data layout in HBM (csrc/layout.hpp has the encoding and the per-occurrence logic, shared with a CPU walk-through in tests/)
//   ids[]   int32, position-stable: every unique word owns a fixed slot range [HDR|wi] b0 b1 ... (one slot per byte), words back
//           to back in reference word order (A3).  A token lives at the slot of its first byte and never moves; a merged token
//           marks its second and last slot with SKIP(length).  Flat position is monotone in the reference's (word, position)
//           scan order: it is the sequence number the host needs (Appendix A14) and what the occurrence lists store.
//   wid[]   uint32 word index per slot; wcnt[] uint64 word counts; woff[] uint64 slot offsets (N+1)
//   pool[]  uint32 occurrence lists: for every pair that can reach the heap, the slots where it occurred when it was created
//           (by the count pass or by the one merge that created its younger token); validated lazily, never updated
//   pair table   open addressing, uint64 key (first<<32|second) -> uint64 freq + serial; lists[serial] = {offset, length}
//   delta table  open addressing scratch, key -> (sum of +/-count, min sequence, occurrences)  (reference FreqChangeMap, bpe.cpp:9-38)
//
// Kernels (reference loop each one replaces)                                               file
//   k_tokenize ........ bpe.cpp:131-153 + hash.cpp:29-53   tokenise, unique-word table     kernels_tokenize.cuh
//   k_hist_words ...... histogram.cpp:30-36                unweighted byte histogram        kernels_ingest.cuh
//   k_scatter/k_sort_buckets  hash.cpp:61-72               word order (djb2 & 4095, first)  kernels_ingest.cuh
//   k_symbolize ....... histogram.cpp:7-27                 bytes -> ids, unk substitution   kernels_ingest.cuh
//   k_count ........... bpe.cpp:187-218                    adjacent pair counts             kernels_count.cuh
//   k_finalize_count .. bpe.cpp:219-227                    fold counts, seed records, reserve lists   kernels_fold.cuh
//   k_fill_lists ...... (no counterpart)                   initial occurrence lists         kernels_count.cuh
//   k_merge ........... bpe.cpp:265-318                    one cooperative launch per merge kernels_merge.cuh
//                       probe the pair's occurrence list, per-occurrence count deltas | grid barrier | deltas folded into the
//                       pair table + records published | grid barrier | in-place rewrite (4 stores per occurrence) + new lists
//   exchange_deltas ... (multi-GPU) per-merge delta exchange over NVLink peer memory        kernels_dist.cuh
//   k_token_freq ...... bpe.cpp:409-415                    final token frequencies          kernels_merge.cuh
#include <cuda_profiler_api.h>
#include <cuda_runtime.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/shred_abi.h"
#include "../charset.hpp"
#include "../engine.hpp"
#include "../layout.hpp"
#include "../shard.hpp"

namespace shred {
namespace {

#include "common.cuh"
#include "kernels_tokenize.cuh"
#include "kernels_scan.cuh"
#include "kernels_ingest.cuh"
#include "kernels_count.cuh"
#include "kernels_fold.cuh"
#include "kernels_dist.cuh"
#include "kernels_merge.cuh"

// ============================================================================================================ engine

inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline uint64_t next_pow2(uint64_t x) { uint64_t p = 1; while (p < x) p <<= 1; return p; }

// Virtual ranks -- TEST MODE for machines with fewer GPUs than ranks (SHRED_VIRTUAL_RANKS=W): the W ranks of one sharded job are
// W trainers of ONE process on ONE GPU, each driven by its own host thread.  The sharded data path is unchanged (same device code,
// same inbox protocol, same host control per rank); only the three kernels in which ranks wait for one another are launched
// differently: kernels that spin on a peer must never be separate launches on one GPU (nothing guarantees they run together), so
// the ranks stage their arguments here and the last one to arrive launches ONE cooperative kernel that holds every rank's CTA group
// (k_*_virtual).  Ranks and job instances are numbered by creation order; all ranks share one stream.
struct VirtualCluster {
  static constexpr int GRID_PER_RANK = 8;
  std::mutex mu;
  std::condition_variable cv;
  int world = 0, created = 0, arrived = 0, rc = 0;
  uint64_t generation = 0;
  cudaStream_t stream = nullptr;
  uint8_t* inbox[MAX_RANKS] = {};
  MergeArgs m[MAX_RANKS]; CountFinArgs c[MAX_RANKS]; SumArgs sm[MAX_RANKS];
  void* d_args = nullptr;
  static VirtualCluster& get() { static VirtualCluster v; return v; }
  // every rank calls this with the same kind; returns once the combined kernel has been launched
  template <class Stage>
  int collective(int kind, Stage stage) {
    std::unique_lock<std::mutex> lk(mu);
    stage(*this);
    const uint64_t gen = generation;
    if (++arrived == world) {
      arrived = 0;
      rc = launch(kind);
      ++generation;
      cv.notify_all();
    } else if (!cv.wait_for(lk, std::chrono::seconds(120), [&] { return generation != gen; })) {
      std::fprintf(stderr, "[ERROR]\t virtual ranks: a rank never reached the collective\n");
      return -1;
    }
    return rc;
  }
  int launch(int kind) {
    for (int r = 0; r < world; r++) for (int k = 0; k < MAX_RANKS; k++) {
      uint8_t* pk = k < world ? inbox[k] : nullptr;
      if (kind == 0) m[r].D.peer[k] = pk; else if (kind == 1) c[r].D.peer[k] = pk; else sm[r].D.peer[k] = pk;
    }
    const size_t bytes = kind == 0 ? sizeof(MergeArgs) : kind == 1 ? sizeof(CountFinArgs) : sizeof(SumArgs);
    const void* src = kind == 0 ? static_cast<const void*>(m) : kind == 1 ? static_cast<const void*>(c) : static_cast<const void*>(sm);
    if (!d_args) CK(cudaMalloc(&d_args, MAX_RANKS * std::max(sizeof(MergeArgs), std::max(sizeof(CountFinArgs), sizeof(SumArgs)))));
    CK(cudaMemcpyAsync(d_args, src, bytes * world, cudaMemcpyHostToDevice, stream));
    uint32_t g = GRID_PER_RANK;
    void* args[] = {&d_args, &g};
    const void* fn = kind == 0 ? reinterpret_cast<const void*>(k_merge_virtual) : kind == 1 ? reinterpret_cast<const void*>(k_dist_count_finalize_virtual) : reinterpret_cast<const void*>(k_dist_sum_u64_virtual);
    CK(cudaLaunchCooperativeKernel(fn, dim3(g * world), dim3(256), args, 0, stream));
    return 0;
  }
};

// Mapped pinned host blocks (control block, record buffer, server command block) are kept for the life of the process and handed
// from one trainer to the next: cudaHostAlloc / cudaFreeHost pin and unpin pages and synchronise the device -- 90 ms per
// create_trainer and 270-370 ms per bpe_trainer_destroy at the 10 GB configuration before this cache.  A block is returned only
// after its trainer's streams have drained, and every user clears it before use (stale pass tags must not look valid).
class PinnedCache {
 public:
  static PinnedCache& get() { static PinnedCache c; return c; }
  void* acquire(size_t bytes) {
    bytes = (bytes + 4095) & ~static_cast<size_t>(4095);
    {
      std::lock_guard<std::mutex> g(mu_);
      size_t best = free_.size();
      for (size_t i = 0; i < free_.size(); i++)
        if (free_[i].second >= bytes && free_[i].second <= 4 * bytes && (best == free_.size() || free_[i].second < free_[best].second)) best = i;
      if (best != free_.size()) { void* p = free_[best].first; free_.erase(free_.begin() + static_cast<std::ptrdiff_t>(best)); return p; }
    }
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
  }
  void release(void* p, size_t bytes) {
    if (!p) return;
    bytes = (bytes + 4095) & ~static_cast<size_t>(4095);
    std::lock_guard<std::mutex> g(mu_);
    free_.emplace_back(p, bytes);
  }

 private:
  std::mutex mu_;
  std::vector<std::pair<void*, size_t>> free_;
};

class CudaEngine : public Engine {
 public:
  CudaEngine(int dev, const cudaDeviceProp& prop) : dev_(dev), n_sm_(prop.multiProcessorCount > 0 ? prop.multiProcessorCount : N_SM_FALLBACK) {
    std::snprintf(name_, sizeof name_, "%s sm_%d%d %d SMs", prop.name, prop.major, prop.minor, n_sm_);
  }
  ~CudaEngine() override { release_all(); }

  int init() {
    CK(cudaSetDevice(dev_));
    CK(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
    void* cp = PinnedCache::get().acquire(sizeof(Ctrl));
    if (!cp) { std::fprintf(stderr, "[ERROR]\t no pinned host memory\n"); return -1; }
    std::memset(cp, 0, sizeof(Ctrl));
    ctrl_ = static_cast<Ctrl*>(cp);
    CK(cudaMallocAsync(reinterpret_cast<void**>(&ctr_), sizeof(DevCounters), st_));
    CK(cudaMemset(ctr_, 0, sizeof(DevCounters)));
    CK(cudaEventCreate(&ev0_));
    CK(cudaEventCreate(&ev1_));
    CK(cudaEventCreate(&evm0_));
    CK(cudaEventCreate(&evm1_));
    CK(cudaEventCreateWithFlags(&ev_gen_, cudaEventDisableTiming));
    cudaMemPool_t pool;  // stream-ordered allocations; keep freed blocks cached so repeated loads do not pay cudaMalloc
    if (cudaDeviceGetDefaultMemPool(&pool, dev_) == cudaSuccess) { uint64_t thr = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr); }
    {  // in-kernel phase timestamps (%globaltimer) of the timed launches
      void* dp = PinnedCache::get().acquire(32 * sizeof(ull));
      if (!dp) { std::fprintf(stderr, "[ERROR]\t no pinned host memory\n"); return -1; }
      dbg_ = static_cast<ull*>(dp);
      std::memset(dp, 0, 32 * sizeof(ull));
      const char* d = std::getenv("SHRED_DEBUG_TIMING");
      dbg_print_ = d && *d && *d != '0';
    }
    CK(cudaFuncSetAttribute(k_sort_buckets, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(SORT_CAP * sizeof(ull))));
    CK(cudaFuncSetAttribute(k_count, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(CountStage))));
    int nb = 0;
    CK(cudaFuncSetAttribute(k_merge_small, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(SmallStage))));
    CK(cudaFuncSetAttribute(k_merge_server, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(SmallStage))));
    {
      void* sp = PinnedCache::get().acquire(sizeof(ServerCmd) + 64);
      if (!sp) { std::fprintf(stderr, "[ERROR]\t no pinned host memory\n"); return -1; }
      std::memset(sp, 0, sizeof(ServerCmd) + 64);
      srv_cmd_ = static_cast<ServerCmd*>(sp);
      srv_done_ = reinterpret_cast<ServerDone*>(static_cast<uint8_t*>(sp) + sizeof(ServerCmd));
      CK(cudaStreamCreateWithFlags(&st_srv_, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&ev_srv_, cudaEventDisableTiming));
      if (const char* ns = std::getenv("SHRED_NO_SERVER")) srv_enabled_ = !(*ns && *ns != '0');
    }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_merge<false>, 256, 0) == cudaSuccess && nb > 0) merge_ctas_per_sm_ = nb < 4 ? nb : 4;
    if (const char* w = std::getenv("SHRED_WORLD")) world_ = std::atoi(w);
    if (const char* vr = std::getenv("SHRED_VIRTUAL_RANKS")) if (std::atoi(vr) > 1) { world_ = std::atoi(vr); virtual_ = true; }
    if (world_ > 1) {
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_merge<true>, 256, 0) == cudaSuccess && nb > 0 && nb < merge_ctas_per_sm_) merge_ctas_per_sm_ = nb;
      const char* r = std::getenv("SHRED_RANK");
      rank_ = r ? std::atoi(r) : 0;
      if (virtual_) {  // ranks by creation order; every rank works on the cluster's stream
        VirtualCluster& vc = VirtualCluster::get();
        std::lock_guard<std::mutex> hold(vc.mu);
        if (vc.world != world_) { vc.world = world_; vc.created = 0; }
        rank_ = vc.created++ % world_;
        if (!vc.stream) CK(cudaStreamCreateWithFlags(&vc.stream, cudaStreamNonBlocking));
        CK(cudaStreamDestroy(st_));
        st_ = vc.stream;
        force_grid_ = VirtualCluster::GRID_PER_RANK;
      }
      if (world_ > MAX_RANKS || rank_ < 0 || rank_ >= world_) { std::fprintf(stderr, "[ERROR]\t bad SHRED_RANK/SHRED_WORLD (%d/%d, at most %d ranks)\n", rank_, world_, MAX_RANKS); return -1; }
      RC(dist_setup());
    } else { world_ = 1; rank_ = 0; }
    const char* e = std::getenv("SHRED_TIMING");
    timing_every_ = e && *e ? std::atoi(e) : 0;
    if (const char* mg = std::getenv("SHRED_MERGE_GRID")) force_grid_ = std::atoi(mg);  // tests: fixed grid for every merge launch
    if (const char* sm = std::getenv("SHRED_SMALL_MAX")) small_max_ = std::min<uint32_t>(SMALL_MAX, static_cast<uint32_t>(std::strtoul(sm, nullptr, 10)));  // longest list k_merge_small takes (0 = never)
    if (std::getenv("SHRED_SMALL_MAX") && std::strtoul(std::getenv("SHRED_SMALL_MAX"), nullptr, 10) == 0) small_max_ = 0;
    if (const char* pm = std::getenv("SHRED_PROFILE_MERGES")) {  // "0,1,2000": cudaProfilerStart/Stop around these merges (ncu --profile-from-start off)
      for (const char* q = pm; *q;) { char* end = nullptr; const unsigned long v = std::strtoul(q, &end, 10); if (end == q) break; profile_merges_.push_back(static_cast<uint32_t>(v)); q = *end ? end + 1 : end; }
    }
    return 0;
  }

  // ------------------------------------------------------------------------------------------------- multi-GPU setup
  // One process per GPU.  Every rank allocates its inbox with cudaMalloc, exports it with CUDA IPC through a small file
  // in the rendezvous directory SHRED_RDV (shared by the ranks of one job) and maps every peer's inbox.
  int dist_setup() {
    if (virtual_) {  // peers are plain device pointers of this process; wait until every rank of this job instance has registered
      const size_t bytes = 2ull * world_ * INBOX_BYTES;
      CK(cudaMalloc(reinterpret_cast<void**>(&inbox_), bytes));
      CK(cudaMemset(inbox_, 0, bytes));
      CK(cudaDeviceSynchronize());
      VirtualCluster& vc = VirtualCluster::get();
      std::lock_guard<std::mutex> hold(vc.mu);
      vc.inbox[rank_] = inbox_;  // the peers' pointers are filled in when a collective is launched: every rank exists by then
      for (int r = 0; r < MAX_RANKS; r++) dist_.peer[r] = nullptr;
      dist_.rank = rank_; dist_.world = world_; dist_.xseq = 0;
      return 0;
    }
    static int instance = 0;  // ranks create their trainers in the same order, so instance numbers agree
    const int inst = instance++;
    const char* rdv = std::getenv("SHRED_RDV");
    if (!rdv || !*rdv) { std::fprintf(stderr, "[ERROR]\t SHRED_WORLD > 1 needs SHRED_RDV (a directory shared by the ranks)\n"); return -1; }
    const size_t bytes = 2ull * world_ * INBOX_BYTES;
    CK(cudaMalloc(reinterpret_cast<void**>(&inbox_), bytes));
    CK(cudaMemset(inbox_, 0, bytes));
    CK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t mine;
    CK(cudaIpcGetMemHandle(&mine, inbox_));
    auto path = [&](int r, const char* ext) { return std::string(rdv) + "/inst" + std::to_string(inst) + "_rank" + std::to_string(r) + ext; };
    {
      const std::string tmp = path(rank_, ".tmp"), fin = path(rank_, ".ipc");
      FILE* f = std::fopen(tmp.c_str(), "wb");
      if (!f) { std::fprintf(stderr, "[ERROR]\t cannot write %s\n", tmp.c_str()); return -1; }
      std::fwrite(&mine, sizeof mine, 1, f);
      std::fclose(f);
      if (std::rename(tmp.c_str(), fin.c_str()) != 0) return -1;
    }
    for (int r = 0; r < MAX_RANKS; r++) dist_.peer[r] = nullptr;
    dist_.rank = rank_; dist_.world = world_; dist_.xseq = 0;
    dist_.peer[rank_] = inbox_;
    const double t0 = now_ms();
    for (int r = 0; r < world_; r++) if (r != rank_) {
      cudaIpcMemHandle_t h;
      for (;;) {
        FILE* f = std::fopen(path(r, ".ipc").c_str(), "rb");
        if (f) { size_t got = std::fread(&h, sizeof h, 1, f); std::fclose(f); if (got == 1) break; }
        if (now_ms() - t0 > 120000.0) { std::fprintf(stderr, "[ERROR]\t rank %d: timed out waiting for rank %d in %s\n", rank_, r, rdv); return -1; }
        usleep(1000);
      }
      void* p = nullptr;
      CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      dist_.peer[r] = static_cast<uint8_t*>(p);
    }
    // nobody may unlink its file before every rank has opened every handle
    { FILE* f = std::fopen(path(rank_, ".ok").c_str(), "wb"); if (f) std::fclose(f); }
    for (int r = 0; r < world_; r++) {
      struct stat st;
      while (stat(path(r, ".ok").c_str(), &st) != 0) {
        if (now_ms() - t0 > 120000.0) { std::fprintf(stderr, "[ERROR]\t rank %d: rank %d never finished its setup\n", rank_, r); return -1; }
        usleep(1000);
      }
    }
    rdv_prefix_ = std::string(rdv) + "/inst" + std::to_string(inst) + "_rank";
    return 0;
  }
  void dist_teardown() {
    if (world_ <= 1 || !inbox_) return;
    if (virtual_) { cudaStreamSynchronize(st_); cudaFree(inbox_); inbox_ = nullptr; return; }  // (the test destroys its trainers after all ranks have finished)
    for (int r = 0; r < world_; r++) if (r != rank_ && dist_.peer[r]) cudaIpcCloseMemHandle(dist_.peer[r]);
    // exported memory must outlive every peer's mapping: free only after all ranks have closed their handles
    if (!rdv_prefix_.empty()) {
      { FILE* f = std::fopen((rdv_prefix_ + std::to_string(rank_) + ".closed").c_str(), "wb"); if (f) std::fclose(f); }
      const double t0 = now_ms();
      for (int r = 0; r < world_; r++) {
        struct stat st;
        while (stat((rdv_prefix_ + std::to_string(r) + ".closed").c_str(), &st) != 0 && now_ms() - t0 < 20000.0) usleep(500);
      }
    }
    cudaFree(inbox_);
    inbox_ = nullptr;
    // the small rendezvous files stay: the job that created the directory removes it
  }
  DistArgs next_exchange() { DistArgs d = dist_; d.xseq = ++dist_.xseq; return d; }

  // ---------------------------------------------------------------------------------------------------------- load
  int load(const uint8_t* text, size_t n, const EngineConfig& cfg, LoadInfo* info) override { return load_impl(text, -1, n, cfg, info); }

  // bpe_load_corpus(path): the file goes to HBM through a ring of pinned staging buffers, one reader thread per buffer
  // (pread from the page cache) while the previous chunks are already in flight over PCIe.
  int load_file(int fd, size_t n, const EngineConfig& cfg, LoadInfo* info) override { return load_impl(nullptr, fd, n, cfg, info); }

  int load_impl(const uint8_t* text, int fd, size_t n, const EngineConfig& cfg, LoadInfo* info) {
    CK(cudaSetDevice(dev_));
    RC(srv_stop());
    cfg_ = cfg;
    vocab_hint_ = cfg.vocab_size < (1ull << 22) ? cfg.vocab_size : (1ull << 22);
    P_.unk_id = cfg.unk_id;
    P_.unk_code = cfg.unk_id >= 0 ? cfg.unk_id : UNK_CODE_NEG;
    P_.min_freq = cfg.min_freq;
    release_corpus();
    std::memset(info, 0, sizeof *info);
    std::memset(&es_, 0, sizeof es_);
    // --- corpus bytes to HBM, padded with spaces so token walks and 16-byte loads stay in bounds
    const uint64_t padded = ((n + 15) & ~15ull) + 64;
    uint8_t* d_text = nullptr;
    const double ta = now_ms();
    CK(cudaMallocAsync(reinterpret_cast<void**>(&d_text), padded, st_));
    PreTok pre;
    double t0 = now_ms();
    if (n && text) CK(cudaMemcpyAsync(d_text, text, n, cudaMemcpyHostToDevice, st_));
    if (n && !text) {
      // file path: the unique-word table exists before the first byte arrives and every chunk is tokenised as soon as it has
      // landed (copy stream | compute stream), so the tokeniser hides behind the PCIe copy
      CK(cudaMemsetAsync(d_text, ' ', padded, st_));
      pre.cap = wt_cap_for(n); pre.seed = 0x5bd1e995u;
      RC(alloc_wt(&pre.wt, pre.cap));
      CK(cudaMemsetAsync(ctr_, 0, sizeof(DevCounters), st_));
      bar_count_ = 0;
      pre.valid = true;
      // First load of a process: the stream-ordered pool is still small and every later allocation of the ingest (symbols, word
      // indices, occurrence lists: ~0.6 B per corpus byte) would grow it on the critical path (0.78 s at 10 GB).  Grow it now, on a
      // helper thread and stream, while the file is on its way over PCIe.
      std::thread warm;
      {
        cudaMemPool_t pool; uint64_t reserved = 0, used = 0; size_t free_b = 0, total_b = 0;
        if (cudaDeviceGetDefaultMemPool(&pool, dev_) == cudaSuccess && cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
          uint64_t want = n - n / 4, idle = reserved > used ? reserved - used : 0;
          if (want > free_b / 3) want = free_b / 3;
          if (want > idle + (256ull << 20)) {
            const uint64_t grow = want - idle;
            const int dev = dev_;
            warm = std::thread([grow, dev]() {
              cudaSetDevice(dev);
              cudaStream_t sa; void* p = nullptr;
              if (cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking) != cudaSuccess) return;
              if (cudaMallocAsync(&p, grow, sa) == cudaSuccess) cudaFreeAsync(p, sa);
              cudaStreamSynchronize(sa);
              cudaStreamDestroy(sa);
            });
          }
        }
      }
      const int src = stream_file(fd, n, d_text, &pre);
      if (warm.joinable()) warm.join();
      if (src != 0) { free_wt(pre.wt); cudaFreeAsync(d_text, st_); return -1; }
    } else {
      CK(cudaMemsetAsync(d_text + n, ' ', padded - n, st_));
    }
    CK(cudaStreamSynchronize(st_));
    es_.h2d_ms += now_ms() - t0; es_.h2d_bytes += n;
    if (dbg_print_) std::fprintf(stderr, "[LOAD]\t text buffer %.1f ms, copy%s %.1f ms\n", t0 - ta, pre.valid ? " + overlapped tokenise" : "", now_ms() - t0);

    CK(cudaEventRecord(ev0_, st_));
    const double ti = now_ms();
    int rc = ingest(d_text, n, info, pre);
    cudaFreeAsync(d_text, st_);
    if (rc != 0) return rc;
    CK(cudaEventRecord(ev1_, st_));
    CK(cudaStreamSynchronize(st_));
    float ms = 0; cudaEventElapsedTime(&ms, ev0_, ev1_);
    es_.ingest_device_ms = ms;
    es_.ingest_bytes = static_cast<double>(n) + 4.0 * info->n_symbols + 12.0 * info->n_words;
    if (dbg_print_) std::fprintf(stderr, "[LOAD]\t rest of the ingest %.1f ms (device %.1f ms)\n", now_ms() - ti, ms);
    return 0;
  }

  // unique-word table of a load attempt, possibly filled while the file was still arriving
  struct PreTok { WordTable wt{}; uint64_t cap = 0; uint32_t seed = 0; bool valid = false; };
  static uint64_t wt_cap_for(uint64_t n) { uint64_t cap = next_pow2(n / 64 + 1); return cap < (1u << 16) ? (1u << 16) : cap; }  // grown 4x and redone if more than half fills up
  int alloc_wt(WordTable* wt, uint64_t cap) {
    CK(cudaMallocAsync(reinterpret_cast<void**>(&wt->tag), cap * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&wt->first), cap * 8, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&wt->count), cap * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&wt->len), cap * 4, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&wt->bucket), cap * 4, st_));
    wt->cap = cap; wt->mask = cap - 1;
    CK(cudaMemsetAsync(wt->tag, 0, cap * 8, st_)); CK(cudaMemsetAsync(wt->first, 0xFF, cap * 8, st_)); CK(cudaMemsetAsync(wt->count, 0, cap * 8, st_));
    return 0;
  }
  void free_wt(WordTable& wt) { cudaFreeAsync(wt.tag, st_); cudaFreeAsync(wt.first, st_); cudaFreeAsync(wt.count, st_); cudaFreeAsync(wt.len, st_); cudaFreeAsync(wt.bucket, st_); }

  // pinned staging ring shared by all trainers of the process (cudaHostAlloc is slow, so it is done once)
  static constexpr int STAGE_BUFS = 12;
  static constexpr size_t STAGE_BYTES = 16u << 20;
  struct StageRing { uint8_t* buf[STAGE_BUFS] = {}; std::mutex mu; };
  static StageRing& stage_ring() { static StageRing r; return r; }

  int stream_file(int fd, size_t n, uint8_t* d_text, PreTok* pre) {
    StageRing& ring = stage_ring();
    std::lock_guard<std::mutex> hold(ring.mu);  // one load at a time uses the ring
    if (!st_copy_) CK(cudaStreamCreateWithFlags(&st_copy_, cudaStreamNonBlocking));
    cudaEvent_t ready, landed;
    CK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&landed, cudaEventDisableTiming));
    CK(cudaEventRecord(ready, st_));          // text buffer pre-filled, word table cleared
    CK(cudaStreamWaitEvent(st_copy_, ready, 0));
    uint64_t cut = 0;                        // everything below it has been handed to the tokeniser
    const size_t n_chunks = (n + STAGE_BYTES - 1) / STAGE_BYTES;
    cudaEvent_t done[STAGE_BUFS];
    for (int b = 0; b < STAGE_BUFS; b++) CK(cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming));
    // state[b]: 0 = reader owns the buffer, 1 = filled (main may copy), 2 = copy issued (reader waits for the event)
    std::atomic<int> state[STAGE_BUFS];
    std::atomic<bool> failed{false};
    for (int b = 0; b < STAGE_BUFS; b++) state[b].store(0);
    std::vector<std::thread> readers;
    for (int b = 0; b < STAGE_BUFS; b++) {
      readers.emplace_back([&, b]() {
        cudaSetDevice(dev_);
        // first load of a process: every reader pins its own buffer (192 MB take the driver 80 ms in one piece), so the first chunks travel while the rest of the ring is still being pinned
        if (!ring.buf[b] && static_cast<size_t>(b) < n_chunks && cudaHostAlloc(reinterpret_cast<void**>(&ring.buf[b]), STAGE_BYTES, cudaHostAllocDefault) != cudaSuccess) { ring.buf[b] = nullptr; failed.store(true); return; }
        for (size_t c = b; c < n_chunks && !failed.load(); c += STAGE_BUFS) {
          const size_t off = c * STAGE_BYTES, len = std::min(STAGE_BYTES, n - off);
          size_t got = 0;
          while (got < len) {
            const ssize_t r = pread(fd, ring.buf[b] + got, len - got, static_cast<off_t>(off + got));
            if (r <= 0) { failed.store(true); break; }
            got += static_cast<size_t>(r);
          }
          state[b].store(1, std::memory_order_release);
          while (state[b].load(std::memory_order_acquire) != 2 && !failed.load()) std::this_thread::yield();
          if (failed.load()) break;
          if (cudaEventSynchronize(done[b]) != cudaSuccess) { failed.store(true); break; }  // the buffer may be overwritten again
          state[b].store(0, std::memory_order_release);
        }
      });
    }
    for (size_t c = 0; c < n_chunks && !failed.load(); c++) {  // issue the copies in file order
      const int b = static_cast<int>(c % STAGE_BUFS);
      const size_t off = c * STAGE_BYTES, len = std::min(STAGE_BYTES, n - off);
      while (state[b].load(std::memory_order_acquire) != 1 && !failed.load()) std::this_thread::yield();
      if (failed.load()) break;
      if (cudaMemcpyAsync(d_text + off, ring.buf[b], len, cudaMemcpyHostToDevice, st_copy_) != cudaSuccess || cudaEventRecord(done[b], st_copy_) != cudaSuccess) { failed.store(true); break; }
      // the chunk's last delimiter (found in the staging buffer, still intact): tokens that start before it also end before it
      size_t k = len;
      while (k > 0) { const uint8_t ch = ring.buf[b][k - 1]; if (ch == ' ' || ch == '\n' || ch == '\t' || ch == '\r') break; --k; }
      state[b].store(2, std::memory_order_release);
      if (k > 0 && off + k > cut) {
        const uint64_t hi = off + k;
        if (cudaStreamWaitEvent(st_, done[b], 0) != cudaSuccess) { failed.store(true); break; }
        k_tokenize<<<grid_for((hi - cut + 15) / 16 + 1, 256), 256, 0, st_>>>(d_text, cut, hi, pre->wt, ctr_, pre->seed); launches_++; es_.ingest_launches++;
        cut = hi;
      }
    }
    for (auto& t : readers) t.join();
    if (!failed.load()) {
      if (cudaEventRecord(landed, st_copy_) != cudaSuccess || cudaStreamWaitEvent(st_, landed, 0) != cudaSuccess) failed.store(true);
      else if (cut < n) { k_tokenize<<<grid_for((n - cut + 15) / 16 + 1, 256), 256, 0, st_>>>(d_text, cut, n, pre->wt, ctr_, pre->seed); launches_++; es_.ingest_launches++; }
    }
    cudaStreamSynchronize(st_copy_);
    cudaEventDestroy(ready); cudaEventDestroy(landed);
    for (int b = 0; b < STAGE_BUFS; b++) cudaEventDestroy(done[b]);
    if (failed.load()) { std::fprintf(stderr, "[ERROR]\t reading the corpus file failed\n"); return -1; }
    return 0;
  }

  int ingest(const uint8_t* d_text, uint64_t n, LoadInfo* info, const PreTok& pre) {
    DevCounters zero; std::memset(&zero, 0, sizeof zero);
    WordTable wt; std::memset(&wt, 0, sizeof wt);
    uint32_t N = 0; ull n_tokens = 0;
    uint64_t cap = pre.valid ? pre.cap : wt_cap_for(n);
    uint32_t seed = pre.valid ? pre.seed : 0x5bd1e995u;
    for (int attempt = 0;; ++attempt) {
      if (attempt > 8) { std::fprintf(stderr, "[ERROR]\t unique-word table did not converge\n"); return -1; }
      if (cap > (1ull << 32)) { std::fprintf(stderr, "[ERROR]\t unique-word table too large\n"); return -1; }
      if (attempt == 0 && pre.valid) wt = pre.wt;  // filled while the file was arriving (load_impl)
      else {
        RC(alloc_wt(&wt, cap));
        CK(cudaMemcpyAsync(ctr_, &zero, sizeof zero, cudaMemcpyHostToDevice, st_));
        bar_count_ = 0;
        if (n) { k_tokenize<<<grid_for((n + 15) / 16, 256), 256, 0, st_>>>(d_text, 0, n, wt, ctr_, seed); launches_++; es_.ingest_launches++; }
      }
      DevCounters c;
      CK(cudaMemcpyAsync(&c, ctr_, sizeof c, cudaMemcpyDeviceToHost, st_));
      CK(cudaStreamSynchronize(st_));
      CK(cudaGetLastError());
      if (c.err & ERR_HAS_NUL) {
        cudaFreeAsync(wt.tag, st_); cudaFreeAsync(wt.first, st_); cudaFreeAsync(wt.count, st_); cudaFreeAsync(wt.len, st_); cudaFreeAsync(wt.bucket, st_);
        return 1;
      }
      const bool too_full = static_cast<uint64_t>(c.n_unique) * 2 > cap;
      if ((c.err & (ERR_WT_FULL | ERR_WT_COLLISION)) || too_full) {
        cudaFreeAsync(wt.tag, st_); cudaFreeAsync(wt.first, st_); cudaFreeAsync(wt.count, st_); cudaFreeAsync(wt.len, st_); cudaFreeAsync(wt.bucket, st_);
        if ((c.err & ERR_WT_FULL) || too_full) cap *= 4;
        if (c.err & ERR_WT_COLLISION) seed = seed * 2654435761u + 12345u;
        continue;
      }
      N = c.n_unique; n_tokens = c.n_tokens;
      break;
    }
    auto free_wt = [&]() { this->free_wt(wt); };
    if (N >= 0x7FFFFFF0u) { free_wt(); std::fprintf(stderr, "[ERROR]\t too many unique words\n"); return -1; }
    n_words_ = N;
    info->n_words = N; info->n_tokens = n_tokens;
    // --- reference word order: bucket = djb2 & 4095 ascending, first occurrence ascending inside a bucket
    uint32_t *u_slot = nullptr, *u_n = nullptr, *bcnt = nullptr, *bstart = nullptr, *cursor = nullptr, *tmp_slot = nullptr, *order_slot = nullptr;
    ull *tmp_first = nullptr, *d_hist = nullptr, *len1 = nullptr, *sums = nullptr;
    uint8_t* d_keep = nullptr;
    uint32_t* wlen = nullptr;  // word lengths: only needed until the offsets exist
    const uint64_t Na = N ? N : 1;
    CK(cudaMallocAsync(reinterpret_cast<void**>(&u_slot), Na * 4, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&tmp_slot), Na * 4, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&order_slot), Na * 4, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&tmp_first), Na * 8, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&len1), Na * 8, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&u_n), 4, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&bcnt), 4096 * 4, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&bstart), 4097 * 4, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&cursor), 4096 * 4, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&d_hist), 256 * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&d_keep), 256, st_));
    const uint32_t nb_scan = static_cast<uint32_t>((Na + SCAN_TILE - 1) / SCAN_TILE);
    CK(cudaMallocAsync(reinterpret_cast<void**>(&sums), (static_cast<uint64_t>(nb_scan) + 1) * 8, st_));
    auto free_tmp = [&]() {
      cudaFreeAsync(u_slot, st_); cudaFreeAsync(tmp_slot, st_); cudaFreeAsync(order_slot, st_); cudaFreeAsync(tmp_first, st_); cudaFreeAsync(len1, st_); cudaFreeAsync(u_n, st_); cudaFreeAsync(bcnt, st_);
      cudaFreeAsync(bstart, st_); cudaFreeAsync(cursor, st_); cudaFreeAsync(d_hist, st_); cudaFreeAsync(d_keep, st_); cudaFreeAsync(sums, st_); cudaFreeAsync(wlen, st_);
    };
    CK(cudaMemsetAsync(u_n, 0, 4, st_)); CK(cudaMemsetAsync(bcnt, 0, 4096 * 4, st_)); CK(cudaMemsetAsync(cursor, 0, 4096 * 4, st_));
    CK(cudaMemsetAsync(d_hist, 0, 256 * 8, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&wcnt_), Na * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&wlen), Na * 4, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&woff_), (Na + 1) * 8, st_));
    ull S1 = 0;  // total slots = symbols + headers
    if (N) {
      k_collect<<<grid_for(cap, 256), 256, 0, st_>>>(wt, u_slot, u_n, bcnt);
      k_scan4096<<<1, 1024, 0, st_>>>(bcnt, bstart);
      k_scatter<<<grid_for(N, 256), 256, 0, st_>>>(wt, u_slot, N, bstart, cursor, tmp_slot, tmp_first);
      k_sort_buckets<<<4096, SORT_THREADS, SORT_CAP * sizeof(ull), st_>>>(tmp_slot, tmp_first, bstart, order_slot);
      k_rank_big<<<grid_for(N, 128), 128, 0, st_>>>(wt, tmp_slot, tmp_first, N, bstart, order_slot); launches_++;  // only buckets above SORT_CAP do work
      k_hist_words<<<grid_for(N, 256), 256, 0, st_>>>(d_text, wt, order_slot, N, d_hist, wcnt_, wlen, len1);
      k_scan_sums<<<nb_scan, SCAN_THREADS, 0, st_>>>(len1, N, sums);
      k_scan_top<<<1, SCAN_THREADS, 0, st_>>>(sums, nb_scan, sums + nb_scan);
      k_scan_apply<<<nb_scan, SCAN_THREADS, 0, st_>>>(len1, N, sums, woff_);
      launches_ += 8; es_.ingest_launches += 8;
      CK(cudaMemcpyAsync(&S1, sums + nb_scan, 8, cudaMemcpyDeviceToHost, st_));
    }
    CK(cudaMemcpyAsync(info->hist, d_hist, 256 * 8, cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    CK(cudaGetLastError());
    es_.d2h_bytes += 256 * 8 + 8 + sizeof(DevCounters);
    charset_keep(info->hist, cfg_.coverage, info->keep, &info->n_distinct, &info->n_keep);
    std::memcpy(keep_, info->keep, 256);
    info->n_symbols = S1 - N;
    uint32_t lo = 0, n_local = N;
    host_counts_.clear();
    if (world_ > 1 && N) {  // keep only this rank's contiguous range of words (shard.hpp); ingest itself is replicated
      std::vector<ull> hoff(static_cast<size_t>(N) + 1);
      CK(cudaMemcpyAsync(hoff.data(), woff_, static_cast<uint64_t>(N) * 8, cudaMemcpyDeviceToHost, st_));
      host_counts_.resize(N);
      CK(cudaMemcpyAsync(host_counts_.data(), wcnt_, static_cast<uint64_t>(N) * 8, cudaMemcpyDeviceToHost, st_));
      CK(cudaStreamSynchronize(st_));
      hoff[N] = S1;
      lo = static_cast<uint32_t>(shard_begin(hoff.data(), N, rank_, world_));
      const uint32_t hi = static_cast<uint32_t>(shard_begin(hoff.data(), N, rank_ + 1, world_));
      n_local = hi - lo;
      const ull base_off = hoff[lo];
      S1 = hoff[hi] - base_off;
      ull *wc = nullptr, *wo = nullptr;
      const uint64_t nl = n_local ? n_local : 1;
      CK(cudaMallocAsync(reinterpret_cast<void**>(&wc), nl * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&wo), (nl + 1) * 8, st_));
      if (n_local) {
        CK(cudaMemcpyAsync(wc, wcnt_ + lo, static_cast<uint64_t>(n_local) * 8, cudaMemcpyDeviceToDevice, st_));
        k_rebase<<<grid_for(n_local, 256), 256, 0, st_>>>(woff_ + lo, n_local, base_off, wo); launches_++;
      }
      cudaFreeAsync(wcnt_, st_); cudaFreeAsync(woff_, st_);
      wcnt_ = wc; woff_ = wo;
      es_.d2h_bytes += static_cast<uint64_t>(N) * 16;
    }
    n_words_ = n_local;
    n_slots_ = S1; n_live_ = S1;
    if (S1 + 64 >= (1ull << 32) || N >= lay::LOW30) { free_tmp(); free_wt(); std::fprintf(stderr, "[ERROR]\t corpus needs more than 2^32 symbol slots or 2^30 words on one GPU\n"); return -1; }
    ids_cap_ = ((S1 + 8 + 1023) / 1024) * 1024;
    CK(cudaMallocAsync(reinterpret_cast<void**>(&ids_), ids_cap_ * 4, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&wid_), ids_cap_ * 4, st_));
    CK(cudaMemsetAsync(wid_, 0, ids_cap_ * 4, st_));
    // occurrence lists: one entry per adjacent pair of the fresh corpus (< S1) + two per rewritten occurrence (< S1 - words)
    pool_cap_ = S1 + 2 * (S1 - n_local) + 1024;
    CK(cudaMallocAsync(reinterpret_cast<void**>(&pool_), pool_cap_ * sizeof(PoolEnt), st_));
    CK(cudaMemcpyAsync(d_keep, info->keep, 256, cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(woff_ + n_local, &S1, 8, cudaMemcpyHostToDevice, st_));
    if (n_local) { k_symbolize<<<grid_for(n_local, 256), 256, 0, st_>>>(d_text, wt, order_slot + lo, n_local, woff_, d_keep, P_.unk_code, ids_, wid_); launches_++; es_.ingest_launches++; }
    k_fill_i32<<<grid_for(ids_cap_ - S1, 256), 256, 0, st_>>>(ids_, S1, ids_cap_, DEAD); launches_++; es_.ingest_launches++;
    const int32_t terminator = lay::make_hdr(n_local);  // the last token of the last word has no right neighbour
    CK(cudaMemcpyAsync(ids_ + S1, &terminator, 4, cudaMemcpyHostToDevice, st_));
    CK(cudaStreamSynchronize(st_));
    CK(cudaGetLastError());
    free_tmp(); free_wt();
    // --- pair/delta tables sized for this trainer
    RC(alloc_tables());
    loaded_ = true;
    fresh_ = true;  // every token is still a single byte
    return 0;
  }

  int alloc_tables() {
    if (!dt_.keys) {
      uint64_t cap = next_pow2(8ull * (256 + vocab_hint_) + 1024); if (cap < (1u << 16)) cap = 1u << 16;
      RC(alloc_dt(cap));
    }
    if (!pt_.ent) RC(alloc_pt(&pt_, 1ull << 20));
    RC(ensure_scratch(2 * SMALL_MAX));  // the one-CTA kernel and the merge server never wait for scratch
    return 0;
  }
  int alloc_dt(uint64_t cap) {
    if (dt_.keys) {
      cudaFreeAsync(dt_.keys, st_); cudaFreeAsync(dt_.delta, st_); cudaFreeAsync(dt_.seq, st_); cudaFreeAsync(dt_.nocc, st_); cudaFreeAsync(dt_.base, st_);
      cudaFreeAsync(dt_.list, st_); cudaFreeAsync(dt_.klist, st_); dt_.keys = nullptr;
    }
    CK(cudaMallocAsync(reinterpret_cast<void**>(&dt_.keys), cap * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&dt_.delta), cap * 8, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&dt_.seq), cap * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&dt_.nocc), cap * 4, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&dt_.base), cap * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&dt_.list), cap * 4, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&dt_.klist), cap * 8, st_));
    dt_.cap = static_cast<uint32_t>(cap); dt_.mask = cap - 1;
    // a key value no pair can produce: high word >= 2^31 that is neither all-ones nor unk_id
    uint32_t hi = 0x80000000u; if (static_cast<uint32_t>(cfg_.unk_id) == hi) hi = 0x80000001u;
    dt_.empty = static_cast<uint64_t>(hi) << 32;
    k_fill_u64<<<grid_for(cap, 256), 256, 0, st_>>>(reinterpret_cast<ull*>(dt_.keys), cap, dt_.empty);
    CK(cudaMemsetAsync(dt_.delta, 0, cap * 8, st_)); CK(cudaMemsetAsync(dt_.seq, 0xFF, cap * 8, st_)); CK(cudaMemsetAsync(dt_.nocc, 0, cap * 4, st_));
    CK(cudaMemsetAsync(dt_.base, 0xFF, cap * 8, st_));
    launches_++;
    // the record buffer holds at most one record per key of a pass; callers fetch recs_ only after the last possible growth
    if (!recs_ || rec_cap_ < dt_.cap) {
      if (recs_) { CK(cudaStreamSynchronize(st_)); PinnedCache::get().release(recs_, static_cast<size_t>(rec_cap_) * sizeof(WireRec)); recs_ = nullptr; }
      rec_cap_ = dt_.cap;
      recs_ = static_cast<WireRec*>(PinnedCache::get().acquire(static_cast<size_t>(rec_cap_) * sizeof(WireRec)));
      if (!recs_) { std::fprintf(stderr, "[ERROR]\t no pinned host memory for %u records\n", rec_cap_); rec_cap_ = 0; return -1; }
      std::memset(recs_, 0, static_cast<size_t>(rec_cap_) * sizeof(WireRec));  // tag 0 = never written (pass tags start at 1)
    }
    return 0;
  }
  int alloc_pt(PairTable* pt, uint64_t cap) {
    CK(cudaMallocAsync(reinterpret_cast<void**>(&pt->ent), cap * sizeof(PairEnt), st_));
    pt->cap = cap; pt->mask = cap - 1;
    pt->lists_cap = cap / 2 + 4096;  // the table stays at most half full, serials are dense
    CK(cudaMallocAsync(reinterpret_cast<void**>(&pt->lists), pt->lists_cap * sizeof(ListRef), st_));
    CK(cudaMemsetAsync(pt->ent, 0xFF, cap * sizeof(PairEnt), st_));  // key = EMPTY; freq is written when the entry is claimed
    CK(cudaMemsetAsync(pt->lists, 0, pt->lists_cap * sizeof(ListRef), st_));  // len 0 = the pair has no list
    return 0;
  }
  int grow_pt(uint64_t need_entries) {
    uint64_t cap = pt_.cap; while (need_entries * 2 > cap) cap *= 2;
    if (cap == pt_.cap) return 0;
    PairTable nt; std::memset(&nt, 0, sizeof nt);
    RC(alloc_pt(&nt, cap));
    k_rehash<<<grid_for(pt_.cap, 256), 256, 0, st_>>>(pt_, nt, ctr_); launches_++;
    CK(cudaMemcpyAsync(nt.lists, pt_.lists, pt_.lists_cap * sizeof(ListRef), cudaMemcpyDeviceToDevice, st_));
    CK(cudaStreamSynchronize(st_));
    cudaFreeAsync(pt_.ent, st_); cudaFreeAsync(pt_.lists, st_);
    pt_ = nt;
    return 0;
  }
  int ensure_scratch(uint64_t n_entries) {  // one scratch entry per occurrence of a merge, at most one per list entry
    if (n_entries <= sc_.cap) return 0;
    uint64_t cap = sc_.cap ? sc_.cap : (1u << 16);
    while (cap < n_entries) cap *= 2;
    if (cap > 0xFFFFFFF0ull) cap = 0xFFFFFFF0ull;
    if (sc_.a) { cudaFreeAsync(sc_.a, st_); cudaFreeAsync(sc_.b, st_); }
    CK(cudaMallocAsync(reinterpret_cast<void**>(&sc_.a), cap * sizeof(uint4), st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&sc_.b), cap * sizeof(uint4), st_));
    sc_.cap = static_cast<uint32_t>(cap);
    return 0;
  }

  // --------------------------------------------------------------------------------------------------------- count
  int count_pairs(const Rec** recs, size_t* n) override {
    CK(cudaSetDevice(dev_));
    RC(srv_stop());
    *recs = out_.data(); *n = 0;
    if (!loaded_) return 0;
    for (int attempt = 0; attempt < 12; ++attempt) {
      CK(cudaMemsetAsync(pt_.ent, 0xFF, pt_.cap * sizeof(PairEnt), st_));
      CK(cudaMemsetAsync(pt_.lists, 0, pt_.lists_cap * sizeof(ListRef), st_));
      CK(cudaMemsetAsync(ctr_, 0, sizeof(DevCounters), st_));
      bar_count_ = 0;
      pt_n_ = 0;
      pass_ = 0;  // the count pass runs on parity 0 of the freshly zeroed counters; the first merge takes parity 1
      dt_.n = &ctr_->dt_n[0];
      ++flag_;
      CK(cudaEventRecord(ev0_, st_));
      const uint32_t n4c = static_cast<uint32_t>((n_slots_ + 3) / 4);
      if (n_words_) {
        // fresh corpus with few distinct bytes: direct-indexed shared-memory tables (k_count_dense); otherwise the hashed ones
        uint32_t K = 0;
        ByteLut lut;
        if (fresh_ && !std::getenv("SHRED_COUNT_HASHED")) {
          for (int bv = 0; bv < 256; bv++) lut.code[bv] = (keep_[bv] && bv != P_.unk_code) ? static_cast<uint8_t>(K++) : static_cast<uint8_t>(0xFF);
        }
        const size_t dense_bytes = static_cast<size_t>(K) * K * 16;
        if (K > 0 && K < 255 && dense_bytes <= (196u << 10)) {
          const int per_sm = std::max<int>(1, std::min<int>(4, static_cast<int>((200u << 10) / (dense_bytes + 1024))));
          CK(cudaFuncSetAttribute(k_count_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dense_bytes)));
          k_count_dense<<<n_sm_ * per_sm, 256, dense_bytes, st_>>>(reinterpret_cast<const int4*>(ids_), reinterpret_cast<const uint4*>(wid_), n4c, wcnt_, lut, K, dt_, ctr_,
                                                                   world_ > 1 ? seq_base(rank_) : 0ull);
        } else {
          k_count<<<n_sm_ * 4, 256, sizeof(CountStage), st_>>>(reinterpret_cast<const int4*>(ids_), reinterpret_cast<const uint4*>(wid_), n4c, wcnt_, P_, dt_, ctr_,
                                                               world_ > 1 ? seq_base(rank_) : 0ull);
        }
        launches_++;
      }
      CK(cudaEventRecord(ev1_, st_));
      // the finalize pass needs dt_n <= cap/2 and room in the pair table: check before consuming the delta table
      DevCounters c;
      CK(cudaMemcpyAsync(&c, ctr_, sizeof c, cudaMemcpyDeviceToHost, st_));
      CK(cudaStreamSynchronize(st_));
      CK(cudaGetLastError());
      float ms = 0; cudaEventElapsedTime(&ms, ev0_, ev1_);
      if ((c.err & ERR_DT_FULL) || static_cast<uint64_t>(c.dt_n[0]) * 2 > dt_.cap) {  // enlarge the scratch table, redo (ids untouched)
        RC(alloc_dt(static_cast<uint64_t>(dt_.cap) * 4));
        continue;
      }
      RC(grow_pt((world_ > 1 ? static_cast<uint64_t>(dt_.cap) / 2 : static_cast<uint64_t>(c.dt_n[0])) + 4ull * (256 + vocab_hint_) + 1024));  // replicas must size identically
      CK(cudaEventRecord(ev0_, st_));
      Ctrl* a_ctrl = ctrl_;
      const uint32_t tag = static_cast<uint32_t>(flag_);
      if (world_ > 1) {
        const int grid = virtual_ ? VirtualCluster::GRID_PER_RANK : n_sm_ * 2;
        CountFinArgs ca;
        ca.dt = dt_; ca.pt = pt_; ca.ctr = ctr_; ca.par = 0; ca.pool_cap = pool_cap_; ca.recs = recs_; ca.rec_cap = rec_cap_; ca.ctrl = a_ctrl; ca.P = P_; ca.tag = tag;
        ca.D = next_exchange(); ca.bar_base = bar_count_;
        bar_count_ += 1u * static_cast<uint32_t>(grid);
        if (virtual_) RC(VirtualCluster::get().collective(1, [&](VirtualCluster& vc) { vc.c[rank_] = ca; }));
        else {
          void* args[] = {&ca};
          CK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_dist_count_finalize), dim3(grid), dim3(256), args, 0, st_));
        }
      } else {
        k_finalize_count<<<1, 1024, 0, st_>>>(dt_, pt_, ctr_, 0u, pool_cap_, recs_, rec_cap_, a_ctrl, P_, tag);
      }
      launches_++;
      if (n_words_) { k_fill_lists<<<n_sm_ * 8, 256, 0, st_>>>(reinterpret_cast<const int4*>(ids_), reinterpret_cast<const uint4*>(wid_), wcnt_, n4c, P_, pt_, pool_, ctr_); launches_++; }
      CK(cudaEventRecord(ev1_, st_));
      RC(wait_flag());
      if (cv_.err) { std::fprintf(stderr, "[ERROR]\t device count pass failed (err=%u)\n", cv_.err); return -1; }
      CK(cudaStreamSynchronize(st_));  // the lists are complete before the first merge reads them (same stream anyway) and before the error check below
      float ms2 = 0; cudaEventElapsedTime(&ms2, ev0_, ev1_);
      DevCounters c2;
      CK(cudaMemcpy(&c2, ctr_, sizeof c2, cudaMemcpyDeviceToHost));
      if (c2.err) { std::fprintf(stderr, "[ERROR]\t device list fill failed (err=%u)\n", c2.err); return -1; }
      es_.count_launches++; es_.count_device_ms += ms; es_.count_bytes += 4.0 * static_cast<double>(n_slots_) + 12.0 * n_words_;  // 4S + 12N; the kernel reads 8 B per slot + counts
      es_.fill_device_ms += ms2; es_.fill_bytes += 8.0 * static_cast<double>(n_slots_) + 8.0 * static_cast<double>(cv_.pool_top);
      pt_n_ = cv_.pt_n;
      RC(fetch_records());
      *recs = out_.data();
      *n = out_.size();
      es_.d2h_bytes += *n * sizeof(Rec) + sizeof(Ctrl);
      return 0;
    }
    return -1;
  }

  // ------------------------------------------------------------------------------------------------- merge server
  // (kernels_merge.cuh k_merge_server.)  The server reads symbols, lists and tables that every other kernel of the engine also
  // writes, so the host keeps them apart: it waits for the server's "phase 3 done" before it launches anything else on the data
  // (srv_quiesce), makes the server wait for the engine's stream when it starts (event), waits for the last general merge kernel
  // to finish before it hands the server a command, and stops the server before buffers move or a run of merges ends.
  int srv_start() {
    if (srv_alive_) return 0;
    CK(cudaEventRecord(ev_srv_, st_));
    CK(cudaStreamWaitEvent(st_srv_, ev_srv_, 0));
    MergeArgs ma; std::memset(&ma, 0, sizeof ma);
    ma.ids = ids_; ma.ids_cap = ids_cap_; ma.wid = wid_; ma.wcnt = wcnt_; ma.pool = pool_; ma.pool_cap = pool_cap_; ma.sc = sc_;
    ma.P = P_; ma.dt = dt_; ma.pt = pt_; ma.ctr = ctr_; ma.recs = recs_; ma.rec_cap = rec_cap_; ma.ctrl = ctrl_; ma.dbg = dbg_; ma.D = dist_;
    std::memset(srv_cmd_, 0, sizeof(ServerCmd));
    srv_done_->w[0] = 0; srv_done_->w[1] = 0;
    ++srv_seq_;  // sequence numbers never repeat, so a stale command block can never be taken for a new command
    k_merge_server<<<1, 1024, sizeof(SmallStage), st_srv_>>>(ma, srv_cmd_, srv_done_, srv_seq_);
    CK(cudaGetLastError());
    launches_++; srv_starts_++;
    srv_cmds_since_start_ = 0;
    srv_alive_ = true; srv_busy_ = false;
    live_servers().fetch_add(1);
    return 0;
  }
  // merge servers resident on the device on behalf of trainers of this process: each holds one SM that a cooperative launch of
  // any of those trainers must not count on
  static std::atomic<int>& live_servers() { static std::atomic<int> n{0}; return n; }
  void srv_write(uint32_t op, int32_t a, int32_t b, int32_t n, uint32_t serial, uint32_t lenA, uint32_t lenB, uint32_t list_len, uint32_t tag, bool timed) {
    volatile ull* w = srv_cmd_->w;
    const ull seq = (srv_seq_ & 0xFFFFFFFFull) << 32;
    const ull d0[4] = {static_cast<uint32_t>(a) | (static_cast<ull>(static_cast<uint32_t>(b)) << 32), static_cast<uint32_t>(n) | (static_cast<ull>(serial) << 32),
                       list_len | (static_cast<ull>(tag) << 32), timed ? 1ull : 0ull};
    const ull d1[4] = {op, lenA, lenB, 0};
    for (int q = 0; q < 4; q++) { w[2 * q] = d0[q]; __atomic_store_n(&w[2 * q + 1], d1[q] | seq, __ATOMIC_RELEASE); }  // a quarter's tag goes last
  }
  bool srv_exited() const { return static_cast<uint32_t>(__atomic_load_n(&srv_done_->w[0], __ATOMIC_ACQUIRE) >> 32) == 2u; }
  // wait until the server has finished phase 3 of the last merge it was given
  int srv_quiesce() {
    if (!srv_alive_ || !srv_busy_) return 0;
    const double t0 = now_ms();
    uint64_t spins = 0;
    for (;;) {
      const ull d = __atomic_load_n(&srv_done_->w[0], __ATOMIC_ACQUIRE);
      if (static_cast<uint32_t>(d) == srv_tag_ && static_cast<uint32_t>(d >> 32) == 1u) break;
      if (static_cast<uint32_t>(d >> 32) == 2u) break;  // it has left: everything it did is complete
      if ((++spins & 0xFFFF) == 0 && now_ms() - t0 > 30000.0) { std::fprintf(stderr, "[ERROR]\t merge server does not answer\n"); return -1; }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    srv_busy_ = false;
    return 0;
  }
  int srv_stop() {
    if (!srv_alive_) return 0;
    int rc = srv_quiesce();
    if (!srv_exited()) srv_write(SRV_OP_QUIT, 0, 0, 0, 0, 0, 0, 0, 0, false);
    if (cudaStreamSynchronize(st_srv_) != cudaSuccess) rc = -1;  // the kernel has left (it also leaves by itself after SERVER_IDLE_NS)
    srv_alive_ = false; srv_busy_ = false;
    live_servers().fetch_sub(1);
    return rc;
  }
  void begin_merges() override {}
  void end_merges() override { srv_stop(); }

  // --------------------------------------------------------------------------------------------------------- merge
  int merge(int32_t a, int32_t b, int32_t new_id, uint32_t serial, uint32_t list_len, const Rec** recs, size_t* n, uint64_t* occurrences) override {
    *recs = out_.data(); *n = 0; *occurrences = 0;
    const double tm0 = now_ms();
    CK(cudaSetDevice(dev_));  // the caller's thread may have another current device
    if (a < 0 || b < 0 || new_id < 0 || static_cast<size_t>(std::max(a, b)) >= tok_len_.size() || serial == REC_NO_SERIAL) {
      std::fprintf(stderr, "[ERROR]\t merge of an unknown token or pair (%d,%d)\n", a, b);
      return -1;
    }
    fresh_ = false;
    if (static_cast<size_t>(new_id) >= tok_len_.size()) tok_len_.resize(static_cast<size_t>(new_id) + 1, 1u);
    const uint32_t lenA = tok_len_[a], lenB = tok_len_[b];
    tok_len_[new_id] = lenA + lenB;
    // keep the pair table at most half full even if this merge creates every key it can (4 per distinct id)
    const uint64_t worst_new = 4ull * (static_cast<uint64_t>(new_id) + 2);
    if ((pt_n_ + worst_new) * 2 > pt_.cap) { RC(srv_stop()); RC(grow_pt(pt_n_ + worst_new)); }   // buffers move: the server holds their addresses
    if (worst_new * 2 > dt_.cap) { RC(srv_stop()); RC(alloc_dt(next_pow2(worst_new * 2))); }
    if (list_len > sc_.cap) { RC(srv_stop()); RC(ensure_scratch(list_len)); }
    const bool timed = timing_every_ > 0 && (merge_seq_++ % timing_every_) == 0;
    const bool profiled = !profile_merges_.empty() && std::find(profile_merges_.begin(), profile_merges_.end(), merge_no_) != profile_merges_.end();
    if (profiled) { cudaStreamSynchronize(st_); cudaProfilerStart(); }
    ++merge_no_;
    // one CTA with shared-memory tables for short lists (kernels_merge.cuh k_merge_small); the general kernel otherwise, and
    // again if the small one had to give up (ERR_RETRY: it has not changed anything then)
    bool small = world_ == 1 && force_grid_ <= 0 && list_len <= small_max_ && ids_cap_ < (1ull << 30);
    int grid = 1;
    double srv_ms = 0;
    for (;;) {
      const double tl0 = now_ms();
      ++flag_;
      if (small && srv_enabled_) {  // hand the merge to the resident CTA: no launch
        if (srv_alive_ && srv_exited()) { cudaStreamSynchronize(st_srv_); srv_alive_ = false; srv_busy_ = false; live_servers().fetch_sub(1); }  // it left after an idle period
        if (general_pending_) { CK(cudaEventSynchronize(ev_gen_)); general_pending_ = false; }  // the last cooperative merge kernel (its phase 3) has finished
        RC(srv_start());
        RC(srv_quiesce());  // one command at a time
        srv_write(SRV_OP_MERGE, a, b, new_id, serial, lenA, lenB, list_len, static_cast<uint32_t>(flag_), timed);
        srv_tag_ = static_cast<uint32_t>(flag_); srv_busy_ = true;
        launch_ms_ += now_ms() - tl0;
        const int wrc = wait_flag();
        if (wrc == 2) {  // the server left just before the command: start it again
          cudaStreamSynchronize(st_srv_); srv_alive_ = false; srv_busy_ = false; live_servers().fetch_sub(1);
          // A server that leaves before its FIRST command, twice in a row, never ran beside the host: kernel launches are being
          // serialised (ncu, compute-sanitizer, cuda-gdb, CUDA_LAUNCH_BLOCKING=1).  Short merges become launches for this trainer.
          if (srv_cmds_since_start_ == 0 && ++srv_stillborn_ >= 2) {
            srv_enabled_ = false;
            std::fprintf(stderr, "[WARN]\t the resident merge server cannot run beside the host here (serialised kernel launches?): using one launch per merge\n");
          }
          continue;
        }
        if (wrc != 0) return wrc;
        ++srv_cmds_since_start_; srv_stillborn_ = 0;
        ++srv_seq_;
        srv_ms = now_ms() - tl0;
        srv_merges_++;
        if (cv_.err & ERR_RETRY) { small = false; small_retries_++; continue; }
        single_launches_++;
        break;
      }
      RC(srv_quiesce());  // a kernel launch on the same data: the server must have finished its rewrite
      if (!small) ++pass_;
      if (timed) CK(cudaEventRecord(ev0_, st_));
      grid = small ? 1 : merge_grid(list_len);
      MergeArgs ma;
      ma.ids = ids_; ma.ids_cap = ids_cap_; ma.wid = wid_; ma.wcnt = wcnt_; ma.pool = pool_; ma.pool_cap = pool_cap_; ma.sc = sc_;
      ma.A = a; ma.B = b; ma.N = new_id; ma.lenA = lenA; ma.lenB = lenB; ma.serial = serial; ma.par = pass_ & 1u;
      ma.P = P_; ma.dt = dt_; ma.dt.n = &ctr_->dt_n[pass_ & 1u]; ma.pt = pt_; ma.ctr = ctr_;
      ma.recs = recs_; ma.rec_cap = rec_cap_; ma.ctrl = ctrl_; ma.tag = static_cast<uint32_t>(flag_);
      ma.bar_base = bar_count_;  // barrier counter before this launch; it only grows (wraps mod 2^32)
      ma.seq_base = world_ > 1 ? seq_base(rank_) : 0ull;
      ma.dbg = timed ? dbg_ : nullptr;
      ma.D = dist_;
      if (small) {
        const uint32_t nt = std::min<uint32_t>(1024u, std::max<uint32_t>(64u, (list_len + 31u) & ~31u));  // one list entry per thread, up to four beyond 1024
        k_merge_small<<<1, nt, sizeof(SmallStage), st_>>>(ma, small_slots_for(list_len));
        CK(cudaGetLastError());
      } else {
        if (world_ > 1) ma.D = next_exchange();
        bar_count_ += (world_ > 1 ? 3u : 2u) * static_cast<uint32_t>(grid);
        if (virtual_) RC(VirtualCluster::get().collective(0, [&](VirtualCluster& vc) { vc.m[rank_] = ma; }));
        else {
          void* args[] = {&ma};
          const void* kfn = world_ > 1 ? reinterpret_cast<const void*>(k_merge<true>) : reinterpret_cast<const void*>(k_merge<false>);
          CK(cudaLaunchCooperativeKernel(kfn, dim3(grid), dim3(256), args, 0, st_));
          if (srv_enabled_) { CK(cudaEventRecord(ev_gen_, st_)); general_pending_ = true; }
        }
      }
      if (timed) CK(cudaEventRecord(ev1_, st_));
      launches_ += 1;
      launch_ms_ += now_ms() - tl0;
      RC(wait_flag());
      if (small && (cv_.err & ERR_RETRY)) { small = false; small_retries_++; continue; }
      if (small) single_launches_++;
      break;
    }
    if (profiled) {
      cudaStreamSynchronize(st_); cudaProfilerStop();
      std::fprintf(stderr, "[PROFILE]\t merge %u pair (%d,%d): live slots %llu (algorithmic %llu bytes), list entries %llu, occurrences %llu, keys %llu, grid %d\n", merge_no_ - 1, a, b,
                   static_cast<ull>(n_live_), 4ull * n_live_, static_cast<ull>(cv_.list_len), static_cast<ull>(cv_.occ), static_cast<ull>(cv_.n_keys), grid);
    }
    if (cv_.err) { std::fprintf(stderr, "[ERROR]\t device merge pass failed (err=%u)\n", cv_.err); return -1; }
    if (timed) {
      float ms = 0;
      if (srv_ms > 0) {  // handled by the resident server: command written -> control block seen, on the host clock; phase timers valid once phase 3 is done
        ms = static_cast<float>(srv_ms);
        RC(srv_quiesce());
      } else {
        CK(cudaEventSynchronize(ev1_));
        cudaEventElapsedTime(&ms, ev0_, ev1_);
      }
      // algorithmic bytes of the scan formulation: 4 B x (live symbols + unique words) -- n_live_ counts both
      const double algo = 4.0 * static_cast<double>(n_live_), touched = 36.0 * static_cast<double>(cv_.list_len) + 88.0 * static_cast<double>(cv_.occ_local);
      es_.scan_launches++; es_.scan_device_ms += ms; es_.scan_bytes += algo; es_.scan_bytes_touched += touched;
      const double p1 = (dbg_[1] - dbg_[0]) * 1e-6, p2 = (dbg_[2] - dbg_[1]) * 1e-6, p3 = (dbg_[3] - dbg_[2]) * 1e-6;  // ms: probe+emit+barrier | fold+publish | rewrite (CTA 0)
      es_.scan_phase_ms += p1; es_.fold_phase_ms += p2; es_.rewrite_phase_ms += p3;
      if (cv_.list_len >= DENSE_LIST) { es_.dense_launches++; es_.dense_device_ms += ms; es_.dense_bytes += algo; es_.dense_phase_ms += p1; }
      dbg_acc_[0] += p1; dbg_acc_[1] += p2; dbg_acc_[2] += p3; dbg_acc_[3] += ms; dbg_n_++;
      if (small) { small_acc_[0] += p1; small_acc_[1] += p2; small_acc_[2] += ms; small_n_++; }
      if (dbg_print_ && (dbg_n_ % 500) == 0)
        std::fprintf(stderr, "[KTIME]\t %llu timed merges: probe+emit+barrier %.1f us, fold+publish %.1f us, rewrite %.1f us | kernel (events) %.1f us (averages)\n",
                     (unsigned long long)dbg_n_, 1e3 * dbg_acc_[0] / dbg_n_, 1e3 * dbg_acc_[1] / dbg_n_, 1e3 * dbg_acc_[2] / dbg_n_, 1e3 * dbg_acc_[3] / dbg_n_);
    }
    list_entries_total_ += cv_.list_len;
    const double tf0 = dbg_print_ ? now_ms() : 0.0;
    RC(fetch_records());
    if (dbg_print_) fetch_ms_ += now_ms() - tf0;
    *recs = out_.data();
    *n = out_.size(); *occurrences = cv_.occ;
    pt_n_ = cv_.pt_n;
    n_live_ -= cv_.occ_local;
    es_.d2h_bytes += *n * sizeof(Rec) + sizeof(Ctrl);
    merge_ms_ += now_ms() - tm0;
    return 0;
  }

  // ---------------------------------------------------------------------------------------------------------- save
  int token_freqs(uint64_t* freq, size_t T) override {
    CK(cudaSetDevice(dev_));
    RC(srv_stop());
    if (!loaded_ || (!n_words_ && world_ == 1) || !T) return 0;
    ull* d = nullptr;
    CK(cudaMallocAsync(reinterpret_cast<void**>(&d), T * 8, st_));
    CK(cudaMemsetAsync(d, 0, T * 8, st_));
    if (n_words_) { k_token_freq<<<grid_for(n_words_, 256), 256, 0, st_>>>(ids_, woff_, wcnt_, n_words_, P_, d, T); launches_++; }
    if (world_ > 1) {
      if (T > INBOX_ENTRIES * 3) { cudaFreeAsync(d, st_); std::fprintf(stderr, "[ERROR]\t vocabulary too large for the exchange buffer\n"); return -1; }
      const int grid = virtual_ ? VirtualCluster::GRID_PER_RANK : n_sm_ * 2;
      SumArgs sa;
      sa.vals = d; sa.T = T; sa.ctr = ctr_; sa.D = next_exchange(); sa.bar_base = bar_count_;
      bar_count_ += 2u * static_cast<uint32_t>(grid);
      if (virtual_) RC(VirtualCluster::get().collective(2, [&](VirtualCluster& vc) { vc.sm[rank_] = sa; }));
      else {
        void* args[] = {&sa};
        CK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_dist_sum_u64), dim3(grid), dim3(256), args, 0, st_));
      }
      launches_++;
    }
    CK(cudaMemcpyAsync(freq, d, T * 8, cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    cudaFreeAsync(d, st_);
    es_.d2h_bytes += T * 8;
    return 0;
  }
  int word_counts(uint64_t* out) override {
    if (world_ > 1) { std::memcpy(out, host_counts_.data(), host_counts_.size() * 8); return 0; }  // global counts, kept from ingest
    if (!n_words_) return 0;
    CK(cudaMemcpyAsync(out, wcnt_, static_cast<uint64_t>(n_words_) * 8, cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    es_.d2h_bytes += static_cast<uint64_t>(n_words_) * 8;
    return 0;
  }
  int get_words(uint64_t* counts, uint64_t* off, int32_t* ids, uint64_t ids_cap) override {
    if (!loaded_) return -1;
    RC(srv_stop());
    CK(cudaStreamSynchronize(st_));
    const uint32_t N = n_words_;
    std::vector<ull> ho(N + 1);
    std::vector<int32_t> hi(n_slots_ + 2);
    if (N) {
      CK(cudaMemcpy(ho.data(), woff_, (static_cast<uint64_t>(N) + 1) * 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hi.data(), ids_, (n_slots_ + 2) * 4, cudaMemcpyDeviceToHost));
      if (counts) CK(cudaMemcpy(counts, wcnt_, static_cast<uint64_t>(N) * 8, cudaMemcpyDeviceToHost));
    }
    uint64_t at = 0;
    for (uint32_t wi = 0; wi < N; wi++) {
      if (off) off[wi] = at;
      if (static_cast<uint32_t>(hi[ho[wi]]) != (HDR_BIT | wi)) return -2;  // layout invariant
      for (uint64_t q = ho[wi] + 1; q < ho[wi + 1]; q = lay::is_skip(hi[q + 1]) ? q + lay::skip_len(hi[q + 1]) : q + 1) {  // layout.hpp next_start
        const int32_t code = hi[q];
        if (code < 0) return -3;  // a token must start here
        if (ids && at < ids_cap) ids[at] = code_to_id(code, P_);
        at++;
      }
    }
    if (off) off[N] = at;
    return 0;
  }
  uint64_t get_pairs(int32_t* ab, uint64_t* freq, uint64_t cap) override {
    if (!pt_.ent) return 0;
    srv_stop();
    cudaStreamSynchronize(st_);
    std::vector<PairEnt> e(pt_.cap);
    if (cudaMemcpy(e.data(), pt_.ent, pt_.cap * sizeof(PairEnt), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    uint64_t n = 0;
    for (uint64_t s = 0; s < pt_.cap; s++) if (e[s].key != PT_EMPTY) {
      if (n < cap) { ab[2 * n] = static_cast<int32_t>(e[s].key >> 32); ab[2 * n + 1] = static_cast<int32_t>(e[s].key & 0xFFFFFFFFu); freq[n] = e[s].freq; }
      n++;
    }
    return n;
  }
  int mark_begin() override { CK(cudaEventRecord(evm0_, st_)); return 0; }
  double mark_end() override {
    if (cudaEventRecord(evm1_, st_) != cudaSuccess || cudaEventSynchronize(evm1_) != cudaSuccess) return -1.0;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, evm0_, evm1_) != cudaSuccess) return -1.0;
    return ms;
  }
  void stats(EngineStats* out) override {
    *out = es_;
    out->n_slots = n_slots_; out->n_symbols_live = n_live_ >= n_words_ ? n_live_ - n_words_ : 0; out->pair_entries = pt_n_;
    out->kernel_launches = launches_; out->wait_ms = wait_ms_; out->launch_ms = launch_ms_; out->merge_ms = merge_ms_;
    out->list_entries = list_entries_total_; out->pool_entries = cv_.pool_top; out->single_launches = single_launches_; out->server_merges = srv_merges_; out->server_starts = srv_starts_;
  }
  const char* name() override { return name_; }

 private:
  int grid_for(uint64_t n, int block) const {
    uint64_t g = (n + block - 1) / block, maxg = static_cast<uint64_t>(n_sm_) * 16;
    if (g < 1) g = 1;
    return static_cast<int>(g < maxg ? g : maxg);
  }
  // grid of a merge launch: one thread per list entry up to the co-resident maximum (cooperative launch); most merges of a
  // run have a few thousand entries and get a handful of CTAs, which keeps the two grid barriers short
  static constexpr uint32_t DENSE_LIST = 1u << 16;
  int merge_grid(uint32_t list_len) const {
    if (force_grid_ > 0) return force_grid_;
    const uint64_t ctas = (static_cast<uint64_t>(list_len) + 255) / 256, maxg = static_cast<uint64_t>(std::max(1, n_sm_ - std::max(live_servers().load(), srv_enabled_ ? 1 : 0))) * merge_ctas_per_sm_;  // SMs held by merge servers are not ours
    return static_cast<int>(ctas < 1 ? 1 : (ctas < maxg ? ctas : maxg));
  }

  // Spins on the control block until its first 16-byte block carries the current pass tag, then reads the other blocks the same
  // way (common.cuh: every block is self-validating, the device issues no fence towards the host).
  int wait_flag() {
    const double t0 = now_ms();
    const uint32_t tag = static_cast<uint32_t>(flag_);
    const volatile ull* w = ctrl_->w;
    uint64_t spins = 0;
    auto fresh = [&](int blk) { return static_cast<uint32_t>(__atomic_load_n(&w[2 * blk], __ATOMIC_ACQUIRE)) == tag; };
    for (int blk = 0; blk < 4; blk++) {
      while (!fresh(blk)) {
        if ((++spins & 0x3FFF) == 0) {
          if (srv_busy_) {  // the pass was handed to the merge server
            if (srv_exited()) {  // it left (idle period) -- before or after taking the command?
              if (fresh(blk)) continue;
              if (static_cast<uint32_t>(__atomic_load_n(&srv_done_->w[1], __ATOMIC_ACQUIRE)) == static_cast<uint32_t>(srv_seq_)) return 2;  // it left still waiting for this command's number: it never saw the command
            }
            if (now_ms() - t0 > 120000.0) { std::fprintf(stderr, "[ERROR]\t merge server timed out\n"); return -1; }
            continue;
          }
          cudaError_t q = cudaStreamQuery(st_);
          if (q == cudaSuccess) { if (fresh(blk)) break; std::fprintf(stderr, "[ERROR]\t device pass finished without publishing its result\n"); return -1; }
          if (q != cudaErrorNotReady) { std::fprintf(stderr, "[ERROR]\t CUDA: %s\n", cudaGetErrorString(q)); return -1; }
          if (now_ms() - t0 > 120000.0) { std::fprintf(stderr, "[ERROR]\t device pass timed out\n"); return -1; }
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
      }
    }
    cv_.n_recs = static_cast<uint32_t>(w[0] >> 32); cv_.err = static_cast<uint32_t>(w[1]); cv_.list_len = static_cast<uint32_t>(w[1] >> 32);
    cv_.occ_local = static_cast<uint32_t>(w[2] >> 32); cv_.occ = w[3];
    cv_.n_keys = static_cast<uint32_t>(w[4] >> 32); cv_.pt_n = w[5]; cv_.pool_top = w[7];
    wait_ms_ += now_ms() - t0;
    return 0;
  }
  // Decodes the records of the pass just published into out_ (engine.hpp Rec), waiting for any that is still in flight.
  int fetch_records() {
    const uint32_t tag = static_cast<uint32_t>(flag_), n = cv_.n_recs;
    const ull lo = static_cast<ull>(tag & 0xFFFFu), hi = static_cast<ull>(tag >> 16);
    out_.resize(n);
    double t0 = 0;
    for (uint32_t i = 0; i < n; i++) {
      const volatile ull* w = recs_[i].w;
      uint64_t spins = 0;
      ull w1, w2;
      for (;;) {
        w1 = __atomic_load_n(&w[1], __ATOMIC_ACQUIRE); w2 = __atomic_load_n(&w[2], __ATOMIC_ACQUIRE);
        if ((w1 >> 48) == lo && (w2 >> 48) == hi) break;
        if ((++spins & 0xFFFF) == 0) {
          if (t0 == 0) t0 = now_ms();
          if (now_ms() - t0 > 20000.0) { std::fprintf(stderr, "[ERROR]\t record %u of %u never arrived\n", i, n); return -1; }
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
      }
      Rec& r = out_[i];
      r.key = w[0];
      const uint32_t kind = static_cast<uint32_t>(w[3]);
      r.val = w1 & MASK48;
      if (rec_kind(kind) == REC_PHANTOM && (r.val >> 47)) r.val |= ~MASK48;  // a phantom's value is a signed net delta
      r.seq = w2 & MASK48;
      r.kind = kind; r.serial = static_cast<uint32_t>(w[3] >> 32);
    }
    return 0;
  }

  void release_corpus() {
    if (ids_) cudaFreeAsync(ids_, st_); ids_ = nullptr;
    if (wid_) cudaFreeAsync(wid_, st_); wid_ = nullptr;
    if (woff_) cudaFreeAsync(woff_, st_); woff_ = nullptr;
    if (wcnt_) cudaFreeAsync(wcnt_, st_); wcnt_ = nullptr;
    if (pool_) cudaFreeAsync(pool_, st_); pool_ = nullptr;
    pool_cap_ = 0;
    n_words_ = 0; n_slots_ = n_live_ = 0; loaded_ = false; pt_n_ = 0;
  }
  void release_all() {
    cudaSetDevice(dev_);
    srv_stop();
    if (dbg_print_ && dbg_n_) {
      if (small_n_) std::fprintf(stderr, "[KTIME]\t of which %llu k_merge_small launches: probe+deltas %.1f us, fold+publish %.1f us | kernel (events) %.1f us\n", (unsigned long long)small_n_,
                                 1e3 * small_acc_[0] / small_n_, 1e3 * small_acc_[1] / small_n_, 1e3 * small_acc_[2] / small_n_);
      std::fprintf(stderr, "[KTIME]\t decoding the records of all merges took the host %.1f ms\n", fetch_ms_);
      std::fprintf(stderr, "[KTIME]\t %llu timed merges: probe+emit+barrier %.1f us, fold+publish %.1f us, rewrite %.1f us | kernel (events) %.1f us (averages)\n",
                   (unsigned long long)dbg_n_, 1e3 * dbg_acc_[0] / dbg_n_, 1e3 * dbg_acc_[1] / dbg_n_, 1e3 * dbg_acc_[2] / dbg_n_, 1e3 * dbg_acc_[3] / dbg_n_);
    }
    const double tr0 = now_ms();
    release_corpus();
    if (dt_.keys) {
      cudaFreeAsync(dt_.keys, st_); cudaFreeAsync(dt_.delta, st_); cudaFreeAsync(dt_.seq, st_); cudaFreeAsync(dt_.nocc, st_); cudaFreeAsync(dt_.base, st_);
      cudaFreeAsync(dt_.list, st_); cudaFreeAsync(dt_.klist, st_);
    }
    if (pt_.ent) { cudaFreeAsync(pt_.ent, st_); cudaFreeAsync(pt_.lists, st_); }
    if (sc_.a) { cudaFreeAsync(sc_.a, st_); cudaFreeAsync(sc_.b, st_); }
    if (ctr_) cudaFreeAsync(ctr_, st_);
    const double tr1 = now_ms();
    if (st_) cudaStreamSynchronize(st_);
    const double tr2 = now_ms();
    if (dbg_print_) std::fprintf(stderr, "[TIMING]\t engine release: frees %.1f ms, stream sync %.1f ms\n", tr1 - tr0, tr2 - tr1);
    if (st_copy_) { cudaStreamDestroy(st_copy_); st_copy_ = nullptr; }
    dist_teardown();
    if (st_srv_) { cudaStreamSynchronize(st_srv_); cudaStreamDestroy(st_srv_); }
    // nothing of this trainer can write to its pinned blocks any more: hand them to the next one
    PinnedCache::get().release(recs_, static_cast<size_t>(rec_cap_) * sizeof(WireRec)); recs_ = nullptr;
    PinnedCache::get().release(ctrl_, sizeof(Ctrl)); ctrl_ = nullptr;
    PinnedCache::get().release(dbg_, 32 * sizeof(ull)); dbg_ = nullptr;
    PinnedCache::get().release(srv_cmd_, sizeof(ServerCmd) + 64); srv_cmd_ = nullptr;
    if (ev_srv_) cudaEventDestroy(ev_srv_);
    if (ev_gen_) cudaEventDestroy(ev_gen_);
    if (ev0_) cudaEventDestroy(ev0_);
    if (ev1_) cudaEventDestroy(ev1_);
    if (evm0_) cudaEventDestroy(evm0_);
    if (evm1_) cudaEventDestroy(evm1_);
    if (st_ && !virtual_) cudaStreamDestroy(st_);
  }

  int dev_, n_sm_;
  char name_[320];
  cudaStream_t st_ = nullptr, st_copy_ = nullptr;
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr, evm0_ = nullptr, evm1_ = nullptr;
  EngineConfig cfg_{};
  Params P_{};
  bool loaded_ = false, fresh_ = false;
  uint8_t keep_[256] = {};
  uint32_t n_words_ = 0;
  uint64_t n_slots_ = 0, n_live_ = 0, ids_cap_ = 0;
  int32_t* ids_ = nullptr;
  uint32_t* wid_ = nullptr;
  ull* woff_ = nullptr;
  ull* wcnt_ = nullptr;
  PoolEnt* pool_ = nullptr;
  uint64_t pool_cap_ = 0;
  OccScratch sc_{};
  std::vector<uint32_t> tok_len_ = std::vector<uint32_t>(256, 1u);  // bytes covered by each token id (span length in slots)
  uint32_t merge_no_ = 0, bar_count_ = 0, pass_ = 0;
  int rank_ = 0, world_ = 1;
  bool virtual_ = false;
  DistArgs dist_{};
  uint8_t* inbox_ = nullptr;
  std::vector<uint64_t> host_counts_;
  std::string rdv_prefix_;
  uint64_t list_entries_total_ = 0;
  DeltaTable dt_{};
  PairTable pt_{};
  uint64_t pt_n_ = 0;
  WireRec* recs_ = nullptr;   // mapped pinned host memory the device writes its records into (common.cuh wire format)
  uint32_t rec_cap_ = 0;
  std::vector<Rec> out_;      // the records of the last pass, decoded for the caller
  Ctrl* ctrl_ = nullptr;
  struct CtrlView { uint32_t n_recs = 0, err = 0, list_len = 0, occ_local = 0, n_keys = 0; uint64_t occ = 0, pt_n = 0, pool_top = 0; } cv_;
  ServerCmd* srv_cmd_ = nullptr;
  ServerDone* srv_done_ = nullptr;
  cudaStream_t st_srv_ = nullptr;
  cudaEvent_t ev_srv_ = nullptr, ev_gen_ = nullptr;
  bool srv_enabled_ = true, srv_alive_ = false, srv_busy_ = false, general_pending_ = false;
  uint64_t srv_cmds_since_start_ = 0;  // commands the running server has taken
  int srv_stillborn_ = 0;              // consecutive servers that left without taking one
  uint64_t srv_seq_ = 0, srv_merges_ = 0, srv_starts_ = 0;
  uint32_t srv_tag_ = 0;
  uint32_t small_max_ = SMALL_MAX;
  uint64_t single_launches_ = 0, small_retries_ = 0;
  DevCounters* ctr_ = nullptr;
  uint64_t flag_ = 0;
  uint64_t vocab_hint_ = 32768;
  EngineStats es_{};
  uint64_t launches_ = 0, merge_seq_ = 0;
  double wait_ms_ = 0, launch_ms_ = 0, merge_ms_ = 0, fetch_ms_ = 0;
  int timing_every_ = 0;
  ull* dbg_ = nullptr;
  bool dbg_print_ = false;
  std::vector<uint32_t> profile_merges_;
  double dbg_acc_[5] = {0, 0, 0, 0, 0};
  double small_acc_[3] = {0, 0, 0};
  uint64_t small_n_ = 0;
  uint64_t dbg_n_ = 0;
  int merge_ctas_per_sm_ = 4, force_grid_ = 0;
};

char g_devname[320] = "no CUDA device";

}  // namespace

Engine* make_device_engine() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    std::fprintf(stderr, "[ERROR]\t CUDA: no usable device (%s)\n", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return nullptr;
  }
  int dev = 0;
  if (const char* s = std::getenv("SHRED_DEVICE")) dev = std::atoi(s);
  else if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  if (dev < 0 || dev >= n) dev = 0;
  // cudaGetDeviceProperties takes tens of milliseconds (it queries clocks and the board): once per device and process
  static std::mutex prop_mu;
  static cudaDeviceProp props[64];
  static bool have[64] = {false};
  cudaDeviceProp prop;
  {
    std::lock_guard<std::mutex> g(prop_mu);
    if (dev >= 64 || !have[dev]) {
      if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return nullptr;
      if (dev < 64) { props[dev] = prop; have[dev] = true; }
    } else prop = props[dev];
  }
  if (prop.major != 10) {
    std::fprintf(stderr, "[ERROR]\t CUDA: device %d (%s, sm_%d%d) is not a Blackwell sm_100 part; this library carries sm_100a code only\n", dev, prop.name,
                 prop.major, prop.minor);
    return nullptr;
  }
  CudaEngine* eng = new CudaEngine(dev, prop);
  if (eng->init() != 0) { delete eng; return nullptr; }
  std::snprintf(g_devname, sizeof g_devname, "%s", eng->name());
  return eng;
}

}  // namespace shred

extern "C" const char* bpe_b200_device_name(void) {
  static char buf[320];
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return "no CUDA device";
  int dev = 0; cudaGetDevice(&dev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return "no CUDA device";
  std::snprintf(buf, sizeof buf, "%s sm_%d%d %d SMs", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  return buf;
}
