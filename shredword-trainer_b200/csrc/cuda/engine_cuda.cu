// engine_cuda.cu -- the B200 (sm_100a) device engine of the BPE trainer: corpus ingest, pair counting, merge
// application and pair-table maintenance as hand-written CUDA kernels.  Implements shred::Engine (../engine.hpp).
// This file holds the host side (buffers, launches, multi-GPU rendezvous); the kernels live in the .cuh fragments
// included below.
//
// Data layout in HBM
//   ids[]   int32, one flat array holding every unique word back to back in reference word order (A3):
//             [HDR|wi] s0 s1 ... s(len-1) [DEAD ...]          HDR|wi < -1, symbols >= 0, DEAD == -1
//           A word keeps its slot range between compactions; merges left-pack its live symbols and fill the tail with
//           DEAD.  Because words are stored in reference scan order, "flat position" is monotone in the reference's
//           (word index, position) order and serves as the sequence number the host needs (Appendix A14).
//   wid[]   uint32 word index per slot; wcnt[] uint64 word counts, woff[] uint64 slot offsets (N+1), wlen[] uint32 lengths
//   planes  tile occurrence index (one bit plane per token id; a merge scans only tiles holding both tokens)
//   pair table   open addressing, uint64 key (first<<32|second) -> uint64 freq + serial  (reference BIMap, hash.cpp:104-130)
//   delta table  open addressing scratch, key -> (sum of +/-count, min sequence)         (reference FreqChangeMap, bpe.cpp:9-38)
//
// Kernels (reference loop each one replaces)                                               file
//   k_tokenize ........ bpe.cpp:131-153 + hash.cpp:29-53   tokenise, unique-word table     kernels_tokenize.cuh
//   k_hist_words ...... histogram.cpp:30-36                unweighted byte histogram        kernels_ingest.cuh
//   k_scatter/k_sort_buckets  hash.cpp:61-72               word order (djb2 & 4095, first)  kernels_ingest.cuh
//   k_symbolize ....... histogram.cpp:7-27                 bytes -> ids, unk substitution   kernels_ingest.cuh
//   k_count ........... bpe.cpp:187-218                    adjacent pair counts             kernels_count.cuh
//   k_finalize_count .. bpe.cpp:219-227                    fold counts, seed records        kernels_fold.cuh
//   k_merge ........... bpe.cpp:265-318                    one cooperative launch per merge kernels_merge.cuh
//                       scan of the candidate tiles with per-occurrence count deltas | grid barrier | deltas folded
//                       into the pair table + records published | in-place rewrite of the touched words
//   exchange_deltas ... (multi-GPU) per-merge delta exchange over NVLink peer memory        kernels_dist.cuh
//   k_token_freq ...... bpe.cpp:409-415                    final token frequencies          kernels_merge.cuh
#include <cuda_profiler_api.h>
#include <cuda_runtime.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
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

class CudaEngine : public Engine {
 public:
  CudaEngine(int dev, const cudaDeviceProp& prop) : dev_(dev), n_sm_(prop.multiProcessorCount > 0 ? prop.multiProcessorCount : N_SM_FALLBACK) {
    std::snprintf(name_, sizeof name_, "%s sm_%d%d %d SMs", prop.name, prop.major, prop.minor, n_sm_);
  }
  ~CudaEngine() override { release_all(); }

  int init() {
    CK(cudaSetDevice(dev_));
    CK(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
    void* cp = nullptr;
    CK(cudaHostAlloc(&cp, sizeof(Ctrl), cudaHostAllocMapped));
    std::memset(cp, 0, sizeof(Ctrl));
    ctrl_ = static_cast<volatile Ctrl*>(cp);
    CK(cudaMallocAsync(reinterpret_cast<void**>(&ctr_), sizeof(DevCounters), st_));
    CK(cudaMemset(ctr_, 0, sizeof(DevCounters)));
    CK(cudaEventCreate(&ev0_));
    CK(cudaEventCreate(&ev1_));
    CK(cudaEventCreate(&evm0_));
    CK(cudaEventCreate(&evm1_));
    cudaMemPool_t pool;  // stream-ordered allocations; keep freed blocks cached so repeated loads do not pay cudaMalloc
    if (cudaDeviceGetDefaultMemPool(&pool, dev_) == cudaSuccess) { uint64_t thr = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr); }
    {  // in-kernel phase timestamps (%globaltimer) of the timed launches
      void* dp = nullptr;
      CK(cudaHostAlloc(&dp, 8 * sizeof(ull), cudaHostAllocMapped));
      dbg_ = static_cast<ull*>(dp);
      std::memset(dp, 0, 8 * sizeof(ull));
      const char* d = std::getenv("SHRED_DEBUG_TIMING");
      dbg_print_ = d && *d && *d != '0';
    }
    CK(cudaFuncSetAttribute(k_sort_buckets, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(SORT_CAP * sizeof(ull))));
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_merge<4, false>, 256, 0) == cudaSuccess && nb > 0) scan_ctas_per_sm_ = nb;
    if (const char* w = std::getenv("SHRED_WORLD")) world_ = std::atoi(w);
    if (world_ > 1) {
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_merge<4, true>, 256, 0) == cudaSuccess && nb > 0 && nb < scan_ctas_per_sm_) scan_ctas_per_sm_ = nb;
      const char* r = std::getenv("SHRED_RANK");
      rank_ = r ? std::atoi(r) : 0;
      if (world_ > MAX_RANKS || rank_ < 0 || rank_ >= world_) { std::fprintf(stderr, "[ERROR]\t bad SHRED_RANK/SHRED_WORLD (%d/%d, at most %d ranks)\n", rank_, world_, MAX_RANKS); return -1; }
      RC(dist_setup());
    } else { world_ = 1; rank_ = 0; }
    if (const char* pl = std::getenv("SHRED_PLAIN_LAUNCH")) plain_launch_ = *pl && *pl != '0';
    const char* e = std::getenv("SHRED_TIMING");
    timing_every_ = e && *e ? std::atoi(e) : 0;
    if (const char* ho = std::getenv("SHRED_HOT_ON_OCC")) hot_on_occ_ = std::strtoull(ho, nullptr, 10);
    if (const char* pm = std::getenv("SHRED_PROFILE_MERGES")) {  // "0,1,2000": cudaProfilerStart/Stop around these merges (ncu --profile-from-start off)
      for (const char* q = pm; *q;) { char* end = nullptr; const unsigned long v = std::strtoul(q, &end, 10); if (end == q) break; profile_merges_.push_back(static_cast<uint32_t>(v)); q = *end ? end + 1 : end; }
    }
    return 0;
  }

  // ------------------------------------------------------------------------------------------------- multi-GPU setup
  // One process per GPU.  Every rank allocates its inbox with cudaMalloc, exports it with CUDA IPC through a small file
  // in the rendezvous directory SHRED_RDV (shared by the ranks of one job) and maps every peer's inbox.
  int dist_setup() {
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
    CK(cudaMallocAsync(reinterpret_cast<void**>(&d_text), padded, st_));
    double t0 = now_ms();
    if (n && text) CK(cudaMemcpyAsync(d_text, text, n, cudaMemcpyHostToDevice, st_));
    if (n && !text && stream_file(fd, n, d_text) != 0) { cudaFreeAsync(d_text, st_); return -1; }
    CK(cudaMemsetAsync(d_text + n, ' ', padded - n, st_));
    CK(cudaStreamSynchronize(st_));
    es_.h2d_ms += now_ms() - t0; es_.h2d_bytes += n;

    CK(cudaEventRecord(ev0_, st_));
    int rc = ingest(d_text, n, info);
    cudaFreeAsync(d_text, st_);
    if (rc != 0) return rc;
    CK(cudaEventRecord(ev1_, st_));
    CK(cudaStreamSynchronize(st_));
    float ms = 0; cudaEventElapsedTime(&ms, ev0_, ev1_);
    es_.ingest_device_ms = ms;
    es_.ingest_bytes = static_cast<double>(n) + 4.0 * info->n_symbols + 12.0 * info->n_words;
    return 0;
  }

  // The tile index is the one GB-sized buffer whose size is close to the corpus text's: handing it back to the stream-ordered
  // pool lets the pool give it to the next trainer's text buffer and then grow again for the next index (a ~1 s stall when
  // it happens).  One spare per process is parked here instead.
  struct PlaneCache { uint32_t* ptr = nullptr; uint64_t bytes = 0; int dev = -1; std::mutex mu; };
  static PlaneCache& plane_cache() { static PlaneCache c; return c; }
  int acquire_planes(uint64_t bytes) {
    PlaneCache& c = plane_cache();
    {
      std::lock_guard<std::mutex> hold(c.mu);
      if (c.ptr && c.dev == dev_ && c.bytes >= bytes && c.bytes <= bytes + bytes / 2) { planes_ = c.ptr; planes_bytes_ = c.bytes; c.ptr = nullptr; c.bytes = 0; return 0; }
    }
    CK(cudaMallocAsync(reinterpret_cast<void**>(&planes_), bytes, st_));
    planes_bytes_ = bytes;
    return 0;
  }
  void release_planes() {
    if (!planes_) return;
    cudaStreamSynchronize(st_);  // nothing queued may still touch it
    PlaneCache& c = plane_cache();
    std::lock_guard<std::mutex> hold(c.mu);
    if (!c.ptr) { c.ptr = planes_; c.bytes = planes_bytes_; c.dev = dev_; }
    else cudaFreeAsync(planes_, st_);
    planes_ = nullptr; planes_bytes_ = 0;
  }

  // pinned staging ring shared by all trainers of the process (cudaHostAlloc is slow, so it is done once)
  static constexpr int STAGE_BUFS = 8;
  static constexpr size_t STAGE_BYTES = 16u << 20;
  struct StageRing { uint8_t* buf[STAGE_BUFS] = {}; std::mutex mu; };
  static StageRing& stage_ring() { static StageRing r; return r; }

  int stream_file(int fd, size_t n, uint8_t* d_text) {
    StageRing& ring = stage_ring();
    std::lock_guard<std::mutex> hold(ring.mu);  // one load at a time uses the ring
    for (int b = 0; b < STAGE_BUFS; b++) if (!ring.buf[b]) CK(cudaHostAlloc(reinterpret_cast<void**>(&ring.buf[b]), STAGE_BYTES, cudaHostAllocDefault));
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
      if (cudaMemcpyAsync(d_text + off, ring.buf[b], len, cudaMemcpyHostToDevice, st_) != cudaSuccess || cudaEventRecord(done[b], st_) != cudaSuccess) { failed.store(true); break; }
      state[b].store(2, std::memory_order_release);
    }
    for (auto& t : readers) t.join();
    for (int b = 0; b < STAGE_BUFS; b++) cudaEventDestroy(done[b]);
    if (failed.load()) { std::fprintf(stderr, "[ERROR]\t reading the corpus file failed\n"); return -1; }
    return 0;
  }

  int ingest(const uint8_t* d_text, uint64_t n, LoadInfo* info) {
    DevCounters zero; std::memset(&zero, 0, sizeof zero);
    WordTable wt; std::memset(&wt, 0, sizeof wt);
    uint32_t N = 0; ull n_tokens = 0;
    uint64_t cap = next_pow2(n / 64 + 1); if (cap < (1u << 16)) cap = 1u << 16;  // grown 4x and redone if more than half fills up
    uint32_t seed = 0x5bd1e995u;
    for (int attempt = 0;; ++attempt) {
      if (attempt > 8) { std::fprintf(stderr, "[ERROR]\t unique-word table did not converge\n"); return -1; }
      if (cap > (1ull << 32)) { std::fprintf(stderr, "[ERROR]\t unique-word table too large\n"); return -1; }
      CK(cudaMallocAsync(reinterpret_cast<void**>(&wt.tag), cap * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&wt.first), cap * 8, st_));
      CK(cudaMallocAsync(reinterpret_cast<void**>(&wt.count), cap * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&wt.len), cap * 4, st_));
      CK(cudaMallocAsync(reinterpret_cast<void**>(&wt.bucket), cap * 4, st_));
      wt.cap = cap; wt.mask = cap - 1;
      CK(cudaMemsetAsync(wt.tag, 0, cap * 8, st_)); CK(cudaMemsetAsync(wt.first, 0xFF, cap * 8, st_)); CK(cudaMemsetAsync(wt.count, 0, cap * 8, st_));
      CK(cudaMemcpyAsync(ctr_, &zero, sizeof zero, cudaMemcpyHostToDevice, st_));
      bar_count_ = 0;
      if (n) { k_tokenize<<<grid_for((n + 15) / 16, 256), 256, 0, st_>>>(d_text, n, wt, ctr_, seed); launches_++; es_.ingest_launches++; }
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
    auto free_wt = [&]() { cudaFreeAsync(wt.tag, st_); cudaFreeAsync(wt.first, st_); cudaFreeAsync(wt.count, st_); cudaFreeAsync(wt.len, st_); cudaFreeAsync(wt.bucket, st_); };
    if (N >= 0x7FFFFFF0u) { free_wt(); std::fprintf(stderr, "[ERROR]\t too many unique words\n"); return -1; }
    n_words_ = N;
    info->n_words = N; info->n_tokens = n_tokens;
    // --- reference word order: bucket = djb2 & 4095 ascending, first occurrence ascending inside a bucket
    uint32_t *u_slot = nullptr, *u_n = nullptr, *bcnt = nullptr, *bstart = nullptr, *cursor = nullptr, *tmp_slot = nullptr, *order_slot = nullptr;
    ull *tmp_first = nullptr, *d_hist = nullptr, *len1 = nullptr, *sums = nullptr;
    uint8_t* d_keep = nullptr;
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
      cudaFreeAsync(bstart, st_); cudaFreeAsync(cursor, st_); cudaFreeAsync(d_hist, st_); cudaFreeAsync(d_keep, st_); cudaFreeAsync(sums, st_);
    };
    CK(cudaMemsetAsync(u_n, 0, 4, st_)); CK(cudaMemsetAsync(bcnt, 0, 4096 * 4, st_)); CK(cudaMemsetAsync(cursor, 0, 4096 * 4, st_));
    CK(cudaMemsetAsync(d_hist, 0, 256 * 8, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&wcnt_), Na * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&wlen_), Na * 4, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&woff_[0]), (Na + 1) * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&woff_[1]), (Na + 1) * 8, st_));
    ull S1 = 0;  // total slots = symbols + headers
    if (N) {
      k_collect<<<grid_for(cap, 256), 256, 0, st_>>>(wt, u_slot, u_n, bcnt);
      k_scan4096<<<1, 1024, 0, st_>>>(bcnt, bstart);
      k_scatter<<<grid_for(N, 256), 256, 0, st_>>>(wt, u_slot, N, bstart, cursor, tmp_slot, tmp_first);
      k_sort_buckets<<<4096, SORT_THREADS, SORT_CAP * sizeof(ull), st_>>>(tmp_slot, tmp_first, bstart, order_slot);
      k_rank_big<<<grid_for(N, 128), 128, 0, st_>>>(wt, tmp_slot, tmp_first, N, bstart, order_slot); launches_++;  // only buckets above SORT_CAP do work
      k_hist_words<<<grid_for(N, 256), 256, 0, st_>>>(d_text, wt, order_slot, N, d_hist, wcnt_, wlen_, len1);
      k_scan_sums<<<nb_scan, SCAN_THREADS, 0, st_>>>(len1, N, sums);
      k_scan_top<<<1, SCAN_THREADS, 0, st_>>>(sums, nb_scan, sums + nb_scan);
      k_scan_apply<<<nb_scan, SCAN_THREADS, 0, st_>>>(len1, N, sums, woff_[0]);
      launches_ += 8; es_.ingest_launches += 8;
      CK(cudaMemcpyAsync(&S1, sums + nb_scan, 8, cudaMemcpyDeviceToHost, st_));
    }
    CK(cudaMemcpyAsync(info->hist, d_hist, 256 * 8, cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    CK(cudaGetLastError());
    es_.d2h_bytes += 256 * 8 + 8 + sizeof(DevCounters);
    charset_keep(info->hist, cfg_.coverage, info->keep, &info->n_distinct, &info->n_keep);
    info->n_symbols = S1 - N;
    uint32_t lo = 0, n_local = N;
    host_counts_.clear();
    if (world_ > 1 && N) {  // keep only this rank's contiguous range of words (shard.hpp); ingest itself is replicated
      std::vector<ull> hoff(static_cast<size_t>(N) + 1);
      CK(cudaMemcpyAsync(hoff.data(), woff_[0], static_cast<uint64_t>(N) * 8, cudaMemcpyDeviceToHost, st_));
      host_counts_.resize(N);
      CK(cudaMemcpyAsync(host_counts_.data(), wcnt_, static_cast<uint64_t>(N) * 8, cudaMemcpyDeviceToHost, st_));
      CK(cudaStreamSynchronize(st_));
      hoff[N] = S1;
      lo = static_cast<uint32_t>(shard_begin(hoff.data(), N, rank_, world_));
      const uint32_t hi = static_cast<uint32_t>(shard_begin(hoff.data(), N, rank_ + 1, world_));
      n_local = hi - lo;
      const ull base_off = hoff[lo];
      S1 = hoff[hi] - base_off;
      ull *wc = nullptr, *wo0 = nullptr, *wo1 = nullptr; uint32_t* wl = nullptr;
      const uint64_t nl = n_local ? n_local : 1;
      CK(cudaMallocAsync(reinterpret_cast<void**>(&wc), nl * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&wl), nl * 4, st_));
      CK(cudaMallocAsync(reinterpret_cast<void**>(&wo0), (nl + 1) * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&wo1), (nl + 1) * 8, st_));
      if (n_local) {
        CK(cudaMemcpyAsync(wc, wcnt_ + lo, static_cast<uint64_t>(n_local) * 8, cudaMemcpyDeviceToDevice, st_));
        CK(cudaMemcpyAsync(wl, wlen_ + lo, static_cast<uint64_t>(n_local) * 4, cudaMemcpyDeviceToDevice, st_));
        k_rebase<<<grid_for(n_local, 256), 256, 0, st_>>>(woff_[0] + lo, n_local, base_off, wo0); launches_++;
      }
      cudaFreeAsync(wcnt_, st_); cudaFreeAsync(wlen_, st_); cudaFreeAsync(woff_[0], st_); cudaFreeAsync(woff_[1], st_);
      wcnt_ = wc; wlen_ = wl; woff_[0] = wo0; woff_[1] = wo1;
      es_.d2h_bytes += static_cast<uint64_t>(N) * 16;
    }
    n_words_ = n_local;
    n_slots_ = S1; n_live_ = S1;
    if (S1 + 64 >= (1ull << 32)) { free_tmp(); free_wt(); std::fprintf(stderr, "[ERROR]\t corpus needs more than 2^32 symbol slots on one GPU\n"); return -1; }
    ids_cap_ = ((S1 + 8 + 1023) / 1024) * 1024;
    CK(cudaMallocAsync(reinterpret_cast<void**>(&ids_[0]), ids_cap_ * 4, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&ids_[1]), ids_cap_ * 4, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&wl_), ids_cap_ * 4, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&wid_[0]), ids_cap_ * 4, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&wid_[1]), ids_cap_ * 4, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&claimed_), Na * 4, st_));
    CK(cudaMemsetAsync(claimed_, 0, Na * 4, st_));
    CK(cudaMemsetAsync(wid_[0], 0, ids_cap_ * 4, st_)); CK(cudaMemsetAsync(wid_[1], 0, ids_cap_ * 4, st_));
    merge_no_ = 0;
    cur_ = 0;
    CK(cudaMemcpyAsync(d_keep, info->keep, 256, cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(woff_[0] + n_local, &S1, 8, cudaMemcpyHostToDevice, st_));
    if (n_local) { k_symbolize<<<grid_for(n_local, 256), 256, 0, st_>>>(d_text, wt, order_slot + lo, n_local, woff_[0], d_keep, P_.unk_code, ids_[0], wid_[0]); launches_++; es_.ingest_launches++; }
    k_fill_i32<<<grid_for(ids_cap_ - S1, 256), 256, 0, st_>>>(ids_[0], S1, ids_cap_, DEAD); launches_++; es_.ingest_launches++;
    CK(cudaStreamSynchronize(st_));
    CK(cudaGetLastError());
    free_tmp(); free_wt();
    RC(build_planes());
    // --- pair/delta tables sized for this trainer
    RC(alloc_tables());
    loaded_ = true;
    return 0;
  }

  // (re)build the tile occurrence index for the current ids buffer
  int build_planes() {
    if (std::getenv("SHRED_NO_TILE_INDEX")) return 0;
    const double tb0 = now_ms();
    const uint32_t id_cap = static_cast<uint32_t>(256 + vocab_hint_ + 64);
    // finest tiling whose bit planes fit the budget (a tile is at least one warp row = 128 slots)
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = 0;
    uint64_t budget = 6ull << 30;
    if (const char* e = std::getenv("SHRED_TILE_INDEX_MB")) budget = static_cast<uint64_t>(std::atoll(e)) << 20;
    if (budget > (free_b + (planes_ ? static_cast<uint64_t>(id_cap_) * plane_words_ * 4 : 0)) / 3) budget = (free_b + (planes_ ? static_cast<uint64_t>(id_cap_) * plane_words_ * 4 : 0)) / 3;
    uint32_t shift = MIN_TILE_SHIFT, W = 0;
    uint64_t bytes = 0;
    for (;; ++shift) {
      const uint32_t n_tiles_cap = static_cast<uint32_t>((ids_cap_ + (1ull << shift) - 1) >> shift);
      W = (n_tiles_cap + 31) / 32;
      bytes = static_cast<uint64_t>(id_cap) * W * 4;
      if (bytes <= budget || shift == MAX_TILE_SHIFT) break;
    }
    if (bytes > budget) {  // even the coarsest tiling is too big: scan every tile
      release_planes();
      plane_words_ = 0; id_cap_ = 0; tile_shift_ = MAX_TILE_SHIFT;
      return 0;
    }
    if (!planes_ || W != plane_words_ || id_cap != id_cap_ || shift != tile_shift_) {
      release_planes();
      RC(acquire_planes(bytes));
      plane_words_ = W; id_cap_ = id_cap; tile_shift_ = shift;
    }
    CK(cudaMemsetAsync(planes_, 0, bytes, st_));
    const double tb1 = now_ms();
    if (n_slots_) { k_build_planes<<<grid_for(n_slots_, 256), 256, 0, st_>>>(ids_[cur_], n_slots_, planes_, plane_words_, id_cap_, tile_shift_); launches_++; }
    if (dbg_print_) {
      cudaStreamSynchronize(st_);
      std::fprintf(stderr, "[PLANES]\t shift %u, %.1f MB: alloc+memset issue %.2f ms, build %.2f ms (free %.1f GB)\n", tile_shift_, bytes / 1048576.0, tb1 - tb0, now_ms() - tb1,
                   free_b / 1073741824.0);
    }
    return 0;
  }

  int alloc_tables() {
    if (!dt_.keys) {
      uint64_t cap = next_pow2(8ull * (256 + vocab_hint_) + 1024); if (cap < (1u << 16)) cap = 1u << 16;
      RC(alloc_dt(cap));
    }
    if (!pt_.ent) RC(alloc_pt(&pt_, 1ull << 20));
    if (!recs_) {
      rec_cap_ = dt_.cap;
      CK(cudaHostAlloc(reinterpret_cast<void**>(&recs_), static_cast<size_t>(rec_cap_) * sizeof(Rec), cudaHostAllocMapped));
    }
    return 0;
  }
  int alloc_dt(uint64_t cap) {
    if (dt_.keys) { cudaFreeAsync(dt_.keys, st_); cudaFreeAsync(dt_.delta, st_); cudaFreeAsync(dt_.seq, st_); cudaFreeAsync(dt_.list, st_); cudaFreeAsync(dt_.klist, st_); dt_.keys = nullptr; }
    CK(cudaMallocAsync(reinterpret_cast<void**>(&dt_.keys), cap * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&dt_.delta), cap * 8, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&dt_.seq), cap * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&dt_.list), cap * 4, st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&dt_.klist), cap * 8, st_));
    dt_.cap = static_cast<uint32_t>(cap); dt_.mask = cap - 1;
    // a key value no pair can produce: high word >= 2^31 that is neither all-ones nor unk_id
    uint32_t hi = 0x80000000u; if (static_cast<uint32_t>(cfg_.unk_id) == hi) hi = 0x80000001u;
    dt_.empty = static_cast<uint64_t>(hi) << 32;
    k_fill_u64<<<grid_for(cap, 256), 256, 0, st_>>>(reinterpret_cast<ull*>(dt_.keys), cap, dt_.empty);
    CK(cudaMemsetAsync(dt_.delta, 0, cap * 8, st_)); CK(cudaMemsetAsync(dt_.seq, 0xFF, cap * 8, st_));
    launches_++;
    if (recs_ && rec_cap_ < dt_.cap) { cudaFreeHost(recs_); recs_ = nullptr; rec_cap_ = dt_.cap; CK(cudaHostAlloc(reinterpret_cast<void**>(&recs_), static_cast<size_t>(rec_cap_) * sizeof(Rec), cudaHostAllocMapped)); }
    return 0;
  }
  int alloc_pt(PairTable* pt, uint64_t cap) {
    CK(cudaMallocAsync(reinterpret_cast<void**>(&pt->ent), cap * sizeof(PairEnt), st_));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&pt->serial), cap * 4, st_));
    pt->cap = cap; pt->mask = cap - 1;
    CK(cudaMemsetAsync(pt->ent, 0xFF, cap * sizeof(PairEnt), st_));  // key = EMPTY; freq is written when the entry is claimed
    return 0;
  }
  int grow_pt(uint64_t need_entries) {
    uint64_t cap = pt_.cap; while (need_entries * 2 > cap) cap *= 2;
    if (cap == pt_.cap) return 0;
    PairTable nt; std::memset(&nt, 0, sizeof nt);
    RC(alloc_pt(&nt, cap));
    k_rehash<<<grid_for(pt_.cap, 256), 256, 0, st_>>>(pt_, nt, ctr_); launches_++;
    CK(cudaStreamSynchronize(st_));
    cudaFreeAsync(pt_.ent, st_); cudaFreeAsync(pt_.serial, st_);
    pt_ = nt;
    return 0;
  }

  // --------------------------------------------------------------------------------------------------------- count
  int count_pairs(const Rec** recs, size_t* n) override {
    CK(cudaSetDevice(dev_));
    *recs = recs_; *n = 0;
    if (!loaded_) return 0;
    last_occ_ = ~0ull;  // a fresh pair table: the first merges are the occurrence-heavy ones
    for (int attempt = 0; attempt < 12; ++attempt) {
      CK(cudaMemsetAsync(pt_.ent, 0xFF, pt_.cap * sizeof(PairEnt), st_));
      CK(cudaMemsetAsync(ctr_, 0, sizeof(DevCounters), st_));
      bar_count_ = 0;
      pt_n_ = 0;
      ++flag_;
      CK(cudaEventRecord(ev0_, st_));
      if (n_words_) {
        const uint32_t n4c = static_cast<uint32_t>((n_slots_ + 3) / 4);
        k_count<<<n_sm_ * 4, 256, 0, st_>>>(reinterpret_cast<const int4*>(ids_[cur_]), reinterpret_cast<const uint4*>(wid_[cur_]), n4c, wcnt_, P_, dt_, ctr_,
                                            world_ > 1 ? seq_base(rank_) : 0ull);
        launches_++;
      }
      CK(cudaEventRecord(ev1_, st_));
      // the finalize pass needs dt_n <= cap/2 and room in the pair table: check before consuming the delta table
      DevCounters c;
      CK(cudaMemcpyAsync(&c, ctr_, sizeof c, cudaMemcpyDeviceToHost, st_));
      CK(cudaStreamSynchronize(st_));
      CK(cudaGetLastError());
      float ms = 0; cudaEventElapsedTime(&ms, ev0_, ev1_);
      if ((c.err & ERR_DT_FULL) || static_cast<uint64_t>(c.dt_n) * 2 > dt_.cap) {  // enlarge the scratch table, redo (ids untouched)
        RC(alloc_dt(static_cast<uint64_t>(dt_.cap) * 4));
        continue;
      }
      RC(grow_pt((world_ > 1 ? static_cast<uint64_t>(dt_.cap) / 2 : static_cast<uint64_t>(c.dt_n)) + 4ull * (256 + vocab_hint_) + 1024));  // replicas must size identically
      if (world_ > 1) {
        DistArgs a_D = next_exchange();
        const int grid = n_sm_ * 2;
        uint32_t a_reccap = rec_cap_, a_bar = bar_count_;
        Ctrl* a_ctrl = const_cast<Ctrl*>(ctrl_);
        uint64_t a_flag = flag_;
        bar_count_ += 1u * static_cast<uint32_t>(grid);
        void* args[] = {&dt_, &pt_, &ctr_, &recs_, &a_reccap, &a_ctrl, &P_, &a_flag, &a_D, &a_bar};
        CK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_dist_count_finalize), dim3(grid), dim3(256), args, 0, st_));
      } else {
        k_finalize_count<<<1, 256, 0, st_>>>(dt_, pt_, ctr_, recs_, rec_cap_, const_cast<Ctrl*>(ctrl_), P_, flag_);
      }
      launches_++;
      RC(wait_flag());
      if (ctrl_->err) { std::fprintf(stderr, "[ERROR]\t device count pass failed (err=%u)\n", ctrl_->err); return -1; }
      es_.count_launches++; es_.count_device_ms += ms; es_.count_bytes += 4.0 * static_cast<double>(n_slots_) + 12.0 * n_words_;  // 4S + 12N; the kernel reads 8 B per slot + counts
      pt_n_ = ctrl_->pt_n;
      *n = ctrl_->n_recs;
      es_.d2h_bytes += *n * sizeof(Rec) + sizeof(Ctrl);
      return 0;
    }
    return -1;
  }

  // --------------------------------------------------------------------------------------------------------- merge
  int merge(int32_t a, int32_t b, int32_t new_id, uint32_t /*serial*/, uint32_t /*list_len*/, const Rec** recs, size_t* n, uint64_t* occurrences) override {
    *recs = recs_; *n = 0; *occurrences = 0;
    const double tm0 = now_ms();
    CK(cudaSetDevice(dev_));  // the caller's thread may have another current device
    // keep the pair table at most half full even if this merge creates every key it can (4 per distinct id)
    const uint64_t worst_new = 4ull * (static_cast<uint64_t>(new_id) + 2);
    if ((pt_n_ + worst_new) * 2 > pt_.cap) RC(grow_pt(pt_n_ + worst_new));
    if (worst_new * 2 > dt_.cap) RC(alloc_dt(next_pow2(worst_new * 2)));
    // reclaim dead slots once a quarter of the scanned array is dead
    if (n_slots_ > 4096 && (n_slots_ - n_live_) * 4 > n_slots_) RC(compact());
    const uint32_t n4 = static_cast<uint32_t>((n_slots_ + 3) / 4);
    const double tl0 = now_ms();
    ++flag_;
    const bool timed = timing_every_ > 0 && (merge_seq_++ % timing_every_) == 0;
    const bool profiled = !profile_merges_.empty() && std::find(profile_merges_.begin(), profile_merges_.end(), merge_no_) != profile_merges_.end();
    if (profiled) { cudaStreamSynchronize(st_); cudaProfilerStart(); }
    if (timed) CK(cudaEventRecord(ev0_, st_));
    ++merge_no_;
    ull* a_dbg = timed ? dbg_ : nullptr;
    const uint32_t n_tiles = static_cast<uint32_t>((n_slots_ + (1ull << tile_shift_) - 1) >> tile_shift_);
    int grid = detect_grid(n4);
    const uint32_t tiles_per_cta = (n_tiles + grid - 1) / grid;
    const bool indexed = planes_ && static_cast<uint32_t>(a) < id_cap_ && static_cast<uint32_t>(b) < id_cap_;
    const uint32_t* pa = indexed ? planes_ + static_cast<uint64_t>(a) * plane_words_ : nullptr;
    const uint32_t* pb = indexed ? planes_ + static_cast<uint64_t>(b) * plane_words_ : nullptr;
    {
      int4* a_ids = reinterpret_cast<int4*>(ids_[cur_]);
      uint32_t a_n4 = n4, a_nt = n_tiles, a_tpc = tiles_per_cta, a_ts = tile_shift_, a_W = plane_words_, a_idcap = id_cap_, a_mno = merge_no_, a_reccap = rec_cap_;
      const uint32_t* a_wid = wid_[cur_]; const ull* a_wcnt = wcnt_; const ull* a_woff = woff_[cur_];
      int32_t a_A = a, a_B = b, a_N = new_id;
      Ctrl* a_ctrl = const_cast<Ctrl*>(ctrl_);
      uint64_t a_flag = flag_;
      uint32_t a_bar = bar_count_;  // barrier counter before this launch; it only grows (wraps mod 2^32)
      DistArgs a_D = dist_;
      if (world_ > 1) a_D = next_exchange();
      bar_count_ += (world_ > 1 ? 2u : 1u) * static_cast<uint32_t>(grid);
      uint32_t a_hot = last_occ_ >= hot_on_occ_ ? 1u : 0u;  // many occurrences last time: CTAs defer and aggregate their emission (kernels_fold.cuh)
      void* args[] = {&a_ids, &a_n4, &a_nt, &a_tpc, &a_ts, &pa, &pb, &planes_, &a_W, &a_idcap, &a_wid, &a_wcnt, &a_woff, &wlen_, &claimed_, &a_mno, &a_A, &a_B, &a_N,
                      &P_, &dt_, &pt_, &ctr_, &wl_, &recs_, &a_reccap, &a_ctrl, &a_flag, &a_bar, &a_dbg, &a_D, &a_hot};
      const void* kfn = world_ > 1 ? reinterpret_cast<const void*>(k_merge<4, true>) : reinterpret_cast<const void*>(k_merge<4, false>);
      if (plain_launch_) CK(cudaLaunchKernel(kfn, dim3(grid), dim3(256), args, 0, st_));  // experiment: same grid, no co-residency check by the driver
      else CK(cudaLaunchCooperativeKernel(kfn, dim3(grid), dim3(256), args, 0, st_));
    }
    if (timed) CK(cudaEventRecord(ev1_, st_));
    launches_ += 1;
    launch_ms_ += now_ms() - tl0;
    RC(wait_flag());
    last_occ_ = ctrl_->occ_local;
    if (profiled) {
      cudaStreamSynchronize(st_); cudaProfilerStop();
      std::fprintf(stderr, "[PROFILE]\t merge %u pair (%d,%d): slots %llu (algorithmic %llu bytes), candidate tiles %llu of %u (touched %llu bytes), occurrences %llu\n", merge_no_ - 1, a, b,
                   static_cast<ull>(n_slots_), 4ull * n_slots_, static_cast<ull>(ctrl_->cand_tiles), n_tiles, (4ull << tile_shift_) * ctrl_->cand_tiles, static_cast<ull>(ctrl_->occ));
    }
    if (ctrl_->err) { std::fprintf(stderr, "[ERROR]\t device merge pass failed (err=%u)\n", ctrl_->err); return -1; }
    if (timed) {
      float ms = 0;
      CK(cudaEventSynchronize(ev1_));
      cudaEventElapsedTime(&ms, ev0_, ev1_);
      const double algo = 4.0 * static_cast<double>(n4) * 4.0, touched = 4.0 * static_cast<double>(1u << tile_shift_) * static_cast<double>(ctrl_->cand_tiles);
      es_.scan_launches++; es_.scan_device_ms += ms; es_.scan_bytes += algo; es_.scan_bytes_touched += touched;
      const double p1 = (dbg_[1] - dbg_[0]) * 1e-6, p2 = (dbg_[2] - dbg_[1]) * 1e-6, p3 = (dbg_[3] - dbg_[2]) * 1e-6;  // ms: scan+emit+barrier | fold+publish | rewrite
      es_.scan_phase_ms += p1;
      if (ctrl_->cand_tiles * 10 >= static_cast<uint64_t>(n_tiles) * 9) { es_.dense_launches++; es_.dense_device_ms += ms; es_.dense_bytes += algo; es_.dense_phase_ms += p1; }  // streams >= 90 % of the array
      dbg_acc_[0] += p1; dbg_acc_[1] += p2; dbg_acc_[2] += p3; dbg_acc_[3] += ms; dbg_n_++;
      if (dbg_print_ && (dbg_n_ % 500) == 0)
        std::fprintf(stderr, "[KTIME]\t %llu timed merges: scan+emit+barrier %.1f us, fold+publish %.1f us, rewrite+rearm %.1f us | kernel (events) %.1f us (averages)\n",
                     (unsigned long long)dbg_n_, 1e3 * dbg_acc_[0] / dbg_n_, 1e3 * dbg_acc_[1] / dbg_n_, 1e3 * dbg_acc_[2] / dbg_n_, 1e3 * dbg_acc_[3] / dbg_n_);
    }
    cand_tiles_total_ += ctrl_->cand_tiles; tiles_total_ += n_tiles;
    *n = ctrl_->n_recs; *occurrences = ctrl_->occ;
    pt_n_ = ctrl_->pt_n;
    n_live_ -= ctrl_->occ_local;
    es_.d2h_bytes += *n * sizeof(Rec) + sizeof(Ctrl);
    merge_ms_ += now_ms() - tm0;
    return 0;
  }

  int compact() {
    const uint32_t N = n_words_;
    if (!N) return 0;
    ull *len1 = nullptr, *sums = nullptr;
    const uint32_t nb_scan = (N + SCAN_TILE - 1) / SCAN_TILE;
    CK(cudaMallocAsync(reinterpret_cast<void**>(&len1), static_cast<uint64_t>(N) * 8, st_)); CK(cudaMallocAsync(reinterpret_cast<void**>(&sums), (static_cast<uint64_t>(nb_scan) + 1) * 8, st_));
    const int nxt = cur_ ^ 1;
    k_len1<<<grid_for(N, 256), 256, 0, st_>>>(wlen_, N, len1);
    k_scan_sums<<<nb_scan, SCAN_THREADS, 0, st_>>>(len1, N, sums);
    k_scan_top<<<1, SCAN_THREADS, 0, st_>>>(sums, nb_scan, sums + nb_scan);
    k_scan_apply<<<nb_scan, SCAN_THREADS, 0, st_>>>(len1, N, sums, woff_[nxt]);
    ull S1 = 0;
    CK(cudaMemcpyAsync(&S1, sums + nb_scan, 8, cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    CK(cudaMemcpyAsync(woff_[nxt] + N, &S1, 8, cudaMemcpyHostToDevice, st_));
    k_compact<<<grid_for(N, 256), 256, 0, st_>>>(ids_[cur_], woff_[cur_], woff_[nxt], wlen_, N, ids_[nxt], wid_[nxt]);
    const uint64_t pad_to = ((S1 + 8 + 1023) / 1024) * 1024;
    k_fill_i32<<<grid_for(pad_to - S1, 256), 256, 0, st_>>>(ids_[nxt], S1, pad_to < ids_cap_ ? pad_to : ids_cap_, DEAD);
    launches_ += 6;
    CK(cudaStreamSynchronize(st_));
    CK(cudaGetLastError());
    cudaFreeAsync(len1, st_); cudaFreeAsync(sums, st_);
    cur_ = nxt; n_slots_ = S1; n_live_ = S1;
    es_.compactions++;
    RC(build_planes());
    return 0;
  }

  // ---------------------------------------------------------------------------------------------------------- save
  int token_freqs(uint64_t* freq, size_t T) override {
    CK(cudaSetDevice(dev_));
    if (!loaded_ || (!n_words_ && world_ == 1) || !T) return 0;
    ull* d = nullptr;
    CK(cudaMallocAsync(reinterpret_cast<void**>(&d), T * 8, st_));
    CK(cudaMemsetAsync(d, 0, T * 8, st_));
    if (n_words_) { k_token_freq<<<grid_for(n_words_, 256), 256, 0, st_>>>(ids_[cur_], woff_[cur_], wlen_, wcnt_, n_words_, P_, d, T); launches_++; }
    if (world_ > 1) {
      if (T > INBOX_ENTRIES * 3) { cudaFreeAsync(d, st_); std::fprintf(stderr, "[ERROR]\t vocabulary too large for the exchange buffer\n"); return -1; }
      DistArgs a_D = next_exchange();
      const int grid = n_sm_ * 2;
      uint64_t a_T = T;
      uint32_t a_bar = bar_count_;
      bar_count_ += 2u * static_cast<uint32_t>(grid);
      void* args[] = {&d, &a_T, &ctr_, &a_D, &a_bar};
      CK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_dist_sum_u64), dim3(grid), dim3(256), args, 0, st_));
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
    CK(cudaStreamSynchronize(st_));
    const uint32_t N = n_words_;
    std::vector<ull> ho(N + 1); std::vector<uint32_t> hl(N ? N : 1); std::vector<int32_t> hi(n_slots_ ? n_slots_ : 1);
    if (N) {
      CK(cudaMemcpy(ho.data(), woff_[cur_], (static_cast<uint64_t>(N) + 1) * 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hl.data(), wlen_, static_cast<uint64_t>(N) * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hi.data(), ids_[cur_], n_slots_ * 4, cudaMemcpyDeviceToHost));
      if (counts) CK(cudaMemcpy(counts, wcnt_, static_cast<uint64_t>(N) * 8, cudaMemcpyDeviceToHost));
    }
    uint64_t at = 0;
    for (uint32_t wi = 0; wi < N; wi++) {
      if (off) off[wi] = at;
      if (static_cast<uint32_t>(hi[ho[wi]]) != (HDR_BIT | wi)) return -2;  // layout invariant
      for (uint32_t j = 0; j < hl[wi]; j++) {
        int32_t code = hi[ho[wi] + 1 + j];
        if (ids && at < ids_cap) ids[at] = (cfg_.unk_id < 0 && code == UNK_CODE_NEG) ? cfg_.unk_id : code;
        at++;
      }
    }
    if (off) off[N] = at;
    return 0;
  }
  uint64_t get_pairs(int32_t* ab, uint64_t* freq, uint64_t cap) override {
    if (!pt_.ent) return 0;
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
    out->cand_tiles = cand_tiles_total_; out->tiles_total = tiles_total_;
  }
  const char* name() override { return name_; }

 private:
  int grid_for(uint64_t n, int block) const {
    uint64_t g = (n + block - 1) / block, maxg = static_cast<uint64_t>(n_sm_) * 16;
    if (g < 1) g = 1;
    return static_cast<int>(g < maxg ? g : maxg);
  }
  int detect_grid(uint32_t n4) const {  // persistent grid: SM count x resident CTAs per SM (occupancy query), never more than the work
    uint64_t warps_needed = (static_cast<uint64_t>(n4) + 127) / 128, ctas = (warps_needed + 7) / 8;
    uint64_t maxg = static_cast<uint64_t>(n_sm_) * scan_ctas_per_sm_;
    if (ctas < 1) ctas = 1;
    return static_cast<int>(ctas < maxg ? ctas : maxg);
  }

  int wait_flag() {
    double t0 = now_ms();
    uint64_t spins = 0;
    while (__atomic_load_n(&ctrl_->flag, __ATOMIC_ACQUIRE) != flag_) {
      if ((++spins & 0x3FFF) == 0) {
        cudaError_t q = cudaStreamQuery(st_);
        if (q == cudaSuccess) { if (__atomic_load_n(&ctrl_->flag, __ATOMIC_ACQUIRE) == flag_) break; std::fprintf(stderr, "[ERROR]\t device pass finished without publishing its result\n"); return -1; }
        if (q != cudaErrorNotReady) { std::fprintf(stderr, "[ERROR]\t CUDA: %s\n", cudaGetErrorString(q)); return -1; }
        if (now_ms() - t0 > 120000.0) { std::fprintf(stderr, "[ERROR]\t device pass timed out\n"); return -1; }
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    wait_ms_ += now_ms() - t0;
    return 0;
  }

  void release_corpus() {
    for (int i = 0; i < 2; i++) { if (ids_[i]) cudaFreeAsync(ids_[i], st_); ids_[i] = nullptr; if (woff_[i]) cudaFreeAsync(woff_[i], st_); woff_[i] = nullptr; }
    if (wcnt_) cudaFreeAsync(wcnt_, st_); wcnt_ = nullptr;
    if (wlen_) cudaFreeAsync(wlen_, st_); wlen_ = nullptr;
    if (wl_) cudaFreeAsync(wl_, st_); wl_ = nullptr;
    for (int i = 0; i < 2; i++) { if (wid_[i]) cudaFreeAsync(wid_[i], st_); wid_[i] = nullptr; }
    if (claimed_) cudaFreeAsync(claimed_, st_); claimed_ = nullptr;
    release_planes(); plane_words_ = 0; id_cap_ = 0; tile_shift_ = 9;
    n_words_ = 0; n_slots_ = n_live_ = 0; loaded_ = false; pt_n_ = 0;
  }
  void release_all() {
    cudaSetDevice(dev_);
    release_corpus();
    if (dt_.keys) { cudaFreeAsync(dt_.keys, st_); cudaFreeAsync(dt_.delta, st_); cudaFreeAsync(dt_.seq, st_); cudaFreeAsync(dt_.list, st_); cudaFreeAsync(dt_.klist, st_); }
    if (pt_.ent) { cudaFreeAsync(pt_.ent, st_); cudaFreeAsync(pt_.serial, st_); }
    if (recs_) cudaFreeHost(recs_);
    if (ctrl_) cudaFreeHost(const_cast<Ctrl*>(ctrl_));
    if (ctr_) cudaFreeAsync(ctr_, st_);
    if (st_) cudaStreamSynchronize(st_);
    dist_teardown();
    if (ev0_) cudaEventDestroy(ev0_);
    if (ev1_) cudaEventDestroy(ev1_);
    if (evm0_) cudaEventDestroy(evm0_);
    if (evm1_) cudaEventDestroy(evm1_);
    if (st_) cudaStreamDestroy(st_);
  }

  int dev_, n_sm_;
  char name_[320];
  cudaStream_t st_ = nullptr;
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr, evm0_ = nullptr, evm1_ = nullptr;
  EngineConfig cfg_{};
  Params P_{};
  bool loaded_ = false;
  uint32_t n_words_ = 0;
  uint64_t n_slots_ = 0, n_live_ = 0, ids_cap_ = 0;
  int32_t* ids_[2] = {nullptr, nullptr};
  ull* woff_[2] = {nullptr, nullptr};
  int cur_ = 0;
  ull* wcnt_ = nullptr;
  uint32_t* wlen_ = nullptr;
  uint32_t* wl_ = nullptr;
  uint32_t* wid_[2] = {nullptr, nullptr};
  uint32_t* claimed_ = nullptr;
  uint32_t merge_no_ = 0, bar_count_ = 0;
  int rank_ = 0, world_ = 1;
  DistArgs dist_{};
  uint8_t* inbox_ = nullptr;
  std::vector<uint64_t> host_counts_;
  std::string rdv_prefix_;
  uint32_t* planes_ = nullptr;
  uint64_t planes_bytes_ = 0;
  uint32_t plane_words_ = 0, id_cap_ = 0, tile_shift_ = 9;
  uint64_t cand_tiles_total_ = 0, tiles_total_ = 0;
  DeltaTable dt_{};
  PairTable pt_{};
  uint64_t pt_n_ = 0;
  Rec* recs_ = nullptr;
  uint32_t rec_cap_ = 0;
  volatile Ctrl* ctrl_ = nullptr;
  DevCounters* ctr_ = nullptr;
  uint64_t flag_ = 0;
  uint64_t vocab_hint_ = 32768;
  EngineStats es_{};
  uint64_t launches_ = 0, merge_seq_ = 0;
  double wait_ms_ = 0, launch_ms_ = 0, merge_ms_ = 0;
  int timing_every_ = 0;
  ull* dbg_ = nullptr;
  bool dbg_print_ = false, plain_launch_ = false;
  std::vector<uint32_t> profile_merges_;
  uint64_t hot_on_occ_ = 8192;  // SHRED_HOT_ON_OCC: 0 = every launch takes the hot-CTA path (tests), huge = never
  uint64_t last_occ_ = ~0ull;  // occurrences of the previous merge on this GPU (first merge of a corpus: assume many)
  double dbg_acc_[5] = {0, 0, 0, 0, 0};
  uint64_t dbg_n_ = 0;
  int scan_ctas_per_sm_ = 4;
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
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return nullptr;
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
