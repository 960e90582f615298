// encoder_cuda.cu -- B200 (sm_100a) BPE encoder / decoder for the model file the trainer writes (SURVEY.md section 8f rank 2:
// the caller-side step right after bpe_save).  The reference has only a pure-Python encoder (shredword/utils/bpe.py:191-225);
// this file is its device replacement behind the bpe_b200_encoder_* / bpe_b200_encode* entry points of include/shred_abi.h.
//
// Pipeline of one bpe_b200_encode call (all on the encoder's stream):
//   text -> HBM (padded with spaces)
//   k_enc_count_starts + scan   occurrences per 4 KB unit -> index of every occurrence in text order
//   k_enc_tokenize        distinct words of the text (the trainer's tokeniser + word table); every occurrence notes its word
//   k_enc_collect + scan  dense list of distinct words, pool offsets
//   k_enc_words           one warp per distinct word: the reference's merge loop, in place            kernels_encode.cuh
//   k_enc_toklen + scan   encoded lengths -> CSR offsets of the output
//   k_expand<int32>       occurrences copy their word's ids; stores coalesced over the output
// Results stay in HBM until bpe_b200_encode_fetch copies them out.  There is no CPU path: without a usable sm_100 device
// bpe_b200_encoder_create fails.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../../include/shred_abi.h"

namespace shred {
namespace {

#include "common.cuh"
#include "kernels_tokenize.cuh"
#include "kernels_scan.cuh"
#include "kernels_encode.cuh"

inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline uint64_t next_pow2(uint64_t x) { uint64_t p = 1; while (p < x) p <<= 1; return p; }

class CudaEncoder {
 public:
  ~CudaEncoder() { release(); }

  // A model is accepted iff row m is {a, b, 256 + m} with 0 <= a, b < 256 + m: what bpe_save can write (bpe.cpp:419-427).
  static bool valid_model(const int32_t* tri, size_t n) {
    for (size_t m = 0; m < n; m++) {
      const int32_t a = tri[3 * m], b = tri[3 * m + 1], c = tri[3 * m + 2];
      if (c != static_cast<int32_t>(256 + m) || a < 0 || b < 0 || a >= c || b >= c) return false;
    }
    return n < (1u << 30);
  }

  int init(const int32_t* tri, size_t n) {
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt <= 0) { std::fprintf(stderr, "[ERROR]\t CUDA: no usable device (%s)\n", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e)); return -1; }
    if (const char* s = std::getenv("SHRED_DEVICE")) dev_ = std::atoi(s);
    else if (cudaGetDevice(&dev_) != cudaSuccess) dev_ = 0;
    if (dev_ < 0 || dev_ >= cnt) dev_ = 0;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev_));
    if (prop.major != 10) { std::fprintf(stderr, "[ERROR]\t CUDA: device %d (%s, sm_%d%d) is not a Blackwell sm_100 part; this library carries sm_100a code only\n", dev_, prop.name, prop.major, prop.minor); return -1; }
    n_sm_ = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : N_SM_FALLBACK;
    CK(cudaSetDevice(dev_));
    CK(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
    for (auto& ev : ev_) CK(cudaEventCreate(&ev));
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev_) == cudaSuccess) { uint64_t thr = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr); }
    CK(cudaMalloc(reinterpret_cast<void**>(&ctr_), sizeof(DevCounters) + 64));
    scal_ = reinterpret_cast<ull*>(reinterpret_cast<uint8_t*>(ctr_) + sizeof(DevCounters));  // 8 scalar slots after the counters

    // merge dict: filled in file order, so a pair listed twice keeps its LAST id (Python dict assignment, utils/bpe.py:150-153)
    n_merges_ = n;
    tri_.assign(tri, tri + 3 * n);
    const uint64_t cap = next_pow2(4 * n + 16);
    std::vector<MergeEnt> ent(cap, MergeEnt{0, 0, 0});
    for (size_t m = 0; m < n; m++) {
      const uint64_t k = ((static_cast<uint64_t>(static_cast<uint32_t>(tri[3 * m])) << 32) | static_cast<uint32_t>(tri[3 * m + 1])) + 1ull;
      uint64_t s = mix64(k) & (cap - 1);
      while (ent[s].key != 0 && ent[s].key != k) s = (s + 1) & (cap - 1);
      ent[s].key = k; ent[s].val = tri[3 * m + 2];
    }
    CK(cudaMalloc(reinterpret_cast<void**>(&d_ent_), cap * sizeof(MergeEnt)));
    CK(cudaMemcpy(d_ent_, ent.data(), cap * sizeof(MergeEnt), cudaMemcpyHostToDevice));
    mt_.ent = d_ent_; mt_.mask = cap - 1;

    // token bytes: vocab[256 + m] = vocab[a] + vocab[b] (build_vocab, utils/bpe.py:74-76)
    const size_t T = 256 + n;
    std::vector<ull> toff(T + 1);
    for (size_t i = 0; i <= 256; i++) toff[i] = i;
    ull total = 256;
    for (size_t m = 0; m < n; m++) {
      const int32_t a = tri[3 * m], b = tri[3 * m + 1];
      total += (toff[a + 1] - toff[a]) + (toff[b + 1] - toff[b]);
      toff[256 + m + 1] = total;
    }
    std::vector<uint8_t> tbytes(total + 1);
    for (int i = 0; i < 256; i++) tbytes[i] = static_cast<uint8_t>(i);
    for (size_t m = 0; m < n; m++) {
      const int32_t a = tri[3 * m], b = tri[3 * m + 1];
      const ull la = toff[a + 1] - toff[a], lb = toff[b + 1] - toff[b];
      std::memcpy(&tbytes[toff[256 + m]], &tbytes[toff[a]], la);
      std::memcpy(&tbytes[toff[256 + m] + la], &tbytes[toff[b]], lb);
    }
    CK(cudaMalloc(reinterpret_cast<void**>(&d_toff_), (T + 1) * sizeof(ull)));
    CK(cudaMalloc(reinterpret_cast<void**>(&d_tbytes_), total + 1));
    CK(cudaMemcpy(d_toff_, toff.data(), (T + 1) * sizeof(ull), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_tbytes_, tbytes.data(), total, cudaMemcpyHostToDevice));
    return 0;
  }

  size_t vocab_size() const { return 256 + n_merges_; }

  int encode(const uint8_t* text, uint64_t n, uint64_t* n_words_out, uint64_t* n_ids_out) {
    CK(cudaSetDevice(dev_));
    have_result_ = false; n_ids_ = 0; n_tok_ = 0;
    std::memset(&stats_, 0, sizeof stats_);
    launches_ = 0;
    const double t_begin = now_ms();
    const uint64_t padded = ((n + 15) & ~15ull) + 64;
    uint8_t* d_text = nullptr;
    RC(need(b_text_, padded, &d_text));
    if (n) CK(cudaMemcpyAsync(d_text, text, n, cudaMemcpyHostToDevice, st_));
    CK(cudaMemsetAsync(d_text + n, ' ', padded - n, st_));
    CK(cudaStreamSynchronize(st_));
    stats_.h2d_ms = now_ms() - t_begin;
    stats_.h2d_bytes = n;
    RC(encode_device(d_text, n));
    have_result_ = true;
    stats_.encode_wall_ms = now_ms() - t_begin;
    stats_.text_bytes = n; stats_.n_words = n_tok_; stats_.n_ids = n_ids_; stats_.kernel_launches = launches_;
    if (n_words_out) *n_words_out = n_tok_;
    if (n_ids_out) *n_ids_out = n_ids_;
    return 0;
  }

  int fetch(int32_t* ids_out, uint64_t* off_out) {
    if (!have_result_) return -1;
    CK(cudaSetDevice(dev_));
    const double t0 = now_ms();
    if (ids_out && n_ids_) CK(cudaMemcpyAsync(ids_out, d_ids_, n_ids_ * sizeof(int32_t), cudaMemcpyDeviceToHost, st_));
    if (off_out) CK(cudaMemcpyAsync(off_out, d_off_, (n_tok_ + 1) * sizeof(ull), cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    stats_.d2h_ms += now_ms() - t0;
    stats_.d2h_bytes += (ids_out ? n_ids_ * 4 : 0) + (off_out ? (n_tok_ + 1) * 8 : 0);
    return 0;
  }

  // decode (utils/bpe.py:214-222): bytes of every id back to back.  Returns the byte count (nothing is written when it exceeds
  // cap), -2 for an id outside the vocab (ValueError in the reference), -1 on a CUDA error.
  int64_t decode(const int32_t* ids, uint64_t n, uint8_t* out, uint64_t cap) {
    if (cudaSetDevice(dev_) != cudaSuccess) return -1;
    if (n == 0) return 0;
    int32_t* d_in = nullptr; ull *d_len = nullptr, *d_src = nullptr; uint8_t* d_out = nullptr;
    uint32_t* bad = reinterpret_cast<uint32_t*>(scal_ + 2);
    auto cleanup = [&]() { afree(d_in); afree(d_len); afree(d_src); afree(d_out); cudaStreamSynchronize(st_); };
    auto body = [&]() -> int64_t {
      CK(cudaMallocAsync(reinterpret_cast<void**>(&d_in), n * 4, st_));
      CK(cudaMallocAsync(reinterpret_cast<void**>(&d_len), (n + 1) * 8, st_));
      CK(cudaMallocAsync(reinterpret_cast<void**>(&d_src), n * 8, st_));
      CK(cudaMemcpyAsync(d_in, ids, n * 4, cudaMemcpyHostToDevice, st_));
      CK(cudaMemsetAsync(bad, 0, 4, st_));
      k_dec_lens<<<grid_for(n, 256), 256, 0, st_>>>(d_in, n, d_toff_, static_cast<uint32_t>(vocab_size()), d_len, d_src, bad);
      RC(scan_in_place(d_len, n));
      ull total = 0; uint32_t is_bad = 0;
      CK(cudaMemcpyAsync(&total, d_len + n, 8, cudaMemcpyDeviceToHost, st_));
      CK(cudaMemcpyAsync(&is_bad, bad, 4, cudaMemcpyDeviceToHost, st_));
      CK(cudaStreamSynchronize(st_));
      if (is_bad) return -2;
      if (total > cap || total == 0) return static_cast<int64_t>(total);
      CK(cudaMallocAsync(reinterpret_cast<void**>(&d_out), total, st_));
      k_expand<uint8_t><<<grid_for((n + 31) / 32 * 32, 256), 256, 0, st_>>>(d_len, d_src, nullptr, nullptr, n, d_tbytes_, d_out);
      CK(cudaMemcpyAsync(out, d_out, total, cudaMemcpyDeviceToHost, st_));
      CK(cudaStreamSynchronize(st_));
      CK(cudaGetLastError());
      return static_cast<int64_t>(total);
    };
    const int64_t r = body();
    cleanup();
    return r;
  }

  void get_stats(shred_encode_stats_t* s) const { *s = stats_; }

 private:
  void afree(void* p) { if (p) cudaFreeAsync(p, st_); }
  int grid_for(uint64_t items, int threads) const {
    uint64_t b = (items + threads - 1) / threads;
    const uint64_t cap = static_cast<uint64_t>(n_sm_) * 16;
    if (b > cap) b = cap;
    return b ? static_cast<int>(b) : 1;
  }

  // Work buffers belong to the encoder and only ever grow: repeated calls on similar inputs allocate nothing (handing GB-sized
  // blocks back to the stream-ordered pool between calls made it re-grow, ~90 ms per call at 1 GB).
  struct Buf { void* p = nullptr; uint64_t cap = 0; };
  template <typename T>
  int need(Buf& b, uint64_t bytes, T** out) {
    if (b.cap < bytes) {
      if (b.p) { CK(cudaStreamSynchronize(st_)); CK(cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
      const uint64_t want = (bytes + (bytes >> 3) + (2ull << 20) - 1) & ~((2ull << 20) - 1);  // 12 % slack, 2 MB granules
      CK(cudaMalloc(&b.p, want));
      b.cap = want;
    }
    *out = static_cast<T*>(b.p);
    return 0;
  }

  // exclusive scan of a[0, n) in place; a[n] receives the total (a holds n + 1 entries)
  int scan_in_place(ull* a, uint64_t n) {
    const uint32_t nb = static_cast<uint32_t>((n + SCAN_TILE - 1) / SCAN_TILE);
    if (nb == 0) { CK(cudaMemsetAsync(a, 0, 8, st_)); return 0; }
    ull* sums = nullptr;
    RC(need(b_sums_, static_cast<uint64_t>(nb) * 8, &sums));
    k_scan_sums<<<nb, SCAN_THREADS, 0, st_>>>(a, n, sums);
    k_scan_top<<<1, SCAN_THREADS, 0, st_>>>(sums, nb, a + n);
    k_scan_apply<<<nb, SCAN_THREADS, 0, st_>>>(a, n, sums, a);
    launches_ += 3;
    return 0;
  }

  int encode_device(const uint8_t* d_text, uint64_t n) {
    WordTable wt; std::memset(&wt, 0, sizeof wt);
    uint32_t *u_slot = nullptr, *enc_len = nullptr, *tok_slot = nullptr, *u_n = reinterpret_cast<uint32_t*>(scal_);
    ull *u_len = nullptr, *enc_off = nullptr, *unit_cnt = nullptr;
    int32_t* pool = nullptr;
    CK(cudaEventRecord(ev_[0], st_));
    // --- occurrences per 4 KB unit -> index of every occurrence in text order
    const uint64_t n_units = (n + UNIT_BYTES - 1) / UNIT_BYTES;
    RC(need(b_unit_, (n_units + 1) * 8, &unit_cnt));
    if (n_units) { k_enc_count_starts<<<grid_for(n_units * 256, 256), 256, 0, st_>>>(d_text, n, n_units, unit_cnt); launches_++; }
    RC(scan_in_place(unit_cnt, n_units));
    ull total_tok = 0;
    CK(cudaMemcpyAsync(&total_tok, unit_cnt + n_units, 8, cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    n_tok_ = total_tok;
    RC(need(b_tok_slot_, (n_tok_ + 1) * 4, &tok_slot));
    // --- distinct words of the text (same table and retry rules as the trainer's ingest); every occurrence notes its word
    DevCounters zero; std::memset(&zero, 0, sizeof zero);
    DevCounters c;
    uint64_t cap = next_pow2(n / 64 + 1); if (cap < (1u << 16)) cap = 1u << 16;
    if (cap < wt_cap_hint_) cap = wt_cap_hint_;  // what the previous call ended with
    uint32_t seed = 0x5bd1e995u;
    for (int attempt = 0;; ++attempt) {
      if (attempt > 8 || cap > (1ull << 32)) { std::fprintf(stderr, "[ERROR]\t encoder: distinct-word table did not converge\n"); return -1; }
      RC(need(b_wt_tag_, cap * 8, &wt.tag)); RC(need(b_wt_first_, cap * 8, &wt.first));
      RC(need(b_wt_len_, cap * 4, &wt.len)); RC(need(b_wt_bucket_, cap * 4, &wt.bucket));
      wt.cap = cap; wt.mask = cap - 1;
      CK(cudaMemsetAsync(wt.tag, 0, cap * 8, st_)); CK(cudaMemsetAsync(wt.first, 0xFF, cap * 8, st_));
      CK(cudaMemcpyAsync(ctr_, &zero, sizeof zero, cudaMemcpyHostToDevice, st_));
      if (n_units) { k_enc_tokenize<<<grid_for(n_units * 256, 256), 256, 0, st_>>>(d_text, n, n_units, unit_cnt, wt, ctr_, seed, tok_slot); launches_++; }
      CK(cudaMemcpyAsync(&c, ctr_, sizeof c, cudaMemcpyDeviceToHost, st_));
      CK(cudaStreamSynchronize(st_));
      CK(cudaGetLastError());
      const bool too_full = static_cast<uint64_t>(c.n_unique) * 2 > cap;
      if ((c.err & (ERR_WT_FULL | ERR_WT_COLLISION)) || too_full) {
        if ((c.err & ERR_WT_FULL) || too_full) cap *= 4;
        if (c.err & ERR_WT_COLLISION) seed = seed * 2654435761u + 12345u;
        continue;
      }
      break;
    }
    wt_cap_hint_ = cap <= (1ull << 26) ? cap : 0;
    const uint32_t N = c.n_unique;
    stats_.n_unique_words = N;
    CK(cudaEventRecord(ev_[1], st_));

    // --- every distinct word once
    RC(need(b_u_slot_, (static_cast<uint64_t>(N) + 1) * 4, &u_slot));
    RC(need(b_u_len_, (static_cast<uint64_t>(N) + 1) * 8, &u_len));
    RC(need(b_enc_len_, cap * 4, &enc_len));
    RC(need(b_enc_off_, cap * 8, &enc_off));
    CK(cudaMemsetAsync(u_n, 0, 4, st_));
    k_enc_collect<<<grid_for(cap, 256), 256, 0, st_>>>(wt, u_slot, u_n, u_len); launches_++;
    RC(scan_in_place(u_len, N));
    ull pool_n = 0;
    CK(cudaMemcpyAsync(&pool_n, u_len + N, 8, cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    stats_.pool_ids = pool_n;
    RC(need(b_pool_, (pool_n + 1) * 4, &pool));
    if (N) { k_enc_words<<<grid_for(static_cast<uint64_t>(N) * 32, ENC_WARPS * 32), ENC_WARPS * 32, 0, st_>>>(d_text, wt, u_slot, u_len, N, mt_, pool, enc_len, enc_off); launches_++; }
    CK(cudaEventRecord(ev_[2], st_));

    // --- occurrences copy their word's ids
    RC(need(b_off_, (n_tok_ + 1) * 8, &d_off_));
    if (n_tok_) { k_enc_toklen<<<grid_for(n_tok_, 256), 256, 0, st_>>>(tok_slot, n_tok_, enc_len, d_off_); launches_++; }
    RC(scan_in_place(d_off_, n_tok_));
    ull n_ids = 0;
    CK(cudaMemcpyAsync(&n_ids, d_off_ + n_tok_, 8, cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    n_ids_ = n_ids;
    RC(need(b_ids_, (n_ids_ + 1) * 4, &d_ids_));
    if (n_tok_) { k_expand<int32_t><<<grid_for((n_tok_ + 31) / 32 * 32, 256), 256, 0, st_>>>(d_off_, nullptr, tok_slot, enc_off, n_tok_, pool, d_ids_); launches_++; }
    CK(cudaEventRecord(ev_[3], st_));
    CK(cudaStreamSynchronize(st_));
    CK(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, ev_[0], ev_[1]); stats_.tokenize_ms = ms;
    cudaEventElapsedTime(&ms, ev_[1], ev_[2]); stats_.words_ms = ms;
    cudaEventElapsedTime(&ms, ev_[2], ev_[3]); stats_.expand_ms = ms;
    cudaEventElapsedTime(&ms, ev_[0], ev_[3]); stats_.device_ms = ms;
    return 0;
  }

  void release() {
    if (!st_) return;
    cudaSetDevice(dev_);
    cudaStreamSynchronize(st_);
    for (Buf* b : {&b_text_, &b_unit_, &b_sums_, &b_tok_slot_, &b_wt_tag_, &b_wt_first_, &b_wt_len_, &b_wt_bucket_, &b_u_slot_, &b_u_len_, &b_enc_len_, &b_enc_off_,
                   &b_pool_, &b_off_, &b_ids_})
      if (b->p) { cudaFree(b->p); b->p = nullptr; b->cap = 0; }
    cudaFree(ctr_); cudaFree(d_ent_); cudaFree(d_toff_); cudaFree(d_tbytes_);
    for (auto& ev : ev_) if (ev) cudaEventDestroy(ev);
    cudaStreamDestroy(st_);
    st_ = nullptr;
  }

  int dev_ = 0, n_sm_ = N_SM_FALLBACK;
  cudaStream_t st_ = nullptr;
  cudaEvent_t ev_[4] = {nullptr, nullptr, nullptr, nullptr};
  DevCounters* ctr_ = nullptr;
  ull* scal_ = nullptr;
  size_t n_merges_ = 0;
  std::vector<int32_t> tri_;
  MergeEnt* d_ent_ = nullptr; MergeTable mt_{};
  ull* d_toff_ = nullptr; uint8_t* d_tbytes_ = nullptr;
  Buf b_text_, b_unit_, b_sums_, b_tok_slot_, b_wt_tag_, b_wt_first_, b_wt_len_, b_wt_bucket_, b_u_slot_, b_u_len_, b_enc_len_, b_enc_off_, b_pool_, b_off_, b_ids_;
  uint64_t wt_cap_hint_ = 0;
  int32_t* d_ids_ = nullptr; ull* d_off_ = nullptr;  // views into b_ids_ / b_off_: the last result
  uint64_t n_ids_ = 0, n_tok_ = 0;
  bool have_result_ = false;
  uint64_t launches_ = 0;
  shred_encode_stats_t stats_{};
};

}  // namespace
}  // namespace shred

// the opaque handle is the CudaEncoder itself
static inline shred::CudaEncoder* impl_of(shred_encoder_t* e) { return reinterpret_cast<shred::CudaEncoder*>(e); }
static inline const shred::CudaEncoder* impl_of(const shred_encoder_t* e) { return reinterpret_cast<const shred::CudaEncoder*>(e); }

extern "C" {

shred_encoder_t* bpe_b200_encoder_create(const int32_t* triples, size_t n_merges) {
  if (n_merges && !triples) return nullptr;
  if (!shred::CudaEncoder::valid_model(triples, n_merges)) { std::fprintf(stderr, "[ERROR]\t encoder: not a BPE model written by bpe_save (row m must be {a, b, 256 + m} with 0 <= a, b < 256 + m)\n"); return nullptr; }
  shred::CudaEncoder* e = new shred::CudaEncoder();
  if (e->init(triples, n_merges) != 0) { delete e; return nullptr; }
  return reinterpret_cast<shred_encoder_t*>(e);
}

shred_encoder_t* bpe_b200_encoder_load(const char* model_path) {
  if (!model_path) return nullptr;
  FILE* f = std::fopen(model_path, "rb");
  if (!f) { std::fprintf(stderr, "[ERROR]\t encoder: cannot open %s\n", model_path); return nullptr; }
  std::fseek(f, 0, SEEK_END);
  const long sz = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  if (sz < 0 || sz % 12 != 0) { std::fclose(f); std::fprintf(stderr, "[ERROR]\t encoder: %s is not a whole number of 12-byte merge records\n", model_path); return nullptr; }
  std::vector<int32_t> tri(static_cast<size_t>(sz) / 4 + 3);
  const size_t got = std::fread(tri.data(), 1, static_cast<size_t>(sz), f);
  std::fclose(f);
  if (got != static_cast<size_t>(sz)) return nullptr;
  return bpe_b200_encoder_create(tri.data(), static_cast<size_t>(sz) / 12);
}

void bpe_b200_encoder_destroy(shred_encoder_t* enc) { delete impl_of(enc); }

size_t bpe_b200_encoder_vocab_size(const shred_encoder_t* enc) { return enc ? impl_of(enc)->vocab_size() : 0; }

int bpe_b200_encode(shred_encoder_t* enc, const uint8_t* text, uint64_t n_bytes, uint64_t* n_words, uint64_t* n_ids) {
  if (!enc || (n_bytes && !text)) return -1;
  return impl_of(enc)->encode(text, n_bytes, n_words, n_ids);
}

int bpe_b200_encode_fetch(shred_encoder_t* enc, int32_t* ids_out, uint64_t* offsets_out) {
  if (!enc) return -1;
  return impl_of(enc)->fetch(ids_out, offsets_out);
}

int64_t bpe_b200_decode(shred_encoder_t* enc, const int32_t* ids, uint64_t n_ids, uint8_t* out, uint64_t cap) {
  if (!enc || (n_ids && !ids)) return -1;
  return impl_of(enc)->decode(ids, n_ids, out, cap);
}

int bpe_b200_encoder_get_stats(const shred_encoder_t* enc, shred_encode_stats_t* out) {
  if (!enc || !out) return -1;
  impl_of(enc)->get_stats(out);
  return 0;
}

}  // extern "C"
