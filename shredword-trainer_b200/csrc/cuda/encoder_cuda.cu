// encoder_cuda.cu -- B200 (sm_100a) BPE encoder / decoder for the model file the trainer writes (SURVEY.md section 8f rank 2:
// the caller-side step right after bpe_save).  The reference has only a pure-Python encoder (shredword/utils/bpe.py:191-225);
// this file is its device replacement behind the bpe_b200_encoder_* / bpe_b200_encode* entry points of include/shred_abi.h.
//
// Pipeline of one call, per piece of the text (bpe_b200_encode: one piece; bpe_b200_encode_to_host: 64 MB pieces cut at
// delimiters, piece k + 1 on its way in and piece k - 1 on its way out while piece k is encoded):
//   k_enc_count_starts + scan   occurrences per 4 KB unit -> index of every occurrence in text order
//   k_enc_tokenize        distinct words (the trainer's tokeniser + word table, kept across pieces); every occurrence notes its
//                         word, newly seen words are listed                                          kernels_tokenize.cuh
//   k_enc_newlens + scan  pool offsets of the new words
//   k_enc_words           one warp per new word: the reference's merge loop, in place                kernels_encode.cuh
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
#include "../layout.hpp"

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
    CK(cudaStreamCreateWithFlags(&st_h2d_, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&st_d2h_, cudaStreamNonBlocking));
    for (auto& ev : ev_) CK(cudaEventCreate(&ev));
    for (int i = 0; i < 2; i++) CK(cudaEventCreateWithFlags(&ev_d2h_[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ev_comp_, cudaEventDisableTiming));
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev_) == cudaSuccess) { uint64_t thr = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr); }
    CK(cudaMalloc(reinterpret_cast<void**>(&ctr_), sizeof(DevCounters) + 64));
    scal_ = reinterpret_cast<ull*>(reinterpret_cast<uint8_t*>(ctr_) + sizeof(DevCounters));  // 8 scalar slots after the counters
    static_assert(sizeof(DevCounters) % 8 == 0, "mirrored as 8-byte words");
    void* hp = nullptr;
    CK(cudaHostAlloc(&hp, 256, cudaHostAllocMapped));
    std::memset(hp, 0, 256);
    mirror_ = static_cast<volatile ull*>(hp);
    void* dp = nullptr;
    CK(cudaHostGetDevicePointer(&dp, hp, 0));
    mirror_dev_ = static_cast<ull*>(dp);

    // merge dict: filled in file order, so a pair listed twice keeps its LAST id (Python dict assignment, utils/bpe.py:150-153)
    n_merges_ = n;
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

  // Resident variant: the whole text is one piece; ids and offsets stay in HBM until fetch().
  int encode(const uint8_t* text, uint64_t n, uint64_t* n_words_out, uint64_t* n_ids_out) {
    return run(text, n, false, nullptr, 0, nullptr, 0, n_words_out, n_ids_out);
  }

  // Streamed variant: the text is cut at delimiters into pieces; while piece k is encoded, piece k + 1 is on its way to the
  // device and the ids and offsets of piece k - 1 are on their way back (PCIe is full duplex; the resident call spends 80 % of
  // its time in the two copies, one after the other).  The text, the word table, the encoded words and the pool persist across
  // the pieces, so every distinct word is still encoded once; only the result buffers are double.
  // Returns 0, -1 on a device error, -3 when ids_cap / off_cap are too small (ids_cap >= n_bytes and off_cap >= n_bytes / 2 + 2
  // always suffice).
  int encode_stream(const uint8_t* text, uint64_t n, int32_t* ids_out, uint64_t ids_cap, uint64_t* off_out, uint64_t off_cap, uint64_t* n_words_out,
                    uint64_t* n_ids_out) {
    return run(text, n, true, ids_out, ids_cap, off_out, off_cap, n_words_out, n_ids_out);
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

  int run(const uint8_t* text, uint64_t n, bool stream, int32_t* ids_out, uint64_t ids_cap, uint64_t* off_out, uint64_t off_cap, uint64_t* n_words_out,
          uint64_t* n_ids_out) {
    CK(cudaSetDevice(dev_));
    have_result_ = false; n_ids_ = 0; n_tok_ = 0;
    std::memset(&stats_, 0, sizeof stats_);
    launches_ = 0;
    const double t_begin = now_ms();
    if (const char* e = std::getenv("SHRED_ENCODE_DEBUG")) debug_ = *e && *e != '0';
    if (stream && off_out && off_cap < 1) return -3;
    // --- pieces: cut just after the last delimiter at or before each nominal boundary, so every piece ends with a delimiter
    //     (or at the end of the text, which is padded with spaces) and token walks never leave the bytes that have landed
    std::vector<uint64_t> cut{0};
    if (!stream) { if (n) cut.push_back(n); }
    else {
      uint64_t piece = 64ull << 20;
      if (const char* e = std::getenv("SHRED_ENCODE_PIECE_BYTES")) { const uint64_t v = std::strtoull(e, nullptr, 10); if (v >= 1) piece = v; }
      auto delim = [](uint8_t c) { return c == 9 || c == 10 || c == 13 || c == 32; };
      while (cut.back() < n) {
        uint64_t pos = cut.back() + piece;
        if (pos >= n) { cut.push_back(n); break; }
        while (pos > cut.back() && !delim(text[pos - 1])) --pos;
        if (pos == cut.back()) {  // one word longer than a piece: extend past its end
          pos = cut.back() + piece;
          while (pos < n && !delim(text[pos])) ++pos;
          if (pos < n) ++pos;
        }
        cut.push_back(pos);
      }
    }
    const size_t K = cut.size() - 1;

    // --- the text lives in HBM as a whole (word representatives are offsets into it), padded with spaces
    const uint64_t padded = ((n + 15) & ~15ull) + 64;
    uint8_t* d_text = nullptr;
    RC(need(b_text_, padded, &d_text));
    CK(cudaMemsetAsync(d_text + n, ' ', padded - n, st_h2d_));
    // every piece is queued on the copy stream right away (the text buffer is whole, nothing waits for compute); one event each
    while (ev_piece_.size() < K) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ev_piece_.push_back(e); }
    for (size_t k = 0; k < K; k++) {
      CK(cudaMemcpyAsync(d_text + cut[k], text + cut[k], cut[k + 1] - cut[k], cudaMemcpyHostToDevice, st_h2d_));
      CK(cudaEventRecord(ev_piece_[k], st_h2d_));
    }
    if (!stream || K == 0) { CK(cudaStreamSynchronize(st_h2d_)); stats_.h2d_ms = now_ms() - t_begin; }
    stats_.h2d_bytes = n;

    // --- the distinct-word table of the call (same sizing and retry rules as the trainer's ingest)
    WordTable wt; std::memset(&wt, 0, sizeof wt);
    uint64_t cap = next_pow2(n / 64 + 1); if (cap < (1u << 16)) cap = 1u << 16;
    if (cap < wt_cap_hint_ && wt_cap_hint_ <= 64 * cap) cap = wt_cap_hint_;  // what a previous call of similar size ended with
    uint32_t seed = 0x5bd1e995u;
    uint32_t *u_n = reinterpret_cast<uint32_t*>(scal_), *new_list = nullptr, *enc_len = nullptr, *tok_slot = nullptr;
    ull *u_len = nullptr, *enc_off = nullptr, *unit_cnt = nullptr;
    int32_t* pool = nullptr;
    DevCounters c;
    uint64_t tok_base = 0, ids_base = 0;

    for (int attempt = 0;; ++attempt) {
      if (attempt > 8 || cap > (1ull << 32)) { std::fprintf(stderr, "[ERROR]\t encoder: distinct-word table did not converge\n"); return -1; }
      RC(need(b_wt_tag_, cap * 8, &wt.tag)); RC(need(b_wt_first_, cap * 8, &wt.first));
      RC(need(b_wt_len_, cap * 4, &wt.len)); RC(need(b_wt_bucket_, cap * 4, &wt.bucket));
      RC(need(b_enc_len_, cap * 4, &enc_len)); RC(need(b_enc_off_, cap * 8, &enc_off));
      wt.cap = cap; wt.mask = cap - 1;
      CK(cudaMemsetAsync(wt.tag, 0, cap * 8, st_)); CK(cudaMemsetAsync(wt.first, 0xFF, cap * 8, st_));
      CK(cudaMemsetAsync(ctr_, 0, sizeof(DevCounters), st_));  // not a host copy: it would queue behind the text pieces
      uint64_t n_unique = 0, pool_top = 0;
      tok_base = 0; ids_base = 0;
      bool again = false;
      shred_encode_stats_t acc; std::memset(&acc, 0, sizeof acc);

      for (size_t k = 0; k < K && !again; k++) {
        const uint64_t lo = cut[k], hi = cut[k + 1];
        const double t_piece = now_ms();
        if (attempt == 0 && stream) CK(cudaStreamWaitEvent(st_, ev_piece_[k], 0));
        Buf& b_off = stream ? b_offp_[k & 1] : b_off_;
        Buf& b_ids = stream ? b_idsp_[k & 1] : b_ids_;
        if (stream && k >= 2) CK(cudaEventSynchronize(ev_d2h_[k & 1]));  // the result buffers of piece k - 2 have left
        CK(cudaEventRecord(ev_[0], st_));
        // occurrences per 4 KB unit -> index of every occurrence of the piece in text order
        const uint64_t u0 = lo / UNIT_BYTES, n_units = hi > lo ? (hi + UNIT_BYTES - 1) / UNIT_BYTES - u0 : 0;
        RC(need(b_unit_, (n_units + 1) * 8, &unit_cnt));
        if (n_units) { k_enc_count_starts<<<grid_for(n_units * 256, 256), 256, 0, st_>>>(d_text, lo, hi, u0, n_units, unit_cnt); launches_++; }
        RC(scan_in_place(unit_cnt, n_units));
        ull nt = 0;
        RC(read_back(unit_cnt + n_units, &nt, 8));
        RC(need(b_tok_slot_, (nt + 1) * 4, &tok_slot));
        RC(need(b_u_slot_, (nt + 1) * 4, &new_list));  // at most every occurrence is a new word
        // distinct words: every occurrence notes its word, new words are listed
        CK(cudaMemsetAsync(u_n, 0, 4, st_));
        if (n_units) { k_enc_tokenize<<<grid_for(n_units * 256, 256), 256, 0, st_>>>(d_text, lo, hi, u0, n_units, unit_cnt, wt, ctr_, seed, tok_slot, new_list, u_n); launches_++; }
        RC(read_back(ctr_, &c, sizeof c));
        CK(cudaGetLastError());
        const bool too_full = static_cast<uint64_t>(c.n_unique) * 2 > cap;
        if ((c.err & (ERR_WT_FULL | ERR_WT_COLLISION)) || too_full) {  // start the call over (the text is already on its way)
          if ((c.err & ERR_WT_FULL) || too_full) cap *= 4;
          if (c.err & ERR_WT_COLLISION) seed = seed * 2654435761u + 12345u;
          again = true;
          break;
        }
        const uint32_t n_new = static_cast<uint32_t>(c.n_unique - n_unique);
        n_unique = c.n_unique;
        CK(cudaEventRecord(ev_[1], st_));
        // every new word once
        RC(need(b_u_len_, (static_cast<uint64_t>(n_new) + 1) * 8, &u_len));
        if (n_new) { k_enc_newlens<<<grid_for(n_new, 256), 256, 0, st_>>>(wt, new_list, n_new, u_len); launches_++; }
        RC(scan_in_place(u_len, n_new));
        ull pool_n = 0;
        RC(read_back(u_len + n_new, &pool_n, 8));
        if ((pool_top + pool_n + 1) * 4 > b_pool_.cap) RC(grow_keep(b_pool_, (pool_top + pool_n + 1) * 4 + (stream ? (pool_top + pool_n) * 2 : 0), pool_top * 4));
        pool = static_cast<int32_t*>(b_pool_.p);
        if (n_new) { k_enc_words<<<grid_for(static_cast<uint64_t>(n_new) * 32, ENC_WARPS * 32), ENC_WARPS * 32, 0, st_>>>(d_text, wt, new_list, u_len, n_new, mt_, pool, pool_top, enc_len, enc_off); launches_++; }
        pool_top += pool_n;
        CK(cudaEventRecord(ev_[2], st_));
        // occurrences copy their word's ids
        RC(need(b_off, (nt + 1) * 8, &d_off_));
        if (nt) { k_enc_toklen<<<grid_for(nt, 256), 256, 0, st_>>>(tok_slot, nt, enc_len, d_off_); launches_++; }
        RC(scan_in_place(d_off_, nt));
        ull ni = 0;
        RC(read_back(d_off_ + nt, &ni, 8));
        RC(need(b_ids, (ni + 1) * 4, &d_ids_));
        if (nt) { k_expand<int32_t><<<grid_for((nt + 31) / 32 * 32, 256), 256, 0, st_>>>(d_off_, nullptr, tok_slot, enc_off, nt, pool, d_ids_); launches_++; }
        CK(cudaEventRecord(ev_[3], st_));
        if (stream) {
          if (ids_base + ni > ids_cap || (off_out && tok_base + nt + 1 > off_cap)) { cudaStreamSynchronize(st_); cudaStreamSynchronize(st_d2h_); cudaStreamSynchronize(st_h2d_); return -3; }
          if (ids_base && off_out) { k_add_u64<<<grid_for(nt + 1, 256), 256, 0, st_>>>(d_off_, nt + 1, ids_base); launches_++; }
          CK(cudaEventRecord(ev_comp_, st_));
          CK(cudaStreamWaitEvent(st_d2h_, ev_comp_, 0));
          const bool last = k + 1 == K;
          if (ni) CK(cudaMemcpyAsync(ids_out + ids_base, d_ids_, ni * 4, cudaMemcpyDeviceToHost, st_d2h_));
          if (off_out) CK(cudaMemcpyAsync(off_out + tok_base, d_off_, (nt + (last ? 1 : 0)) * 8, cudaMemcpyDeviceToHost, st_d2h_));
          CK(cudaEventRecord(ev_d2h_[k & 1], st_d2h_));
        }
        CK(cudaStreamSynchronize(st_));
        CK(cudaGetLastError());
        float ms = 0;
        cudaEventElapsedTime(&ms, ev_[0], ev_[1]); acc.tokenize_ms += ms;
        cudaEventElapsedTime(&ms, ev_[1], ev_[2]); acc.words_ms += ms;
        cudaEventElapsedTime(&ms, ev_[2], ev_[3]); acc.expand_ms += ms;
        cudaEventElapsedTime(&ms, ev_[0], ev_[3]); acc.device_ms += ms;
        if (debug_) {
          float a = 0, b = 0, d = 0;
          cudaEventElapsedTime(&a, ev_[0], ev_[1]); cudaEventElapsedTime(&b, ev_[1], ev_[2]); cudaEventElapsedTime(&d, ev_[2], ev_[3]);
          std::fprintf(stderr, "[encode] piece %zu [%llu, %llu) start %.2f ms  end %.2f ms | tokenize %.3f words %.3f expand %.3f | %llu words %u new %llu ids\n", k,
                       static_cast<ull>(lo), static_cast<ull>(hi), t_piece - t_begin, now_ms() - t_begin, a, b, d, nt, n_new, ni);
        }
        tok_base += nt; ids_base += ni;
        if (!stream) { n_tok_ = nt; n_ids_ = ni; }
      }
      if (again) {  // every copy in flight targets buffers this call owns: drain, then start over
        CK(cudaStreamSynchronize(st_d2h_));
        CK(cudaStreamSynchronize(st_h2d_));
        continue;
      }
      if (stream) CK(cudaStreamSynchronize(st_d2h_));
      const double keep_h2d = stats_.h2d_ms;
      stats_ = acc;
      stats_.h2d_ms = keep_h2d;
      stats_.n_unique_words = n_unique; stats_.pool_ids = pool_top;
      break;
    }
    if (K == 0) {  // empty text: one offset, no ids
      if (stream && off_out) off_out[0] = 0;
      else { RC(need(b_off_, 8, &d_off_)); RC(need(b_ids_, 4, &d_ids_)); CK(cudaMemsetAsync(d_off_, 0, 8, st_)); CK(cudaStreamSynchronize(st_)); }
    }
    wt_cap_hint_ = cap <= (1ull << 26) ? cap : 0;
    have_result_ = !stream;
    stats_.text_bytes = n; stats_.n_words = tok_base; stats_.n_ids = ids_base; stats_.kernel_launches = launches_;
    stats_.h2d_bytes = n; stats_.d2h_bytes = stream ? ids_base * 4 + (off_out ? (tok_base + 1) * 8 : 0) : 0;
    stats_.encode_wall_ms = now_ms() - t_begin;
    if (n_words_out) *n_words_out = tok_base;
    if (n_ids_out) *n_ids_out = ids_base;
    return 0;
  }

  // device words -> host through the mapped mirror (no copy engine involved); returns after the stream has drained
  int read_back(const void* d_src, void* h_dst, uint32_t bytes) {
    k_mirror<<<1, 32, 0, st_>>>(static_cast<const ull*>(d_src), mirror_dev_, bytes / 8);
    CK(cudaStreamSynchronize(st_));
    for (uint32_t i = 0; i < bytes / 8; i++) static_cast<ull*>(h_dst)[i] = mirror_[i];
    return 0;
  }

  // grow a buffer whose first `keep` bytes are live (the pool of encoded words while a streamed call is running)
  int grow_keep(Buf& b, uint64_t bytes, uint64_t keep) {
    const uint64_t want = (bytes + (bytes >> 3) + (2ull << 20) - 1) & ~((2ull << 20) - 1);
    void* np = nullptr;
    CK(cudaStreamSynchronize(st_));
    CK(cudaMalloc(&np, want));
    if (keep && b.p) CK(cudaMemcpy(np, b.p, keep, cudaMemcpyDeviceToDevice));
    if (b.p) CK(cudaFree(b.p));
    b.p = np; b.cap = want;
    return 0;
  }

  void release() {
    if (!st_) return;
    cudaSetDevice(dev_);
    cudaStreamSynchronize(st_);
    for (Buf* b : {&b_text_, &b_unit_, &b_sums_, &b_tok_slot_, &b_wt_tag_, &b_wt_first_, &b_wt_len_, &b_wt_bucket_, &b_u_slot_, &b_u_len_, &b_enc_len_, &b_enc_off_,
                   &b_pool_, &b_off_, &b_ids_, &b_offp_[0], &b_offp_[1], &b_idsp_[0], &b_idsp_[1]})
      if (b->p) { cudaFree(b->p); b->p = nullptr; b->cap = 0; }
    if (mirror_) cudaFreeHost(const_cast<ull*>(mirror_));
    cudaFree(ctr_); cudaFree(d_ent_); cudaFree(d_toff_); cudaFree(d_tbytes_);
    for (auto& ev : ev_) if (ev) cudaEventDestroy(ev);
    for (int i = 0; i < 2; i++) if (ev_d2h_[i]) cudaEventDestroy(ev_d2h_[i]);
    for (cudaEvent_t e : ev_piece_) cudaEventDestroy(e);
    if (ev_comp_) cudaEventDestroy(ev_comp_);
    cudaStreamDestroy(st_h2d_); cudaStreamDestroy(st_d2h_);
    cudaStreamDestroy(st_);
    st_ = nullptr;
  }

  int dev_ = 0, n_sm_ = N_SM_FALLBACK;
  cudaStream_t st_ = nullptr, st_h2d_ = nullptr, st_d2h_ = nullptr;
  cudaEvent_t ev_d2h_[2] = {nullptr, nullptr}, ev_comp_ = nullptr;
  std::vector<cudaEvent_t> ev_piece_;  // one per piece of a streamed call: its bytes have landed
  cudaEvent_t ev_[4] = {nullptr, nullptr, nullptr, nullptr};
  DevCounters* ctr_ = nullptr;
  volatile ull* mirror_ = nullptr;  // mapped pinned words the device publishes small results to
  ull* mirror_dev_ = nullptr;
  ull* scal_ = nullptr;
  size_t n_merges_ = 0;
  MergeEnt* d_ent_ = nullptr; MergeTable mt_{};
  ull* d_toff_ = nullptr; uint8_t* d_tbytes_ = nullptr;
  Buf b_text_, b_unit_, b_sums_, b_tok_slot_, b_wt_tag_, b_wt_first_, b_wt_len_, b_wt_bucket_, b_u_slot_, b_u_len_, b_enc_len_, b_enc_off_, b_pool_, b_off_, b_ids_;
  Buf b_offp_[2], b_idsp_[2];  // double result buffers of the streamed path
  uint64_t wt_cap_hint_ = 0;
  int32_t* d_ids_ = nullptr; ull* d_off_ = nullptr;  // views into b_ids_ / b_off_: the last result
  uint64_t n_ids_ = 0, n_tok_ = 0;
  bool have_result_ = false, debug_ = false;
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

int bpe_b200_encode_to_host(shred_encoder_t* enc, const uint8_t* text, uint64_t n_bytes, int32_t* ids_out, uint64_t ids_cap, uint64_t* offsets_out,
                            uint64_t offsets_cap, uint64_t* n_words, uint64_t* n_ids) {
  if (!enc || (n_bytes && !text) || (ids_cap && !ids_out)) return -1;
  return impl_of(enc)->encode_stream(text, n_bytes, ids_out, ids_cap, offsets_out, offsets_cap, n_words, n_ids);
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
