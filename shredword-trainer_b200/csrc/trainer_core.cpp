// trainer_core.cpp -- sequential host control of the B200 BPE trainer.
//
// What the reference does per merge (shredword/csrc/bpe/bpe.cpp:232-323) splits into
//   * data-parallel work over the whole word table (find and rewrite every occurrence, collect the +/- count deltas,
//     update the pair table)  ->  Engine::merge, CUDA kernels in cuda/engine_cuda.cu;
//   * strictly sequential work that fixes the *order* of merges: the lazy max-heap (heap.cpp:53-114), the order in
//     which changed pairs are pushed (FreqChangeMap iteration, bpe.cpp:9-38,297-313), version-based invalidation
//     (bpe.cpp:247-250) and the knock-out of pairs that touch unk_id (recompute_freq, bpe.cpp:52-53,251-257).
// This file is the second part.  It receives the touched pair keys UNORDERED from the engine, each with the sequence
// number of its first sighting in reference scan order, re-creates the reference's push order from that
//  and replays the heap exactly.
//
// Equivalences used (each checked against the unmodified reference by tests/test_host_logic.py):
//   * recompute_freq (bpe.cpp:52-65) returns the table frequency for every pair without unk_id, because the deltas
//     keep the table exact; only "contains unk_id => 0" is observable.  Pairs containing unk_id ("phantoms",
//     Appendix A12) are tracked here on the host with the reference's clamped arithmetic.
//   * A pair whose frequency drops below min_pair_freq keeps its version in the reference and its heap entry is
//     dropped when popped (bpe.cpp:258).  Bumping the host-side version at the moment of the drop makes the same
//     entry stale instead; either way the pop has no side effect.
#include "trainer_core.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

// NVTX ranges per phase: only in the product build (build.py defines SHRED_NVTX; nvtx3 is header-only and
// costs a null-pointer check per call unless a profiler has injected itself)
#ifdef SHRED_NVTX
#include <nvtx3/nvToolsExt.h>
namespace { struct NvtxRange { explicit NvtxRange(const char* n) { nvtxRangePushA(n); } ~NvtxRange() { nvtxRangePop(); } }; }
#define SHRED_RANGE(name) NvtxRange shred_nvtx_range_(name)
#else
#define SHRED_RANGE(name) do { } while (0)
#endif

namespace shred {

static inline double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static inline uint64_t pack_key(int32_t a, int32_t b) { return (static_cast<uint64_t>(static_cast<uint32_t>(a)) << 32) | static_cast<uint32_t>(b); }
static inline PairKey unpack_key(uint64_t k) {  // bpe.cpp:301
  PairKey p; p.first = static_cast<int32_t>(k >> 32); p.second = static_cast<int32_t>(k & 0xFFFFFFFFu); return p;
}

static inline uint32_t fnv1a_pair(PairKey k) {  // hash.cpp:7-16 over the 8 little-endian bytes of {first, second}
  uint32_t h = 2166136261u, w[2] = {static_cast<uint32_t>(k.first), static_cast<uint32_t>(k.second)};
  for (int j = 0; j < 2; j++) for (int i = 0; i < 4; i++) { h ^= (w[j] >> (8 * i)) & 255u; h *= 16777619u; }
  return h;
}

TrainerCore::TrainerCore(Trainer* abi, Engine* eng) : abi_(abi), eng_(eng) {
  std::memset(&info_, 0, sizeof info_);
  std::memset(&placeholder_, 0, sizeof placeholder_);
  merge_cap_ = abi_->config.target_vocab_size ? abi_->config.target_vocab_size : 1;
  abi_->merge_ops = static_cast<PairKey*>(std::calloc(merge_cap_, sizeof(PairKey)));  // bpe.cpp:81
  abi_->num_merges = 0;
  const char* e = std::getenv("SHRED_LOG_MERGES");
  log_merges_ = e && *e && *e != '0';
  e = std::getenv("SHRED_QUIET");
  quiet_ = e && *e && *e != '0';
  e = std::getenv("SHRED_HEAP_TRACE");
  if (e && *e) { trace_file_ = std::fopen(e, "wb"); heap_.set_trace(trace_file_); }
  sync_mirrors();
}

TrainerCore::~TrainerCore() {
  const char* dbg = std::getenv("SHRED_DEBUG_TIMING");
  const bool timing = dbg && *dbg && *dbg != '0';
  const double t0 = now_ms();
  if (trace_file_) std::fclose(trace_file_);
  std::free(abi_->merge_ops); abi_->merge_ops = nullptr;
  std::free(abi_->corpus.words); abi_->corpus.words = nullptr;
  std::free(abi_->corpus.word_counts); abi_->corpus.word_counts = nullptr;
  const double t1 = now_ms();
  delete eng_;
  if (timing) std::fprintf(stderr, "[TIMING]\t destroy: host mirrors %.1f ms, engine %.1f ms (the heap and version tables follow)\n", t1 - t0, now_ms() - t1);
}

void TrainerCore::sync_mirrors() {  // Trainer.heap in the reference's layout (rebuilt only if the heap changed)
  size_t cap = 0;
  abi_->heap.data = heap_.materialize(&cap);
  abi_->heap.size = heap_.size();
  abi_->heap.cap = cap;
}

// ---------------------------------------------------------------------------------------------- load (bpe.cpp:110-185)

int TrainerCore::load_file(const char* path) {
  SHRED_RANGE("bpe_load_corpus");
  int fd = ::open(path, O_RDONLY);
  if (fd < 0) { std::fprintf(stderr, "[ERROR]\t Couldn't open file: %s\n", path); return -1; }  // bpe.cpp:118-122
  struct stat st;
  if (fstat(fd, &st) != 0) { ::close(fd); return -1; }
  const size_t n = static_cast<size_t>(st.st_size);
  int rc;
  if (n == 0) {
    rc = load_buffer(reinterpret_cast<const uint8_t*>(""), 0);
  } else {
    // fast path: the engine streams the file itself (pinned staging ring, read and PCIe copy overlapped)
    const double t0 = now_ms();
    EngineConfig ec = engine_config();
    rc = eng_->load_file(fd, n, ec, &info_);
    if (rc == 0) rc = finish_load(n, t0);
    else if (rc > 0) {  // 1: NUL bytes (needs the host-side blanking pass), 2: engine without a file path -> map the file
      void* p = mmap(nullptr, n, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
      if (p == MAP_FAILED) { std::fprintf(stderr, "[ERROR]\t Couldn't map file: %s\n", path); ::close(fd); return -1; }
      madvise(p, n, MADV_SEQUENTIAL);
      rc = load_buffer(static_cast<const uint8_t*>(p), n);
      munmap(p, n);
    }
  }
  ::close(fd);
  return rc;
}

// The reference reads the file with fgets into a growing buffer and measures each line with strlen
// (bpe.cpp:129-147), so a NUL byte hides the rest of its fgets chunk.  This replays that loop over the in-memory
// image and blanks every byte the reference would not see; the result has no NUL and tokenises identically.
static std::vector<uint8_t> blank_hidden_spans(const uint8_t* text, size_t n) {
  std::vector<uint8_t> out(n, static_cast<uint8_t>(' '));
  size_t cap = 4096, pos = 0;  // bpe.cpp:123,129
  while (pos < n) {
    const size_t start = pos;
    size_t raw = 0;
    while (raw < cap - 1 && pos < n) { uint8_t c = text[pos++]; raw++; if (c == '\n') break; }  // fgets(line, cap)
    const void* z = std::memchr(text + start, 0, raw);
    size_t len = z ? static_cast<size_t>(static_cast<const uint8_t*>(z) - (text + start)) : raw;  // strlen
    while (len == cap - 1 && text[start + len - 1] != '\n') {  // bpe.cpp:133-145
      cap *= 2;
      if (pos >= n) break;
      size_t room = cap - len, got = 0;
      while (got < room - 1 && pos < n) { uint8_t c = text[pos++]; got++; if (c == '\n') break; }
      const void* z2 = std::memchr(text + start + len, 0, got);
      len = z2 ? static_cast<size_t>(static_cast<const uint8_t*>(z2) - (text + start)) : len + got;
    }
    std::memcpy(out.data() + start, text + start, len);
  }
  return out;
}

EngineConfig TrainerCore::engine_config() const {
  EngineConfig ec;
  ec.unk_id = abi_->config.unk_id; ec.coverage = abi_->config.character_coverage; ec.min_freq = abi_->config.min_pair_freq;
  ec.vocab_size = abi_->config.target_vocab_size;
  return ec;
}

int TrainerCore::load_buffer(const uint8_t* text, size_t n) {
  SHRED_RANGE("bpe_b200_load_buffer");
  double t0 = now_ms();
  EngineConfig ec = engine_config();
  int rc = eng_->load(text, n, ec, &info_);
  if (rc == 1) {  // NUL bytes present
    std::vector<uint8_t> visible = blank_hidden_spans(text, n);
    rc = eng_->load(visible.data(), n, ec, &info_);
  }
  if (rc != 0) return -1;
  return finish_load(n, t0);
}

// host mirrors of Corpus (bpe.h:37-41) after a successful engine load
int TrainerCore::finish_load(size_t n, double t0) {
  corpus_bytes_ = n;
  loaded_ = true;
  // counts are real, words[] are non-NULL placeholders (the symbols live in HBM)
  std::free(abi_->corpus.words); std::free(abi_->corpus.word_counts);
  size_t N = info_.n_words;
  abi_->corpus.vocab_size = N;
  abi_->corpus.words = static_cast<Symbol**>(std::malloc((N ? N : 1) * sizeof(Symbol*)));
  abi_->corpus.word_counts = static_cast<uint64_t*>(std::malloc((N ? N : 1) * sizeof(uint64_t)));
  if (!abi_->corpus.words || !abi_->corpus.word_counts) return -1;
  // two 8-byte-per-word host arrays (131 MB each at 10 GB): fill one on a helper thread while the counts come back from the device
  Symbol** words = abi_->corpus.words;
  Symbol* ph = &placeholder_;
  std::thread filler([words, ph, N]() { for (size_t i = 0; i < N; i++) words[i] = ph; });
  const int wrc = N ? eng_->word_counts(abi_->corpus.word_counts) : 0;
  filler.join();
  if (wrc != 0) return -1;
  // bpe.cpp:183 re-initialises the pair table; the heap is left alone (it is reset by bpe_init)
  version_.clear(); phantom_.clear(); ver_.clear();
  load_wall_ms_ = now_ms() - t0;
  if (!quiet_) std::printf("[DEBUG]\t Character histogram built with %u unique characters.\n", info_.n_distinct);  // bpe.cpp:168
  return 0;
}

// ------------------------------------------------------------------------------------- count (bpe.cpp:187-230), init

void TrainerCore::count_bigrams() {
  SHRED_RANGE("bpe_count_bigrams");
  if (!loaded_) { sync_mirrors(); return; }
  const Rec* recs = nullptr; size_t n = 0;
  if (!quiet_) std::printf("[INFO]\t Counting bigrams from %zu words...\n", static_cast<size_t>(info_.n_words));
  if (eng_->count_pairs(&recs, &n) != 0) { std::fprintf(stderr, "[ERROR]\t device bigram count failed\n"); return; }
  version_.clear(); phantom_.clear(); ver_.clear();
  // BIMap iteration order (bpe.cpp:219-227): bucket = fnv1a32(pair) & 4095 ascending, chain = creation order, and a
  // pair is created at its first sighting in scan order (hash.cpp:126-129)  ->  sort by (bucket, seq).
  order_.assign(recs, recs + n);
  std::sort(order_.begin(), order_.end(), [](const Rec& x, const Rec& y) {
    uint32_t bx = fnv1a_pair(unpack_key(x.key)) & 4095u, by = fnv1a_pair(unpack_key(y.key)) & 4095u;
    if (bx != by) return bx < by;
    return x.seq < y.seq;
  });
  for (const Rec& r : order_) {
    ver_of(r.serial, r.key) = 0;
    if (r.serial != REC_NO_SERIAL) meta_of(r.serial).list_len = rec_list_len(r.kind);
    heap_.push(unpack_key(r.key), r.val, 0, r.serial);
  }
  if (!quiet_) std::printf("[INFO]\t Added %zu pairs to heap (freq >= %llu)\n", n, static_cast<unsigned long long>(abi_->config.min_pair_freq));
  sync_mirrors();
}

void TrainerCore::init() {  // bpe.cpp:98-108
  heap_.clear();
  count_bigrams();
}

// ------------------------------------------------------------------------------------------ merge (bpe.cpp:232-323)

void TrainerCore::apply_records(const Rec* recs, size_t n) {
  // FreqChangeMap iteration (bpe.cpp:297-298): bucket = key % 1024 ascending; inside a bucket the chain is LIFO by
  // first insertion (prepend at bpe.cpp:36-37)  ->  sort by (key & 1023, seq descending).
  // Bucket the records by key & 1023 with intrusive lists (counting sort: a few hundred records over 1024 buckets),
  // then order each bucket's handful of records by descending sequence.
  if (bucket_head_.empty()) bucket_head_.assign(1024, -1);
  next_in_bucket_.resize(n);
  uint64_t used[16] = {0};  // non-empty buckets: a merge touches ~55 of the 1024, so walk set bits instead of all heads
  for (size_t i = 0; i < n; i++) {
    if (recs[i].serial == REC_NO_SERIAL) version_.prefetch(recs[i].key); else if (recs[i].serial < ver_.size()) __builtin_prefetch(&ver_[recs[i].serial]);
    const uint32_t b = static_cast<uint32_t>(recs[i].key & 1023u);
    next_in_bucket_[i] = bucket_head_[b];
    bucket_head_[b] = static_cast<int32_t>(i);
    used[b >> 6] |= 1ull << (b & 63u);
  }
  order_idx_.clear();
  for (uint32_t w = 0; w < 16; w++) for (uint64_t bits = used[w]; bits; bits &= bits - 1) {
    const uint32_t b = (w << 6) + static_cast<uint32_t>(__builtin_ctzll(bits));
    int32_t i = bucket_head_[b];
    bucket_head_[b] = -1;
    const size_t start = order_idx_.size();
    for (; i >= 0; i = next_in_bucket_[i]) {  // insertion sort, descending sequence
      size_t at = order_idx_.size();
      order_idx_.push_back(static_cast<uint32_t>(i));
      while (at > start && recs[order_idx_[at - 1]].seq < recs[i].seq) { order_idx_[at] = order_idx_[at - 1]; --at; }
      order_idx_[at] = static_cast<uint32_t>(i);
    }
  }
  const uint64_t min_freq = abi_->config.min_pair_freq;
  for (const uint32_t ri : order_idx_) {
    const Rec& r = recs[ri];
    PairKey pk = unpack_key(r.key);
    switch (rec_kind(r.kind)) {
      case REC_PUSH: {  // bpe.cpp:308-311
        if (rec_list_len(r.kind) && r.serial != REC_NO_SERIAL) meta_of(r.serial).list_len = rec_list_len(r.kind);  // the pass that created the pair
        uint32_t v = ++ver_of(r.serial, r.key);
        heap_.push(pk, r.val, v, r.serial);
        break;
      }
      case REC_DEMOTE:  // bpe.cpp:258 seen from the other side: invalidate now instead of discarding at pop
        ++ver_of(r.serial, r.key);
        break;
      case REC_PHANTOM: {  // bpe.cpp:303-311 on a pair the device table does not hold
        uint64_t& f = phantom_[r.key];
        int64_t d = static_cast<int64_t>(r.val);
        if (d < 0) { uint64_t ad = static_cast<uint64_t>(-d); f = f >= ad ? f - ad : 0; } else { f += static_cast<uint64_t>(d); }
        if (f >= min_freq) { uint32_t v = ++version_[r.key]; heap_.push(pk, f, v, REC_NO_SERIAL); }
        break;
      }
      default: break;
    }
  }
}

int TrainerCore::merge_batch(int batch_size) {  // the ABI's step-wise entry: C consumers read Trainer.heap after every call
  heap_.set_tracking(true);
  return merge_loop(batch_size);
}

int TrainerCore::merge_loop(int batch_size) {
  SHRED_RANGE("bpe_merge_batch");
  if (heap_.empty()) {  // bpe.cpp:237-240
    if (!quiet_) std::printf("[INFO]\t Heap is empty, no more merges possible\n");
    return 0;
  }
  int done = 0;
  struct MergeRun { Engine* e; explicit MergeRun(Engine* x) : e(x) { e->begin_merges(); } ~MergeRun() { e->end_merges(); } } run(eng_);
  // host time between two device merges (pops until a current entry, then the records of the merge): two clock reads per
  // merge -- one per pop (~56 stale pops per merge) cost 2.5 us per merge in clock calls alone
  double h_open = now_ms();
  while (done < batch_size && !heap_.empty()) {
    {  // start the version lookup of the entry about to be popped: its miss overlaps the sift-down
      const HeapEnt t = heap_.top();
      if (t.serial == REC_NO_SERIAL) version_.prefetch(pack_key(t.key.first, t.key.second)); else if (t.serial < ver_.size()) __builtin_prefetch(&ver_[t.serial]);
    }
    HeapEnt top = heap_.pop();
    const uint64_t k = pack_key(top.key.first, top.key.second);
    uint32_t cur;
    if (top.serial == REC_NO_SERIAL) { uint32_t* vp = version_.find(k); cur = vp ? *vp : 0; } else cur = top.serial < ver_.size() ? ver_[top.serial].ver : 0;
    if (top.version != cur) continue;  // stale, bpe.cpp:247-250
    if (is_phantom(top.key.first, top.key.second)) {  // recompute_freq == 0, bpe.cpp:53,252-257
      uint64_t* f = phantom_.find(k);
      if (f && *f != 0) { *f = 0; ++version_[k]; }
      continue;
    }
    { const double dt = now_ms() - h_open; host_heap_ms_ += dt; host_pop_ms_ += dt; }
    // tie statistics: a device argmax could only pick this merge if no other mergeable pair shares its
    // frequency.  Upper bound: the entry now at the root has the same frequency (it may be stale); lower bound: the previous
    // merge had the same frequency.
    if (!heap_.empty() && heap_.top().freq == top.freq) ++tie_root_equal_;
    if (top.freq == last_merge_freq_) ++tie_same_as_prev_;
    last_merge_freq_ = top.freq;
    // a current entry of a pair without unk_id carries the exact table frequency, which is >= min_pair_freq
    const int32_t new_id = static_cast<int32_t>(256 + abi_->num_merges);  // bpe.cpp:259
    if (abi_->num_merges < merge_cap_) abi_->merge_ops[abi_->num_merges] = top.key;  // bpe.cpp:261
    const Rec* recs = nullptr; size_t n = 0; uint64_t occ = 0;
    const uint32_t list_len = top.serial != REC_NO_SERIAL && top.serial < ver_.size() ? ver_[top.serial].list_len : 0;
    if (eng_->merge(top.key.first, top.key.second, new_id, top.serial, list_len, &recs, &n, &occ) != 0) {
      std::fprintf(stderr, "[ERROR]\t device merge failed\n");
      sync_mirrors();
      return -1;
    }
    if (log_merges_) std::printf("[MERGE]\t Merging (%d,%d) freq=%llu -> new_id=%d (merge %zu)\n", top.key.first, top.key.second,
                                 static_cast<unsigned long long>(top.freq), new_id, abi_->num_merges + 1);
    h_open = now_ms();
    apply_records(recs, n);
    ++ver_of(top.serial, k);  // bpe.cpp:315-316
    { const double t_applied = now_ms(); host_apply_ms_ += t_applied - h_open; host_heap_ms_ += t_applied - h_open; h_open = t_applied; }
    occurrences_ += occ;
    abi_->num_merges++;
    done++;
  }
  host_heap_ms_ += now_ms() - h_open;
  sync_mirrors();
  return done;
}

int TrainerCore::train() {  // bpe.cpp:345-386
  SHRED_RANGE("bpe_train");
  double t0 = now_ms();
  eng_->mark_begin();
  host_heap_ms_ = 0; host_pop_ms_ = 0; host_apply_ms_ = 0; occurrences_ = 0;
  heap_.pushes = heap_.pops = 0;
  tie_root_equal_ = tie_same_as_prev_ = 0; last_merge_freq_ = ~0ull;
  heap_.set_tracking(false);  // one rebuild of the Trainer.heap mirror at the end instead of noting every written slot
  if (!quiet_) std::printf("[INFO]\t Starting BPE training (target vocab size: %zu)\n", abi_->config.target_vocab_size);
  init();
  int total = 0;
  const int target = static_cast<int>(abi_->config.target_vocab_size) - 256;  // bpe.cpp:353
  if (!quiet_) std::printf("[INFO]\t Need to perform %d merges to reach target vocab size\n", target);
  while (total < target) {
    if (heap_.empty()) { if (!quiet_) std::printf("[INFO]\t Heap exhausted, stopping training\n"); break; }
    // the reference sizes its batches from the top frequency (bpe.cpp:362-368); batch boundaries have no effect on the
    // result, so one call covers the remaining merges
    int merged = merge_loop(target - total);
    if (merged <= 0) { if (!quiet_) std::printf("[WARNING]\t No merges performed, stopping\n"); break; }
    total += merged;
  }
  train_device_ms_ = eng_->mark_end();
  train_wall_ms_ = now_ms() - t0;
  merges_last_ = static_cast<uint64_t>(total);
  if (!quiet_) std::printf("[INFO]\t Training completed. Performed %d merges\n", total);
  if (const char* dbg = std::getenv("SHRED_DEBUG_TIMING")) if (*dbg && *dbg != '0') {
    EngineStats es; std::memset(&es, 0, sizeof es); eng_->stats(&es);
    std::fprintf(stderr, "[TIMING]\t train %.1f ms (device %.1f): engine.merge %.1f (launch %.1f, wait %.1f), host heap %.1f (pops until a current entry %.1f, records applied %.1f), merges %d, pushes %llu pops %llu\n",
                 train_wall_ms_, train_device_ms_, es.merge_ms, es.launch_ms, es.wait_ms, host_heap_ms_, host_pop_ms_, host_apply_ms_, total,
                 static_cast<unsigned long long>(heap_.pushes), static_cast<unsigned long long>(heap_.pops));
  }
  sync_mirrors();
  return total;
}

// ------------------------------------------------------------------------------------------- save (bpe.cpp:388-432)

void TrainerCore::save(const char* model_path, const char* vocab_path) {
  SHRED_RANGE("bpe_save");
  double t0 = now_ms();
  size_t M = abi_->num_merges < merge_cap_ ? abi_->num_merges : merge_cap_;
  size_t T = 256 + M;
  // token strings by C-string concatenation (bpe.cpp:395-408): token 0 is the empty string
  std::vector<std::string> tok(T);
  for (size_t i = 1; i < 256; i++) tok[i].assign(1, static_cast<char>(i));
  for (size_t m = 0; m < M; m++) {
    PairKey op = abi_->merge_ops[m];
    const std::string empty;
    const std::string& a = (op.first >= 0 && static_cast<size_t>(op.first) < 256 + m) ? tok[op.first] : empty;
    const std::string& b = (op.second >= 0 && static_cast<size_t>(op.second) < 256 + m) ? tok[op.second] : empty;
    tok[256 + m] = a + b;
  }
  std::vector<uint64_t> freq(T, 0);
  if (loaded_ && eng_->token_freqs(freq.data(), T) != 0) std::fprintf(stderr, "[ERROR]\t device token count failed\n");
  FILE* vf = std::fopen(vocab_path, "w");
  if (vf) {
    for (size_t i = 0; i < T; i++) {
      std::fwrite(tok[i].data(), 1, tok[i].size(), vf);
      std::fprintf(vf, " %llu\n", static_cast<unsigned long long>(freq[i]));
    }
    std::fclose(vf);
  } else std::fprintf(stderr, "[ERROR]\t Couldn't open file: %s\n", vocab_path);
  FILE* mf = std::fopen(model_path, "wb");
  if (mf) {
    for (size_t m = 0; m < M; m++) {
      int32_t rec[3] = {abi_->merge_ops[m].first, abi_->merge_ops[m].second, static_cast<int32_t>(256 + m)};
      std::fwrite(rec, sizeof(int32_t), 3, mf);
    }
    std::fclose(mf);
  } else std::fprintf(stderr, "[ERROR]\t Couldn't open file: %s\n", model_path);
  save_wall_ms_ = now_ms() - t0;
  if (!quiet_) std::printf("[INFO]\tSaved %zu-token vocab to %s and %zu merges to %s\n", T, vocab_path, M, model_path);
}

void TrainerCore::get_stats(shred_stats_t* s) {
  std::memset(s, 0, sizeof *s);
  EngineStats es; std::memset(&es, 0, sizeof es);
  eng_->stats(&es);
  s->n_words = info_.n_words; s->n_symbols_initial = info_.n_symbols; s->n_tokens = info_.n_tokens; s->corpus_bytes = corpus_bytes_;
  s->n_symbols_live = es.n_symbols_live; s->n_slots = es.n_slots; s->pair_entries = es.pair_entries;
  s->list_entries = es.list_entries; s->pool_entries = es.pool_entries; s->fill_device_ms = es.fill_device_ms; s->fill_bytes = es.fill_bytes;
  s->heap_size = heap_.size(); s->heap_pushes = heap_.pushes; s->heap_pops = heap_.pops;
  s->merges = merges_last_; s->occurrences = occurrences_;
  s->scan_launches = es.scan_launches; s->scan_device_ms = es.scan_device_ms; s->scan_bytes = es.scan_bytes;
  s->scan_bytes_touched = es.scan_bytes_touched; s->dense_launches = es.dense_launches; s->dense_device_ms = es.dense_device_ms; s->dense_bytes = es.dense_bytes;
  s->scan_phase_ms = es.scan_phase_ms; s->dense_phase_ms = es.dense_phase_ms;
  s->count_launches = es.count_launches; s->count_device_ms = es.count_device_ms; s->count_bytes = es.count_bytes;
  s->ingest_launches = es.ingest_launches; s->ingest_device_ms = es.ingest_device_ms; s->ingest_bytes = es.ingest_bytes;
  s->kernel_launches = es.kernel_launches;
  s->load_wall_ms = load_wall_ms_; s->h2d_ms = es.h2d_ms; s->train_wall_ms = train_wall_ms_; s->train_device_ms = train_device_ms_; s->host_heap_ms = host_heap_ms_;
  s->wait_ms = es.wait_ms; s->launch_ms = es.launch_ms; s->save_wall_ms = save_wall_ms_;
  s->h2d_bytes = es.h2d_bytes; s->d2h_bytes = es.d2h_bytes;
  s->tie_root_equal = tie_root_equal_; s->tie_same_as_prev = tie_same_as_prev_;
  s->fold_phase_ms = es.fold_phase_ms; s->rewrite_phase_ms = es.rewrite_phase_ms; s->single_launches = es.single_launches; s->server_merges = es.server_merges; s->server_starts = es.server_starts;
}

}  // namespace shred
