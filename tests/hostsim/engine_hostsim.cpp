// tests/hostsim/engine_hostsim.cpp -- TEST INFRASTRUCTURE, never part of libtrainer.so.
//
// A deliberately naive CPU stand-in for the device engine (shredword-trainer_b200/csrc/engine.hpp).  It exists so
// that the HOST control logic of the product (trainer_core.cpp: push ordering, versions, phantom pairs, exact heap
// replay) can be checked against the reference on a machine without a GPU.  It produces the same kind of records
// the CUDA kernels produce -- and shuffles them, because the kernels emit them in nondeterministic order.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/shred_abi.h"
#include "../../shredword-trainer_b200/csrc/charset.hpp"
#include "../../shredword-trainer_b200/csrc/engine.hpp"
#include "../../shredword-trainer_b200/csrc/shard.hpp"

namespace shred {

// Test-only transport for the sharded mode: an allgather of byte blobs supplied by the test (torch.distributed/gloo).
// recv receives the blobs in rank order back to back, sizes[r] their lengths.  Returns 0 on success.
typedef int (*hostsim_allgather_fn)(const void* send, uint64_t send_bytes, void* recv, uint64_t recv_cap, uint64_t* sizes);
static hostsim_allgather_fn g_allgather = nullptr;

namespace {
struct Triple { uint64_t key; int64_t delta; uint64_t seq; };

inline bool is_delim(uint8_t c) { return c == '\t' || c == '\r' || c == '\n' || c == ' '; }
inline uint64_t fc_key(int32_t a, int32_t b) { return (static_cast<uint64_t>(static_cast<int64_t>(a)) << 32) | static_cast<uint64_t>(static_cast<int64_t>(b)); }

class HostSimEngine : public Engine {
 public:
  HostSimEngine() {
    if (const char* w = std::getenv("SHRED_WORLD")) world_ = std::atoi(w) > 1 ? std::atoi(w) : 1;
    if (const char* r = std::getenv("SHRED_RANK")) rank_ = std::atoi(r);
    if (world_ == 1) rank_ = 0;
  }
  // every rank contributes its (key, delta, seq) list; returns the union (sum delta, min seq) -- identical on all ranks
  int exchange(std::unordered_map<uint64_t, std::pair<int64_t, uint64_t>>& agg) {
    if (world_ == 1) return 0;
    if (!g_allgather) { std::fprintf(stderr, "hostsim: sharded mode needs hostsim_set_allgather\n"); return -1; }
    std::vector<Triple> mine;
    for (auto& kv : agg) mine.push_back(Triple{kv.first, kv.second.first, kv.second.second});
    std::vector<uint64_t> sizes(world_);
    std::vector<uint8_t> recv(1u << 26);
    if (g_allgather(mine.data(), mine.size() * sizeof(Triple), recv.data(), recv.size(), sizes.data()) != 0) return -1;
    agg.clear();
    size_t at = 0;
    for (int r = 0; r < world_; r++) {
      const Triple* t = reinterpret_cast<const Triple*>(recv.data() + at);
      for (size_t i = 0; i < sizes[r] / sizeof(Triple); i++) {
        auto it = agg.find(t[i].key);
        if (it == agg.end()) agg.emplace(t[i].key, std::make_pair(t[i].delta, t[i].seq));
        else { it->second.first += t[i].delta; it->second.second = std::min(it->second.second, t[i].seq); }
      }
      at += sizes[r];
    }
    return 0;
  }

  int load(const uint8_t* text, size_t n, const EngineConfig& cfg, LoadInfo* info) override {
    cfg_ = cfg;
    if (n && std::memchr(text, 0, n)) return 1;
    std::unordered_map<std::string, size_t> idx;
    struct W { std::string s; uint64_t count; uint32_t bucket; size_t first; };
    std::vector<W> ws;
    size_t i = 0, ntok = 0;
    while (i < n) {
      while (i < n && is_delim(text[i])) i++;
      size_t s = i;
      while (i < n && !is_delim(text[i])) i++;
      if (i > s) {
        std::string w(reinterpret_cast<const char*>(text + s), i - s);
        ntok++;
        auto it = idx.find(w);
        if (it == idx.end()) {
          uint64_t dj = 5381;
          for (unsigned char c : w) dj = dj * 33 + c;
          idx.emplace(w, ws.size());
          ws.push_back(W{w, 1, static_cast<uint32_t>(dj & 4095), s});
        } else ws[it->second].count++;
      }
    }
    std::stable_sort(ws.begin(), ws.end(), [](const W& a, const W& b) { return a.bucket != b.bucket ? a.bucket < b.bucket : a.first < b.first; });
    std::memset(info, 0, sizeof *info);
    for (auto& w : ws) for (unsigned char c : w.s) info->hist[c]++;
    charset_keep(info->hist, cfg.coverage, info->keep, &info->n_distinct, &info->n_keep);
    words_.clear(); counts_.clear(); all_counts_.clear();
    uint64_t S = 0;
    std::vector<uint64_t> off(ws.size() + 1, 0);
    for (size_t i = 0; i < ws.size(); i++) { off[i + 1] = off[i] + ws[i].s.size() + 1; S += ws[i].s.size(); all_counts_.push_back(ws[i].count); }
    const uint64_t lo = shard_begin(off.data(), ws.size(), rank_, world_), hi = shard_begin(off.data(), ws.size(), rank_ + 1, world_);
    for (uint64_t i = lo; i < hi; i++) {  // this rank's contiguous range of words (all of them when world == 1)
      std::vector<int32_t> ids;
      for (unsigned char c : ws[i].s) ids.push_back(info->keep[c] ? static_cast<int32_t>(c) : cfg.unk_id);
      words_.push_back(std::move(ids)); counts_.push_back(ws[i].count);
    }
    info->n_words = ws.size(); info->n_symbols = S; info->n_tokens = ntok;
    table_.clear(); next_serial_ = 0;
    return 0;
  }

  int count_pairs(const Rec** recs, size_t* n) override {
    table_.clear(); next_serial_ = 0;
    std::unordered_map<uint64_t, std::pair<int64_t, uint64_t>> agg;
    uint64_t p = 0;
    for (size_t wi = 0; wi < words_.size(); wi++) {
      p++;  // header slot
      auto& s = words_[wi];
      for (size_t j = 0; j < s.size(); j++, p++) {
        if (j + 1 >= s.size() || s[j] == cfg_.unk_id || s[j + 1] == cfg_.unk_id) continue;
        uint64_t k = fc_key(s[j], s[j + 1]);
        auto it = agg.find(k);
        if (it == agg.end()) agg.emplace(k, std::make_pair(static_cast<int64_t>(counts_[wi]), seq_base(rank_) | p));
        else it->second.first += static_cast<int64_t>(counts_[wi]);
      }
    }
    if (exchange(agg) != 0) return -1;
    out_.clear();
    for (auto& kv : agg) {
      Ent& e = ent(kv.first);
      e.freq = static_cast<uint64_t>(kv.second.first);
      if (e.freq >= cfg_.min_freq) out_.push_back(Rec{kv.first, e.freq, kv.second.second, REC_PUSH, e.serial});
    }
    std::shuffle(out_.begin(), out_.end(), rng_);
    *recs = out_.data(); *n = out_.size();
    return 0;
  }

  int merge(int32_t A, int32_t B, int32_t N, uint32_t /*serial*/, uint32_t /*list_len*/, const Rec** recs, size_t* n, uint64_t* occurrences) override {
    std::unordered_map<uint64_t, std::pair<int64_t, uint64_t>> agg;
    auto add = [&](uint64_t k, int64_t d, uint64_t seq) {
      seq |= seq_base(rank_);
      auto it = agg.find(k);
      if (it == agg.end()) agg.emplace(k, std::make_pair(d, seq)); else { it->second.first += d; it->second.second = std::min(it->second.second, seq); }
    };
    uint64_t occ = 0, p = 0;
    for (size_t wi = 0; wi < words_.size(); wi++) {
      auto& s = words_[wi];
      uint64_t base = p + 1;  // flat position of the word's first symbol in the (uncompacted) layout of this pass
      p += 1 + s.size();
      int64_t c = static_cast<int64_t>(counts_[wi]);
      size_t r = 0, w = 0;
      while (r < s.size()) {
        if (r + 1 < s.size() && s[r] == A && s[r + 1] == B) {
          uint64_t seq = (base + r) * 4;
          occ++;
          if (w > 0) { add(fc_key(s[w - 1], A), -c, seq + 0); add(fc_key(s[w - 1], N), c, seq + 1); }
          if (r + 2 < s.size()) { add(fc_key(B, s[r + 2]), -c, seq + 2); add(fc_key(N, s[r + 2]), c, seq + 3); }
          s[w++] = N; r += 2;
        } else s[w++] = s[r++];
      }
      s.resize(w);
    }
    if (world_ > 1) {  // global aggregate of the deltas; the occurrence count is global too (separate exchange: any key value is legal)
      if (exchange(agg) != 0) return -1;
      std::unordered_map<uint64_t, std::pair<int64_t, uint64_t>> o;
      o.emplace(0ull, std::make_pair(static_cast<int64_t>(occ), 0ull));
      if (exchange(o) != 0) return -1;
      occ = static_cast<uint64_t>(o[0ull].first);
    }
    out_.clear();
    for (auto& kv : agg) {
      int32_t pa = static_cast<int32_t>(kv.first >> 32), pb = static_cast<int32_t>(kv.first & 0xFFFFFFFFu);
      if (pa == A && pb == B) continue;
      if (pa == cfg_.unk_id || pb == cfg_.unk_id) { out_.push_back(Rec{kv.first, static_cast<uint64_t>(kv.second.first), kv.second.second, REC_PHANTOM, REC_NO_SERIAL}); continue; }
      Ent& e = ent(kv.first);
      uint64_t& f = e.freq;
      uint64_t old = f;
      int64_t d = kv.second.first;
      if (d < 0) { uint64_t ad = static_cast<uint64_t>(-d); f = f >= ad ? f - ad : 0; } else f += static_cast<uint64_t>(d);
      if (f >= cfg_.min_freq) out_.push_back(Rec{kv.first, f, kv.second.second, REC_PUSH, e.serial});
      else if (old >= cfg_.min_freq) out_.push_back(Rec{kv.first, f, kv.second.second, REC_DEMOTE, e.serial});
    }
    ent(fc_key(A, B)).freq = 0;
    std::shuffle(out_.begin(), out_.end(), rng_);
    *recs = out_.data(); *n = out_.size(); *occurrences = occ;
    return 0;
  }

  int token_freqs(uint64_t* freq, size_t T) override {
    std::unordered_map<uint64_t, std::pair<int64_t, uint64_t>> part;
    for (size_t wi = 0; wi < words_.size(); wi++) for (int32_t id : words_[wi]) if (id >= 0 && static_cast<size_t>(id) < T) {
      auto it = part.find(static_cast<uint64_t>(id));
      if (it == part.end()) part.emplace(static_cast<uint64_t>(id), std::make_pair(static_cast<int64_t>(counts_[wi]), 0ull)); else it->second.first += static_cast<int64_t>(counts_[wi]);
    }
    if (exchange(part) != 0) return -1;
    for (auto& kv : part) freq[kv.first] += static_cast<uint64_t>(kv.second.first);
    return 0;
  }
  int word_counts(uint64_t* out) override { std::copy(all_counts_.begin(), all_counts_.end(), out); return 0; }
  int get_words(uint64_t* counts, uint64_t* off, int32_t* ids, uint64_t cap) override {
    uint64_t at = 0;
    for (size_t wi = 0; wi < words_.size(); wi++) {
      if (counts) counts[wi] = counts_[wi];
      if (off) off[wi] = at;
      for (int32_t id : words_[wi]) { if (ids && at < cap) ids[at] = id; at++; }
    }
    if (off) off[words_.size()] = at;
    return 0;
  }
  uint64_t get_pairs(int32_t* ab, uint64_t* freq, uint64_t cap) override {
    uint64_t i = 0;
    for (auto& kv : table_) { if (i < cap) { ab[2 * i] = static_cast<int32_t>(kv.first >> 32); ab[2 * i + 1] = static_cast<int32_t>(kv.first & 0xFFFFFFFFu); freq[i] = kv.second.freq; } i++; }
    return i;
  }
  void stats(EngineStats* out) override { std::memset(out, 0, sizeof *out); out->pair_entries = table_.size(); }
  int mark_begin() override { return 0; }
  double mark_end() override { return 0.0; }
  const char* name() override { return "hostsim (CPU stand-in, tests only)"; }

 private:
  EngineConfig cfg_{};
  std::vector<std::vector<int32_t>> words_;
  std::vector<uint64_t> counts_, all_counts_;
  int rank_ = 0, world_ = 1;
  struct Ent { uint64_t freq = 0; uint32_t serial = REC_NO_SERIAL; };
  std::unordered_map<uint64_t, Ent> table_;
  Ent& ent(uint64_t k) { Ent& e = table_[k]; if (e.serial == REC_NO_SERIAL) e.serial = next_serial_++; return e; }
  uint32_t next_serial_ = 0;
  std::vector<Rec> out_;
  std::mt19937_64 rng_{12345};
};
}  // namespace

Engine* make_listsim_engine();  // engine_listsim.cpp
Engine* make_device_engine() {
  const char* e = std::getenv("SHRED_HOSTSIM_ENGINE");
  if (e && std::string(e) == "lists") return make_listsim_engine();
  return new HostSimEngine();
}

}  // namespace shred

extern "C" const char* bpe_b200_device_name(void) { return "hostsim"; }
extern "C" __attribute__((visibility("default"))) void hostsim_set_allgather(shred::hostsim_allgather_fn fn) { shred::g_allgather = fn; }
