// tests/hostsim/engine_listsim.cpp -- TEST INFRASTRUCTURE, never part of libtrainer.so.
//
// A sequential CPU walk-through of the data structures of the CUDA engine (position-stable symbol array, per-pair
// occurrence lists validated lazily, delta aggregation, list creation for the pairs a merge creates) built on the SAME
// header the kernels use (shredword-trainer_b200/csrc/layout.hpp: probe_occurrence / rewrite_occurrence / navigation).
// It lets the layout rules and the list life cycle be checked against the reference's golden vectors without a GPU;
// what it cannot check is the kernels' concurrency.  Selected with SHRED_HOSTSIM_ENGINE=lists.
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
#include "../../shredword-trainer_b200/csrc/layout.hpp"

namespace shred {
namespace {
using namespace lay;

inline bool is_delim_b(uint8_t c) { return c == '\t' || c == '\r' || c == '\n' || c == ' '; }

class ListSimEngine : public Engine {
 public:
  int load(const uint8_t* text, size_t n, const EngineConfig& cfg, LoadInfo* info) override {
    cfg_ = cfg;
    P_.unk_id = cfg.unk_id; P_.unk_code = cfg.unk_id >= 0 ? cfg.unk_id : UNK_CODE_NEG; P_.min_freq = cfg.min_freq;
    if (n && std::memchr(text, 0, n)) return 1;
    std::unordered_map<std::string, size_t> idx;
    struct W { std::string s; uint64_t count; uint32_t bucket; size_t first; };
    std::vector<W> ws;
    size_t i = 0, ntok = 0;
    while (i < n) {
      while (i < n && is_delim_b(text[i])) i++;
      size_t s = i;
      while (i < n && !is_delim_b(text[i])) i++;
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
    ids_.clear(); wid_.clear(); wcnt_.clear(); woff_.clear();
    uint64_t S = 0;
    for (size_t wi = 0; wi < ws.size(); wi++) {
      woff_.push_back(ids_.size());
      ids_.push_back(make_hdr(static_cast<uint32_t>(wi))); wid_.push_back(static_cast<uint32_t>(wi));
      for (unsigned char c : ws[wi].s) { ids_.push_back(info->keep[c] ? static_cast<int32_t>(c) : P_.unk_code); wid_.push_back(static_cast<uint32_t>(wi)); }
      wcnt_.push_back(ws[wi].count);
      S += ws[wi].s.size();
    }
    woff_.push_back(ids_.size());
    n_slots_ = ids_.size();
    ids_.push_back(make_hdr(static_cast<uint32_t>(ws.size())));  // terminator: the last token has no right neighbour
    for (int k = 0; k < 8; k++) ids_.push_back(DEAD);
    wid_.resize(ids_.size(), 0);
    info->n_words = ws.size(); info->n_symbols = S; info->n_tokens = ntok;
    table_.clear(); lists_.clear(); pool_.clear();
    return 0;
  }

  struct Agg { int64_t delta = 0; uint64_t seq = ~0ull; std::vector<uint32_t> pos; };

  int count_pairs(const Rec** recs, size_t* n) override {
    table_.clear(); lists_.clear(); pool_.clear();
    std::unordered_map<uint64_t, Agg> agg;
    auto ld = [&](uint64_t p) { return ids_[p]; };
    for (uint64_t p = 0; p < n_slots_; p++) {  // general form: works on an already merged corpus too (bpe_train called twice)
      const int32_t v = ids_[p];
      if (v < 0) continue;
      const uint64_t q = next_start(ld, p);
      const int32_t y = ids_[q];
      if (y < 0 || v == P_.unk_code || y == P_.unk_code) continue;  // bpe.cpp:201
      Agg& a = agg[fc_key(v, y)];
      a.delta += static_cast<int64_t>(wcnt_[wid_[p]]);
      a.seq = std::min<uint64_t>(a.seq, p);
      a.pos.push_back(static_cast<uint32_t>(p));
    }
    out_.clear();
    for (auto& kv : agg) {
      Ent& e = ent(kv.first);
      e.freq = static_cast<uint64_t>(kv.second.delta);
      new_list(e.serial, kv.second.pos);
      if (e.freq >= cfg_.min_freq) out_.push_back(Rec{kv.first, e.freq, kv.second.seq, rec_pack(REC_PUSH, kv.second.pos.size()), e.serial});
    }
    std::shuffle(out_.begin(), out_.end(), rng_);
    *recs = out_.data(); *n = out_.size();
    return 0;
  }

  int merge(int32_t A, int32_t B, int32_t N, uint32_t serial, uint32_t list_len, const Rec** recs, size_t* n, uint64_t* occurrences) override {
    if (static_cast<size_t>(std::max(A, B)) >= tok_len_.size() || N < 0) return -1;
    if (static_cast<size_t>(N) >= tok_len_.size()) tok_len_.resize(N + 1, 1);
    const uint32_t lenA = tok_len_[A], lenB = tok_len_[B];
    tok_len_[N] = lenA + lenB;
    if (serial >= lists_.size() || lists_[serial].len != list_len) { std::fprintf(stderr, "listsim: bad serial/list_len %u/%u\n", serial, list_len); return -1; }
    auto it = table_.find(fc_key(A, B));
    if (it == table_.end() || it->second.serial != serial) { std::fprintf(stderr, "listsim: serial does not belong to the pair\n"); return -1; }
    const ListRef lr = lists_[serial];
    auto ld = [&](uint64_t p) { return ids_[p]; };
    auto st = [&](uint64_t p, int32_t v) { ids_[p] = v; };
    // P1: probe every list entry against the pre-merge state
    std::unordered_map<uint64_t, Agg> agg;
    std::vector<uint32_t> occ_pos;
    auto add = [&](uint64_t k, int64_t d, uint64_t seq, bool list, uint32_t pos) {
      Agg& a = agg[k];
      a.delta += d; a.seq = std::min(a.seq, seq);
      if (list && !key_has_unk(k, P_)) a.pos.push_back(pos);
    };
    for (uint64_t i = 0; i < lr.len; i++) {
      const uint64_t p = pool_[lr.off + i];
      Occ o;
      if (!probe_occurrence(ld, p, A, B, lenA, lenB, N, P_, &o)) continue;
      const int64_t c = static_cast<int64_t>(wcnt_[wid_[p]]);
      const uint64_t seq = p * 4ull;
      if (o.has_l) { add(fc_key(o.lid, A), -c, seq + 0, false, 0); add(fc_key(o.lid, N), c, seq + 1, true, o.pl); }
      if (o.has_r) { add(fc_key(B, o.rid), -c, seq + 2, false, 0); add(fc_key(N, o.rid), c, seq + 3, true, static_cast<uint32_t>(p)); }
      occ_pos.push_back(static_cast<uint32_t>(p));
    }
    // P2: fold into the pair table, records, lists for the pairs this merge creates
    out_.clear();
    for (auto& kv : agg) {
      const int32_t pa = static_cast<int32_t>(kv.first >> 32), pb = static_cast<int32_t>(kv.first & 0xFFFFFFFFu);
      if (pa == A && pb == B) continue;
      if (pa == P_.unk_id || pb == P_.unk_id) { out_.push_back(Rec{kv.first, static_cast<uint64_t>(kv.second.delta), kv.second.seq, REC_PHANTOM, REC_NO_SERIAL}); continue; }
      Ent& e = ent(kv.first);
      uint64_t& f = e.freq;
      const uint64_t old = f;
      const int64_t d = kv.second.delta;
      if (d < 0) { const uint64_t ad = static_cast<uint64_t>(-d); f = f >= ad ? f - ad : 0; } else f += static_cast<uint64_t>(d);
      if (!kv.second.pos.empty()) {
        if (lists_[e.serial].len != 0) { std::fprintf(stderr, "listsim: pair (%d,%d) received occurrences in two passes\n", pa, pb); return -1; }
        new_list(e.serial, kv.second.pos);
      }
      if (f >= cfg_.min_freq) out_.push_back(Rec{kv.first, f, kv.second.seq, rec_pack(REC_PUSH, kv.second.pos.size()), e.serial});
      else if (old >= cfg_.min_freq) out_.push_back(Rec{kv.first, f, kv.second.seq, REC_DEMOTE, e.serial});
    }
    ent(fc_key(A, B)).freq = 0;
    // P3: rewrite
    for (uint32_t p : occ_pos) rewrite_occurrence(st, p, lenA, lenB, N);
    std::shuffle(out_.begin(), out_.end(), rng_);
    *recs = out_.data(); *n = out_.size(); *occurrences = occ_pos.size();
    return 0;
  }

  template <class F>
  void walk(size_t wi, F f) {
    auto ld = [&](uint64_t p) { return ids_[p]; };
    for (uint64_t p = woff_[wi] + 1; p < woff_[wi + 1]; p = next_start(ld, p)) f(ids_[p]);
  }
  int token_freqs(uint64_t* freq, size_t T) override {
    for (size_t wi = 0; wi < wcnt_.size(); wi++) walk(wi, [&](int32_t code) { const int32_t id = code_to_id(code, P_); if (id >= 0 && static_cast<size_t>(id) < T) freq[id] += wcnt_[wi]; });
    return 0;
  }
  int word_counts(uint64_t* out) override { std::copy(wcnt_.begin(), wcnt_.end(), out); return 0; }
  int get_words(uint64_t* counts, uint64_t* off, int32_t* ids, uint64_t cap) override {
    uint64_t at = 0;
    for (size_t wi = 0; wi < wcnt_.size(); wi++) {
      if (counts) counts[wi] = wcnt_[wi];
      if (off) off[wi] = at;
      walk(wi, [&](int32_t code) { if (ids && at < cap) ids[at] = code_to_id(code, P_); at++; });
    }
    if (off) off[wcnt_.size()] = at;
    return 0;
  }
  uint64_t get_pairs(int32_t* ab, uint64_t* freq, uint64_t cap) override {
    uint64_t i = 0;
    for (auto& kv : table_) { if (i < cap) { ab[2 * i] = static_cast<int32_t>(kv.first >> 32); ab[2 * i + 1] = static_cast<int32_t>(kv.first & 0xFFFFFFFFu); freq[i] = kv.second.freq; } i++; }
    return i;
  }
  void stats(EngineStats* out) override { std::memset(out, 0, sizeof *out); out->pair_entries = table_.size(); out->n_slots = n_slots_; }
  int mark_begin() override { return 0; }
  double mark_end() override { return 0.0; }
  const char* name() override { return "listsim (CPU walk-through of the CUDA engine's data structures, tests only)"; }

 private:
  struct Ent { uint64_t freq = 0; uint32_t serial = REC_NO_SERIAL; };
  struct ListRef { uint64_t off = 0; uint32_t len = 0; };
  Ent& ent(uint64_t k) { Ent& e = table_[k]; if (e.serial == REC_NO_SERIAL) { e.serial = static_cast<uint32_t>(lists_.size()); lists_.push_back(ListRef{}); } return e; }
  void new_list(uint32_t serial, std::vector<uint32_t>& pos) {  // the kernels fill lists through atomically reserved ranks: any order
    std::shuffle(pos.begin(), pos.end(), rng_);
    lists_[serial].off = pool_.size(); lists_[serial].len = static_cast<uint32_t>(pos.size());
    pool_.insert(pool_.end(), pos.begin(), pos.end());
  }
  EngineConfig cfg_{};
  Params P_{};
  std::vector<int32_t> ids_;
  std::vector<uint32_t> wid_, pool_;
  std::vector<uint64_t> wcnt_, woff_;
  uint64_t n_slots_ = 0;
  std::vector<uint32_t> tok_len_ = std::vector<uint32_t>(256, 1);
  std::unordered_map<uint64_t, Ent> table_;
  std::vector<ListRef> lists_;
  std::vector<Rec> out_;
  std::mt19937_64 rng_{4242};
};
}  // namespace

Engine* make_listsim_engine() { return new ListSimEngine(); }

}  // namespace shred
