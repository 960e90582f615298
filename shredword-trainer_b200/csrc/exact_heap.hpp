// exact_heap.hpp -- the replay heap.
//
// The reference picks the next merge with an array binary max-heap that compares `freq` only
// (reference shredword/csrc/bpe/heap.cpp:53-114), so which of several equal-frequency pairs wins is decided by
// the heap's structural history.  Bit-exact merge lists therefore require replaying the same sift rules on the same
// push/pop sequence.
//
// Same algorithm, different memory layout.  The sift decisions only ever read frequencies, and a pop is a chain of ~21
// dependent cache accesses through an array of millions of entries (6 M pushes and 2.6 M pops at the 10 GB configuration, 80 %
// of the host's time per merge).  So an entry of the heap array is ONE 8-byte word, frequency << shift | entry id, and the
// payloads (pair, version, serial: 16 bytes) sit in an append-only side array indexed by the id: sifting moves 8 bytes where
// the reference moves 24, the whole array is a third of the reference's (it stays in the last-level cache), and a payload is
// touched exactly twice -- when it is pushed (sequential write) and when it reaches the root.  Replaying the recorded
// operation trace of the 10 GB configuration (tests/bench_heap.cpp): 25 % less time per pop than with separate frequency and
// payload arrays, 45 % less than with the reference's 24-byte entries.
// Storage is 1-based (the reference's entry i lives in slot i + 1) so that the 2^d descendants d levels below a node form one
// aligned block: the walk prefetches the 16 great-great-grandchildren (two cache lines) four levels ahead.
// The split of the word adapts: ids are renumbered (live entries only) when they run out, and the shift shrinks if a frequency
// ever needs more bits.  materialize() writes the reference's 24-byte array-of-structs (heap.h:17-21) for C consumers that
// read Trainer.heap.data.
#pragma once
#include <sys/mman.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/shred_abi.h"

namespace shred {

struct HeapPayload { PairKey key; uint32_t version; uint32_t serial; };  // serial: dense device-side id of the pair (engine.hpp)
struct HeapEnt { PairKey key; uint64_t freq; uint32_t version; uint32_t serial; };
static_assert(sizeof(HeapPayload) == 16, "payload packs into 16 bytes");

class ExactHeap {
 public:
  ~ExactHeap() { std::free(w_); std::free(f_); std::free(pay_); }
  void clear() { n_ = 0; np_ = 0; dirty_ = true; touched_.clear(); }
  // Step-wise C consumers read Trainer.heap.data after every bpe_merge_batch call: while tracking is on, push/pop note the
  // slots they write so that materialize() patches those few entries instead of rebuilding the whole mirror (O(heap)).
  void set_tracking(bool on) {
    if (on == track_) return;
    track_ = on;
    touched_.clear();
    dirty_ = true;  // the slots written while the mode changed are unknown: next materialize() rebuilds once
  }
  size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }
  HeapEnt top() const { return entry(1); }

  // heap.cpp:70-79: append, then swap upwards while the parent's freq is strictly smaller.
  void push(PairKey key, uint64_t freq, uint32_t version, uint32_t serial) {
    if (trace_) { const uint64_t t = freq; std::fwrite(&t, 8, 1, trace_); }
    if (n_ + 2 >= cap_) grow();
    if (!wide_ && (np_ >= (1ull << sh_) || (freq >> (64 - sh_)) != 0)) repack(freq);
    if (np_ + 1 >= pcap_) grow_pay();
    pay_[np_] = HeapPayload{key, version, serial};
    if (wide_) push_impl<true>(freq); else push_impl<false>(freq);
    ++np_;
    ++pushes;
  }

  // heap.cpp:97-111: last entry to the root, then swap with the left child if it is strictly larger, with the right
  // child if it is strictly larger than the better of the two, until neither is -- i.e. the hole sinks to the larger child
  // (the left one on a tie) as long as that child is strictly larger than the entry being placed.
  HeapEnt pop() {
    if (trace_) { const uint64_t t = ~0ull; std::fwrite(&t, 8, 1, trace_); }
    const HeapEnt out = top();
    if (wide_) pop_impl<true>(); else pop_impl<false>();
    if (!track_) dirty_ = true;
    ++pops;
    // the payload side array only grows between renumberings: renumber when dead payloads outnumber the live ones 3:1
    if (np_ > 4 * n_ + (1u << 16)) repack(0);
    return out;
  }

  // The reference's array of 24-byte entries, rebuilt only when the heap changed since the last call.
  BPEHeapEntry* materialize(size_t* cap_out) {
    if (dirty_ || !mirror_valid_) {
      mirror_.resize(n_ ? n_ : 1);
      HeapEnt* m = reinterpret_cast<HeapEnt*>(mirror_.data());
      for (size_t i = 0; i < n_; i++) m[i] = entry(i + 1);
    } else if (!touched_.empty()) {  // patch only the slots written since the last call
      if (mirror_.size() < (n_ ? n_ : 1)) mirror_.resize(std::max(n_ ? n_ : 1, mirror_.size() * 2));
      HeapEnt* m = reinterpret_cast<HeapEnt*>(mirror_.data());
      for (const size_t s : touched_) if (s <= n_) m[s - 1] = entry(s);
    }
    dirty_ = false; mirror_valid_ = true; touched_.clear();
    if (cap_out) *cap_out = mirror_.capacity();
    return mirror_.data();
  }

  uint64_t pushes = 0, pops = 0;
  // development aid: SHRED_HEAP_TRACE=<file> records the operation sequence (push: the frequency, pop: ~0, 8 bytes each), which
  // is all the heap's structure depends on; tests/bench_heap.cpp replays it to time heap layouts in isolation
  void set_trace(FILE* f) { trace_ = f; }

 private:
  uint64_t id_mask() const { return (1ull << sh_) - 1; }
  // WIDE = false: w_[s] = freq << sh_ | id.  WIDE = true (frequencies and ids that do not fit one word together): w_[s] = id, f_[s] = freq.
  template <bool WIDE> uint64_t freq_at(size_t s) const { return WIDE ? f_[s] : w_[s] >> sh_; }
  uint64_t id_at(size_t s) const { return wide_ ? w_[s] : w_[s] & id_mask(); }
  HeapEnt entry(size_t s) const {
    const HeapPayload& p = pay_[id_at(s)];
    return HeapEnt{p.key, wide_ ? f_[s] : w_[s] >> sh_, p.version, p.serial};
  }
  template <bool WIDE> void push_impl(uint64_t freq) {
    size_t s = ++n_;  // slot of the reference's index n_-1
    while (s > 1) {
      const size_t p = s >> 1;
      if (freq_at<WIDE>(p) >= freq) break;
      w_[s] = w_[p];
      if (WIDE) f_[s] = f_[p];
      if (track_) note(s);
      s = p;
    }
    w_[s] = WIDE ? np_ : ((freq << sh_) | np_);
    if (WIDE) f_[s] = freq;
    if (track_) note(s); else dirty_ = true;
  }
  template <bool WIDE> void pop_impl() {
    const uint64_t x = w_[n_], xf = freq_at<WIDE>(n_);
    --n_;
    size_t s = 1;
    for (;;) {
      const size_t l = s << 1;
      if (l > n_) break;
      const size_t g = s << 4;  // the 16 descendants four levels down: one aligned 128-byte block
      if (g <= n_) { const uint64_t* q = WIDE ? f_ : w_; __builtin_prefetch(q + g); __builtin_prefetch(q + g + 8); }
      const uint64_t fl = freq_at<WIDE>(l), fr = l + 1 <= n_ ? freq_at<WIDE>(l + 1) : 0;
      const bool right = fr > fl;
      if ((right ? fr : fl) <= xf) break;
      const size_t c = l + (right ? 1 : 0);
      w_[s] = w_[c];
      if (WIDE) f_[s] = f_[c];
      if (track_) note(s);
      s = c;
    }
    if (n_) {
      w_[s] = x;
      if (WIDE) f_[s] = xf;
      if (track_) note(s);
      __builtin_prefetch(&pay_[id_at(1)]);  // the next top's payload: read by the caller's version check
    }
  }
  void note(size_t s) {
    if (dirty_) return;
    if (touched_.size() * 8 > n_ + 1024) { dirty_ = true; touched_.clear(); return; }  // cheaper to rebuild everything
    touched_.push_back(s);
  }
  // 2 MB aligned, transparent huge pages requested: the arrays are tens of MB and accessed at random
  static void* acquire(size_t bytes) {
    void* p = nullptr;
    const size_t rounded = (bytes + (2u << 20) - 1) & ~static_cast<size_t>((2u << 20) - 1);
    if (posix_memalign(&p, 2u << 20, rounded) != 0 || !p) std::abort();
#ifdef MADV_HUGEPAGE
    madvise(p, rounded, MADV_HUGEPAGE);
#endif
    return p;
  }
  void grow() {
    const size_t nc = cap_ ? cap_ * 2 : 4096;  // bpe.h:19 MIN_HEAP_SIZE, doubling as heap.cpp:59-68
    uint64_t* nw = static_cast<uint64_t*>(acquire(nc * sizeof(uint64_t)));
    if (n_) std::memcpy(nw, w_, (n_ + 1) * sizeof(uint64_t));
    std::free(w_);
    w_ = nw;
    if (wide_) {
      uint64_t* nf = static_cast<uint64_t*>(acquire(nc * sizeof(uint64_t)));
      if (n_) std::memcpy(nf, f_, (n_ + 1) * sizeof(uint64_t));
      std::free(f_);
      f_ = nf;
    }
    cap_ = nc;
  }
  void grow_pay() {
    const size_t nc = pcap_ ? pcap_ * 2 : 4096;
    HeapPayload* np = static_cast<HeapPayload*>(acquire(nc * sizeof(HeapPayload)));
    if (np_) std::memcpy(np, pay_, np_ * sizeof(HeapPayload));
    std::free(pay_);
    pay_ = np; pcap_ = nc;
  }
  static int bits_of(uint64_t v) { int b = 0; while (v) { ++b; v >>= 1; } return b; }
  // Renumber the ids (live entries only, in slot order) and, if `incoming` or a stored frequency needs it, move the split --
  // or, when frequency bits + id bits exceed one word, switch to the wide layout (separate frequency array) for good.
  void repack(uint64_t incoming) {
    uint64_t fmax = incoming;
    if (n_) fmax = std::max(fmax, wide_ ? f_[1] : w_[1] >> sh_);  // the root holds the largest frequency
    const int fbits = std::max(bits_of(fmax), 1), ibits = bits_of(2 * n_ + (1u << 17));  // room for as many pushes again
    const bool to_wide = wide_ || fbits + ibits > 64;
    const int nsh = to_wide ? sh_ : std::max(ibits, std::min(64 - fbits, 40));  // as many id bits as the frequencies leave (renumbering gets rarer)
    const size_t need = n_ + 1, ncap = std::max<size_t>(need * 2, 4096);
    HeapPayload* np = static_cast<HeapPayload*>(acquire(ncap * sizeof(HeapPayload)));
    uint64_t* nf = to_wide && !wide_ ? static_cast<uint64_t*>(acquire(cap_ * sizeof(uint64_t))) : nullptr;
    for (size_t s = 1; s <= n_; s++) {
      const uint64_t f = wide_ ? f_[s] : w_[s] >> sh_;
      np[s - 1] = pay_[id_at(s)];
      if (to_wide) { if (nf) nf[s] = f; w_[s] = s - 1; } else w_[s] = (f << nsh) | (s - 1);
    }
    std::free(pay_);
    pay_ = np; pcap_ = ncap; np_ = n_;
    if (nf) { f_ = nf; wide_ = true; }
    sh_ = nsh;
  }

  uint64_t* w_ = nullptr;       // slot -> frequency << sh_ | payload id   (wide layout: payload id only)
  uint64_t* f_ = nullptr;       // wide layout only: slot -> frequency
  bool wide_ = false;
  HeapPayload* pay_ = nullptr;  // payload id -> (pair, version, serial); append-only between renumberings
  size_t n_ = 0, cap_ = 0, np_ = 0, pcap_ = 0;
  int sh_ = 27;
  std::vector<BPEHeapEntry> mirror_;
  std::vector<size_t> touched_;
  bool dirty_ = true, mirror_valid_ = false, track_ = false;
  FILE* trace_ = nullptr;
};

static_assert(sizeof(HeapEnt) == sizeof(BPEHeapEntry) && sizeof(HeapEnt) == 24, "heap entry layout (reference heap.h:17-21)");

}  // namespace shred
