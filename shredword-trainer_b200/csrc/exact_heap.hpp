// exact_heap.hpp -- the replay heap.
//
// The reference picks the next merge with an array binary max-heap that compares `freq` only
// (reference shredword/csrc/bpe/heap.cpp:53-114), so which of several equal-frequency pairs wins is decided by
// the heap's structural history.  Bit-exact merge lists therefore require replaying the same sift rules on the same
// push/pop sequence.
//
// Same algorithm, different memory layout: the sift decisions only ever read frequencies, so the frequencies live in
// their own dense array (8 bytes per entry instead of 24) and the payloads (pair, version, serial) in a second one.
// With millions of entries a pop is a chain of ~20 dependent cache misses; walking the 3x smaller frequency array keeps
// the upper ~17 levels cache resident, and the payload moves along the found path are independent of each other, so
// their misses overlap.  Storage is 1-based (the reference's entry i lives in slot i + 1) so that the 2^d descendants d
// levels below a node form one aligned block: the walk prefetches the 16 great-great-grandchildren (two cache lines) four
// levels ahead, which turns the chain of full memory latencies into a pipelined one.  materialize() writes the
// reference's 24-byte array-of-structs (heap.h:17-21) for C consumers that read Trainer.heap.data.
#pragma once
#include <sys/mman.h>

#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/shred_abi.h"

namespace shred {

struct HeapPayload { PairKey key; uint32_t version; uint32_t serial; };  // serial: dense device-side id of the pair (engine.hpp)
struct HeapEnt { PairKey key; uint64_t freq; uint32_t version; uint32_t serial; };
static_assert(sizeof(HeapPayload) == 16, "payload packs into 16 bytes");

class ExactHeap {
 public:
  ~ExactHeap() { release(freq_, cap_ * sizeof(uint64_t)); release(pay_, cap_ * sizeof(HeapPayload)); }
  void clear() { n_ = 0; dirty_ = true; touched_.clear(); }
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
  HeapEnt top() const { return HeapEnt{pay_[1].key, freq_[1], pay_[1].version, pay_[1].serial}; }

  // heap.cpp:70-79: append, then swap upwards while the parent's freq is strictly smaller.
  void push(PairKey key, uint64_t freq, uint32_t version, uint32_t serial) {
    if (trace_) { const uint64_t t = freq; std::fwrite(&t, 8, 1, trace_); }
    if (n_ + 1 >= cap_) grow();
    size_t s = ++n_;  // slot of the reference's index n_-1
    while (s > 1) {
      const size_t p = s >> 1;
      if (freq_[p] >= freq) break;
      freq_[s] = freq_[p];
      pay_[s] = pay_[p];
      if (track_) note(s);
      s = p;
    }
    freq_[s] = freq;
    pay_[s] = HeapPayload{key, version, serial};
    if (track_) note(s); else dirty_ = true;
    ++pushes;
  }

  // heap.cpp:97-111: last entry to the root, then swap with the left child if it is strictly larger, with the right
  // child if it is strictly larger than the better of the two, until neither is.
  HeapEnt pop() {
    if (trace_) { const uint64_t t = ~0ull; std::fwrite(&t, 8, 1, trace_); }
    const HeapEnt out = top();
    const uint64_t xf = freq_[n_];
    const HeapPayload xp = pay_[n_];
    --n_;
    // 1. the path: frequencies only
    size_t path[72];
    int depth = 0;
    size_t s = 1;
    for (;;) {
      const size_t g = s << 4;  // the 16 descendants four levels down: one aligned 128-byte block
      if (g <= n_) { __builtin_prefetch(freq_ + g); __builtin_prefetch(freq_ + g + 8); }
      const size_t l = s << 1, r = l + 1;
      size_t best = s;
      uint64_t bf = xf;
      if (l <= n_ && freq_[l] > bf) { best = l; bf = freq_[l]; }
      if (r <= n_ && freq_[r] > bf) { best = r; }
      if (best == s) break;
      __builtin_prefetch(pay_ + best);  // needed in step 2, independent of the rest of the walk
      path[depth++] = best;
      s = best;
    }
    // 2. shift the entries one level up along the path
    size_t at = 1;
    for (int k = 0; k < depth; k++) {
      freq_[at] = freq_[path[k]];
      pay_[at] = pay_[path[k]];
      if (track_) note(at);
      at = path[k];
    }
    if (n_) { freq_[at] = xf; pay_[at] = xp; if (track_) note(at); }
    if (!track_) dirty_ = true;
    ++pops;
    return out;
  }

  // The reference's array of 24-byte entries, rebuilt only when the heap changed since the last call.
  BPEHeapEntry* materialize(size_t* cap_out) {
    if (dirty_ || !mirror_valid_) {
      mirror_.resize(n_ ? n_ : 1);
      HeapEnt* m = reinterpret_cast<HeapEnt*>(mirror_.data());
      for (size_t i = 0; i < n_; i++) m[i] = HeapEnt{pay_[i + 1].key, freq_[i + 1], pay_[i + 1].version, pay_[i + 1].serial};
    } else if (!touched_.empty()) {  // patch only the slots written since the last call
      if (mirror_.size() < (n_ ? n_ : 1)) mirror_.resize(std::max(n_ ? n_ : 1, mirror_.size() * 2));
      HeapEnt* m = reinterpret_cast<HeapEnt*>(mirror_.data());
      for (const size_t s : touched_) if (s <= n_) m[s - 1] = HeapEnt{pay_[s].key, freq_[s], pay_[s].version, pay_[s].serial};
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
  FILE* trace_ = nullptr;
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
  static void release(void* p, size_t) { std::free(p); }
  void grow() {
    const size_t nc = cap_ ? cap_ * 2 : 4096;  // bpe.h:19 MIN_HEAP_SIZE, doubling as heap.cpp:59-68
    uint64_t* nf = static_cast<uint64_t*>(acquire(nc * sizeof(uint64_t)));
    HeapPayload* np = static_cast<HeapPayload*>(acquire(nc * sizeof(HeapPayload)));
    if (n_) { std::memcpy(nf, freq_, (n_ + 1) * sizeof(uint64_t)); std::memcpy(np, pay_, (n_ + 1) * sizeof(HeapPayload)); }
    release(freq_, 0); release(pay_, 0);
    freq_ = nf; pay_ = np; cap_ = nc;
  }

  void note(size_t s) {
    if (dirty_) return;
    if (touched_.size() * 8 > n_ + 1024) { dirty_ = true; touched_.clear(); return; }  // cheaper to rebuild everything
    touched_.push_back(s);
  }
  uint64_t* freq_ = nullptr;
  HeapPayload* pay_ = nullptr;
  size_t n_ = 0, cap_ = 0;
  std::vector<BPEHeapEntry> mirror_;
  std::vector<size_t> touched_;
  bool dirty_ = true, mirror_valid_ = false, track_ = false;
};

static_assert(sizeof(HeapEnt) == sizeof(BPEHeapEntry) && sizeof(HeapEnt) == 24, "heap entry layout (reference heap.h:17-21)");

}  // namespace shred
