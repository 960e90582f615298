// tests/heap/bench_heap.cpp -- TEST INFRASTRUCTURE / development tool, never part of libtrainer.so.
//
// (1) check: the product's replay heap (csrc/exact_heap.hpp: packed 8-byte words, payload side array, renumbering, adaptive
//     split) against a literal restatement of the reference's heap rules (24-byte entries, heap.cpp:53-114) on random operation
//     sequences, including frequencies above 2^37 and long runs that force renumbering;
// (2) replay: time the product heap on a recorded operation trace (SHRED_HEAP_TRACE=<file> on any trainer run).
// usage: bench_heap check [seed]        bench_heap replay <trace.bin>        bench_heap replay-loop <trace.bin>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../shredword-trainer_b200/csrc/exact_heap.hpp"
#include "../../shredword-trainer_b200/csrc/flat_map.hpp"
using namespace shred;

struct RefHeap {  // heap.cpp:53-114 restated literally: 0-based array of {key, freq, version}, compares freq only
  std::vector<HeapEnt> a;
  void push(HeapEnt e) {
    a.push_back(e);
    size_t i = a.size() - 1;
    while (i > 0) { size_t p = (i - 1) / 2; if (a[p].freq >= a[i].freq) break; std::swap(a[p], a[i]); i = p; }
  }
  HeapEnt pop() {
    HeapEnt top = a[0];
    a[0] = a.back(); a.pop_back();
    size_t i = 0, n = a.size();
    for (;;) {
      size_t l = 2 * i + 1, r = 2 * i + 2, best = i;
      if (l < n && a[l].freq > a[best].freq) best = l;
      if (r < n && a[r].freq > a[best].freq) best = r;
      if (best == i) break;
      std::swap(a[i], a[best]); i = best;
    }
    return top;
  }
};

static uint64_t rng_state = 88172645463325252ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static bool same(const HeapEnt& x, const HeapEnt& y) { return x.freq == y.freq && x.serial == y.serial && x.version == y.version && x.key.first == y.key.first && x.key.second == y.key.second; }

static int check(uint64_t seed) {
  rng_state ^= seed * 0x9E3779B97F4A7C15ull;
  for (int round = 0; round < 6; round++) {
    ExactHeap h; RefHeap r;
    if (round & 1) h.set_tracking(true);
    const uint64_t fmask = round == 2 ? (1ull << 45) - 1 : round == 3 ? 7 : (1ull << 20) - 1;  // huge frequencies / massive ties / ordinary
    const int n_ops = round == 4 ? 1500000 : 200000;
    uint32_t next = 0;
    for (int i = 0; i < n_ops; i++) {
      const bool do_pop = !r.a.empty() && (rnd() % 100) < (round == 4 ? 49 : 45);
      if (do_pop) {
        if (!same(h.top(), r.a[0])) { std::printf("top differs at op %d round %d\n", i, round); return 1; }
        HeapEnt x = h.pop(), y = r.pop();
        if (!same(x, y)) { std::printf("pop differs at op %d round %d\n", i, round); return 1; }
      } else {
        uint64_t f = rnd() & fmask;
        if (round == 5 && (i % 1000) == 0) f = (1ull << 50) + (rnd() & 0xFFFF);  // occasional giant: the split must move
        HeapEnt e{PairKey{(int32_t)(rnd() & 0xFFFFF), (int32_t)(rnd() & 0xFFFFF)}, f, (uint32_t)(rnd() & 0xFF), next++};
        h.push(e.key, e.freq, e.version, e.serial); r.push(e);
      }
      if ((i % 20011) == 0) {  // the mirror for C consumers equals the reference's array, slot by slot
        size_t cap = 0;
        const BPEHeapEntry* m = h.materialize(&cap);
        if (h.size() != r.a.size()) { std::printf("size differs\n"); return 1; }
        for (size_t k = 0; k < r.a.size(); k++)
          if (m[k].freq != r.a[k].freq || m[k].key.first != r.a[k].key.first || m[k].key.second != r.a[k].key.second || m[k].version != r.a[k].version) { std::printf("mirror differs at slot %zu op %d round %d\n", k, i, round); return 1; }
      }
    }
  }
  std::printf("heap check ok\n");
  return 0;
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static int replay(const char* path) {
  FILE* f = std::fopen(path, "rb");
  if (!f) { std::perror(path); return 1; }
  std::vector<uint64_t> ops; uint64_t buf[4096]; size_t n;
  while ((n = std::fread(buf, 8, 4096, f)) > 0) ops.insert(ops.end(), buf, buf + n);
  std::fclose(f);
  for (int rep = 0; rep < 3; rep++) {
    ExactHeap h; uint64_t chk = 0, np = 0, nq = 0; double tp = 0, tq = 0; size_t i = 0;
    while (i < ops.size()) {
      size_t j = i; while (j < ops.size() && ops[j] != ~0ull) j++;
      const double t0 = now();
      for (size_t k = i; k < j; k++) h.push(PairKey{(int32_t)k, 0}, ops[k], 0, (uint32_t)k);
      const double t1 = now();
      size_t e = j; while (e < ops.size() && ops[e] == ~0ull) e++;
      for (size_t k = j; k < e; k++) { HeapEnt en = h.pop(); chk = chk * 1000003u + en.serial + en.freq; }
      const double t2 = now();
      tq += t1 - t0; tp += t2 - t1; nq += j - i; np += e - j; i = e;
    }
    std::printf("rep %d: %llu pushes %.3f s (%.0f ns each), %llu pops %.3f s (%.0f ns each), checksum %llx, final size %zu\n", rep, (unsigned long long)nq, tq, 1e9 * tq / (nq ? nq : 1),
                (unsigned long long)np, tp, 1e9 * tp / (np ? np : 1), (unsigned long long)chk, h.size());
  }
  return 0;
}

// (3) replay-loop: the trainer's pop loop around the heap (trainer_core.cpp merge_loop): every popped entry is looked up in a table
//     indexed by its serial (scattered here, 16 M entries of 8 bytes).  On the GPU box's host (Sapphire Rapids VM) 137 ns per pop
//     on the 10 GB trace -- prefetching the table lines of the root's children one pop ahead changed nothing (138 ns).
static int replay_loop(const char* path) {
  FILE* f = std::fopen(path, "rb");
  if (!f) { std::perror(path); return 1; }
  std::vector<uint64_t> ops; uint64_t buf[4096]; size_t n;
  while ((n = std::fread(buf, 8, 4096, f)) > 0) ops.insert(ops.end(), buf, buf + n);
  std::fclose(f);
  const size_t VN = 1u << 24;
  HugeArray<uint64_t> ver;  // as the trainer's table: 2 MB aligned, transparent huge pages requested
  ver.ensure(VN);
  for (size_t i = 0; i < VN; i++) ver[i] = i * 0x9E3779B97F4A7C15ull;
  auto slot = [](uint32_t serial) { return (serial * 2654435761u) & ((1u << 24) - 1); };
  for (int rep = 0; rep < 3; rep++) {
    ExactHeap h; uint64_t chk = 0, np = 0; double tp = 0; size_t i = 0;
    while (i < ops.size()) {
      size_t j = i; while (j < ops.size() && ops[j] != ~0ull) j++;
      for (size_t k = i; k < j; k++) h.push(PairKey{(int32_t)k, 0}, ops[k], 0, (uint32_t)k);
      size_t e = j; while (e < ops.size() && ops[e] == ~0ull) e++;
      const double t1 = now();
      for (size_t k = j; k < e; k++) {
        const HeapEnt t = h.top();
        __builtin_prefetch(&ver[slot(t.serial)]);
        HeapEnt en = h.pop();
        if (ver[slot(en.serial)] == en.version + 12345u) break;  // never true: a predictable branch on the looked-up value, like the version check
        chk = chk * 1000003u + en.serial + en.freq;
      }
      tp += now() - t1; np += e - j; i = e;
    }
    std::printf("rep %d: %llu pops %.3f s (%.0f ns each), checksum %llx\n", rep, (unsigned long long)np, tp, 1e9 * tp / (np ? np : 1), (unsigned long long)chk);
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 3 && !std::strcmp(argv[1], "replay-loop")) return replay_loop(argv[2]);
  if (argc >= 2 && !std::strcmp(argv[1], "check")) return check(argc > 2 ? std::strtoull(argv[2], nullptr, 10) : 1);
  if (argc >= 3 && !std::strcmp(argv[1], "replay")) return replay(argv[2]);
  std::fprintf(stderr, "usage: bench_heap check [seed] | bench_heap replay <trace.bin>\n");
  return 2;
}
