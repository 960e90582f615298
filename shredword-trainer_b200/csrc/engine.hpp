// engine.hpp -- the seam between the sequential host control (exact heap replay, trainer_core.cpp) and the
// data-parallel device work (CUDA kernels, cuda/engine_cuda.cu).
//
// The product library links exactly one implementation: the CUDA engine.  tests/hostsim/ holds a CPU stand-in of
// this interface that exists only so the *host* logic (ordering, versions, phantom pairs, heap replay) can be
// checked against the reference on machines without a GPU; it is never linked into libtrainer.so.
#pragma once
#include <cstddef>
#include <cstdint>

namespace shred {

// One record per pair key touched by a device pass, returned UNORDERED; the host orders them.
//   count pass : kind=PUSH,    val = freq            seq = flat position of the pair's first sighting   (Appendix A6/A7)
//   merge pass : kind=PUSH,    val = new table freq  (>= min_pair_freq)                                   (A11)
//                kind=DEMOTE,  val = new table freq  (fell from >= min to < min; heap entry must die)     (A9 "< min")
//                kind=PHANTOM, val = net delta (two's complement int64) of a key that contains unk_id     (A12)
//                seq = smallest (flat position*4 + slot) at which the reference's FreqChangeMap first saw the key (A14)
struct Rec {
  uint64_t key;  // ((uint64)first << 32) | (uint64)second with both int32 sign-extended (reference bpe.cpp:277-278)
  uint64_t val;
  uint64_t seq;
  uint32_t kind;
  uint32_t serial;  // dense id of the pair in the device table (stable until the next count pass); REC_NO_SERIAL for phantoms
};
enum : uint32_t { REC_PUSH = 0, REC_DEMOTE = 1, REC_PHANTOM = 2 };
constexpr uint32_t REC_NO_SERIAL = 0xFFFFFFFFu;
// Rec.kind carries, above its two kind bits, the length of the occurrence list the device created for the key in this very
// pass (0 when the key already existed: a pair's occurrences are all created by one pass, afterwards they only disappear).
constexpr uint32_t REC_KIND_MASK = 3u, REC_LEN_SHIFT = 2u, REC_LEN_MAX = 0x3FFFFFFFu;
#if defined(__CUDACC__)
#define SHRED_REC_FN __host__ __device__ inline
#else
#define SHRED_REC_FN inline
#endif
SHRED_REC_FN uint32_t rec_kind(uint32_t k) { return k & REC_KIND_MASK; }
SHRED_REC_FN uint32_t rec_list_len(uint32_t k) { return k >> REC_LEN_SHIFT; }
SHRED_REC_FN uint32_t rec_pack(uint32_t kind, uint64_t list_len) { return kind | (static_cast<uint32_t>(list_len < REC_LEN_MAX ? list_len : REC_LEN_MAX) << REC_LEN_SHIFT); }

struct EngineConfig {
  int32_t unk_id;
  float coverage;     // already normalised
  uint64_t min_freq;  // already normalised
  uint64_t vocab_size;  // target vocabulary (sizes scratch tables; not a limit)
};

struct LoadInfo {
  uint64_t n_words, n_symbols, n_tokens;
  uint64_t hist[256];
  uint8_t keep[256];
  uint32_t n_distinct, n_keep;
};

struct EngineStats {
  uint64_t n_slots, n_symbols_live, pair_entries;
  // timed merge launches (every SHRED_TIMING-th): CUDA-event duration, algorithmic bytes of the scan formulation (4 B x (live
  // symbols + words)) and an estimate of the bytes the launch really touches (list entries, probes, rewrites)
  uint64_t scan_launches; double scan_device_ms; double scan_bytes; double scan_bytes_touched;
  uint64_t dense_launches; double dense_device_ms; double dense_bytes;  // timed launches whose occurrence list has >= 65536 entries
  double scan_phase_ms, dense_phase_ms;  // in-kernel %globaltimer: kernel start -> end of phase 1 (all timed / dense timed launches)
  double fold_phase_ms, rewrite_phase_ms;  // same timer: phase 1 end -> published (phase 2) -> CTA 0 done (phase 3), all timed launches
  uint64_t list_entries, pool_entries;   // occurrence-list entries probed by all merges / entries allocated in the pool
  uint64_t single_launches;              // merges handled by the one-CTA variant of the kernel
  uint64_t server_merges, server_starts; // of those: taken by the resident merge server (no launch) / times the server was started
  uint64_t count_launches; double count_device_ms; double count_bytes;
  double fill_device_ms, fill_bytes;     // count pass, second half: fold + initial occurrence lists
  uint64_t ingest_launches; double ingest_device_ms; double ingest_bytes;
  uint64_t kernel_launches;
  double h2d_ms, wait_ms, launch_ms, merge_ms;
  uint64_t h2d_bytes, d2h_bytes;
};

class Engine {
 public:
  virtual ~Engine() {}
  // Tokenise `text`, build the unique-word table in reference order, apply character coverage, lay the symbols out.
  // A second load replaces the first.  Returns 0, -1 on failure, or 1 if the text contains NUL bytes (nothing is
  // loaded then: the caller blanks the spans the reference would not see and calls load again).
  virtual int load(const uint8_t* text, size_t n, const EngineConfig& cfg, LoadInfo* info) = 0;
  // Same from an open file of n bytes (lets the engine pipeline the read with the host-to-device copy).  Returns like
  // load(), or 2 if the engine has no file path (the caller maps the file and calls load()).
  virtual int load_file(int /*fd*/, size_t /*n*/, const EngineConfig& /*cfg*/, LoadInfo* /*info*/) { return 2; }
  // Reset the pair table, count all adjacent non-unk pairs; *recs = PUSH records for entries with freq >= min.
  virtual int count_pairs(const Rec** recs, size_t* n) = 0;
  // Rewrite every leftmost non-overlapping (a,b) -> new_id, update the pair table, return the touched keys.
  // serial / list_len: the pair's dense id and the length of its occurrence list, both as reported by the PUSH record that
  // created its heap entry (rec_list_len()); they let the device find the list with one load and size its grid.
  virtual int merge(int32_t a, int32_t b, int32_t new_id, uint32_t serial, uint32_t list_len, const Rec** recs, size_t* n, uint64_t* occurrences) = 0;
  // Brackets of a run of merge() calls (bpe_merge_batch / bpe_train): lets an engine keep state resident between the merges
  // of a run (the CUDA engine's merge server) and guarantees it is gone when the run's entry point returns.
  virtual void begin_merges() {}
  virtual void end_merges() {}
  // freq[id] += word_count over all live symbols with 0 <= id < n_tokens (reference bpe.cpp:409-415).
  virtual int token_freqs(uint64_t* freq, size_t n_tokens) = 0;
  virtual int word_counts(uint64_t* out) = 0;  // host mirror of Corpus.word_counts
  // parity/debug getters
  virtual int get_words(uint64_t* counts, uint64_t* off, int32_t* ids, uint64_t ids_cap) = 0;
  virtual uint64_t get_pairs(int32_t* ab, uint64_t* freq, uint64_t cap) = 0;
  virtual void stats(EngineStats* out) = 0;
  // device-side stopwatch on the engine's stream (CUDA events): mark_begin(), ..., mark_end() -> elapsed ms
  virtual int mark_begin() = 0;
  virtual double mark_end() = 0;
  virtual const char* name() = 0;
};

// Implemented by cuda/engine_cuda.cu.  Fails loudly (message on stderr, returns nullptr) without a usable GPU.
Engine* make_device_engine();

}  // namespace shred
