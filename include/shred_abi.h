/* include/shred_abi.h -- the drop-in boundary of the B200-native BPE trainer.
 *
 * libtrainer.so exports exactly the C ABI that the reference's ctypes binding resolves at import time
 * (reference shredword/cbase.py:50-71): the 8 BPE entry points of shredword/csrc/bpe/bpe.h:62-72 with identical
 * signatures, argument meaning and return conventions, plus the 13 Unigram symbols of
 * shredword/csrc/unigram/unigram.h:50-68 as inert stubs (cbase.py binds them eagerly, so they must resolve; the
 * Unigram trainer itself is out of scope).  Plain pointers and sizes only; no CUDA or torch
 * types cross this boundary.
 *
 * The structs below re-express the reference's ABI-visible layouts (verified there with
 * sizeof/offsetof on the reference build).  C/C++ consumers of the reference read some Trainer fields directly
 * (shredword/csrc/trainer.cpp:100, test/bpe_test.cpp:69-72,112-121,151-161,186-196) -- those fields sit at the same
 * offsets here and are kept up to date as host mirrors; everything else lives behind `impl`.
 */
#ifndef SHRED_ABI_H
#define SHRED_ABI_H

#include <stddef.h>
#include <stdint.h>
#ifndef __cplusplus
#include <stdbool.h>
#endif

#if defined(__GNUC__)
#define SHRED_API __attribute__((visibility("default")))
#else
#define SHRED_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* reference bpe.h:43-48 -- 24 bytes: size_t@0, int32@8, float@12, uint64@16 */
typedef struct BPEConfig {
  size_t target_vocab_size;
  int32_t unk_id;            /* id substituted for bytes dropped by character_coverage */
  float character_coverage;  /* outside (0,1) => 0.995 (bpe.cpp:78) */
  uint64_t min_pair_freq;    /* 0 => 2000 (bpe.cpp:79) */
} BPEConfig;

/* reference hash.h:27-29 */
typedef struct PairKey { int32_t first, second; } PairKey;

/* reference heap.h:17-21 -- 24 bytes */
typedef struct BPEHeapEntry { PairKey key; uint64_t freq; uint32_t version; } BPEHeapEntry;

/* reference heap.h:23-27 */
typedef struct MaxHeap { BPEHeapEntry* data; size_t size; size_t cap; } MaxHeap;

/* reference bpe.h:25-30 -- 32 bytes.  The symbols live in HBM here; corpus.words[i] points at a host placeholder. */
typedef struct Symbol { int32_t id; struct Symbol* prev; struct Symbol* next; bool deleted; } Symbol;

/* reference bpe.h:37-41 */
typedef struct Corpus { Symbol** words; uint64_t* word_counts; size_t vocab_size; } Corpus;

/* reference hash.h:42-45 (opaque here: the pair table is an open-addressing table in HBM) */
typedef struct BIMap { void* buckets; size_t nbuckets; } BIMap;

/* reference bpe.h:50-60 -- 128 bytes: config@0 heap@24 corpus@48 bigram_map@72 next_token@88 num_merges@96
 * merge_ops@104 token_strs@112 token_freq@120.  `impl` is appended after the reference's last field. */
typedef struct Trainer {
  BPEConfig config;
  MaxHeap heap;          /* host mirror of the exact replay heap (same 24-byte entries as the reference) */
  Corpus corpus;         /* vocab_size = unique words; word_counts = host mirror; words[i] = non-NULL placeholder */
  BIMap bigram_map;      /* nbuckets = live pair-table entries; buckets = NULL */
  size_t next_token;     /* unused by the reference */
  size_t num_merges;
  PairKey* merge_ops;    /* merge m produced id 256+m from (first, second) */
  char** token_strs;     /* unused by the reference */
  uint64_t* token_freq;  /* unused by the reference */
  void* impl;            /* B200 state (device buffers, streams, replay heap) */
} Trainer;

/* ---- BPE entry points: reference bpe.h:62-72 / bpe.cpp ------------------------------------------------------ */

/* bpe.cpp:67-85.  Never returns NULL; exit(1) on a NULL config.  Copies the config, normalises coverage/min freq. */
SHRED_API Trainer* create_trainer(const BPEConfig* config);
/* bpe.cpp:87-96.  Releases host and device state.  Safe without a loaded corpus. */
SHRED_API void bpe_trainer_destroy(Trainer* trainer);
/* bpe.cpp:110-185.  0 on success, -1 on NULL arguments / unreadable file / allocation failure / CUDA failure.
 * Tokens are maximal runs of bytes not in {\t,\r,\n,space}; a second load replaces the first. */
SHRED_API int bpe_load_corpus(Trainer* trainer, const char* input_path);
/* bpe.cpp:98-108.  Resets pair table and heap, then counts bigrams. */
SHRED_API void bpe_init(Trainer* trainer);
/* bpe.cpp:187-230.  Counts adjacent pairs into the pair table and seeds the heap (freq >= min_pair_freq). */
SHRED_API void bpe_count_bigrams(Trainer* trainer);
/* bpe.cpp:232-323.  Performs up to batch_size merges; returns merges done, 0 if the heap is empty, -1 on NULL. */
SHRED_API int bpe_merge_batch(Trainer* trainer, int batch_size);
/* bpe.cpp:345-386.  bpe_init + merge loop until vocab_size-256 merges or no pair reaches min_pair_freq.
 * Returns merges performed by this call, -1 on NULL. */
SHRED_API int bpe_train(Trainer* trainer);
/* bpe.cpp:388-432.  vocab: "<token bytes> <freq>\n" x (256+M); model: M x {int32 a, int32 b, int32 256+m}. */
SHRED_API void bpe_save(const Trainer* trainer, const char* model_path, const char* vocab_path);

/* ---- extensions (not part of the reference ABI; never required by a drop-in consumer) ------------------------ */

/* Same as bpe_load_corpus but from a host buffer (used by bench.py's end-to-end leg and the tests). */
SHRED_API int bpe_b200_load_buffer(Trainer* trainer, const uint8_t* text, size_t n_bytes);
/* Phase timers and work counters of the last load/train, see shred_stats_t. */
typedef struct shred_stats_t {
  uint64_t n_words, n_symbols_initial, n_symbols_live, n_slots, n_tokens, corpus_bytes;
  uint64_t pair_entries, heap_size, heap_pushes, heap_pops;
  uint64_t merges, occurrences;
  uint64_t list_entries;         /* occurrence-list entries probed by all merges of the last train (occurrences of them were live) */
  uint64_t pool_entries;         /* occurrence-list entries allocated so far (initial lists + lists of the pairs merges created) */
  uint64_t scan_launches;        /* TIMED per-merge kernel launches in the last train (every SHRED_TIMING-th) */
  double scan_device_ms;         /* sum of their CUDA-event durations */
  double scan_bytes;             /* algorithmic bytes of the scan formulation they stand for: sum 4*(live symbols + words), SURVEY 8d */
  uint64_t count_launches; double count_device_ms; double count_bytes;   /* k_count: CUDA events, 4S + 12N */
  double fill_device_ms, fill_bytes;                                     /* fold + k_fill_lists: the initial occurrence lists */
  uint64_t ingest_launches; double ingest_device_ms; double ingest_bytes;
  uint64_t kernel_launches;      /* every kernel launch since create */
  double load_wall_ms, h2d_ms, train_wall_ms, host_heap_ms, wait_ms, save_wall_ms;
  double train_device_ms;        /* CUDA-event time of the last bpe_train on the engine's stream */
  double launch_ms;              /* host time spent issuing the per-merge kernel launches */
  double scan_bytes_touched;     /* estimate of the bytes the timed launches really touch (list entries, probes, scratch, rewrites) */
  uint64_t dense_launches; double dense_device_ms; double dense_bytes; /* timed launches with >= 65536 list entries */
  double scan_phase_ms, dense_phase_ms; /* in-kernel timer: kernel start -> end of phase 1 (probe + deltas), all / dense timed launches */
  uint64_t h2d_bytes, d2h_bytes;
  /* tie statistics of the last train: merges whose frequency equals that of the entry left at the heap
   * root (upper bound on "another pair shares the maximum") / equals the previous merge's frequency (lower bound) */
  uint64_t tie_root_equal, tie_same_as_prev;
  double fold_phase_ms, rewrite_phase_ms; /* in-kernel timer of the timed launches: end of phase 1 -> published | -> CTA 0 done */
  uint64_t single_launches;      /* merges of the last trainer handled by the one-CTA variant of the merge kernel */
  uint64_t server_merges, server_starts; /* of those: taken by the resident merge server (no launch) / times it was started */
} shred_stats_t;
SHRED_API int bpe_b200_get_stats(const Trainer* trainer, shred_stats_t* out);
/* Debug/parity getters: copy the current word table out of HBM.  word order = reference StrMap iteration order.
 * ids_out receives the live symbol ids of all words back to back; off_out[n_words+1] their offsets. */
SHRED_API int bpe_b200_get_words(const Trainer* trainer, uint64_t* counts_out, uint64_t* off_out, int32_t* ids_out, uint64_t ids_cap);
/* Copies the keep mask (256 flags) and the unweighted byte histogram (256 counters) of the last load. */
SHRED_API int bpe_b200_get_charset(const Trainer* trainer, uint8_t* keep_out, uint64_t* hist_out);
/* Copies up to cap live pair-table entries as (first, second) + freq; returns the number of live entries. */
SHRED_API uint64_t bpe_b200_get_pairs(const Trainer* trainer, int32_t* ab_out, uint64_t* freq_out, uint64_t cap);
/* Human-readable description of the device the trainer runs on ("NVIDIA B200 sm_100 148 SMs"). */
SHRED_API const char* bpe_b200_device_name(void);

/* ---- encoder: the caller-side step after bpe_save ---------------------------------
 * The reference has no FFI for this: its encoder is the pure-Python BPETokenizer (shredword/utils/bpe.py).  These entry
 * points replace, for the model file bpe_save writes (bpe.cpp:419-427):
 *   bpe_b200_encoder_load / _create   BaseTokenizer.load            utils/bpe.py:140-155  (merges[(a, b)] = id in file order)
 *   bpe_b200_encode (+ _fetch)        BPETokenizer.encode           utils/bpe.py:205-212  with _encode_chunk :191-203 per word
 *   bpe_b200_decode                   BPETokenizer.decode           utils/bpe.py:214-225  (bytes; an unknown id is an error)
 * Words are the trainer's: maximal runs of bytes outside {9, 10, 13, 32} (bpe.cpp:131-152); delimiters produce no ids.
 * A model is accepted iff row m is {a, b, 256 + m} with 0 <= a, b < 256 + m; otherwise create/load return NULL. */
typedef struct shred_encoder shred_encoder_t;
typedef struct shred_encode_stats {
  uint64_t text_bytes, n_words, n_unique_words, n_ids;
  uint64_t pool_ids;          /* symbols of the distinct words before merging */
  uint64_t kernel_launches, h2d_bytes, d2h_bytes;
  double h2d_ms;              /* text -> HBM */
  double tokenize_ms, words_ms, expand_ms; /* CUDA events: distinct words | merge loop per distinct word | occurrences -> ids */
  double device_ms;           /* CUDA events around the three phases */
  double encode_wall_ms;      /* host clock around bpe_b200_encode */
  double d2h_ms;              /* bpe_b200_encode_fetch */
} shred_encode_stats_t;
SHRED_API shred_encoder_t* bpe_b200_encoder_create(const int32_t* triples, size_t n_merges);
SHRED_API shred_encoder_t* bpe_b200_encoder_load(const char* model_path);
SHRED_API void bpe_b200_encoder_destroy(shred_encoder_t* enc);
SHRED_API size_t bpe_b200_encoder_vocab_size(const shred_encoder_t* enc);
/* Encodes a host buffer; the ids of all words back to back and their CSR offsets (n_words + 1) stay in HBM.  0 / -1. */
SHRED_API int bpe_b200_encode(shred_encoder_t* enc, const uint8_t* text, uint64_t n_bytes, uint64_t* n_words, uint64_t* n_ids);
/* Copies the last result out: ids_out[n_ids], offsets_out[n_words + 1] (either may be NULL).  0 / -1. */
SHRED_API int bpe_b200_encode_fetch(shred_encoder_t* enc, int32_t* ids_out, uint64_t* offsets_out);
/* Streamed encode straight into host buffers (pinned for full speed): the text is cut at delimiters into pieces and piece k + 1
 * travels to the device while piece k is encoded and piece k - 1 travels back.  Same ids and (global) offsets as
 * bpe_b200_encode + _fetch.  ids_cap >= n_bytes and offsets_cap >= n_bytes / 2 + 2 always suffice; offsets_out may be NULL
 * (ids only, like the reference's encode(): 40 % fewer bytes come back).
 * 0 ok, -1 device error / NULL, -3 a capacity is too small. */
SHRED_API int bpe_b200_encode_to_host(shred_encoder_t* enc, const uint8_t* text, uint64_t n_bytes, int32_t* ids_out, uint64_t ids_cap,
                                      uint64_t* offsets_out, uint64_t offsets_cap, uint64_t* n_words, uint64_t* n_ids);
/* Bytes of the ids back to back into out[cap].  Returns the byte count (nothing written if > cap), -2 for an id outside
 * [0, vocab_size), -1 on a device error. */
SHRED_API int64_t bpe_b200_decode(shred_encoder_t* enc, const int32_t* ids, uint64_t n_ids, uint8_t* out, uint64_t cap);
SHRED_API int bpe_b200_encoder_get_stats(const shred_encoder_t* enc, shred_encode_stats_t* out);

/* ---- Unigram symbols: reference unigram.h:50-68.  Stubs; every call reports failure. -------------------------- */
typedef struct UnigramTrainer UnigramTrainer;
SHRED_API UnigramTrainer* trainerCreate(int vocab_size, float character_coverage, int max_len, int seed_size); /* returns NULL */
SHRED_API void trainerDestroy(UnigramTrainer* trainer);
SHRED_API bool addTextToTrainer(UnigramTrainer* trainer, const char* text);
SHRED_API bool preprocessTexts(UnigramTrainer* trainer);
SHRED_API bool extractInitialSubwords(UnigramTrainer* trainer);
SHRED_API float computeLoss(UnigramTrainer* trainer, const char** texts, int text_count);
SHRED_API double computeTokenLoss(UnigramTrainer* trainer, const char* token, const char** texts, int text_count);
SHRED_API bool pruneVocabStep(UnigramTrainer* trainer, const char** texts, int text_count, double reduction_ratio);
SHRED_API bool updateTokenScores(UnigramTrainer* trainer, const char** texts, int text_count);
SHRED_API bool trainUnigram(UnigramTrainer* trainer, const char** texts, int text_count, int num_iterations);
SHRED_API bool getVocab(UnigramTrainer* trainer, char*** tokens, double** scores, int* count);
SHRED_API bool saveVocab(UnigramTrainer* trainer, const char* filepath);
SHRED_API bool loadVocab(UnigramTrainer* trainer, const char* filepath);

#ifdef __cplusplus
}
#endif
#endif /* SHRED_ABI_H */
