// abi.cpp -- the extern "C" layer of libtrainer.so: the reference's C ABI (shredword/csrc/bpe/bpe.h:62-72 as bound by
// shredword/cbase.py:50-57) on top of TrainerCore + the CUDA engine.  Return codes and error behaviour follow the
// reference functions cited next to each entry point.  There is no CPU fallback: without a usable sm_100 GPU
// create_trainer reports the CUDA error and exits, exactly like the reference exits on allocation failure.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/shred_abi.h"
#include "engine.hpp"
#include "trainer_core.hpp"

using shred::TrainerCore;

static_assert(sizeof(BPEConfig) == 24, "BPEConfig layout (reference bpe.h:43-48)");
static_assert(sizeof(BPEHeapEntry) == 24, "BPEHeapEntry layout (reference heap.h:17-21)");
static_assert(sizeof(Symbol) == 32, "Symbol layout (reference bpe.h:25-30)");
static_assert(offsetof(Trainer, heap) == 24 && offsetof(Trainer, corpus) == 48 && offsetof(Trainer, bigram_map) == 72 &&
                  offsetof(Trainer, num_merges) == 96 && offsetof(Trainer, merge_ops) == 104 && offsetof(Trainer, impl) == 128,
              "Trainer prefix layout (reference bpe.h:50-60)");

static inline TrainerCore* core(const Trainer* t) { return static_cast<TrainerCore*>(t->impl); }

extern "C" {

Trainer* create_trainer(const BPEConfig* config) {  // bpe.cpp:67-85
  if (config == NULL) {
    std::fprintf(stderr, "[ERROR]\t Config pointer is NULL\n");
    std::exit(EXIT_FAILURE);
  }
  Trainer* t = static_cast<Trainer*>(std::calloc(1, sizeof(Trainer)));
  if (!t) {
    std::fprintf(stderr, "[ERROR]\t Couldn't allocate Memory to Trainer\n");
    std::exit(EXIT_FAILURE);
  }
  t->config = *config;
  if (t->config.character_coverage <= 0.0 || t->config.character_coverage >= 1.0) t->config.character_coverage = 0.995;  // bpe.cpp:78
  if (t->config.min_pair_freq == 0) t->config.min_pair_freq = 2000;  // bpe.cpp:79, bpe.h:23
  shred::Engine* eng = shred::make_device_engine();
  if (!eng) {
    std::fprintf(stderr, "[ERROR]\t B200 BPE trainer needs a CUDA device (sm_100); there is no CPU fallback\n");
    std::exit(EXIT_FAILURE);
  }
  t->impl = new TrainerCore(t, eng);
  const char* q = std::getenv("SHRED_QUIET");
  if (!(q && *q && *q != '0')) std::printf("[INFO]\t BPE trainer initialized. Heap initialized successfully.\n");
  return t;
}

void bpe_trainer_destroy(Trainer* trainer) {  // bpe.cpp:87-96
  if (!trainer) {
    std::fprintf(stderr, "[ERROR]\t No Trainer pointer found to destroy!\n");
    std::exit(EXIT_FAILURE);
  }
  delete core(trainer);
  std::free(trainer);
}

int bpe_load_corpus(Trainer* trainer, const char* input_path) {  // bpe.cpp:110-185
  if (!trainer || !input_path) {
    std::fprintf(stderr, "[ERROR]\t NULL trainer or input path pointers\n");
    return -1;
  }
  return core(trainer)->load_file(input_path);
}

int bpe_b200_load_buffer(Trainer* trainer, const uint8_t* text, size_t n_bytes) {
  if (!trainer || (!text && n_bytes)) {
    std::fprintf(stderr, "[ERROR]\t NULL trainer or buffer pointers\n");
    return -1;
  }
  return core(trainer)->load_buffer(text ? text : reinterpret_cast<const uint8_t*>(""), n_bytes);
}

void bpe_init(Trainer* trainer) {  // bpe.cpp:98-108
  if (!trainer) {
    std::fprintf(stderr, "[ERROR]\t NULL trainer pointer\n");
    std::exit(EXIT_FAILURE);
  }
  core(trainer)->init();
}

void bpe_count_bigrams(Trainer* trainer) {  // bpe.cpp:187-230
  if (!trainer) {
    std::fprintf(stderr, "[ERROR]\t NULL trainer pointer\n");
    std::exit(EXIT_FAILURE);
  }
  core(trainer)->count_bigrams();
}

int bpe_merge_batch(Trainer* trainer, int batch_size) {  // bpe.cpp:232-323
  if (!trainer) {
    std::fprintf(stderr, "[ERROR]\t Trainer pointer is NULL!\n");
    return -1;
  }
  return core(trainer)->merge_batch(batch_size);
}

int bpe_train(Trainer* trainer) {  // bpe.cpp:345-386
  if (!trainer) {
    std::fprintf(stderr, "[ERROR]\t Trainer pointer is NULL!\n");
    return -1;
  }
  return core(trainer)->train();
}

void bpe_save(const Trainer* trainer, const char* model_path, const char* vocab_path) {  // bpe.cpp:388-432
  if (!trainer) {
    std::fprintf(stderr, "[ERROR]\t Trainer pointer is NULL!\n");
    std::exit(EXIT_FAILURE);
  }
  if (!model_path || !vocab_path) return;
  core(trainer)->save(model_path, vocab_path);
}

int bpe_b200_get_stats(const Trainer* trainer, shred_stats_t* out) {
  if (!trainer || !out) return -1;
  core(trainer)->get_stats(out);
  return 0;
}

int bpe_b200_get_words(const Trainer* trainer, uint64_t* counts_out, uint64_t* off_out, int32_t* ids_out, uint64_t ids_cap) {
  if (!trainer || !core(trainer)->loaded()) return -1;
  return core(trainer)->engine()->get_words(counts_out, off_out, ids_out, ids_cap);
}

int bpe_b200_get_charset(const Trainer* trainer, uint8_t* keep_out, uint64_t* hist_out) {
  if (!trainer || !core(trainer)->loaded()) return -1;
  const shred::LoadInfo& li = core(trainer)->load_info();
  if (keep_out) std::memcpy(keep_out, li.keep, 256);
  if (hist_out) std::memcpy(hist_out, li.hist, sizeof li.hist);
  return 0;
}

uint64_t bpe_b200_get_pairs(const Trainer* trainer, int32_t* ab_out, uint64_t* freq_out, uint64_t cap) {
  if (!trainer || !core(trainer)->loaded()) return 0;
  return core(trainer)->engine()->get_pairs(ab_out, freq_out, cap);
}

// ---- Unigram stubs (reference unigram.h:50-68; bound eagerly by cbase.py:59-71, trainer itself out of scope) ----
struct UnigramTrainer { int unused; };
static void unigram_unavailable(const char* fn) {
  std::fprintf(stderr, "[ERROR]\t %s: the Unigram trainer is not part of the B200 BPE library\n", fn);
}
UnigramTrainer* trainerCreate(int, float, int, int) { unigram_unavailable("trainerCreate"); return NULL; }
void trainerDestroy(UnigramTrainer*) {}
bool addTextToTrainer(UnigramTrainer*, const char*) { return false; }
bool preprocessTexts(UnigramTrainer*) { return false; }
bool extractInitialSubwords(UnigramTrainer*) { return false; }
float computeLoss(UnigramTrainer*, const char**, int) { return 0.0f; }
double computeTokenLoss(UnigramTrainer*, const char*, const char**, int) { return 0.0; }
bool pruneVocabStep(UnigramTrainer*, const char**, int, double) { return false; }
bool updateTokenScores(UnigramTrainer*, const char**, int) { return false; }
bool trainUnigram(UnigramTrainer*, const char**, int, int) { unigram_unavailable("trainUnigram"); return false; }
bool getVocab(UnigramTrainer*, char***, double**, int*) { return false; }
bool saveVocab(UnigramTrainer*, const char*) { return false; }
bool loadVocab(UnigramTrainer*, const char*) { return false; }

}  // extern "C"
