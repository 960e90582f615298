// tests/c_abi/consumer_test.cpp -- a C++ consumer of the trainer's C ABI that, like the reference's own
// test/bpe_test.cpp, reads Trainer fields directly (config, corpus.vocab_size / words / word_counts, heap.size /
// heap.data[i].freq, num_merges, merge_ops).  It re-expresses the intent of that file's eight cases
// (reference test/bpe_test.cpp:59-330) against include/shred_abi.h; it is compiled by tests/test_c_consumer.py and
// linked against libtrainer.so (GPU) or the hostsim test library (CPU).
//
// usage: consumer_test <corpus> <tmpdir>     exit code = number of failed checks
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/shred_abi.h"

static int failures = 0;
#define CHECK(cond, what)                                                   \
  do {                                                                      \
    if (!(cond)) { std::printf("[FAIL] %s (%s)\n", what, #cond); ++failures; } \
    else std::printf("[PASS] %s\n", what);                                  \
  } while (0)

static long file_size(const std::string& p) {
  FILE* f = std::fopen(p.c_str(), "rb");
  if (!f) return -1;
  std::fseek(f, 0, SEEK_END);
  long n = std::ftell(f);
  std::fclose(f);
  return n;
}
static long count_lines(const std::string& p) {
  FILE* f = std::fopen(p.c_str(), "rb");
  if (!f) return -1;
  long n = 0;
  for (int c; (c = std::fgetc(f)) != EOF;) n += c == '\n';
  std::fclose(f);
  return n;
}

int main(int argc, char** argv) {
  if (argc < 3) return 99;
  const char* corpus = argv[1];
  const std::string tmp = argv[2];

  {  // 1. creation copies the config (bpe_test.cpp:59-77)
    BPEConfig cfg = {300, -1, 0.99f, 2};
    Trainer* t = create_trainer(&cfg);
    CHECK(t != NULL, "create_trainer returns a trainer");
    CHECK(t->config.target_vocab_size == 300 && t->config.unk_id == -1 && t->config.min_pair_freq == 2, "config copied");
    CHECK(std::fabs(t->config.character_coverage - 0.99f) < 1e-6f, "coverage copied");
    CHECK(t->num_merges == 0 && t->merge_ops != NULL, "fresh trainer has no merges");
    bpe_trainer_destroy(t);
  }
  {  // 2. defaults (bpe_test.cpp:79-95): coverage outside (0,1) -> 0.995, min_pair_freq 0 -> 2000
    BPEConfig cfg = {1000, 0, 0.0f, 0};
    Trainer* t = create_trainer(&cfg);
    CHECK(std::fabs(t->config.character_coverage - 0.995f) < 1e-6f, "default coverage");
    CHECK(t->config.min_pair_freq == 2000, "default min_pair_freq");
    bpe_trainer_destroy(t);
  }
  {  // 3. corpus loading (bpe_test.cpp:97-131)
    BPEConfig cfg = {300, -1, 0.99f, 2};
    Trainer* t = create_trainer(&cfg);
    CHECK(bpe_load_corpus(t, corpus) == 0, "load_corpus succeeds");
    CHECK(t->corpus.vocab_size > 0, "unique words found");
    bool ok = true;
    for (size_t i = 0; i < t->corpus.vocab_size; i++) ok = ok && t->corpus.words[i] != NULL && t->corpus.word_counts[i] > 0;
    CHECK(ok, "every word has a symbol pointer and a positive count");
    // 4. bigram counting seeds the heap (bpe_test.cpp:133-168)
    bpe_count_bigrams(t);
    CHECK(t->heap.size > 1, "heap seeded");
    CHECK(t->heap.data[0].freq >= t->heap.data[1].freq && t->heap.data[0].freq >= 2, "heap top is the maximum");
    // 5. a single merge (bpe_test.cpp:170-204)
    size_t before = t->num_merges;
    CHECK(bpe_merge_batch(t, 1) == 1, "merge_batch(1) performs one merge");
    CHECK(t->num_merges == before + 1, "num_merges advanced");
    CHECK(t->merge_ops[0].first >= 0 && t->merge_ops[0].second >= 0, "merge operands recorded");
    bpe_trainer_destroy(t);
  }
  {  // 6. full training (bpe_test.cpp:206-235) and 7. save (bpe_test.cpp:237-299)
    BPEConfig cfg = {300, 0, 0.995f, 2};
    Trainer* t = create_trainer(&cfg);
    CHECK(bpe_load_corpus(t, corpus) == 0, "load for training");
    int merges = bpe_train(t);
    CHECK(merges > 0 && merges <= 300 - 256, "train performs at most vocab-256 merges");
    CHECK(t->num_merges == static_cast<size_t>(merges), "num_merges matches the return value");
    const std::string model = tmp + "/c_model.bin", vocab = tmp + "/c_vocab.txt";
    bpe_save(t, model.c_str(), vocab.c_str());
    CHECK(file_size(model) == static_cast<long>(merges) * 12, "model file holds merges x 3 int32");
    CHECK(count_lines(vocab) == 256 + merges + 1, "vocab file holds 256 + merges tokens (token 10 is a newline)");
    bpe_trainer_destroy(t);
  }
  {  // 8. missing file (bpe_test.cpp:301-318)
    BPEConfig cfg = {300, 0, 0.995f, 2};
    Trainer* t = create_trainer(&cfg);
    CHECK(bpe_load_corpus(t, "/nonexistent/definitely/missing.txt") == -1, "missing corpus -> -1");
    CHECK(bpe_load_corpus(NULL, corpus) == -1 && bpe_train(NULL) == -1 && bpe_merge_batch(NULL, 1) == -1, "NULL trainer -> -1");
    bpe_trainer_destroy(t);
  }
  std::printf("%d failed\n", failures);
  return failures;
}
