// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE (not product code).
//
// Drives an implementation of the shredword BPE C ABI (reference
// shredword/csrc/bpe/bpe.h:62-72) that is given as a shared library on the
// command line, times bpe_load_corpus / bpe_train separately and writes the
// artefacts parity is judged on.  It is used with oracle/_ref/libtrainer_ref.so
// (the UNMODIFIED reference compiled from /root/reference by oracle/Makefile,
// run under the zero-fill malloc shim) to produce golden vectors and the CPU
// baseline, and it can equally be pointed at the product libtrainer.so.
//
// Only the leading, ABI-visible part of the reference Trainer struct is
// mirrored here (offsets verified with offsetof on the
// reference build): config@0, heap@24, corpus@48, num_merges@96, merge_ops@104.
//
// usage: ref_harness <lib.so> <corpus> <vocab_size> <unk_id> <coverage> <min_pair_freq>
//                    [--merges out.bin] [--model m.bin] [--vocab v.txt]
//                    [--max-merges K] [--json out.json]
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

struct AbiConfig { size_t target_vocab_size; int32_t unk_id; float character_coverage; uint64_t min_pair_freq; };
struct AbiPair { int32_t first, second; };
struct AbiTrainerView {
  AbiConfig config;               // @0
  void *heap_data; size_t heap_size, heap_cap;  // @24
  void **words; uint64_t *word_counts; size_t n_words;  // @48
  void *bimap_buckets; size_t bimap_n;  // @72
  size_t next_token;              // @88
  size_t num_merges;              // @96
  AbiPair *merge_ops;             // @104
};

static double now_s() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

int main(int argc, char **argv) {
  if (argc < 7) { fprintf(stderr, "usage: %s lib corpus vocab unk cov minfreq [opts]\n", argv[0]); return 2; }
  const char *libp = argv[1], *corpus = argv[2];
  AbiConfig cfg; cfg.target_vocab_size = (size_t)atoll(argv[3]); cfg.unk_id = (int32_t)atoi(argv[4]);
  cfg.character_coverage = (float)atof(argv[5]); cfg.min_pair_freq = (uint64_t)atoll(argv[6]);
  const char *merges_out = 0, *model_out = 0, *vocab_out = 0, *json_out = 0; long max_merges = -1;
  for (int i = 7; i + 1 < argc; i += 2) {
    if (!strcmp(argv[i], "--merges")) merges_out = argv[i + 1];
    else if (!strcmp(argv[i], "--model")) model_out = argv[i + 1];
    else if (!strcmp(argv[i], "--vocab")) vocab_out = argv[i + 1];
    else if (!strcmp(argv[i], "--json")) json_out = argv[i + 1];
    else if (!strcmp(argv[i], "--max-merges")) max_merges = atol(argv[i + 1]);
  }
  void *h = dlopen(libp, RTLD_NOW | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  typedef void *(*create_t)(const AbiConfig *);
  typedef int (*load_t)(void *, const char *);
  typedef int (*train_t)(void *);
  typedef void (*init_t)(void *);
  typedef int (*batch_t)(void *, int);
  typedef void (*save_t)(const void *, const char *, const char *);
  create_t create = (create_t)dlsym(h, "create_trainer");
  load_t load = (load_t)dlsym(h, "bpe_load_corpus");
  train_t train = (train_t)dlsym(h, "bpe_train");
  init_t init = (init_t)dlsym(h, "bpe_init");
  batch_t batch = (batch_t)dlsym(h, "bpe_merge_batch");
  save_t save = (save_t)dlsym(h, "bpe_save");
  if (!create || !load || !train || !init || !batch || !save) { fprintf(stderr, "missing ABI symbol\n"); return 2; }

  // The implementations log one line per merge on stdout; keep our own report on a dup of it.
  fflush(stdout);
  int report_fd = dup(1);
  if (!freopen("/dev/null", "w", stdout)) return 2;
  FILE *report = fdopen(report_fd, "w");

  void *t = create(&cfg);
  double t0 = now_s();
  int rc = load(t, corpus);
  double t1 = now_s();
  if (rc != 0) { fprintf(report, "{\"error\": \"load rc=%d\"}\n", rc); return 1; }
  AbiTrainerView *v = (AbiTrainerView *)t;
  size_t n_words = v->n_words;
  int merges = 0;
  double t2 = now_s();
  double t_init = 0;
  if (max_merges < 0) {
    merges = train(t);
  } else {
    // Bounded sample (bpe_train = bpe_init + repeated bpe_merge_batch, reference bpe.cpp:345-386).
    init(t);
    t_init = now_s() - t2;
    long target = (long)cfg.target_vocab_size - 256;
    if (max_merges < target) target = max_merges;
    while (merges < target) { int m = batch(t, 1); if (m <= 0) break; merges += m; }
  }
  double t3 = now_s();
  size_t M = v->num_merges;
  if (merges_out) {
    FILE *f = fopen(merges_out, "wb");
    for (size_t m = 0; m < M; m++) { int32_t rec[3] = {v->merge_ops[m].first, v->merge_ops[m].second, (int32_t)(256 + m)}; fwrite(rec, 4, 3, f); }
    fclose(f);
  }
  if (model_out && vocab_out) {
    // The reference indexes freq[unk_id] unchecked (bpe.cpp:413): with unk_id outside [0, 256+M) it corrupts the
    // heap and aborts in free() AFTER both files are complete.  Isolate that in a child.
    fflush(report);
    pid_t pid = fork();
    if (pid == 0) { save(t, model_out, vocab_out); fflush(NULL); _exit(0); }
    int st = 0; waitpid(pid, &st, 0);
  }
  char buf[1024];
  snprintf(buf, sizeof buf,
           "{\"n_words\": %zu, \"merges\": %d, \"num_merges\": %zu, \"load_s\": %.6f, \"train_s\": %.6f, \"init_s\": %.6f}\n",
           n_words, merges, M, t1 - t0, t3 - t2, t_init);
  fputs(buf, report); fflush(report);
  if (json_out) { FILE *f = fopen(json_out, "w"); fputs(buf, f); fclose(f); }
  _exit(0);  // skip the reference's destroy path (it frees uninitialised pointers in corner cases)
}
