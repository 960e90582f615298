// cli_main.cpp -- trainer.exe: drop-in for the reference CLI (shredword/csrc/trainer.cpp) for model_type=bpe.
//
// Same key=value arguments and defaults (trainer.cpp:43-68): vocab_size 32000, character_coverage 0.9995,
// min_pair_freq 2000, and -- because the reference CLI has no unk_id key -- unk_id = -1.  Unknown keys and arguments
// without '=' are ignored; missing required keys print the usage and exit 1; no arguments prints the usage and exits 0
// (trainer.cpp:188-200).  The reference aborts after writing its files when unk_id = -1 (it indexes freq[-1],
// bpe.cpp:413); this CLI writes the same files and exits 0.  model_type=unigram is out of scope and exits 1.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/shred_abi.h"

static void print_usage(const char* prog) {
  std::printf("Usage: %s <args>\n\n", prog);
  std::printf("Arguments (use: key=value format):\n");
  std::printf("  input=<path>              Input corpus file\n");
  std::printf("  model_type=<bpe|unigram>  Model type (this build: bpe)\n");
  std::printf("  output_model=<path>       Output model file\n");
  std::printf("  output_vocab=<path>       Output vocab file\n");
  std::printf("  vocab_size=<int>          Target vocab size (default: 32000)\n");
  std::printf("  character_coverage=<float> Coverage 0.0-1.0 (default: 0.9995)\n");
  std::printf("  min_pair_freq=<int>       Min pair freq BPE (default: 2000)\n");
}

int main(int argc, char** argv) {
  std::printf("Tokenizer Trainer CLI v1.0 (B200)\n=================================\n");
  if (argc < 2) { print_usage(argv[0]); return 0; }
  std::string input, model_type, out_model, out_vocab;
  bool has_input = false, has_type = false, has_model = false, has_vocab = false;
  int vocab_size = 32000;
  float coverage = 0.9995f;
  uint64_t min_pair_freq = 2000;
  const int32_t unk_id = -1;  // trainer.cpp:47, not overridable from the command line
  for (int i = 1; i < argc; i++) {
    char* eq = std::strchr(argv[i], '=');
    if (!eq) continue;
    std::string key(argv[i], eq - argv[i]);
    const char* value = eq + 1;
    if (key == "input") { input = value; has_input = true; }
    else if (key == "model_type") { model_type = value; has_type = true; }
    else if (key == "output_model") { out_model = value; has_model = true; }
    else if (key == "output_vocab") { out_vocab = value; has_vocab = true; }
    else if (key == "vocab_size") vocab_size = std::atoi(value);
    else if (key == "character_coverage") coverage = static_cast<float>(std::atof(value));
    else if (key == "min_pair_freq") min_pair_freq = static_cast<uint64_t>(std::atoll(value));
    // num_iterations / seed_size / max_piece_length are Unigram-only keys: accepted and ignored
  }
  if (!has_input || !has_type || !has_model || !has_vocab) {
    std::fprintf(stderr, "[ERROR] Missing required arguments\n\n");
    print_usage(argv[0]);
    return 1;
  }
  if (model_type != "bpe" && model_type != "unigram") {
    std::fprintf(stderr, "[ERROR] Invalid model_type. Must be 'bpe' or 'unigram'\n");
    return 1;
  }
  if (model_type == "unigram") {
    std::fprintf(stderr, "[ERROR] model_type=unigram is not part of the B200 BPE trainer\n");
    return 1;
  }
  std::printf("\n========== BPE Training ==========\n");
  std::printf("[CONFIG] Vocab Size: %d\n", vocab_size);
  std::printf("[CONFIG] Character Coverage: %.4f\n", coverage);
  std::printf("[CONFIG] Min Pair Freq: %llu\n", static_cast<unsigned long long>(min_pair_freq));
  BPEConfig cfg;
  cfg.target_vocab_size = static_cast<size_t>(vocab_size); cfg.unk_id = unk_id; cfg.character_coverage = coverage; cfg.min_pair_freq = min_pair_freq;
  Trainer* t = create_trainer(&cfg);
  if (!t) { std::fprintf(stderr, "[ERROR] Failed to create BPE trainer\n"); return 1; }
  std::printf("\n[STEP 1] Loading corpus from: %s\n", input.c_str());
  if (bpe_load_corpus(t, input.c_str()) != 0) {
    std::fprintf(stderr, "[ERROR] Failed to load corpus\n");
    bpe_trainer_destroy(t);
    return 255;  // the reference returns -1 from main here
  }
  std::printf("[INFO] Corpus loaded successfully. Vocabulary: %zu words\n", t->corpus.vocab_size);
  std::printf("\n[STEP 2] Training BPE model...\n");
  int merges = bpe_train(t);
  if (merges < 0) { std::fprintf(stderr, "[ERROR] Training failed\n"); bpe_trainer_destroy(t); return 255; }
  std::printf("[SUCCESS] Training completed with %d merges\n", merges);
  std::printf("\n[STEP 3] Saving model and vocabulary...\n");
  bpe_save(t, out_model.c_str(), out_vocab.c_str());
  std::printf("[SUCCESS] Saved to:\n  Model: %s\n  Vocab: %s\n", out_model.c_str(), out_vocab.c_str());
  bpe_trainer_destroy(t);
  std::printf("\n========== Training Complete ==========\n");
  return 0;
}
