/* encoder_consumer.c -- a plain C consumer of the encoder entry points of include/shred_abi.h (no Python, no C++):
 * train -> save -> bpe_b200_encoder_load -> encode (resident and streamed) -> decode, checking what a C caller can check
 * without an oracle: both entry points agree, offsets are a CSR over the words of the text, decode gives the text back
 * without its delimiters, error codes.  usage: encoder_consumer <corpus> <scratch dir>; exit code = number of failures. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/shred_abi.h"

static int failures = 0;
#define CHECK(cond, what) do { if (cond) printf("[PASS] %s\n", what); else { printf("[FAIL] %s\n", what); failures++; } } while (0)

static int is_delim(unsigned char c) { return c == 9 || c == 10 || c == 13 || c == 32; }

int main(int argc, char** argv) {
  if (argc < 3) return 100;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 101;
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  uint8_t* text = (uint8_t*)malloc((size_t)n + 1);
  if (fread(text, 1, (size_t)n, f) != (size_t)n) return 102;
  fclose(f);

  char model[4096], vocab[4096];
  snprintf(model, sizeof model, "%s/enc_model.bin", argv[2]);
  snprintf(vocab, sizeof vocab, "%s/enc_vocab.txt", argv[2]);
  BPEConfig cfg = {400, 0, 0.995f, 2};
  Trainer* t = create_trainer(&cfg);
  CHECK(bpe_load_corpus(t, argv[1]) == 0, "trainer loads the corpus");
  const int merges = bpe_train(t);
  CHECK(merges > 0, "trainer learns merges");
  bpe_save(t, model, vocab);
  bpe_trainer_destroy(t);

  shred_encoder_t* e = bpe_b200_encoder_load(model);
  CHECK(e != NULL, "encoder loads the model file bpe_save wrote");
  if (!e) return failures;
  CHECK(bpe_b200_encoder_vocab_size(e) == (size_t)(256 + merges), "vocab size = 256 + merges");

  uint64_t n_words = 0, n_ids = 0;
  CHECK(bpe_b200_encode(e, text, (uint64_t)n, &n_words, &n_ids) == 0, "bpe_b200_encode");
  uint64_t words = 0;
  for (long i = 0; i < n; i++) if (!is_delim(text[i]) && (i == 0 || is_delim(text[i - 1]))) words++;
  CHECK(n_words == words, "one row per whitespace-delimited word");
  int32_t* ids = (int32_t*)malloc((size_t)(n_ids + 1) * 4);
  uint64_t* off = (uint64_t*)malloc((size_t)(n_words + 1) * 8);
  CHECK(bpe_b200_encode_fetch(e, ids, off) == 0, "bpe_b200_encode_fetch");
  int csr = off[0] == 0 && off[n_words] == n_ids;
  for (uint64_t w = 0; w < n_words && csr; w++) csr = off[w] < off[w + 1];
  CHECK(csr, "offsets are a strictly increasing CSR ending at n_ids");
  int in_range = 1;
  for (uint64_t i = 0; i < n_ids; i++) if (ids[i] < 0 || ids[i] >= 256 + merges) in_range = 0;
  CHECK(in_range, "every id is inside the vocabulary");
  CHECK(n_ids < (uint64_t)n / 2, "merges were applied (fewer ids than half the bytes)");

  /* streamed entry point: worst-case capacities, then exact ones, then too small */
  int32_t* ids2 = (int32_t*)malloc((size_t)n * 4 + 4);
  uint64_t* off2 = (uint64_t*)malloc(((size_t)n / 2 + 2) * 8);
  uint64_t w2 = 0, i2 = 0;
  CHECK(bpe_b200_encode_to_host(e, text, (uint64_t)n, ids2, (uint64_t)n, off2, (uint64_t)n / 2 + 2, &w2, &i2) == 0, "bpe_b200_encode_to_host");
  CHECK(w2 == n_words && i2 == n_ids && memcmp(ids, ids2, (size_t)n_ids * 4) == 0 && memcmp(off, off2, (size_t)(n_words + 1) * 8) == 0,
        "streamed result == resident result");
  CHECK(bpe_b200_encode_to_host(e, text, (uint64_t)n, ids2, n_ids, off2, n_words + 1, &w2, &i2) == 0, "exact capacities suffice");
  CHECK(bpe_b200_encode_to_host(e, text, (uint64_t)n, ids2, n_ids - 1, off2, n_words + 1, &w2, &i2) == -3, "-3 when ids_cap is too small");
  CHECK(bpe_b200_encode_to_host(e, text, (uint64_t)n, ids2, n_ids, NULL, 0, &w2, &i2) == 0 && i2 == n_ids, "offsets_out may be NULL");

  /* decode: size query, then the bytes = the text without its delimiters */
  const int64_t need = bpe_b200_decode(e, ids, n_ids, NULL, 0);
  uint8_t* back = (uint8_t*)malloc((size_t)(need > 0 ? need : 1));
  CHECK(need > 0 && bpe_b200_decode(e, ids, n_ids, back, (uint64_t)need) == need, "bpe_b200_decode");
  int same = 1; int64_t j = 0;
  for (long i = 0; i < n && same; i++) if (!is_delim(text[i])) { if (j >= need || back[j] != text[i]) same = 0; j++; }
  CHECK(same && j == need, "decode(encode(text)) == text without delimiters");
  int32_t bad_id = 256 + merges;
  CHECK(bpe_b200_decode(e, &bad_id, 1, back, (uint64_t)need) == -2, "-2 for an id outside the vocabulary");

  shred_encode_stats_t st;
  CHECK(bpe_b200_encoder_get_stats(e, &st) == 0 && st.kernel_launches > 0, "stats report kernel launches");
  bpe_b200_encoder_destroy(e);
  const int32_t bad_model[3] = {97, 98, 300};
  CHECK(bpe_b200_encoder_create(bad_model, 1) == NULL, "a model bpe_save cannot have written is rejected");
  printf("%d failed\n", failures);
  return failures;
}
