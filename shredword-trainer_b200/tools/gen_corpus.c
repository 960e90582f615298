/* gen_corpus.c -- deterministic, integer-only synthetic corpus generator.
 *
 * Produces the benchmark inputs BASELINE.json names: "synthetic Zipfian text" (mode zipf) and the
 * "synthetic multilingual-codepoint corpus" (mode multi).  Everything is derived from splitmix64, so the
 * same (bytes, seed, w, mode) gives the same file on every machine and golden checksums stay valid.
 *
 *   type spelling, rank r in [1, 2^w):  h0 = sm(seed ^ r*GOLD); L = 1 + popcount(h0 & 0x7FFF)  (1..16)
 *       zipf : letter j = TAB[sm(h0 + j) >> 58]          (64-entry table over a-z, roughly k^-0.8)
 *       multi: script = r mod 7, L = 1 + (h0 & 7), codepoint j uniform in the script's block, UTF-8
 *   token stream ("octave Zipf", s ~ 1):  u = sm(seed' + i); k = (u >> 32) mod w; r = 2^k + (u & (2^k - 1))
 *   tokens are separated by one space, every 16th by '\n'; generation stops at the first multiple of
 *   4096 tokens whose byte total reaches the requested size.
 *
 * usage: gen_corpus <out> <bytes> <seed> <w> [zipf|multi] [threads]
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define GOLD 0x9E3779B97F4A7C15ULL
#define BLOCK_TOKENS (1u << 20) /* work unit of one thread */
#define CUT_TOKENS 4096u         /* granularity of the end of file */

static inline uint64_t sm(uint64_t x) {
  x += GOLD;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

static const char TAB[65] = "eeeeeeee" "tttttt" "aaaaa" "oooo" "iiii" "nnn" "sss" "hhh" "rrr"
                            "dd" "ll" "cc" "uu" "mm" "ww" "ff" "gypbvkjxqz" "e";

static const struct { uint32_t base, span; } SCRIPTS[7] = {
    {0x61, 26},     /* Latin a-z */
    {0x430, 32},    /* Cyrillic */
    {0x3B1, 25},    /* Greek */
    {0x627, 36},    /* Arabic */
    {0x905, 53},    /* Devanagari */
    {0x4E00, 4096}, /* CJK */
    {0xAC00, 2048}, /* Hangul */
};

static int g_multi = 0, g_w = 20;
static uint64_t g_seed = 1, g_seed2;

static inline size_t put_utf8(uint8_t *p, uint32_t cp) {
  if (cp < 0x80) { p[0] = (uint8_t)cp; return 1; }
  if (cp < 0x800) { p[0] = 0xC0 | (cp >> 6); p[1] = 0x80 | (cp & 63); return 2; }
  p[0] = 0xE0 | (cp >> 12); p[1] = 0x80 | ((cp >> 6) & 63); p[2] = 0x80 | (cp & 63); return 3;
}

static inline size_t spell(uint8_t *p, uint64_t r) {
  uint64_t h0 = sm(g_seed ^ (r * GOLD));
  if (!g_multi) {
    int L = 1 + __builtin_popcountll(h0 & 0x7FFF);
    for (int j = 0; j < L; j++) p[j] = (uint8_t)TAB[sm(h0 + (uint64_t)j) >> 58];
    return (size_t)L;
  }
  int sc = (int)(r % 7), L = 1 + (int)(h0 & 7);
  size_t n = 0;
  for (int j = 0; j < L; j++) n += put_utf8(p + n, SCRIPTS[sc].base + (uint32_t)(sm(h0 + (uint64_t)j) % SCRIPTS[sc].span));
  return n;
}

/* one block = BLOCK_TOKENS tokens starting at token index blk * BLOCK_TOKENS; at most 25 bytes per token.
 * cut[c] = bytes written after (c + 1) * CUT_TOKENS tokens of the block. */
static size_t gen_block(uint8_t *out, uint64_t blk, size_t *cut) {
  size_t n = 0;
  uint64_t i0 = blk * (uint64_t)BLOCK_TOKENS;
  for (uint64_t i = i0; i < i0 + BLOCK_TOKENS; i++) {
    if (i > i0 && ((i - i0) % CUT_TOKENS) == 0) cut[(i - i0) / CUT_TOKENS - 1] = n;
    uint64_t u = sm(g_seed2 + i);
    unsigned k = (unsigned)((u >> 32) % (uint64_t)g_w);
    uint64_t r = (1ULL << k) + (u & ((1ULL << k) - 1));
    n += spell(out + n, r);
    out[n++] = ((i & 15) == 15) ? '\n' : ' ';
  }
  cut[BLOCK_TOKENS / CUT_TOKENS - 1] = n;
  return n;
}

typedef struct { uint8_t *buf; size_t n; uint64_t blk; size_t cut[BLOCK_TOKENS / CUT_TOKENS]; } Job;
static void *worker(void *a) { Job *j = (Job *)a; j->n = gen_block(j->buf, j->blk, j->cut); return NULL; }

int main(int argc, char **argv) {
  if (argc < 5) { fprintf(stderr, "usage: %s out bytes seed w [zipf|multi] [threads]\n", argv[0]); return 2; }
  const char *out = argv[1];
  uint64_t want = strtoull(argv[2], 0, 10);
  g_seed = strtoull(argv[3], 0, 10);
  g_w = atoi(argv[4]);
  if (g_w < 1 || g_w > 40) { fprintf(stderr, "w out of range\n"); return 2; }
  if (argc > 5 && !strcmp(argv[5], "multi")) g_multi = 1;
  int nt = argc > 6 ? atoi(argv[6]) : 8;
  if (nt < 1) nt = 1;
  if (nt > 64) nt = 64;
  g_seed2 = sm(g_seed ^ 0xC0FFEEULL);
  FILE *f = fopen(out, "wb");
  if (!f) { perror("fopen"); return 1; }
  Job *jobs = (Job *)calloc((size_t)nt, sizeof(Job));
  pthread_t *th = (pthread_t *)calloc((size_t)nt, sizeof(pthread_t));
  for (int t = 0; t < nt; t++) jobs[t].buf = (uint8_t *)malloc((size_t)BLOCK_TOKENS * 26);
  uint64_t total = 0, blk = 0, tokens = 0;
  int done = 0;
  while (!done) {
    for (int t = 0; t < nt; t++) { jobs[t].blk = blk + (uint64_t)t; pthread_create(&th[t], NULL, worker, &jobs[t]); }
    for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
    for (int t = 0; t < nt && !done; t++) {
      size_t take = jobs[t].n;
      uint64_t ntok = BLOCK_TOKENS;
      for (unsigned c = 0; c < BLOCK_TOKENS / CUT_TOKENS; c++)
        if (total + jobs[t].cut[c] >= want) { take = jobs[t].cut[c]; ntok = (uint64_t)(c + 1) * CUT_TOKENS; done = 1; break; }
      fwrite(jobs[t].buf, 1, take, f);
      total += take; tokens += ntok;
    }
    blk += (uint64_t)nt;
  }
  fclose(f);
  printf("{\"bytes\": %llu, \"tokens\": %llu, \"seed\": %llu, \"w\": %d, \"mode\": \"%s\"}\n", (unsigned long long)total,
         (unsigned long long)tokens, (unsigned long long)g_seed, g_w, g_multi ? "multi" : "zipf");
  return 0;
}
