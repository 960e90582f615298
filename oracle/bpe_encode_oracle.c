/* bpe_encode_oracle.c -- TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg).
 *
 * CPU restatement of the reference's BPE *encoder* for the model file the BPE trainer writes (SURVEY.md section 8f,
 * rank 2).  The reference has no native encoder; its algorithm is the pure-Python
 *     BPETokenizer._encode_chunk  shredword/utils/bpe.py:191-203   (+ get_stats :10-21, merge :23-38)
 *     build_vocab                 shredword/utils/bpe.py:61-79     (token bytes, used by decode :214-225)
 * applied here to the trainer's own artefacts:
 *     model file = M x {int32 a, int32 b, int32 256+m}, native endian    shredword/csrc/bpe/bpe.cpp:419-427
 *     words      = maximal runs of bytes outside {9, 10, 13, 32}         shredword/csrc/bpe/bpe.cpp:131-152
 * (the Python encoder's regex pre-tokeniser belongs to the Python trainer's own model format, not to model.bin).
 *
 * PARITY PINNED: tests/golden/encode_golden.json is produced by tests/golden/make_encode_golden.py, which imports the
 * unmodified reference module shredword/utils/bpe.py, fills BPETokenizer.merges from model files written by the pinned
 * reference trainer and calls its _encode_chunk / decode; tests/test_encode_oracle.py checks this file against it.
 *
 * Semantics restated
 *   merges : dict (a, b) -> new id, filled in file order, so a repeated pair keeps the LAST id (Python dict assignment)
 *   encode : ids = the word's bytes; while len >= 2: take the adjacent pair with the smallest merge id (:196); stop if no
 *            adjacent pair is in the dict (:197); replace every occurrence left to right, non-overlapping (:23-38)
 *   decode : concatenation of vocab[id] (:217-220); an id outside the vocab is an error (ValueError there, -1 here)
 *   A model is accepted iff triple m has new == 256 + m and 0 <= a, b < new (what the trainer can write).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  size_t n_merges;
  int32_t* tri;       /* 3 * n_merges */
  uint64_t* mkey;     /* open addressing: (a << 32 | b) + 1, 0 = empty */
  int32_t* mval;
  uint64_t mmask;
  /* token bytes */
  uint64_t* tok_off;  /* 256 + n_merges + 1 */
  uint8_t* tok_bytes;
  /* result of the last encode */
  int32_t* ids; uint64_t n_ids, ids_cap;
  uint64_t* off; uint64_t n_tok, off_cap;
  /* cache of encoded unique words: purely a speed-up of this checker */
  uint64_t* ckey; uint64_t* cpos; uint64_t cmask, cused;
  uint8_t* arena; uint64_t arena_n, arena_cap;
} EncOracle;

static uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

static int32_t merge_lookup(const EncOracle* e, int32_t a, int32_t b) {
  const uint64_t k = (((uint64_t)(uint32_t)a << 32) | (uint32_t)b) + 1;
  for (uint64_t s = mix(k) & e->mmask;; s = (s + 1) & e->mmask) {
    if (e->mkey[s] == 0) return -1;
    if (e->mkey[s] == k) return e->mval[s];
  }
}

void enc_oracle_destroy(EncOracle* e) {
  if (!e) return;
  free(e->tri); free(e->mkey); free(e->mval); free(e->tok_off); free(e->tok_bytes); free(e->ids); free(e->off);
  free(e->ckey); free(e->cpos); free(e->arena); free(e);
}

EncOracle* enc_oracle_create(const int32_t* triples, size_t n) {
  for (size_t m = 0; m < n; m++) {
    const int32_t a = triples[3 * m], b = triples[3 * m + 1], c = triples[3 * m + 2];
    if (c != (int32_t)(256 + m) || a < 0 || b < 0 || a >= c || b >= c) return NULL;
  }
  EncOracle* e = (EncOracle*)calloc(1, sizeof *e);
  e->n_merges = n;
  e->tri = (int32_t*)malloc(3 * (n + 1) * sizeof(int32_t));
  memcpy(e->tri, triples, 3 * n * sizeof(int32_t));
  uint64_t cap = 16; while (cap < 2 * n + 2) cap <<= 1;
  e->mmask = cap - 1;
  e->mkey = (uint64_t*)calloc(cap, 8); e->mval = (int32_t*)calloc(cap, 4);
  for (size_t m = 0; m < n; m++) {  /* dict assignment in file order: the last id of a repeated pair stays */
    const uint64_t k = (((uint64_t)(uint32_t)triples[3 * m] << 32) | (uint32_t)triples[3 * m + 1]) + 1;
    uint64_t s = mix(k) & e->mmask;
    while (e->mkey[s] != 0 && e->mkey[s] != k) s = (s + 1) & e->mmask;
    e->mkey[s] = k; e->mval[s] = triples[3 * m + 2];
  }
  /* build_vocab (utils/bpe.py:74-76): vocab[idx] = vocab[p0] + vocab[p1] */
  const size_t T = 256 + n;
  e->tok_off = (uint64_t*)malloc((T + 1) * 8);
  uint64_t total = 256;
  for (size_t i = 0; i <= 256; i++) e->tok_off[i] = i;
  for (size_t m = 0; m < n; m++) {
    const int32_t a = triples[3 * m], b = triples[3 * m + 1];
    total += (e->tok_off[a + 1] - e->tok_off[a]) + (e->tok_off[b + 1] - e->tok_off[b]);
    e->tok_off[256 + m + 1] = total;
  }
  e->tok_bytes = (uint8_t*)malloc(total + 1);
  for (int i = 0; i < 256; i++) e->tok_bytes[i] = (uint8_t)i;
  for (size_t m = 0; m < n; m++) {
    const int32_t a = triples[3 * m], b = triples[3 * m + 1];
    uint8_t* dst = e->tok_bytes + e->tok_off[256 + m];
    const uint64_t la = e->tok_off[a + 1] - e->tok_off[a], lb = e->tok_off[b + 1] - e->tok_off[b];
    memcpy(dst, e->tok_bytes + e->tok_off[a], la);
    memcpy(dst + la, e->tok_bytes + e->tok_off[b], lb);
  }
  e->cmask = (1u << 16) - 1;
  e->ckey = (uint64_t*)calloc(e->cmask + 1, 8); e->cpos = (uint64_t*)calloc(e->cmask + 1, 8);
  return e;
}

EncOracle* enc_oracle_load(const char* model_path) {
  FILE* f = fopen(model_path, "rb");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  if (sz < 0 || sz % 12 != 0) { fclose(f); return NULL; }
  int32_t* t = (int32_t*)malloc((size_t)sz + 12);
  const size_t got = fread(t, 1, (size_t)sz, f);
  fclose(f);
  EncOracle* e = got == (size_t)sz ? enc_oracle_create(t, (size_t)sz / 12) : NULL;
  free(t);
  return e;
}

/* _encode_chunk (utils/bpe.py:191-203), literally: ids in place in `w` (len entries), returns the new length */
static size_t encode_chunk(const EncOracle* e, int32_t* w, size_t len) {
  while (len >= 2) {
    int32_t best = -1;  /* smallest merge id among the adjacent pairs (:196) */
    for (size_t i = 0; i + 1 < len; i++) {
      const int32_t r = merge_lookup(e, w[i], w[i + 1]);
      if (r >= 0 && (best < 0 || r < best)) best = r;
    }
    if (best < 0) break;  /* :197 */
    const int32_t a = e->tri[3 * (best - 256)], b = e->tri[3 * (best - 256) + 1];  /* the dict key that owns this id */
    size_t o = 0, i = 0;  /* merge (:23-38) */
    while (i < len) {
      if (i + 1 < len && w[i] == a && w[i + 1] == b) { w[o++] = best; i += 2; }
      else { w[o++] = w[i]; i += 1; }
    }
    len = o;
  }
  return len;
}

size_t enc_oracle_encode_word(EncOracle* e, const uint8_t* word, size_t len, int32_t* out) {
  for (size_t i = 0; i < len; i++) out[i] = word[i];
  return encode_chunk(e, out, len);
}

static int is_delim(uint8_t c) { return c == 9 || c == 10 || c == 13 || c == 32; }

static void cache_grow(EncOracle* e) {
  const uint64_t ncap = (e->cmask + 1) * 2;
  uint64_t* nk = (uint64_t*)calloc(ncap, 8); uint64_t* np = (uint64_t*)calloc(ncap, 8);
  for (uint64_t s = 0; s <= e->cmask; s++) if (e->ckey[s]) {
    uint64_t d = e->ckey[s] & (ncap - 1);
    while (nk[d]) d = (d + 1) & (ncap - 1);
    nk[d] = e->ckey[s]; np[d] = e->cpos[s];
  }
  free(e->ckey); free(e->cpos);
  e->ckey = nk; e->cpos = np; e->cmask = ncap - 1;
}

/* arena record of a cached word: u32 len, u32 n_ids, bytes[len], pad to 4, int32 ids[n_ids] */
int enc_oracle_encode(EncOracle* e, const uint8_t* text, uint64_t n) {
  e->n_ids = 0; e->n_tok = 0;
  int32_t* scratch = NULL; size_t scratch_cap = 0;
  uint64_t i = 0;
  for (;;) {
    while (i < n && is_delim(text[i])) i++;
    if (e->n_tok + 2 > e->off_cap) { e->off_cap = e->off_cap ? e->off_cap * 2 : 1024; e->off = (uint64_t*)realloc(e->off, e->off_cap * 8); }
    e->off[e->n_tok] = e->n_ids;
    if (i >= n) break;
    const uint64_t st = i;
    uint64_t h = 1469598103934665603ULL;
    while (i < n && !is_delim(text[i])) { h = (h ^ text[i]) * 1099511628211ULL; i++; }
    const uint32_t len = (uint32_t)(i - st);
    h = mix(h ^ len) | 1;
    uint64_t s = h & e->cmask;
    const uint8_t* rec = NULL;
    while (e->ckey[s]) {
      if (e->ckey[s] == h) {
        const uint8_t* r = e->arena + e->cpos[s];
        uint32_t rl; memcpy(&rl, r, 4);
        if (rl == len && memcmp(r + 8, text + st, len) == 0) { rec = r; break; }
      }
      s = (s + 1) & e->cmask;
    }
    if (!rec) {
      if (len > scratch_cap) { scratch_cap = (size_t)len * 2; scratch = (int32_t*)realloc(scratch, scratch_cap * 4); }
      const uint32_t k = (uint32_t)enc_oracle_encode_word(e, text + st, len, scratch);
      const uint64_t need = 8 + (((uint64_t)len + 3) & ~3ull) + 4ull * k;
      if (e->arena_n + need > e->arena_cap) { e->arena_cap = (e->arena_cap + need) * 2; e->arena = (uint8_t*)realloc(e->arena, e->arena_cap); }
      uint8_t* r = e->arena + e->arena_n;
      memcpy(r, &len, 4); memcpy(r + 4, &k, 4); memcpy(r + 8, text + st, len);
      memcpy(r + 8 + (((uint64_t)len + 3) & ~3ull), scratch, 4ull * k);
      e->ckey[s] = h; e->cpos[s] = e->arena_n;
      e->arena_n += need;
      rec = e->arena + e->cpos[s];
      if (++e->cused * 2 > e->cmask) cache_grow(e);
    }
    uint32_t rl, k; memcpy(&rl, rec, 4); memcpy(&k, rec + 4, 4);
    if (e->n_ids + k > e->ids_cap) { e->ids_cap = (e->ids_cap + k) * 2; e->ids = (int32_t*)realloc(e->ids, e->ids_cap * 4); }
    memcpy(e->ids + e->n_ids, rec + 8 + (((uint64_t)rl + 3) & ~3ull), 4ull * k);
    e->n_ids += k;
    e->n_tok++;
  }
  free(scratch);
  return 0;
}

uint64_t enc_oracle_n_tokens(const EncOracle* e) { return e->n_tok; }
uint64_t enc_oracle_n_ids(const EncOracle* e) { return e->n_ids; }
const int32_t* enc_oracle_ids(const EncOracle* e) { return e->ids; }
const uint64_t* enc_oracle_offsets(const EncOracle* e) { return e->off; }
uint64_t enc_oracle_vocab_size(const EncOracle* e) { return 256 + e->n_merges; }

/* decode (utils/bpe.py:214-222): bytes of every id back to back; returns the byte count (also when it exceeds cap: nothing
 * beyond cap is written), or -1 for an id outside the vocab */
int64_t enc_oracle_decode(const EncOracle* e, const int32_t* ids, uint64_t n, uint8_t* out, uint64_t cap) {
  const int64_t T = 256 + (int64_t)e->n_merges;
  uint64_t o = 0;
  for (uint64_t i = 0; i < n; i++) {
    if (ids[i] < 0 || ids[i] >= T) return -1;
    const uint64_t a = e->tok_off[ids[i]], l = e->tok_off[ids[i] + 1] - a;
    if (o + l <= cap) memcpy(out + o, e->tok_bytes + a, l);
    o += l;
  }
  return (int64_t)o;
}
