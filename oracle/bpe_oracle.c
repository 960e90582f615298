/* oracle/bpe_oracle.c -- TEST INFRASTRUCTURE.  NOT product code, never linked into libtrainer.so.
 *
 * A CPU restatement, in plain C, of what the reference BPE trainer (shivendrra/shredword-trainer,
 * shredword/csrc/bpe/) computes, written from the result contract in SURVEY.md Appendix A.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it, and only as the checker.
 *
 * PARITY PINNED: this file is checked against the UNMODIFIED reference (oracle/_ref/libtrainer_ref.so, built by
 * oracle/Makefile from /root/reference and run under the zero-fill malloc shim oracle/zmalloc.c) on the
 * known-answer corpora of the reference's own tests and on seeded random corpora x configs; the resulting vectors
 * are committed under tests/golden/ (tests/golden/make_golden.py regenerates them).
 *
 * The data structures are ours (flat CSR symbol arrays, open-addressing tables, a pair -> words inverted index so a
 * 1 GB corpus trains in minutes); the *observable* behaviour follows the reference line by line:
 *   tokenisation / fgets+strtok quirks ...... bpe.cpp:131-153
 *   word order (djb2 & 4095, first seen) ..... hash.cpp:29-53,61-72
 *   byte histogram + coverage keep-set ....... histogram.cpp:30-53, bpe.cpp:156-171
 *   symbols with unk substitution ............ histogram.cpp:7-27
 *   bigram count + initial heap fill ......... bpe.cpp:187-230, hash.cpp:7-16,104-130
 *   binary max-heap (freq-only compares) ..... heap.cpp:53-114
 *   merge step, FreqChangeMap order .......... bpe.cpp:9-50,232-323
 *   recompute_freq ("unk => 0") .............. bpe.cpp:52-65
 *   train loop / stop conditions ............. bpe.cpp:345-386
 *   save (vocab text + model int32 triples) .. bpe.cpp:388-432
 */
#define _GNU_SOURCE
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#define ORA_BASE_VOCAB 256      /* bpe.h:20 INITIAL_VOCAB_SIZE */
#define ORA_WORD_BUCKETS 4096   /* bpe.h:21 INITIAL_STR_BUFFER, used as StrMap bucket count (bpe.cpp:116) */
#define ORA_PAIR_BUCKETS 4096   /* bpe.h:19 MIN_HEAP_SIZE, used as BIMap bucket count (bpe.cpp:104,183) */
#define ORA_FC_BUCKETS 1024     /* bpe.cpp:15 FREQ_CHANGE_BUCKETS */

typedef struct { int32_t a, b; uint64_t freq; uint32_t version; } HeapEnt;  /* heap.h:17-21 */

typedef struct {
  int32_t a, b;
  uint64_t freq;
  uint32_t version;
  uint32_t *words; /* inverted index: words that (may) contain the pair; ours, not the reference's */
  uint32_t nw, cw;
} PairEnt;

typedef struct {
  /* config after create_trainer normalisation (bpe.cpp:77-79) */
  uint64_t vocab_size;
  int32_t unk_id;
  float coverage;
  uint64_t min_freq;
  int verify_recompute; /* 1: recompute_freq by a real scan (bpe.cpp:52-65) and abort on mismatch */
  /* corpus */
  int loaded;
  size_t n_words;
  uint64_t *wcount;
  uint64_t *wstart; /* offset of the word's slot in ids[] */
  uint32_t *wlen;   /* live symbols in the slot */
  uint32_t *wblen;  /* original byte length */
  uint8_t *wbytes;  /* concatenated original spellings, in word order */
  uint64_t *wboff;
  int32_t *ids;
  size_t n_ids;
  uint64_t hist[256];
  uint8_t keep[256];
  size_t n_distinct, n_keep;
  /* pair table */
  PairEnt *ent;
  size_t n_ent, c_ent;
  uint32_t *ht; /* open addressing: index+1 into ent, 0 = empty */
  size_t ht_cap;
  /* heap */
  HeapEnt *heap;
  size_t hn, hc;
  /* merges */
  int32_t *merge_ops; /* pairs, capacity vocab_size (bpe.cpp:81) */
  size_t num_merges;
  uint64_t stat_pops, stat_pushes, stat_occ;
} Oracle;

static void *xmalloc(size_t n) { void *p = malloc(n ? n : 1); if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); } return p; }
static void *xcalloc(size_t n, size_t m) { void *p = calloc(n ? n : 1, m ? m : 1); if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); } return p; }
static void *xrealloc(void *q, size_t n) { void *p = realloc(q, n ? n : 1); if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); } return p; }

/* ---------------------------------------------------------------- create / destroy */

void *oracle_create(uint64_t vocab_size, int32_t unk_id, float coverage, uint64_t min_pair_freq) {
  Oracle *o = (Oracle *)xcalloc(1, sizeof(Oracle));
  o->vocab_size = vocab_size;
  o->unk_id = unk_id;
  o->coverage = coverage;
  /* bpe.cpp:78 compares the float against double literals and stores (float)0.995 */
  if ((double)coverage <= 0.0 || (double)coverage >= 1.0) o->coverage = (float)0.995;
  o->min_freq = min_pair_freq ? min_pair_freq : 2000; /* bpe.cpp:79, bpe.h:23 */
  o->merge_ops = (int32_t *)xcalloc(vocab_size ? vocab_size : 1, 2 * sizeof(int32_t));
  return o;
}

static void pairs_free(Oracle *o) {
  for (size_t i = 0; i < o->n_ent; i++) free(o->ent[i].words);
  free(o->ent); free(o->ht);
  o->ent = NULL; o->ht = NULL; o->n_ent = o->c_ent = 0; o->ht_cap = 0;
}

static void corpus_free(Oracle *o) {
  free(o->wcount); free(o->wstart); free(o->wlen); free(o->wblen); free(o->wbytes); free(o->wboff); free(o->ids);
  o->wcount = NULL; o->wstart = NULL; o->wlen = NULL; o->wblen = NULL; o->wbytes = NULL; o->wboff = NULL; o->ids = NULL;
  o->n_words = 0; o->n_ids = 0; o->loaded = 0;
}

void oracle_destroy(void *h) {
  Oracle *o = (Oracle *)h;
  if (!o) return;
  pairs_free(o); corpus_free(o);
  free(o->heap); free(o->merge_ops); free(o);
}

void oracle_set_verify(void *h, int on) { ((Oracle *)h)->verify_recompute = on; }

/* ---------------------------------------------------------------- ingest (bpe.cpp:110-185) */

static inline int is_delim(uint8_t c) { return c == '\t' || c == '\r' || c == '\n' || c == ' '; } /* bpe.cpp:148 */

typedef struct { uint64_t off; uint32_t len; uint32_t bucket; uint64_t count; } WordRec;
typedef struct {
  const uint8_t *text;
  WordRec *rec; size_t n, cap;   /* in first-seen order */
  uint32_t *ht; size_t ht_cap;   /* index+1 */
} WordTab;

static uint64_t fnv64(const uint8_t *p, size_t n) {
  uint64_t h = 1469598103934665603ULL;
  for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ULL; }
  return h;
}

static void wt_grow(WordTab *t) {
  size_t nc = t->ht_cap ? t->ht_cap * 2 : (1u << 16);
  uint32_t *nh = (uint32_t *)xcalloc(nc, sizeof(uint32_t));
  for (size_t i = 0; i < t->n; i++) {
    size_t s = fnv64(t->text + t->rec[i].off, t->rec[i].len) & (nc - 1);
    while (nh[s]) s = (s + 1) & (nc - 1);
    nh[s] = (uint32_t)(i + 1);
  }
  free(t->ht); t->ht = nh; t->ht_cap = nc;
}

/* strmap_increment (hash.cpp:29-53): count the token, remember first-seen order and its djb2 bucket */
static void wt_add(WordTab *t, uint64_t off, uint32_t len) {
  if ((t->n + 1) * 2 > t->ht_cap) wt_grow(t);
  const uint8_t *p = t->text + off;
  size_t s = fnv64(p, len) & (t->ht_cap - 1);
  while (t->ht[s]) {
    WordRec *r = &t->rec[t->ht[s] - 1];
    if (r->len == len && memcmp(t->text + r->off, p, len) == 0) { r->count++; return; }
    s = (s + 1) & (t->ht_cap - 1);
  }
  if (t->n == t->cap) { t->cap = t->cap ? t->cap * 2 : 4096; t->rec = (WordRec *)xrealloc(t->rec, t->cap * sizeof(WordRec)); }
  uint64_t dj = 5381; /* hash.cpp:35-38: size_t h = 5381; h = h*33 + byte */
  for (uint32_t i = 0; i < len; i++) dj = ((dj << 5) + dj) + p[i];
  WordRec *r = &t->rec[t->n];
  r->off = off; r->len = len; r->count = 1; r->bucket = (uint32_t)(dj & (ORA_WORD_BUCKETS - 1));
  t->ht[s] = (uint32_t)(++t->n);
}

/* strtok(line, "\t\r\n ") over text[b, e) (bpe.cpp:148-152) */
static void tokenize_span(WordTab *t, const uint8_t *text, size_t b, size_t e) {
  size_t i = b;
  while (i < e) {
    while (i < e && is_delim(text[i])) i++;
    size_t s = i;
    while (i < e && !is_delim(text[i])) i++;
    if (i > s) wt_add(t, s, (uint32_t)(i - s));
  }
}

/* The fgets/strlen/realloc line loop of bpe.cpp:130-147, replayed over an in-memory image of the file.  For NUL-free
 * input the loop is plain whitespace splitting (tokenize_span over the whole buffer); a NUL byte hides the rest of
 * its fgets chunk because strlen stops there.  This variant computes, for every fgets "line", the visible span
 * [start, start+len) in the original text. */
static void ingest_lines_nul(WordTab *t, const uint8_t *text, size_t n) {
  size_t cap = 4096, pos = 0;
  while (pos < n) {
    size_t start = pos, raw = 0;
    /* first fgets: up to cap-1 bytes, stops after '\n' */
    while (raw < cap - 1 && pos < n) { uint8_t c = text[pos++]; raw++; if (c == '\n') break; }
    const uint8_t *z = (const uint8_t *)memchr(text + start, 0, raw);
    size_t len = z ? (size_t)(z - (text + start)) : raw;
    while (len == cap - 1 && text[start + len - 1] != '\n') {
      cap *= 2;
      if (pos >= n) break;
      size_t room = cap - len, got = 0;
      while (got < room - 1 && pos < n) { uint8_t c = text[pos++]; got++; if (c == '\n') break; }
      /* strlen restarts from the line start: the earlier part had no NUL (len == raw so far) */
      const uint8_t *z2 = (const uint8_t *)memchr(text + start + len, 0, got);
      len = z2 ? (size_t)(z2 - (text + start)) : len + got;
    }
    tokenize_span(t, text, start, start + len); /* a trailing '\n' is a delimiter anyway */
  }
}

static int cmp_hist(const void *x, const void *y) { /* histogram.cpp:47-53, made total by the iteration rank */
  const uint64_t *a = (const uint64_t *)x, *b = (const uint64_t *)y; /* {count, rank, byte} */
  if (b[0] > a[0]) return 1;
  if (b[0] < a[0]) return -1;
  return a[1] < b[1] ? -1 : (a[1] > b[1]);
}

int oracle_load_buffer(void *h, const uint8_t *text, size_t n) {
  Oracle *o = (Oracle *)h;
  if (!o) return -1;
  WordTab t; memset(&t, 0, sizeof t); t.text = text;
  if (n) { if (memchr(text, 0, n)) ingest_lines_nul(&t, text, n); else tokenize_span(&t, text, 0, n); }
  /* a second load replaces the first (bpe.cpp:176-178,183) */
  corpus_free(o);
  /* word order: StrMap iteration = bucket ascending, chain order = first seen (hash.cpp:45-52,61-72) */
  size_t N = t.n;
  size_t *bstart = (size_t *)xcalloc(ORA_WORD_BUCKETS + 1, sizeof(size_t));
  for (size_t i = 0; i < N; i++) bstart[t.rec[i].bucket + 1]++;
  for (int b = 0; b < ORA_WORD_BUCKETS; b++) bstart[b + 1] += bstart[b];
  uint32_t *order = (uint32_t *)xmalloc(N * sizeof(uint32_t));
  for (size_t i = 0; i < N; i++) order[bstart[t.rec[i].bucket]++] = (uint32_t)i;
  free(bstart);
  /* char_hist (histogram.cpp:30-36): every byte of every UNIQUE word counts once, not weighted by the word count */
  memset(o->hist, 0, sizeof o->hist);
  size_t S = 0;
  for (size_t i = 0; i < N; i++) { const uint8_t *p = text + t.rec[i].off; for (uint32_t j = 0; j < t.rec[i].len; j++) o->hist[p[j]]++; S += t.rec[i].len; }
  /* collect_char iterates a 256-bucket StrMap of 1-byte keys: bucket = (5381*33 + b) & 255 = (b + 165) & 255 */
  uint64_t cc[256][3]; size_t c = 0;
  for (int b = 0; b < 256; b++) if (o->hist[b]) { cc[c][0] = o->hist[b]; cc[c][1] = (uint64_t)((b + 165) & 255); cc[c][2] = (uint64_t)b; c++; }
  qsort(cc, c, sizeof cc[0], cmp_hist); /* stable order made explicit through the rank (glibc qsort = stable mergesort) */
  volatile float prod = (float)c * o->coverage; /* bpe.cpp:169: size_t * float in float32 */
  size_t keep = (size_t)prod;
  memset(o->keep, 0, sizeof o->keep);
  for (size_t i = 0; i < keep && i < c; i++) o->keep[cc[i][2]] = 1; /* bpe.cpp:170-171 */
  o->n_distinct = c; o->n_keep = keep;
  /* build_symbol_cb (histogram.cpp:7-27) */
  o->n_words = N;
  o->wcount = (uint64_t *)xmalloc(N * sizeof(uint64_t));
  o->wstart = (uint64_t *)xmalloc(N * sizeof(uint64_t));
  o->wlen = (uint32_t *)xmalloc(N * sizeof(uint32_t));
  o->wblen = (uint32_t *)xmalloc(N * sizeof(uint32_t));
  o->wboff = (uint64_t *)xmalloc(N * sizeof(uint64_t));
  o->wbytes = (uint8_t *)xmalloc(S);
  o->ids = (int32_t *)xmalloc(S * sizeof(int32_t));
  o->n_ids = S;
  size_t at = 0;
  for (size_t wi = 0; wi < N; wi++) {
    WordRec *r = &t.rec[order[wi]];
    o->wcount[wi] = r->count; o->wstart[wi] = at; o->wlen[wi] = r->len; o->wblen[wi] = r->len; o->wboff[wi] = at;
    const uint8_t *p = text + r->off;
    memcpy(o->wbytes + at, p, r->len);
    for (uint32_t j = 0; j < r->len; j++) o->ids[at + j] = o->keep[p[j]] ? (int32_t)p[j] : o->unk_id;
    at += r->len;
  }
  free(order); free(t.rec); free(t.ht);
  pairs_free(o); /* bimap_init at bpe.cpp:183 */
  o->loaded = 1;
  return 0;
}

int oracle_load(void *h, const char *path) {
  if (!h || !path) return -1; /* bpe.cpp:111-114 */
  int fd = open(path, O_RDONLY);
  if (fd < 0) return -1; /* bpe.cpp:118-122 */
  struct stat st;
  if (fstat(fd, &st) != 0) { close(fd); return -1; }
  size_t n = (size_t)st.st_size;
  int rc;
  if (n == 0) rc = oracle_load_buffer(h, (const uint8_t *)"", 0);
  else { /* map instead of read: a 50 GB corpus then costs page cache, not heap */
    void *p = mmap(NULL, n, PROT_READ, MAP_PRIVATE, fd, 0);
    if (p == MAP_FAILED) { close(fd); return -1; }
    rc = oracle_load_buffer(h, (const uint8_t *)p, n);
    munmap(p, n);
  }
  close(fd);
  return rc;
}

/* ---------------------------------------------------------------- pair table (hash.cpp:104-130) */

static inline uint64_t pk64(int32_t a, int32_t b) { return ((uint64_t)(uint32_t)a << 32) | (uint32_t)b; }
static inline uint64_t mix64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

static uint32_t fnv1a_pair(int32_t a, int32_t b) { /* hash.cpp:7-16 over the 8 little-endian bytes of {first, second} */
  uint32_t h = 2166136261u, w[2] = {(uint32_t)a, (uint32_t)b};
  for (int k = 0; k < 2; k++) for (int i = 0; i < 4; i++) { h ^= (w[k] >> (8 * i)) & 255u; h *= 16777619u; }
  return h;
}

static void pt_grow(Oracle *o) {
  size_t nc = o->ht_cap ? o->ht_cap * 2 : (1u << 12);
  uint32_t *nh = (uint32_t *)xcalloc(nc, sizeof(uint32_t));
  for (size_t i = 0; i < o->n_ent; i++) {
    size_t s = mix64(pk64(o->ent[i].a, o->ent[i].b)) & (nc - 1);
    while (nh[s]) s = (s + 1) & (nc - 1);
    nh[s] = (uint32_t)(i + 1);
  }
  free(o->ht); o->ht = nh; o->ht_cap = nc;
}

/* bimap_get: get-or-create, new entries zeroed (hash.cpp:122) and appended in creation order */
static PairEnt *pt_get(Oracle *o, int32_t a, int32_t b) {
  if ((o->n_ent + 1) * 2 > o->ht_cap) pt_grow(o);
  size_t s = mix64(pk64(a, b)) & (o->ht_cap - 1);
  while (o->ht[s]) {
    PairEnt *e = &o->ent[o->ht[s] - 1];
    if (e->a == a && e->b == b) return e;
    s = (s + 1) & (o->ht_cap - 1);
  }
  if (o->n_ent == o->c_ent) { o->c_ent = o->c_ent ? o->c_ent * 2 : 1024; o->ent = (PairEnt *)xrealloc(o->ent, o->c_ent * sizeof(PairEnt)); }
  PairEnt *e = &o->ent[o->n_ent];
  memset(e, 0, sizeof *e);
  e->a = a; e->b = b;
  o->ht[s] = (uint32_t)(++o->n_ent);
  return e;
}

static void pe_add_word(PairEnt *e, uint32_t wi) {
  if (e->nw && e->words[e->nw - 1] == wi) return;
  if (e->nw == e->cw) { e->cw = e->cw ? e->cw * 2 : 4; e->words = (uint32_t *)xrealloc(e->words, e->cw * sizeof(uint32_t)); }
  e->words[e->nw++] = wi;
}

/* ---------------------------------------------------------------- heap (heap.cpp:53-114) */

static void heap_push(Oracle *o, int32_t a, int32_t b, uint64_t freq, uint32_t version) {
  if (o->hn == o->hc) { o->hc = o->hc ? o->hc * 2 : 4096; o->heap = (HeapEnt *)xrealloc(o->heap, o->hc * sizeof(HeapEnt)); }
  size_t i = o->hn++;
  HeapEnt x = {a, b, freq, version};
  o->heap[i] = x;
  while (i > 0) { /* heap.cpp:74-79: stop when parent.freq >= child.freq */
    size_t p = (i - 1) >> 1;
    if (o->heap[p].freq >= o->heap[i].freq) break;
    HeapEnt tmp = o->heap[p]; o->heap[p] = o->heap[i]; o->heap[i] = tmp;
    i = p;
  }
  o->stat_pushes++;
}

static HeapEnt heap_pop(Oracle *o) {
  HeapEnt top = o->heap[0];
  o->heap[0] = o->heap[--o->hn]; /* heap.cpp:98 */
  size_t i = 0;
  for (;;) { /* heap.cpp:101-111: strict > on both children, left first */
    size_t l = 2 * i + 1, r = l + 1, best = i;
    if (l < o->hn && o->heap[l].freq > o->heap[best].freq) best = l;
    if (r < o->hn && o->heap[r].freq > o->heap[best].freq) best = r;
    if (best == i) break;
    HeapEnt tmp = o->heap[i]; o->heap[i] = o->heap[best]; o->heap[best] = tmp;
    i = best;
  }
  o->stat_pops++;
  return top;
}

/* ---------------------------------------------------------------- count (bpe.cpp:187-230) */

typedef struct { uint32_t bucket; uint32_t idx; } OrdRec;
static int cmp_ord(const void *x, const void *y) {
  const OrdRec *a = (const OrdRec *)x, *b = (const OrdRec *)y;
  if (a->bucket != b->bucket) return a->bucket < b->bucket ? -1 : 1;
  return a->idx < b->idx ? -1 : (a->idx > b->idx);
}

void oracle_count_bigrams(void *h) {
  Oracle *o = (Oracle *)h;
  for (size_t wi = 0; wi < o->n_words; wi++) {
    const int32_t *s = o->ids + o->wstart[wi];
    uint64_t c = o->wcount[wi];
    for (uint32_t j = 0; j + 1 < o->wlen[wi]; j++) {
      if (s[j] == o->unk_id || s[j + 1] == o->unk_id) continue; /* bpe.cpp:201 */
      PairEnt *e = pt_get(o, s[j], s[j + 1]);
      if (e->freq == 0) e->version = 0; /* bpe.cpp:207-210 */
      e->freq += c;
      pe_add_word(e, (uint32_t)wi);
    }
  }
  /* push every entry with freq >= min in BIMap iteration order: bucket = fnv1a & 4095 ascending, chain = creation
   * order (bpe.cpp:219-227, hash.cpp:126-129) */
  OrdRec *ord = (OrdRec *)xmalloc(o->n_ent * sizeof(OrdRec));
  for (size_t i = 0; i < o->n_ent; i++) { ord[i].bucket = fnv1a_pair(o->ent[i].a, o->ent[i].b) & (ORA_PAIR_BUCKETS - 1); ord[i].idx = (uint32_t)i; }
  qsort(ord, o->n_ent, sizeof(OrdRec), cmp_ord);
  for (size_t i = 0; i < o->n_ent; i++) {
    PairEnt *e = &o->ent[ord[i].idx];
    if (e->freq >= o->min_freq) heap_push(o, e->a, e->b, e->freq, e->version);
  }
  free(ord);
}

void oracle_init(void *h) { /* bpe_init, bpe.cpp:98-108 */
  Oracle *o = (Oracle *)h;
  pairs_free(o);
  o->hn = 0;
  oracle_count_bigrams(o);
}

/* ---------------------------------------------------------------- merge (bpe.cpp:232-323) */

typedef struct { uint64_t h; int64_t delta; uint32_t ins; } FcRec;
typedef struct {
  FcRec *rec; size_t n, cap;
  uint32_t *ht; size_t ht_cap; /* index+1 */
  uint32_t *touched; size_t nt, ct;
} FcMap;

static void fc_add(FcMap *m, uint64_t hkey, int64_t delta) { /* freq_change_add, bpe.cpp:25-38 */
  if ((m->n + 1) * 2 > m->ht_cap) {
    size_t nc = m->ht_cap ? m->ht_cap * 2 : 1024;
    uint32_t *nh = (uint32_t *)xcalloc(nc, sizeof(uint32_t));
    for (size_t i = 0; i < m->n; i++) { size_t s = mix64(m->rec[i].h) & (nc - 1); while (nh[s]) s = (s + 1) & (nc - 1); nh[s] = (uint32_t)(i + 1); }
    free(m->ht); m->ht = nh; m->ht_cap = nc;
  }
  size_t s = mix64(hkey) & (m->ht_cap - 1);
  while (m->ht[s]) {
    FcRec *r = &m->rec[m->ht[s] - 1];
    if (r->h == hkey) { r->delta += delta; return; }
    s = (s + 1) & (m->ht_cap - 1);
  }
  if (m->n == m->cap) { m->cap = m->cap ? m->cap * 2 : 1024; m->rec = (FcRec *)xrealloc(m->rec, m->cap * sizeof(FcRec)); }
  m->rec[m->n].h = hkey; m->rec[m->n].delta = delta; m->rec[m->n].ins = (uint32_t)m->n;
  m->ht[s] = (uint32_t)(++m->n);
}

static int cmp_fc(const void *x, const void *y) { /* bucket ascending; inside a bucket the chain is LIFO (prepend, bpe.cpp:36-37) */
  const FcRec *a = (const FcRec *)x, *b = (const FcRec *)y;
  uint64_t ba = a->h % ORA_FC_BUCKETS, bb = b->h % ORA_FC_BUCKETS;
  if (ba != bb) return ba < bb ? -1 : 1;
  return a->ins > b->ins ? -1 : (a->ins < b->ins);
}

static int cmp_u32(const void *x, const void *y) { uint32_t a = *(const uint32_t *)x, b = *(const uint32_t *)y; return a < b ? -1 : (a > b); }

static inline uint64_t fc_key(int32_t first, int32_t second) { /* bpe.cpp:277-278,286-287: both operands sign-extend */
  return ((uint64_t)first << 32) | (uint64_t)second;
}

static uint64_t recompute_scan(Oracle *o, int32_t a, int32_t b) { /* bpe.cpp:52-65, the literal scan */
  if (a == o->unk_id || b == o->unk_id) return 0;
  uint64_t f = 0;
  for (size_t wi = 0; wi < o->n_words; wi++) {
    const int32_t *s = o->ids + o->wstart[wi];
    for (uint32_t j = 0; j + 1 < o->wlen[wi]; j++) if (s[j] == a && s[j + 1] == b) f += o->wcount[wi];
  }
  return f;
}

typedef struct { uint64_t h; uint32_t wi; } AdjRec;

int oracle_merge_batch(void *h, int batch_size) {
  Oracle *o = (Oracle *)h;
  if (!o) return -1;
  if (o->hn == 0) return 0; /* bpe.cpp:237-240 */
  int merges_done = 0;
  const uint64_t min_freq = o->min_freq;
  FcMap fc; memset(&fc, 0, sizeof fc);
  AdjRec *adj = NULL; size_t nadj = 0, cadj = 0;
  uint32_t *cand = NULL; size_t ccand = 0;
  while (merges_done < batch_size && o->hn > 0) {
    HeapEnt top = heap_pop(o);
    PairEnt *info = pt_get(o, top.a, top.b);
    if (top.version != info->version) continue; /* stale, bpe.cpp:247-250 */
    uint64_t actual;
    if (top.a == o->unk_id || top.b == o->unk_id) actual = 0; /* bpe.cpp:53 */
    else if (o->verify_recompute) {
      actual = recompute_scan(o, top.a, top.b);
      if (actual != info->freq) { fprintf(stderr, "oracle: recompute mismatch (%d,%d) scan=%llu table=%llu\n", top.a, top.b, (unsigned long long)actual, (unsigned long long)info->freq); abort(); }
    } else actual = info->freq; /* the scan re-derives what the deltas already maintain */
    if (actual != info->freq) { /* bpe.cpp:252-257 */
      info->freq = actual; info->version++;
      if (actual >= min_freq) heap_push(o, top.a, top.b, actual, info->version);
      continue;
    }
    if (actual < min_freq) continue; /* bpe.cpp:258 */
    const int32_t A = top.a, B = top.b, N = (int32_t)(ORA_BASE_VOCAB + o->num_merges); /* bpe.cpp:259 */
    if (o->num_merges < o->vocab_size) { o->merge_ops[2 * o->num_merges] = A; o->merge_ops[2 * o->num_merges + 1] = B; } /* bpe.cpp:261 */
    /* candidate words in ascending word order (the reference scans all words in order, bpe.cpp:265) */
    size_t ncand = info->nw;
    if (ncand > ccand) { ccand = ncand * 2; cand = (uint32_t *)xrealloc(cand, ccand * sizeof(uint32_t)); }
    memcpy(cand, info->words, ncand * sizeof(uint32_t));
    qsort(cand, ncand, sizeof(uint32_t), cmp_u32);
    fc.n = 0; if (fc.ht_cap) memset(fc.ht, 0, fc.ht_cap * sizeof(uint32_t));
    nadj = 0;
    uint32_t last = UINT32_MAX;
    for (size_t ci = 0; ci < ncand; ci++) {
      uint32_t wi = cand[ci];
      if (wi == last) continue;
      last = wi;
      int32_t *s = o->ids + o->wstart[wi];
      uint32_t len = o->wlen[wi], r = 0, w = 0;
      int64_t c = (int64_t)o->wcount[wi];
      while (r < len) {
        if (r + 1 < len && s[r] == A && s[r + 1] == B) {
          o->stat_occ++;
          if (w > 0) { /* left neighbour = current id there (N if just merged), bpe.cpp:274-281 */
            int32_t L = s[w - 1];
            fc_add(&fc, fc_key(L, A), -c);
            uint64_t hk = fc_key(L, N);
            fc_add(&fc, hk, c);
            if (nadj == cadj) { cadj = cadj ? cadj * 2 : 1024; adj = (AdjRec *)xrealloc(adj, cadj * sizeof(AdjRec)); }
            adj[nadj].h = hk; adj[nadj++].wi = wi;
          }
          if (r + 2 < len) { /* right neighbour = raw next-next id, bpe.cpp:282-290 */
            int32_t R = s[r + 2];
            fc_add(&fc, fc_key(B, R), -c);
            uint64_t hk = fc_key(N, R);
            fc_add(&fc, hk, c);
            if (nadj == cadj) { cadj = cadj ? cadj * 2 : 1024; adj = (AdjRec *)xrealloc(adj, cadj * sizeof(AdjRec)); }
            adj[nadj].h = hk; adj[nadj++].wi = wi;
          }
          s[w++] = N; r += 2; /* bpe.cpp:291-294 */
        } else { s[w++] = s[r++]; }
      }
      o->wlen[wi] = w;
    }
    /* apply in FreqChangeMap order (bpe.cpp:297-313) */
    qsort(fc.rec, fc.n, sizeof(FcRec), cmp_fc);
    for (size_t i = 0; i < fc.n; i++) {
      int32_t pa = (int32_t)(fc.rec[i].h >> 32), pb = (int32_t)(fc.rec[i].h & 0xFFFFFFFFu); /* bpe.cpp:301 */
      if (pa == A && pb == B) continue; /* bpe.cpp:302 */
      PairEnt *e = pt_get(o, pa, pb);
      int64_t d = fc.rec[i].delta;
      if (d < 0) { uint64_t ad = (uint64_t)(-d); e->freq = e->freq >= ad ? e->freq - ad : 0; } else e->freq += (uint64_t)d;
      if (e->freq >= min_freq) { e->version++; heap_push(o, pa, pb, e->freq, e->version); }
    }
    for (size_t i = 0; i < nadj; i++) { /* inverted-index upkeep (ours) */
      int32_t pa = (int32_t)(adj[i].h >> 32), pb = (int32_t)(adj[i].h & 0xFFFFFFFFu);
      pe_add_word(pt_get(o, pa, pb), adj[i].wi);
    }
    info = pt_get(o, A, B); /* the entry array may have moved */
    info->freq = 0; info->version++; /* bpe.cpp:315-316 */
    o->num_merges++; merges_done++;
  }
  free(fc.rec); free(fc.ht); free(adj); free(cand);
  return merges_done;
}

int oracle_train(void *h) { /* bpe_train, bpe.cpp:345-386 */
  Oracle *o = (Oracle *)h;
  if (!o) return -1;
  oracle_init(o);
  int total = 0, target = (int)o->vocab_size - ORA_BASE_VOCAB;
  while (total < target) {
    if (o->hn == 0) break;
    uint64_t top_freq = o->heap[0].freq;
    int batch = top_freq > 50000 ? 10 : top_freq > 20000 ? 5 : top_freq > 10000 ? 3 : top_freq > 5000 ? 2 : 1; /* bpe.cpp:362-367 */
    if (batch > target - total) batch = target - total;
    int merged = oracle_merge_batch(o, batch);
    if (merged <= 0) break;
    total += merged;
  }
  return total;
}

/* ---------------------------------------------------------------- save (bpe.cpp:388-432) */

void oracle_save(void *h, const char *model_path, const char *vocab_path) {
  Oracle *o = (Oracle *)h;
  size_t M = o->num_merges, T = ORA_BASE_VOCAB + M;
  char **tok = (char **)xcalloc(T, sizeof(char *));
  for (size_t i = 0; i < ORA_BASE_VOCAB; i++) { tok[i] = (char *)xmalloc(2); tok[i][0] = (char)i; tok[i][1] = 0; }
  for (size_t m = 0; m < M; m++) {
    const char *a = tok[o->merge_ops[2 * m]], *b = tok[o->merge_ops[2 * m + 1]];
    size_t la = strlen(a), lb = strlen(b);
    tok[ORA_BASE_VOCAB + m] = (char *)xmalloc(la + lb + 1);
    memcpy(tok[ORA_BASE_VOCAB + m], a, la); memcpy(tok[ORA_BASE_VOCAB + m] + la, b, lb + 1);
  }
  uint64_t *freq = (uint64_t *)xcalloc(T, sizeof(uint64_t));
  for (size_t wi = 0; wi < o->n_words; wi++) {
    const int32_t *s = o->ids + o->wstart[wi];
    for (uint32_t j = 0; j < o->wlen[wi]; j++) {
      /* the reference indexes freq[id] unchecked (bpe.cpp:413); ids outside [0,T) (an out-of-range unk_id) are
       * undefined behaviour there and do not reach the file, so they are skipped here */
      if (s[j] >= 0 && (size_t)s[j] < T) freq[s[j]] += o->wcount[wi];
    }
  }
  FILE *vf = fopen(vocab_path, "w");
  if (vf) { for (size_t i = 0; i < T; i++) fprintf(vf, "%s %llu\n", tok[i], (unsigned long long)freq[i]); fclose(vf); }
  FILE *mf = fopen(model_path, "wb");
  if (mf) {
    for (size_t m = 0; m < M; m++) { int32_t rec[3] = {o->merge_ops[2 * m], o->merge_ops[2 * m + 1], (int32_t)(ORA_BASE_VOCAB + m)}; fwrite(rec, sizeof(int32_t), 3, mf); }
    fclose(mf);
  }
  for (size_t i = 0; i < T; i++) free(tok[i]);
  free(tok); free(freq);
}

/* ---------------------------------------------------------------- introspection for staged parity tests */

uint64_t oracle_num_words(void *h) { return ((Oracle *)h)->n_words; }
uint64_t oracle_num_merges(void *h) { return ((Oracle *)h)->num_merges; }
uint64_t oracle_heap_size(void *h) { return ((Oracle *)h)->hn; }
uint64_t oracle_pair_entries(void *h) { return ((Oracle *)h)->n_ent; }
uint64_t oracle_word_count(void *h, uint64_t wi) { return ((Oracle *)h)->wcount[wi]; }
uint64_t oracle_min_freq(void *h) { return ((Oracle *)h)->min_freq; }
float oracle_coverage(void *h) { return ((Oracle *)h)->coverage; }
void oracle_stats(void *h, uint64_t *out) { Oracle *o = (Oracle *)h; out[0] = o->stat_pops; out[1] = o->stat_pushes; out[2] = o->stat_occ; }

uint64_t oracle_num_symbols(void *h) { /* live symbols */
  Oracle *o = (Oracle *)h; uint64_t s = 0;
  for (size_t i = 0; i < o->n_words; i++) s += o->wlen[i];
  return s;
}
uint32_t oracle_word_bytes(void *h, uint64_t wi, uint8_t *out, uint32_t cap) {
  Oracle *o = (Oracle *)h; uint32_t n = o->wblen[wi];
  if (out) memcpy(out, o->wbytes + o->wboff[wi], n < cap ? n : cap);
  return n;
}
uint32_t oracle_word_ids(void *h, uint64_t wi, int32_t *out, uint32_t cap) {
  Oracle *o = (Oracle *)h; uint32_t n = o->wlen[wi];
  if (out) memcpy(out, o->ids + o->wstart[wi], (size_t)(n < cap ? n : cap) * sizeof(int32_t));
  return n;
}
void oracle_keep_mask(void *h, uint8_t *out256, uint64_t *hist256) {
  Oracle *o = (Oracle *)h;
  if (out256) memcpy(out256, o->keep, 256);
  if (hist256) memcpy(hist256, o->hist, sizeof o->hist);
}
void oracle_get_merges(void *h, int32_t *out) { /* (a, b, new_id) triples */
  Oracle *o = (Oracle *)h;
  for (size_t m = 0; m < o->num_merges && m < o->vocab_size; m++) { out[3 * m] = o->merge_ops[2 * m]; out[3 * m + 1] = o->merge_ops[2 * m + 1]; out[3 * m + 2] = (int32_t)(ORA_BASE_VOCAB + m); }
}
/* dump the pair table as (a, b, freq) in creation order; returns the number of entries */
uint64_t oracle_get_pairs(void *h, int32_t *ab, uint64_t *freq, uint64_t cap) {
  Oracle *o = (Oracle *)h;
  for (size_t i = 0; i < o->n_ent && i < cap; i++) { ab[2 * i] = o->ent[i].a; ab[2 * i + 1] = o->ent[i].b; freq[i] = o->ent[i].freq; }
  return o->n_ent;
}
/* dump the heap array (a, b, version) + freq in array order */
uint64_t oracle_get_heap(void *h, int32_t *abv, uint64_t *freq, uint64_t cap) {
  Oracle *o = (Oracle *)h;
  for (size_t i = 0; i < o->hn && i < cap; i++) { abv[3 * i] = o->heap[i].a; abv[3 * i + 1] = o->heap[i].b; abv[3 * i + 2] = (int32_t)o->heap[i].version; freq[i] = o->heap[i].freq; }
  return o->hn;
}
