// common.cuh -- macros, constants, shared structs and the two open-addressing tables' device helpers.
// Fragment of engine_cuda.cu: included inside `namespace shred { namespace {`, in the order listed there.
#pragma once

#define CK(call)                                                                                              \
  do {                                                                                                        \
    cudaError_t e_ = (call);                                                                                  \
    if (e_ != cudaSuccess) {                                                                                  \
      std::fprintf(stderr, "[ERROR]\t CUDA: %s -> %s (%s:%d)\n", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return -1;                                                                                              \
    }                                                                                                         \
  } while (0)

#define RC(call)            \
  do {                      \
    int rc_ = (call);       \
    if (rc_ != 0) return rc_; \
  } while (0)

typedef unsigned long long ull;

using lay::DEAD;
using lay::UNK_CODE_NEG;
using lay::Params;
using lay::code_to_id;
using lay::fc_key;
constexpr uint32_t HDR_BIT = lay::HDR_TAG;
constexpr uint32_t NONE32 = lay::NONE32;
constexpr uint64_t PT_EMPTY = ~0ull;
constexpr uint64_t SEQ_MAX = ~0ull;
constexpr uint64_t NO_LIST = ~0ull;
constexpr int N_SM_FALLBACK = 148;

enum : uint32_t { ERR_DT_FULL = 1, ERR_PT_FULL = 2, ERR_WT_FULL = 4, ERR_WT_COLLISION = 8, ERR_REC_FULL = 16, ERR_HAS_NUL = 32, ERR_BARRIER = 64,
                  ERR_PEER_TIMEOUT = 128, ERR_INBOX_FULL = 256, ERR_POOL_FULL = 512, ERR_SCRATCH_FULL = 1024, ERR_BAD_LIST = 2048 };

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
__device__ __forceinline__ bool is_delim(uint32_t c) { return c <= 32u && ((0x100002600ull >> c) & 1ull); }  // \t \n \r space

// ---- device -> host wire formats (mapped pinned host memory) -------------------------------------------------------
// The device never fences towards the host.  Everything it publishes is made of 16-byte blocks that carry the pass's 32-bit
// tag themselves (written two at a time with one 256-bit store = one PCIe write), and the host accepts a block only when it sees the current tag in it: no
// ordering between different stores, threads or CTAs is assumed, so no __threadfence_system() sits on the per-merge path.
//   WireRec  w0 = key | w1 = val (48 bits) + tag low half << 48 || w2 = seq (48 bits) + tag high half << 48 | w3 = kind/list length + serial << 32
//   Ctrl     block 0: tag + n_recs << 32 | err + list_len << 32     block 1: tag + occ_local << 32 | occ_global
//            block 2: tag + n_keys << 32 | pt_n                       block 3: tag | pool_top
struct WireRec { ull w[4]; };
struct Ctrl { ull w[8]; };
constexpr ull MASK48 = (1ull << 48) - 1ull;
__device__ __forceinline__ void st_wire(void* p, ull x, ull y) { asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(x), "l"(y) : "memory"); }
// 32 bytes with one 256-bit store (sm_100): one PCIe write per record instead of two -- the host still checks both halves
__device__ __forceinline__ void st_wire32(void* p, ull x, ull y, ull z, ull w) {
  asm volatile("st.volatile.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(x), "l"(y), "l"(z), "l"(w) : "memory");
}
__device__ __forceinline__ void wire_rec(WireRec* r, uint32_t tag, uint64_t key, uint64_t val, uint64_t seq, uint32_t kind, uint32_t serial) {
  st_wire32(&r->w[0], key, (val & MASK48) | (static_cast<ull>(tag & 0xFFFFu) << 48), (seq & MASK48) | (static_cast<ull>(tag >> 16) << 48),
            static_cast<ull>(kind) | (static_cast<ull>(serial) << 32));
}
__device__ __forceinline__ void wire_ctrl(Ctrl* c, uint32_t tag, uint32_t n_recs, uint32_t err, uint32_t list_len, uint32_t occ_local, ull occ_global, uint32_t n_keys, ull pt_n,
                                          ull pool_top) {
  st_wire32(&c->w[4], tag | (static_cast<ull>(n_keys) << 32), pt_n, tag, pool_top);
  st_wire32(&c->w[0], tag | (static_cast<ull>(n_recs) << 32), err | (static_cast<ull>(list_len) << 32), tag | (static_cast<ull>(occ_local) << 32), occ_global);
}

// Device counters.  The per-pass counters exist twice: pass number e uses [e & 1] and zeroes [(e + 1) & 1] for its successor
// (the successor is a later launch on the same stream), so no pass ends with a "last CTA re-arms" step.
struct DevCounters {
  uint32_t n_occ[2], dt_n[2], rec_n[2];
  uint32_t bar, err;
  ull pt_n;
  ull pool_top;
  ull n_tokens;
  uint32_t n_unique, sent_ctas;
  ull sent_epoch;  // multi-GPU: number of the last exchange all of whose local CTAs have sent (and so have read their list length)
};

struct DeltaTable {
  uint64_t* keys; ull* delta; ull* seq;
  uint32_t* nocc;  // occurrences that will enter the key's list (unweighted); doubles as the rank dispenser
  ull* base;       // start of the key's list in the pool, written by the fold (NO_LIST: the key gets none)
  uint32_t* list; uint64_t* klist;  // slot and key of every used slot, dense
  uint32_t* n;     // -> DevCounters::dt_n[parity] of the running pass
  uint64_t mask; uint64_t empty; uint32_t cap;
};
// One 32-byte sector per pair: key + frequency (one 16-byte load) and, beside them, the serial (a second load of the same sector):
// a fold touches exactly one DRAM sector per existing pair.
struct PairEnt { uint64_t key; uint64_t freq; uint32_t serial; uint32_t pad0; uint64_t pad1; };
struct PoolEnt { uint32_t pos; uint32_t cnt; };  // occurrence-list entry: slot of the pair's first token + its word's count (CNT_SAT: look it up)
constexpr uint32_t CNT_SAT = 0xFFFFFFFFu;
// Occurrence list of a pair: positions (slot of the pair's first token) in pool[off, off + len), created by ONE pass -- the
// count pass, or the merge that created the younger of the two tokens -- and only ever validated lazily afterwards.
struct ListRef { ull off; uint32_t len; uint32_t fill; };
struct PairTable {
  PairEnt* ent;      // .serial: dense id per entry = number of entries that existed when it was created (the host indexes by it)
  ListRef* lists;    // indexed by serial (survives rehashing)
  uint64_t mask; uint64_t cap; uint64_t lists_cap;
};

// ------------------------------------------------------------------------------------------------ hash-table helpers

// Adds (delta, seq) to `key`, claiming a slot if the key is new; returns the slot (NONE32 if the table is full).
__device__ __forceinline__ uint32_t dt_add(const DeltaTable& dt, DevCounters* ctr, uint64_t key, int64_t delta, uint64_t seq) {
  uint64_t slot = mix64(key) & dt.mask;
  for (uint32_t probe = 0; probe < dt.cap; ++probe) {
    uint64_t cur = dt.keys[slot];
    if (cur == dt.empty) {
      uint64_t prev = atomicCAS(reinterpret_cast<ull*>(&dt.keys[slot]), static_cast<ull>(dt.empty), static_cast<ull>(key));
      if (prev == dt.empty) {
        uint32_t idx = atomicAdd(dt.n, 1u);
        if (idx < dt.cap) { dt.list[idx] = static_cast<uint32_t>(slot); dt.klist[idx] = key; }
        cur = key;
      } else cur = prev;
    }
    if (cur == key) {
      atomicAdd(&dt.delta[slot], static_cast<ull>(delta));
      atomicMin(&dt.seq[slot], static_cast<ull>(seq));
      return static_cast<uint32_t>(slot);
    }
    slot = (slot + 1) & dt.mask;
  }
  atomicOr(&ctr->err, ERR_DT_FULL);
  return NONE32;
}

// The four delta-table updates of one occurrence.  Each update alone is a chain home slot -> claim -> list reservation -> add
// of ~0.5 us links, and a GPU thread issues in order: so the four home slots are resolved together, and then EVERY atomic of the
// occurrence -- the list reservation for newly claimed keys, the four (delta, seq) reductions, the rank draws of the two keys
// this merge creates -- is issued before the first result is consumed.  A key whose home slot holds another key takes the
// probing dt_add afterwards.  cas_first: claim without looking first (one round trip less; for launches with few occurrences,
// where no key is hot).  want_rank: keys that also draw a rank in their future occurrence list (dt.nocc).
// slot_out[j] / rank_out[j]: the key's slot and rank (NONE32 / 0 when not applicable).
__device__ __forceinline__ void dt_emit4(const DeltaTable& dt, DevCounters* ctr, const uint64_t (&key)[4], const int64_t (&delta)[4], const uint64_t (&seq)[4],
                                         uint32_t valid, uint32_t want_rank, bool cas_first, uint32_t (&slot_out)[4], uint32_t (&rank_out)[4]) {
  uint64_t slot[4], cur[4];
  uint32_t claimed = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) { slot[j] = mix64(key[j]) & dt.mask; slot_out[j] = NONE32; rank_out[j] = 0; cur[j] = 0ull; }
  if (cas_first) {
#pragma unroll
    for (int j = 0; j < 4; j++) if ((valid >> j) & 1u)
      cur[j] = atomicCAS(reinterpret_cast<ull*>(&dt.keys[slot[j]]), static_cast<ull>(dt.empty), static_cast<ull>(key[j]));
#pragma unroll
    for (int j = 0; j < 4; j++) if (((valid >> j) & 1u) && cur[j] == dt.empty) { claimed |= 1u << j; cur[j] = key[j]; }
  } else {
#pragma unroll
    for (int j = 0; j < 4; j++) if ((valid >> j) & 1u) cur[j] = dt.keys[slot[j]];
    uint64_t prev[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      prev[j] = cur[j];
      if (((valid >> j) & 1u) && cur[j] == dt.empty)
        prev[j] = atomicCAS(reinterpret_cast<ull*>(&dt.keys[slot[j]]), static_cast<ull>(dt.empty), static_cast<ull>(key[j]));
    }
#pragma unroll
    for (int j = 0; j < 4; j++) if (((valid >> j) & 1u) && cur[j] == dt.empty) {
      if (prev[j] == dt.empty) { claimed |= 1u << j; cur[j] = key[j]; } else cur[j] = prev[j];
    }
  }
  // everything below is issued before anything is consumed
  uint32_t idx = 0;
  if (claimed) idx = atomicAdd(dt.n, static_cast<uint32_t>(__popc(claimed)));
#pragma unroll
  for (int j = 0; j < 4; j++) if (((valid >> j) & 1u) && cur[j] == key[j]) {
    atomicAdd(&dt.delta[slot[j]], static_cast<ull>(delta[j]));
    atomicMin(&dt.seq[slot[j]], static_cast<ull>(seq[j]));
    slot_out[j] = static_cast<uint32_t>(slot[j]);
    if ((want_rank >> j) & 1u) rank_out[j] = atomicAdd(&dt.nocc[slot[j]], 1u);
  }
  if (claimed) {
#pragma unroll
    for (int j = 0; j < 4; j++) if ((claimed >> j) & 1u) {
      if (idx < dt.cap) { dt.list[idx] = static_cast<uint32_t>(slot[j]); dt.klist[idx] = key[j]; }
      ++idx;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; j++) if (((valid >> j) & 1u) && cur[j] != key[j]) {  // home slot taken by another key
    slot_out[j] = dt_add(dt, ctr, key[j], delta[j], seq[j]);
    if (((want_rank >> j) & 1u) && slot_out[j] != NONE32) rank_out[j] = atomicAdd(&dt.nocc[slot_out[j]], 1u);
  }
}

__device__ __forceinline__ ull gtime() { ull t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ ulonglong2 ld_ent(const PairEnt* e) { return *reinterpret_cast<const ulonglong2*>(e); }

// Finds `key` or claims an empty slot for it WITHOUT assigning the serial (the caller batches the serial counter with its other
// counters).  Returns the slot; *is_new tells which; *old_freq = its frequency (0 for a new entry).
__device__ __forceinline__ uint64_t pt_find_or_claim(const PairTable& pt, DevCounters* ctr, uint64_t key, ulonglong2 first, uint64_t* old_freq, bool* is_new) {
  uint64_t slot = mix64(key) & pt.mask;
  ulonglong2 e = first;
  *is_new = false;
  for (uint64_t probe = 0; probe < pt.cap; ++probe) {
    if (e.x == key) { *old_freq = e.y; return slot; }
    if (e.x == PT_EMPTY) {
      const uint64_t prev = atomicCAS(reinterpret_cast<ull*>(&pt.ent[slot].key), static_cast<ull>(PT_EMPTY), static_cast<ull>(key));
      if (prev == PT_EMPTY) { *old_freq = 0; *is_new = true; return slot; }
      if (prev == key) { *old_freq = pt.ent[slot].freq; return slot; }
    }
    slot = (slot + 1) & pt.mask;
    e = ld_ent(&pt.ent[slot]);
  }
  atomicOr(&ctr->err, ERR_PT_FULL);
  *old_freq = 0;
  return 0;
}

// Finds `key` (inserting it if absent) and returns its slot; *old_freq = its frequency (0 for a new entry).
// `first` is the already-loaded entry at the home slot (lets the caller issue several home loads back to back).
// Used for the merged pair's own entry (one thread per pass); the folds use pt_find_or_claim and batch the serial counter.
__device__ __forceinline__ uint64_t pt_find_or_insert(const PairTable& pt, DevCounters* ctr, uint64_t key, ulonglong2 first, uint64_t* old_freq) {
  uint64_t slot = mix64(key) & pt.mask;
  ulonglong2 e = first;
  for (uint64_t probe = 0; probe < pt.cap; ++probe) {
    if (e.x == key) { *old_freq = e.y; return slot; }
    if (e.x == PT_EMPTY) {
      const uint64_t prev = atomicCAS(reinterpret_cast<ull*>(&pt.ent[slot].key), static_cast<ull>(PT_EMPTY), static_cast<ull>(key));
      if (prev == PT_EMPTY) { pt.ent[slot].serial = static_cast<uint32_t>(atomicAdd(&ctr->pt_n, 1ull)); *old_freq = 0; return slot; }
      if (prev == key) { *old_freq = pt.ent[slot].freq; return slot; }
    }
    slot = (slot + 1) & pt.mask;
    e = ld_ent(&pt.ent[slot]);
  }
  atomicOr(&ctr->err, ERR_PT_FULL);
  *old_freq = 0;
  return 0;
}
