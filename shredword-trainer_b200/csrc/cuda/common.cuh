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

constexpr int32_t DEAD = -1;
constexpr uint32_t HDR_BIT = 0x80000000u;
constexpr int32_t UNK_CODE_NEG = 0x7FFFFFFF;  // stored code of unk symbols when unk_id < 0
constexpr uint64_t PT_EMPTY = ~0ull;
constexpr uint64_t SEQ_MAX = ~0ull;
constexpr int N_SM_FALLBACK = 148;

enum : uint32_t { ERR_DT_FULL = 1, ERR_PT_FULL = 2, ERR_WT_FULL = 4, ERR_WT_COLLISION = 8, ERR_REC_FULL = 16, ERR_HAS_NUL = 32, ERR_BARRIER = 64,
                  ERR_PEER_TIMEOUT = 128, ERR_INBOX_FULL = 256 };

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
__device__ __forceinline__ bool is_delim(uint32_t c) { return c <= 32u && ((0x100002600ull >> c) & 1ull); }  // \t \n \r space
__device__ __forceinline__ uint64_t fc_key(int32_t a, int32_t b) {  // bpe.cpp:277-278: both operands sign-extend
  return (static_cast<uint64_t>(static_cast<int64_t>(a)) << 32) | static_cast<uint64_t>(static_cast<int64_t>(b));
}

struct Ctrl {  // mapped pinned host memory, written by finalize_block
  volatile uint64_t flag;
  uint64_t n_recs, occ, occ_local, pt_n, n_leaders, n_keys, cand_tiles;
  uint32_t err, pad;
};

struct DevCounters {  // device memory
  uint32_t wl_n, dt_n, rec_n, blocks_done;
  ull occ;
  ull pt_n;
  uint32_t err, pad;
  ull n_tokens;
  uint32_t n_unique, blocks_done2;
  uint32_t cand_tiles, bar;
  uint32_t sent_ctas, pad4;
};

struct DeltaTable {
  uint64_t* keys; ull* delta; ull* seq; uint32_t* list; uint64_t* klist;  // list/klist: slot and key of every used slot, dense
  uint64_t mask; uint64_t empty; uint32_t cap;
};
struct PairEnt { uint64_t key; uint64_t freq; };  // one 16-byte load fetches both
struct PairTable {
  PairEnt* ent;
  uint32_t* serial;  // dense id per entry = number of entries that existed when it was created (the host indexes by it)
  uint64_t mask; uint64_t cap;
};

// ------------------------------------------------------------------------------------------------ hash-table helpers

__device__ __forceinline__ void dt_add(const DeltaTable& dt, DevCounters* ctr, uint64_t key, int64_t delta, uint64_t seq) {
  uint64_t slot = mix64(key) & dt.mask;
  for (uint32_t probe = 0; probe < dt.cap; ++probe) {
    uint64_t cur = dt.keys[slot];
    if (cur == dt.empty) {
      uint64_t prev = atomicCAS(reinterpret_cast<ull*>(&dt.keys[slot]), static_cast<ull>(dt.empty), static_cast<ull>(key));
      if (prev == dt.empty) {
        uint32_t idx = atomicAdd(&ctr->dt_n, 1u);
        if (idx < dt.cap) { dt.list[idx] = static_cast<uint32_t>(slot); dt.klist[idx] = key; }
        cur = key;
      } else cur = prev;
    }
    if (cur == key) {
      atomicAdd(&dt.delta[slot], static_cast<ull>(delta));
      atomicMin(&dt.seq[slot], static_cast<ull>(seq));
      return;
    }
    slot = (slot + 1) & dt.mask;
  }
  atomicOr(&ctr->err, ERR_DT_FULL);
}

// The four delta-table updates of one occurrence with their memory round trips overlapped (each update alone is a chain
// load -> CAS -> list reservation of ~0.5 us links): home-slot loads together, claims together, one list reservation
// for all newly claimed keys.  A key whose home slot holds another key falls back to the probing dt_add.
__device__ __forceinline__ void dt_add4(const DeltaTable& dt, DevCounters* ctr, const uint64_t (&key)[4], const int64_t (&delta)[4], const uint64_t (&seq)[4],
                                        uint32_t valid) {
  uint64_t slot[4], cur[4], prev[4];
#pragma unroll
  for (int j = 0; j < 4; j++) { slot[j] = mix64(key[j]) & dt.mask; cur[j] = (valid >> j) & 1u ? dt.keys[slot[j]] : 0ull; }
#pragma unroll
  for (int j = 0; j < 4; j++) {
    prev[j] = cur[j];
    if (((valid >> j) & 1u) && cur[j] == dt.empty)
      prev[j] = atomicCAS(reinterpret_cast<ull*>(&dt.keys[slot[j]]), static_cast<ull>(dt.empty), static_cast<ull>(key[j]));
  }
  uint32_t claimed = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) if (((valid >> j) & 1u) && cur[j] == dt.empty) {
    if (prev[j] == dt.empty) { claimed |= 1u << j; cur[j] = key[j]; } else cur[j] = prev[j];
  }
  if (claimed) {
    uint32_t idx = atomicAdd(&ctr->dt_n, static_cast<uint32_t>(__popc(claimed)));
#pragma unroll
    for (int j = 0; j < 4; j++) if ((claimed >> j) & 1u) {
      if (idx < dt.cap) { dt.list[idx] = static_cast<uint32_t>(slot[j]); dt.klist[idx] = key[j]; }
      ++idx;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; j++) if ((valid >> j) & 1u) {
    if (cur[j] == key[j]) {
      atomicAdd(&dt.delta[slot[j]], static_cast<ull>(delta[j]));
      atomicMin(&dt.seq[slot[j]], static_cast<ull>(seq[j]));
    } else dt_add(dt, ctr, key[j], delta[j], seq[j]);
  }
}

__device__ __forceinline__ ull gtime() { ull t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ ulonglong2 ld_ent(const PairEnt* e) { return *reinterpret_cast<const ulonglong2*>(e); }

// Finds `key` (inserting it if absent) and returns its slot; *old_freq = its frequency (0 for a new entry).
// `first` is the already-loaded entry at the home slot (lets the caller issue several home loads back to back).
// Only finalize_block calls this, with distinct keys per pass, so a claimed entry has exactly one writer.
__device__ __forceinline__ uint64_t pt_find_or_insert(const PairTable& pt, DevCounters* ctr, uint64_t key, ulonglong2 first, uint64_t* old_freq) {
  uint64_t slot = mix64(key) & pt.mask;
  ulonglong2 e = first;
  for (uint64_t probe = 0; probe < pt.cap; ++probe) {
    if (e.x == key) { *old_freq = e.y; return slot; }
    if (e.x == PT_EMPTY) {
      const uint64_t prev = atomicCAS(reinterpret_cast<ull*>(&pt.ent[slot].key), static_cast<ull>(PT_EMPTY), static_cast<ull>(key));
      if (prev == PT_EMPTY) { pt.serial[slot] = static_cast<uint32_t>(atomicAdd(&ctr->pt_n, 1ull)); *old_freq = 0; return slot; }
      if (prev == key) { *old_freq = pt.ent[slot].freq; return slot; }
    }
    slot = (slot + 1) & pt.mask;
    e = ld_ent(&pt.ent[slot]);
  }
  atomicOr(&ctr->err, ERR_PT_FULL);
  *old_freq = 0;
  return 0;
}
