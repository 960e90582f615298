// trainer_core.hpp -- host control of the B200 BPE trainer: everything in the reference's train loop that is
// inherently sequential (exact heap replay, push ordering, lazy-invalidation versions, phantom unk pairs), driving an
// Engine for everything that is data parallel.  See trainer_core.cpp for the reference file:line map.
#pragma once
#include <cstdint>
#include <cstdio>
#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "../../include/shred_abi.h"
#include "engine.hpp"
#include "exact_heap.hpp"
#include "flat_map.hpp"

namespace shred {

class TrainerCore {
 public:
  TrainerCore(Trainer* abi, Engine* eng);
  ~TrainerCore();

  int load_file(const char* path);
  int load_buffer(const uint8_t* text, size_t n);
  void count_bigrams();
  void init();
  int merge_batch(int batch_size);
  int merge_loop(int batch_size);
  int train();
  void save(const char* model_path, const char* vocab_path);
  void get_stats(shred_stats_t* out);
  Engine* engine() { return eng_; }
  const LoadInfo& load_info() const { return info_; }
  bool loaded() const { return loaded_; }

 private:
  void sync_mirrors();
  EngineConfig engine_config() const;
  int finish_load(size_t n_bytes, double t0_ms);
  bool is_phantom(int32_t a, int32_t b) const { return a == abi_->config.unk_id || b == abi_->config.unk_id; }
  void apply_records(const Rec* recs, size_t n);

  Trainer* abi_;
  Engine* eng_;
  ExactHeap heap_;
  FlatMap<uint32_t> version_;   // PHANTOM pair key -> current version (absent = 0); phantoms have no device serial
  struct PairMeta { uint32_t ver, list_len; };  // list_len: entries of the pair's occurrence list on the device (engine.hpp rec_list_len)
  HugeArray<PairMeta> ver_;     // pair serial -> current version + list length (dense, no hashing on the replay path)
  PairMeta& meta_of(uint32_t serial) {
    if (serial >= ver_.size()) ver_.ensure(std::max<size_t>(serial + 1, ver_.size() * 2 + 1024));
    return ver_[serial];
  }
  uint32_t& ver_of(uint32_t serial, uint64_t key) {
    if (serial == REC_NO_SERIAL) return version_[key];
    return meta_of(serial).ver;
  }
  FlatMap<uint64_t> phantom_;   // pair keys containing unk_id -> freq as the reference's table would hold it
  std::vector<Rec> order_;      // scratch: records in application order
  std::vector<int32_t> bucket_head_, next_in_bucket_;  // scratch of apply_records
  std::vector<uint32_t> order_idx_;
  LoadInfo info_;
  bool loaded_ = false;
  size_t merge_cap_ = 0;
  Symbol placeholder_;
  // statistics of the last load/train
  uint64_t occurrences_ = 0, merges_last_ = 0, corpus_bytes_ = 0;
  uint64_t tie_root_equal_ = 0, tie_same_as_prev_ = 0, last_merge_freq_ = ~0ull;
  double load_wall_ms_ = 0, train_wall_ms_ = 0, train_device_ms_ = 0, host_heap_ms_ = 0, host_pop_ms_ = 0, host_apply_ms_ = 0, save_wall_ms_ = 0;
  bool log_merges_ = false, quiet_ = false;
  FILE* trace_file_ = nullptr;
};

}  // namespace shred
