// kernels_scan.cuh -- device-wide exclusive scan of uint64 (three passes); shared by engine_cuda.cu and encoder_cuda.cu.
// Fragment: included inside `namespace shred { namespace {` after common.cuh.
#pragma once

// ---- device-wide exclusive scan of uint64 (three passes; 2048 items per block) used at load and at compaction
constexpr int SCAN_ITEMS = 8, SCAN_THREADS = 256, SCAN_TILE = SCAN_ITEMS * SCAN_THREADS;

__device__ __forceinline__ ull block_excl_scan(ull v, ull* total) {  // 256 threads
  __shared__ ull wsum[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  ull x = v;
  for (int o = 1; o < 32; o <<= 1) { ull y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  if (warp == 0) {
    ull s = lane < 8 ? wsum[lane] : 0;
    for (int o = 1; o < 8; o <<= 1) { ull y = __shfl_up_sync(0xFFFFFFFFu, s, o); if (lane >= o) s += y; }
    if (lane < 8) wsum[lane] = s;
  }
  __syncthreads();
  const ull before = warp ? wsum[warp - 1] : 0;
  *total = wsum[7];
  __syncthreads();
  return before + x - v;
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(const ull* in, uint64_t n, ull* sums) {
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * SCAN_TILE + static_cast<uint64_t>(threadIdx.x) * SCAN_ITEMS;
  ull s = 0;
  for (int i = 0; i < SCAN_ITEMS; i++) if (base + i < n) s += in[base + i];
  ull total;
  block_excl_scan(s, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_top(ull* sums, uint32_t nb, ull* grand_total) {  // one block
  ull carry = 0;
  for (uint32_t base = 0; base < nb; base += SCAN_THREADS) {
    const uint32_t i = base + threadIdx.x;
    ull v = i < nb ? sums[i] : 0, total;
    ull ex = block_excl_scan(v, &total);
    if (i < nb) sums[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) *grand_total = carry;
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const ull* in, uint64_t n, const ull* sums, ull* out) {
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * SCAN_TILE + static_cast<uint64_t>(threadIdx.x) * SCAN_ITEMS;
  ull v[SCAN_ITEMS], s = 0;
  for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = base + i < n ? in[base + i] : 0; s += v[i]; }
  ull total;
  ull run = sums[blockIdx.x] + block_excl_scan(s, &total);
  for (int i = 0; i < SCAN_ITEMS; i++) { if (base + i < n) out[base + i] = run; run += v[i]; }
}
