// Device-wide exclusive scan of int32 values produced by a functor (shared by graph.cu and sampler.cu).
// Three launches: per-tile sums, one-CTA scan of the tile sums, down-sweep.  Integer: exact and order-independent.
#pragma once
#include "common.cuh"

namespace bmkg {

// ---------------------------------------------------------------------------
// device-wide exclusive scan of int32 values produced by a functor
// ---------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096

struct LoadI32 {
  const int32_t* p;
  __device__ int operator()(int64_t i) const { return p[i]; }
};
struct LoadU8 {
  const uint8_t* p;
  __device__ int operator()(int64_t i) const { return p[i]; }
};
__device__ __forceinline__ int block_exclusive_scan(int thread_sum, int* smem_warp, int& block_total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = thread_sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) smem_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = lane < (kScanThreads / 32) ? smem_warp[lane] : 0;
    int vi = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, vi, o);
      if (lane >= o) vi += t;
    }
    if (lane < (kScanThreads / 32)) smem_warp[lane] = vi - v;  // exclusive warp offsets
    if (lane == 31) smem_warp[32] = vi;                         // block total
  }
  __syncthreads();
  block_total = smem_warp[32];
  return smem_warp[warp] + incl - thread_sum;
}

template <class F>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(F f, int64_t n, int* __restrict__ block_sums) {
  __shared__ int sw[33];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int64_t i = base + j;
    if (i < n) s += f(i);
  }
  int total;
  block_exclusive_scan(s, sw, total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// in-place exclusive scan of block sums by one CTA; writes the grand total to total_out
static __global__ void __launch_bounds__(1024) scan_block_sums_kernel(int* __restrict__ sums, int nb, int* __restrict__ total_out) {
  __shared__ int sw[33];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    int i = base + threadIdx.x;
    int v = i < nb ? sums[i] : 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) sw[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = sw[lane];
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      sw[lane] = wi - w;
      if (lane == 31) sw[32] = wi;
    }
    __syncthreads();
    int carry = carry_s;
    if (i < nb) sums[i] = carry + sw[warp] + incl - v;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + sw[32];
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry_s;
}

// out has n+1 entries: out[i] = sum_{j<i} f(j), out[n] = total
template <class F>
__global__ void __launch_bounds__(kScanThreads) scan_down_kernel(F f, int64_t n, const int* __restrict__ block_offs,
                                                                 int* __restrict__ out) {
  __shared__ int sw[33];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int64_t i = base + j;
    v[j] = (i < n) ? f(i) : 0;
    s += v[j];
  }
  int total;
  int run = block_exclusive_scan(s, sw, total) + block_offs[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int64_t i = base + j;
    if (i <= n) out[i] = run;  // i == n writes the grand total exactly once
    run += v[j];
  }
}

inline size_t scan_ws_ints(int64_t n) { return (size_t)ceil_div(n + 1, kScanTile) + 1; }

// two-functor form: `fr` feeds the tile sums (and may record what it computed), `fd` feeds the down-sweep
template <class FR, class FD>
static int exclusive_scan2(FR fr, FD fd, int64_t n, int* out, int* ws_block_sums, cudaStream_t st) {
  const int nb = (int)ceil_div(n + 1, kScanTile);
  scan_reduce_kernel<FR><<<nb, kScanThreads, 0, st>>>(fr, n, ws_block_sums);
  scan_block_sums_kernel<<<1, 1024, 0, st>>>(ws_block_sums, nb, nullptr);
  scan_down_kernel<FD><<<nb, kScanThreads, 0, st>>>(fd, n, ws_block_sums, out);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

template <class F>
static int exclusive_scan(F f, int64_t n, int* out, int* ws_block_sums, cudaStream_t st) {
  const int nb = (int)ceil_div(n + 1, kScanTile);
  scan_reduce_kernel<F><<<nb, kScanThreads, 0, st>>>(f, n, ws_block_sums);
  scan_block_sums_kernel<<<1, 1024, 0, st>>>(ws_block_sums, nb, nullptr);
  scan_down_kernel<F><<<nb, kScanThreads, 0, st>>>(f, n, ws_block_sums, out);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}


}  // namespace bmkg
