// Shared device/host helpers for the bmkg_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define BMKG_OK 0
#define BMKG_ERR_BAD_ARG (-1)
#define BMKG_ERR_MISALIGNED (-2)
#define BMKG_ERR_WORKSPACE (-3)
#define BMKG_ERR_LAUNCH (-4)
#define BMKG_ERR_UNSUPPORTED (-5)
#define BMKG_ERR_DRIVER (-6)

#define BMKG_CHECK_LAUNCH()                          \
  do {                                               \
    if (cudaGetLastError() != cudaSuccess) return BMKG_ERR_LAUNCH; \
  } while (0)

#define BMKG_REQUIRE(cond, code) \
  do {                           \
    if (!(cond)) return (code);  \
  } while (0)

namespace bmkg {

constexpr int kWarp = 32;
constexpr int kNumSMs = 148;  // B200
// split-row aggregation: rows longer than kHubThreshold edges are reduced chunk-wise (kHubSeg <= kHubThreshold / 2)
#ifndef BMKG_HUB_THRESHOLD
#define BMKG_HUB_THRESHOLD 1024
#define BMKG_HUB_SEG 512
#endif
constexpr int kHubThreshold = BMKG_HUB_THRESHOLD;
constexpr int kHubSeg = BMKG_HUB_SEG;
static_assert(kHubSeg * 2 <= kHubThreshold, "a chunk may meet at most two hub rows");

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// carve typed arrays out of a caller-provided workspace
struct WsCarver {
  char* base;
  size_t off;
  __host__ explicit WsCarver(void* p) : base(static_cast<char*>(p)), off(0) {}
  template <typename T>
  __host__ T* take(size_t n) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return r;
  }
  __host__ size_t used() const { return align_up(off, 256); }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Split-row pre-pass set-up shared by the GCN / GAT hub kernels.  hub_info = [#hub rows | first row of chunk 0 | chunk 1 | ...]
// (written by bmkg_csr_filter).  Finds the (at most two) hub rows the chunk blockIdx.x meets: s_rows[0] = the one continuing
// into the chunk, s_rows[1] = the one starting inside it.  Returns false when the CTA has nothing to do.
__device__ __forceinline__ bool hub_chunk_rows(const int32_t* __restrict__ rowptr, int64_t N, const int32_t* __restrict__ hub_info,
                                               int* s_rows, int& cs, int& ce) {
  if (hub_info == nullptr || hub_info[0] == 0) return false;
  const int r0 = hub_info[1 + blockIdx.x];
  if (r0 < 0) return false;
  const int nnz = rowptr[N];
  cs = blockIdx.x * kHubSeg;
  ce = min(cs + kHubSeg, nnz);
  const int nxt = (ce < nnz) ? hub_info[2 + blockIdx.x] : (int)N - 1;   // row holding the next chunk's first edge
  const int r1 = nxt < 0 ? (int)N - 1 : nxt;
  if (threadIdx.x == 0) { s_rows[0] = -1; s_rows[1] = -1; }
  __syncthreads();
  for (int r = r0 + (int)threadIdx.x; r <= r1; r += blockDim.x) {
    const int b = rowptr[r], e = rowptr[r + 1];
    if (e - b > kHubThreshold && b < ce && e > cs) s_rows[(b < cs) ? 0 : 1] = r;   // at most one row per slot
  }
  __syncthreads();
  return (s_rows[0] >= 0) || (s_rows[1] >= 0);
}

// 8 bf16 packed in a uint4 <-> 8 floats
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
}

// streaming 128-bit global load that does not allocate in L1 (read-once rows)
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// gather load that may be re-used across CTAs: default caching
__device__ __forceinline__ uint4 ldg_cached(const void* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}

// Counter-based keep/drop stream shared by every stochastic kernel: a
// Philox-free 2-round multiply-xorshift hash of (seed, stream, index).  The
// Python host mirrors it exactly (biomedkg_b200/draws.py) so tests can replay
// the same masks through the oracle.
__host__ __device__ __forceinline__ uint32_t hash_u32(uint64_t seed, uint64_t idx) {
  uint64_t z = idx * 0x9E3779B97F4A7C15ull + seed;
  z ^= z >> 30;
  z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27;
  z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  return static_cast<uint32_t>(z >> 32);
}
// keep iff u >= p with u = hash * 2^-32 in [0,1)
__host__ __device__ __forceinline__ bool hash_keep(uint64_t seed, uint64_t idx, uint32_t p_threshold) {
  return hash_u32(seed, idx) >= p_threshold;
}

}  // namespace bmkg
