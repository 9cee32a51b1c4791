// F2 - ReDAF fusion epilogue (biomedkg/utils/fusion.py:70-90).
//
// The reference runs  relu -> * modal_weights * zeta_r -> dropout(0.1) -> relu -> mean over modalities  as five
// elementwise passes over [N,M,E] fp32 tensors.  Here the transform GEMM writes a bias-free bf16 [N,M,E] once and one
// kernel does the rest: per (node, 8 columns) it reads the M modality rows (one 128-bit load each), adds the Linear bias,
// applies ReLU, the per-(modality, column) gate = modal_weights * sigmoid(relational_context_layer(0.2)), the dropout
// decision (counter-based hash or an explicit keep mask), the second ReLU and the mean.  The backward recomputes the
// same decisions from t and writes dt (bf16) plus per-CTA partial sums of d gate (reduced by bmkg_colsum: deterministic).
//
// HBM-bound.  Algorithmic bytes: forward N*M*E*2 + N*E*4;  backward N*E*4 + N*M*E*2 (t) + N*M*E*2 (dt).
#include "common.cuh"
#include "../../include/bmkg_b200.h"

namespace bmkg {

struct RedafDrop {
  float scale;          // 1/(1-p) or 1
  uint32_t threshold;   // p * 2^16; 0 = no hashed dropout
  uint64_t seed;
  const uint8_t* keep;  // explicit [N,M,E] keep mask or null
};

__device__ __forceinline__ void redaf_keep8(const RedafDrop& d, int64_t base, float* k) {
  if (d.keep) {
    const uint2 m = *reinterpret_cast<const uint2*>(d.keep + base);
    const uint32_t mm[2] = {m.x, m.y};
#pragma unroll
    for (int i = 0; i < 8; ++i) k[i] = ((mm[i >> 2] >> (8 * (i & 3))) & 0xff) ? d.scale : 0.f;
  } else if (d.threshold) {
    // one 64-bit hash decides two consecutive elements (16 bits each): with M*E decisions per node the hash arithmetic,
    // not HBM, was the limit (ncu: 70 % issue-active at 0.40 of the HBM roof with one hash per element)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t h = hash_u32(d.seed, (uint64_t)(base >> 1) + j);
      k[2 * j] = (h & 0xffffu) >= d.threshold ? d.scale : 0.f;
      k[2 * j + 1] = (h >> 16) >= d.threshold ? d.scale : 0.f;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) k[i] = 1.f;
  }
}

__device__ __forceinline__ void load8f(const float* p, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// grid (row blocks, column slabs); thread = one group of 8 columns, rows strided by gridDim.x
template <int M>
__global__ void __launch_bounds__(128) redaf_fwd_kernel(const __nv_bfloat16* __restrict__ t, const float* __restrict__ bias,
                                                        const float* __restrict__ gate, int64_t N, int E, RedafDrop drop,
                                                        float* __restrict__ out) {
  const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * 8;
  if (c0 >= E) return;
  float b[8], g[M][8];
  load8f(bias + c0, b);
#pragma unroll
  for (int m = 0; m < M; ++m) load8f(gate + (int64_t)m * E + c0, g[m]);
  for (int64_t n = blockIdx.x; n < N; n += gridDim.x) {
    uint4 u[M];
#pragma unroll
    for (int m = 0; m < M; ++m) u[m] = ldg_stream(t + (n * M + m) * E + c0);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int m = 0; m < M; ++m) {
      float f[8], k[8];
      unpack8(u[m], f);
      redaf_keep8(drop, (n * M + m) * E + c0, k);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += fmaxf(fmaxf(f[i] + b[i], 0.f) * g[m][i] * k[i], 0.f);
    }
    float* o = out + n * E + c0;
    constexpr float inv = 1.0f / M;
    *reinterpret_cast<float4*>(o) = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
    *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
  }
}

template <int M>
__global__ void __launch_bounds__(128) redaf_bwd_kernel(const __nv_bfloat16* __restrict__ t, const float* __restrict__ bias,
                                                        const float* __restrict__ gate, const float* __restrict__ dout, int64_t N,
                                                        int E, RedafDrop drop, __nv_bfloat16* __restrict__ dt,
                                                        float* __restrict__ dgate_partial /*[gridDim.x][M][E]*/) {
  const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * 8;
  if (c0 >= E) return;
  float b[8], g[M][8], dg[M][8];
  load8f(bias + c0, b);
#pragma unroll
  for (int m = 0; m < M; ++m) {
    load8f(gate + (int64_t)m * E + c0, g[m]);
#pragma unroll
    for (int i = 0; i < 8; ++i) dg[m][i] = 0.f;
  }
  constexpr float inv = 1.0f / M;
  for (int64_t n = blockIdx.x; n < N; n += gridDim.x) {
    uint4 u[M];
#pragma unroll
    for (int m = 0; m < M; ++m) u[m] = ldg_stream(t + (n * M + m) * E + c0);
    float go[8];
    load8f(dout + n * E + c0, go);
#pragma unroll
    for (int m = 0; m < M; ++m) {
      float f[8], k[8], r[8];
      unpack8(u[m], f);
      redaf_keep8(drop, (n * M + m) * E + c0, k);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float pre = f[i] + b[i];
        const float a = fmaxf(pre, 0.f);                    // first ReLU
        const float v = a * g[m][i] * k[i];                 // gated, dropped
        const float dv = (v > 0.f) ? go[i] * inv * k[i] : 0.f;   // through the second ReLU and the dropout scale
        dg[m][i] += dv * a;
        r[i] = (pre > 0.f) ? dv * g[m][i] : 0.f;
      }
      *reinterpret_cast<uint4*>(dt + (n * M + m) * E + c0) = pack8(r);
    }
  }
#pragma unroll
  for (int m = 0; m < M; ++m) {
    float* p = dgate_partial + ((int64_t)blockIdx.x * M + m) * E + c0;
    *reinterpret_cast<float4*>(p) = make_float4(dg[m][0], dg[m][1], dg[m][2], dg[m][3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(dg[m][4], dg[m][5], dg[m][6], dg[m][7]);
  }
}

static int redaf_threads(int E) {
  const int groups = (E / 8 + 31) / 32 * 32;
  return groups < 128 ? groups : 128;
}
static int redaf_row_blocks(int64_t N, int E) {
  const int64_t slabs = ceil_div(E / 8, 128);
  const int64_t blocks = (int64_t)kNumSMs * (2048 / redaf_threads(E)) / slabs;   // one full wave of resident CTAs
  return (int)(N < blocks ? N : (blocks < 1 ? 1 : blocks));
}

static RedafDrop make_drop(float p, uint64_t seed, const uint8_t* keep) {
  RedafDrop d;
  d.scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  d.keep = p > 0.f ? keep : nullptr;
  d.threshold = (p > 0.f && !keep) ? (uint32_t)((double)p * 65536.0) : 0u;   // p < 2^-16 degenerates to no dropout
  d.seed = seed;
  return d;
}

}  // namespace bmkg

using namespace bmkg;

extern "C" int64_t bmkg_redaf_partial_rows(int64_t N, int E) { return (N > 0 && E >= 8) ? redaf_row_blocks(N, E) : 0; }

#define BMKG_REDAF_CHECKS()                                                                             \
  BMKG_REQUIRE(t_bf16 && bias && gate, BMKG_ERR_BAD_ARG);                                               \
  BMKG_REQUIRE(N > 0 && M >= 1 && M <= 4 && E >= 8 && E % 8 == 0 && E <= 8192, BMKG_ERR_BAD_ARG);       \
  BMKG_REQUIRE(drop_p >= 0.f && drop_p < 1.f, BMKG_ERR_BAD_ARG);                                        \
  BMKG_REQUIRE(aligned16(t_bf16) && aligned16(bias) && aligned16(gate) && (!drop_keep || aligned16(drop_keep)), BMKG_ERR_MISALIGNED)

extern "C" int bmkg_redaf_fwd(const void* t_bf16, const float* bias, const float* gate, int64_t N, int M, int E, float drop_p,
                              uint64_t drop_seed, const uint8_t* drop_keep, float* out, void* stream) {
  BMKG_REDAF_CHECKS();
  BMKG_REQUIRE(out && aligned16(out), BMKG_ERR_BAD_ARG);
  const RedafDrop d = make_drop(drop_p, drop_seed, drop_keep);
  const dim3 grid(redaf_row_blocks(N, E), (unsigned)ceil_div(E / 8, 128));
  const int threads = redaf_threads(E);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* t = static_cast<const __nv_bfloat16*>(t_bf16);
  switch (M) {
    case 1: redaf_fwd_kernel<1><<<grid, threads, 0, st>>>(t, bias, gate, N, E, d, out); break;
    case 2: redaf_fwd_kernel<2><<<grid, threads, 0, st>>>(t, bias, gate, N, E, d, out); break;
    case 3: redaf_fwd_kernel<3><<<grid, threads, 0, st>>>(t, bias, gate, N, E, d, out); break;
    default: redaf_fwd_kernel<4><<<grid, threads, 0, st>>>(t, bias, gate, N, E, d, out); break;
  }
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

extern "C" int bmkg_redaf_bwd(const void* t_bf16, const float* bias, const float* gate, const float* dout, int64_t N, int M, int E,
                              float drop_p, uint64_t drop_seed, const uint8_t* drop_keep, void* dt_bf16, float* dgate_partial,
                              void* stream) {
  BMKG_REDAF_CHECKS();
  BMKG_REQUIRE(dout && dt_bf16 && dgate_partial && aligned16(dout) && aligned16(dt_bf16) && aligned16(dgate_partial),
               BMKG_ERR_BAD_ARG);
  const RedafDrop d = make_drop(drop_p, drop_seed, drop_keep);
  const dim3 grid(redaf_row_blocks(N, E), (unsigned)ceil_div(E / 8, 128));
  const int threads = redaf_threads(E);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* t = static_cast<const __nv_bfloat16*>(t_bf16);
  __nv_bfloat16* dt = static_cast<__nv_bfloat16*>(dt_bf16);
  switch (M) {
    case 1: redaf_bwd_kernel<1><<<grid, threads, 0, st>>>(t, bias, gate, dout, N, E, d, dt, dgate_partial); break;
    case 2: redaf_bwd_kernel<2><<<grid, threads, 0, st>>>(t, bias, gate, dout, N, E, d, dt, dgate_partial); break;
    case 3: redaf_bwd_kernel<3><<<grid, threads, 0, st>>>(t, bias, gate, dout, N, E, d, dt, dgate_partial); break;
    default: redaf_bwd_kernel<4><<<grid, threads, 0, st>>>(t, bias, gate, dout, N, E, d, dt, dgate_partial); break;
  }
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}
