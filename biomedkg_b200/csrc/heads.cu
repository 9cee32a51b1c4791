// D1-D3 - DGI / GGD discriminator heads as fused, deterministic reductions.
//
//   DGI  (model/gcl.py:19-27 + PyGCL SingleBranchContrast(JSD,"G2L"), gcl_module.py:127,142):
//        summary = sigmoid(mean_0 z)                          -> bmkg_colmean_sigmoid
//        s+ = z g^T, s- = zn g^T                              -> bmkg_rowdot
//        loss = mean softplus(-s+) + mean softplus(s-) - 2ln2 -> bmkg_softplus_pair_sum
//   GGD  (model/gcl.py:83-91, gcl_module.py:229-234):
//        (z W^T + b).sum(1) == z . (sum_rows W) + sum(b)      -> bmkg_rowdot  (GEMM collapses to a GEMV)
//        BCEWithLogits(cat(pos,neg), cat(1,0)) == (sum softplus(-pos) + sum softplus(neg)) / 2N
// HBM-bound: z and zn are read exactly once (2*N*C*4 bytes).
#include "common.cuh"
#include "../../include/bmkg_b200.h"

namespace bmkg {

__device__ __forceinline__ float softplusf(float x) {  // log(1 + e^x), stable
  return fmaxf(x, 0.f) + log1pf(__expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + __expf(-x)); }

// out[i] = z[i,:] . v   (one warp per row)
__global__ void __launch_bounds__(256) rowdot_kernel(const float* __restrict__ z, const float* __restrict__ v, int64_t N, int C,
                                                     float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= N) return;
  float s = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 a = *reinterpret_cast<const float4*>(z + row * C + c);
    const float4 b = *reinterpret_cast<const float4*>(v + c);
    s += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
  }
  s = warp_sum(s);
  if (lane == 0) out[row] = s;
}

// dz[i,:] = g[i] * v
__global__ void __launch_bounds__(256) rowdot_bwd_kernel(const float* __restrict__ g, const float* __restrict__ v, int64_t N,
                                                         int C4, float* __restrict__ dz) {
  const int64_t total = N * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i / C4];
    const float4 b = reinterpret_cast<const float4*>(v)[i % C4];
    reinterpret_cast<float4*>(dz)[i] = make_float4(gi * b.x, gi * b.y, gi * b.z, gi * b.w);
  }
}

// partial[b] = sum over the CTA's slice of softplus(-sp) + softplus(sn); fixed-order tree inside the CTA
__global__ void __launch_bounds__(256) softplus_pair_partial_kernel(const float* __restrict__ sp, const float* __restrict__ sn,
                                                                    int64_t N, float* __restrict__ partial) {
  __shared__ float red[8];
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
    s += softplusf(-sp[i]) + softplusf(sn[i]);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}
__global__ void sum_finish_kernel(const float* __restrict__ partial, int nb, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float t = 0.f;
    for (int b = 0; b < nb; ++b) t += partial[b];
    *out = t;
  }
}

// d/dsp = -sigmoid(-sp) * g ; d/dsn = sigmoid(sn) * g   (g: device scalar)
__global__ void __launch_bounds__(256) softplus_pair_bwd_kernel(const float* __restrict__ sp, const float* __restrict__ sn,
                                                                const float* __restrict__ g, int64_t N, float* __restrict__ dsp,
                                                                float* __restrict__ dsn) {
  const float gs = *g;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    dsp[i] = -sigmoidf(-sp[i]) * gs;
    dsn[i] = sigmoidf(sn[i]) * gs;
  }
}

__global__ void sigmoid_mean_kernel(const float* __restrict__ colsum, int C, float inv_n, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) out[c] = sigmoidf(colsum[c] * inv_n);
}

inline int pair_ctas(int64_t N) {
  int64_t b = ceil_div(N, 256 * 4);
  return (int)(b < 1 ? 1 : (b > kNumSMs * 2 ? kNumSMs * 2 : b));
}

}  // namespace bmkg

using namespace bmkg;

extern "C" {

int bmkg_rowdot(const float* z, const float* v, int64_t N, int C, float* out, void* stream) {
  BMKG_REQUIRE(z && v && out && N > 0 && C > 0 && C % 4 == 0, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(aligned16(z) && aligned16(v), BMKG_ERR_MISALIGNED);
  rowdot_kernel<<<(unsigned)ceil_div(N, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(z, v, N, C, out);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

int bmkg_rowdot_bwd(const float* g, const float* v, int64_t N, int C, float* dz, void* stream) {
  BMKG_REQUIRE(g && v && dz && N > 0 && C > 0 && C % 4 == 0, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(aligned16(dz) && aligned16(v), BMKG_ERR_MISALIGNED);
  int64_t grid = ceil_div(N * (C / 4), 256);
  if (grid > kNumSMs * 16) grid = kNumSMs * 16;
  rowdot_bwd_kernel<<<(unsigned)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(g, v, N, C / 4, dz);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

size_t bmkg_softplus_pair_workspace_bytes(int64_t N) { return (size_t)pair_ctas(N) * sizeof(float); }

int bmkg_softplus_pair_sum(const float* sp, const float* sn, int64_t N, float* out, void* ws, size_t ws_bytes, void* stream) {
  BMKG_REQUIRE(sp && sn && out && N > 0, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(ws && ws_bytes >= bmkg_softplus_pair_workspace_bytes(N), BMKG_ERR_WORKSPACE);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nb = pair_ctas(N);
  softplus_pair_partial_kernel<<<nb, 256, 0, st>>>(sp, sn, N, static_cast<float*>(ws));
  sum_finish_kernel<<<1, 32, 0, st>>>(static_cast<const float*>(ws), nb, out);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

int bmkg_softplus_pair_bwd(const float* sp, const float* sn, const float* gscale, int64_t N, float* dsp, float* dsn,
                           void* stream) {
  BMKG_REQUIRE(sp && sn && gscale && dsp && dsn && N > 0, BMKG_ERR_BAD_ARG);
  softplus_pair_bwd_kernel<<<pair_ctas(N), 256, 0, static_cast<cudaStream_t>(stream)>>>(sp, sn, gscale, N, dsp, dsn);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

int bmkg_colmean_sigmoid(const float* z, int64_t N, int C, float* summary, void* ws, size_t ws_bytes, void* stream) {
  // column sums (two-stage, deterministic) then sigmoid(mean); ws layout: [partials | colsum[C]]
  BMKG_REQUIRE(z && summary && N > 0 && C > 0, BMKG_ERR_BAD_ARG);
  const size_t need = bmkg_colsum_workspace_bytes(N, C) + (size_t)C * sizeof(float);
  BMKG_REQUIRE(ws && ws_bytes >= need, BMKG_ERR_WORKSPACE);
  float* colsum = reinterpret_cast<float*>(static_cast<char*>(ws) + bmkg_colsum_workspace_bytes(N, C));
  int rc = bmkg_colsum(z, nullptr, N, C, colsum, ws, bmkg_colsum_workspace_bytes(N, C), stream);
  if (rc != BMKG_OK) return rc;
  sigmoid_mean_kernel<<<(unsigned)ceil_div(C, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(colsum, C, 1.0f / (float)N,
                                                                                                  summary);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

}  // extern "C"
