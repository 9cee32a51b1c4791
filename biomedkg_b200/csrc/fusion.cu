// F1 - modality-fusion attention core (biomedkg/utils/fusion.py:17-31).
//
// AttentionFusion = three Linear(768,768) projections (one fused [N*M,768]x[768,2304]
// GEMM on the host side), then single-head SDPA over the M modality tokens of each node
// and a mean over M.  This file is the part after the GEMM: per node an M x M softmax
// (M <= 4, held in registers) and
//      out = mean_i sum_j p_ij v_j = sum_j (mean_i p_ij) v_j.
// One warp per node; 128-bit bf16 loads of q/k/v; dot products by warp shuffle.
// HBM-bound on the qkv read: N*M*3E*2 bytes in, N*E*4 out.
#include "common.cuh"
#include "../../include/bmkg_b200.h"

namespace bmkg {

constexpr int kMaxOct = 4;  // octets (8 bf16) per lane per row: E <= 1024

__device__ __forceinline__ void add_bias8(float* f, const float* __restrict__ b) {
  if (!b) return;
  const float4 b0 = *reinterpret_cast<const float4*>(b), b1 = *reinterpret_cast<const float4*>(b + 4);
  f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
}

// qkv holds the bias-free projections x W^T; the [3E] bias (q|k|v) is added on load when given
template <int M>
__global__ void __launch_bounds__(256) fusion_attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ bias,
                                                              int64_t N, int E, float* __restrict__ out, float* __restrict__ probs) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= N) return;
  const int noct = E / 8;
  const float inv_sqrt = rsqrtf((float)E);
  float s[M][M];
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) s[i][j] = 0.f;

  for (int o = lane; o < noct; o += 32) {
    float q[M][8], k[M][8];
#pragma unroll
    for (int m = 0; m < M; ++m) {
      const __nv_bfloat16* rowp = qkv + ((n * M + m) * 3) * (int64_t)E;
      unpack8(ldg_stream(rowp + o * 8), q[m]);
      unpack8(ldg_stream(rowp + E + o * 8), k[m]);
      add_bias8(q[m], bias ? bias + o * 8 : nullptr);
      add_bias8(k[m], bias ? bias + E + o * 8 : nullptr);
    }
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int e = 0; e < 8; ++e) s[i][j] = fmaf(q[i][e], k[j][e], s[i][j]);
  }
  float wbar[M];
#pragma unroll
  for (int j = 0; j < M; ++j) wbar[j] = 0.f;
#pragma unroll
  for (int i = 0; i < M; ++i) {
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < M; ++j) {
      s[i][j] = warp_sum(s[i][j]) * inv_sqrt;
      mx = fmaxf(mx, s[i][j]);
    }
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < M; ++j) {
      s[i][j] = __expf(s[i][j] - mx);
      den += s[i][j];
    }
    const float inv = 1.0f / den;
#pragma unroll
    for (int j = 0; j < M; ++j) {
      s[i][j] *= inv;
      wbar[j] += s[i][j] * (1.0f / (float)M);
    }
  }
  if (lane < M * M) {
    float pv = 0.f;
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j)
        if (lane == i * M + j) pv = s[i][j];
    probs[n * (M * M) + lane] = pv;
  }
  for (int o = lane; o < noct; o += 32) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < M; ++j) {
      float v[8];
      unpack8(ldg_stream(qkv + ((n * M + j) * 3 + 2) * (int64_t)E + o * 8), v);
      add_bias8(v, bias ? bias + 2 * E + o * 8 : nullptr);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(wbar[j], v[e], acc[e]);
    }
    float* op = out + n * E + o * 8;
    *reinterpret_cast<float4*>(op) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(op + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// dv_j = wbar_j dout ; dp_ij = (dout . v_j)/M ; ds_ij = p_ij (dp_ij - sum_j' p_ij' dp_ij')
// dq_i = sum_j ds_ij k_j / sqrt(E) ; dk_j = sum_i ds_ij q_i / sqrt(E)
template <int M>
__global__ void __launch_bounds__(256) fusion_attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ bias,
                                                              const float* __restrict__ probs, const float* __restrict__ dout,
                                                              int64_t N, int E, __nv_bfloat16* __restrict__ dqkv) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= N) return;
  const int noct = E / 8;
  const float inv_sqrt = rsqrtf((float)E);
  float p[M][M], wbar[M], dv_dot[M];
#pragma unroll
  for (int j = 0; j < M; ++j) { wbar[j] = 0.f; dv_dot[j] = 0.f; }
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) {
      p[i][j] = probs[n * (M * M) + i * M + j];
      wbar[j] += p[i][j] * (1.0f / (float)M);
    }
  for (int o = lane; o < noct; o += 32) {
    const float4 g0 = *reinterpret_cast<const float4*>(dout + n * E + o * 8);
    const float4 g1 = *reinterpret_cast<const float4*>(dout + n * E + o * 8 + 4);
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int j = 0; j < M; ++j) {
      float v[8], dv[8];
      unpack8(ldg_stream(qkv + ((n * M + j) * 3 + 2) * (int64_t)E + o * 8), v);
      add_bias8(v, bias ? bias + 2 * E + o * 8 : nullptr);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        dv_dot[j] = fmaf(g[e], v[e], dv_dot[j]);
        dv[e] = wbar[j] * g[e];
      }
      *reinterpret_cast<uint4*>(dqkv + ((n * M + j) * 3 + 2) * (int64_t)E + o * 8) = pack8(dv);
    }
  }
  float ds[M][M];
#pragma unroll
  for (int j = 0; j < M; ++j) dv_dot[j] = warp_sum(dv_dot[j]) * (1.0f / (float)M);  // dp_ij (same for all i)
#pragma unroll
  for (int i = 0; i < M; ++i) {
    float dotp = 0.f;
#pragma unroll
    for (int j = 0; j < M; ++j) dotp += p[i][j] * dv_dot[j];
#pragma unroll
    for (int j = 0; j < M; ++j) ds[i][j] = p[i][j] * (dv_dot[j] - dotp) * inv_sqrt;
  }
  for (int o = lane; o < noct; o += 32) {
    float q[M][8], k[M][8];
#pragma unroll
    for (int m = 0; m < M; ++m) {
      const __nv_bfloat16* rowp = qkv + ((n * M + m) * 3) * (int64_t)E;
      unpack8(ldg_stream(rowp + o * 8), q[m]);
      unpack8(ldg_stream(rowp + E + o * 8), k[m]);
      add_bias8(q[m], bias ? bias + o * 8 : nullptr);
      add_bias8(k[m], bias ? bias + E + o * 8 : nullptr);
    }
#pragma unroll
    for (int m = 0; m < M; ++m) {
      float dq[8], dk[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { dq[e] = 0.f; dk[e] = 0.f; }
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          dq[e] = fmaf(ds[m][j], k[j][e], dq[e]);
          dk[e] = fmaf(ds[j][m], q[j][e], dk[e]);
        }
      __nv_bfloat16* rowp = dqkv + ((n * M + m) * 3) * (int64_t)E;
      *reinterpret_cast<uint4*>(rowp + o * 8) = pack8(dq);
      *reinterpret_cast<uint4*>(rowp + E + o * 8) = pack8(dk);
    }
  }
}

}  // namespace bmkg

using namespace bmkg;

extern "C" {

int bmkg_fusion_attn_fwd(const void* qkv_bf16, const float* qkv_bias, int64_t N, int M, int E, float* out, float* probs,
                         void* stream) {
  BMKG_REQUIRE(qkv_bf16 && out && probs && N > 0, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(M >= 1 && M <= 4 && E % 8 == 0 && E > 0, BMKG_ERR_UNSUPPORTED);
  BMKG_REQUIRE(aligned16(qkv_bf16) && aligned16(out) && (!qkv_bias || aligned16(qkv_bias)), BMKG_ERR_MISALIGNED);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(qkv_bf16);
  const unsigned grid = (unsigned)ceil_div(N, 8);
  switch (M) {
    case 1: fusion_attn_fwd_kernel<1><<<grid, 256, 0, st>>>(q, qkv_bias, N, E, out, probs); break;
    case 2: fusion_attn_fwd_kernel<2><<<grid, 256, 0, st>>>(q, qkv_bias, N, E, out, probs); break;
    case 3: fusion_attn_fwd_kernel<3><<<grid, 256, 0, st>>>(q, qkv_bias, N, E, out, probs); break;
    default: fusion_attn_fwd_kernel<4><<<grid, 256, 0, st>>>(q, qkv_bias, N, E, out, probs); break;
  }
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

int bmkg_fusion_attn_bwd(const void* qkv_bf16, const float* qkv_bias, const float* probs, const float* dout, int64_t N, int M,
                         int E, void* dqkv_bf16, void* stream) {
  BMKG_REQUIRE(qkv_bf16 && probs && dout && dqkv_bf16 && N > 0, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(M >= 1 && M <= 4 && E % 8 == 0 && E > 0, BMKG_ERR_UNSUPPORTED);
  BMKG_REQUIRE(aligned16(qkv_bf16) && aligned16(dout) && aligned16(dqkv_bf16), BMKG_ERR_MISALIGNED);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(qkv_bf16);
  __nv_bfloat16* dq = static_cast<__nv_bfloat16*>(dqkv_bf16);
  const unsigned grid = (unsigned)ceil_div(N, 8);
  switch (M) {
    case 1: fusion_attn_bwd_kernel<1><<<grid, 256, 0, st>>>(q, qkv_bias, probs, dout, N, E, dq); break;
    case 2: fusion_attn_bwd_kernel<2><<<grid, 256, 0, st>>>(q, qkv_bias, probs, dout, N, E, dq); break;
    case 3: fusion_attn_bwd_kernel<3><<<grid, 256, 0, st>>>(q, qkv_bias, probs, dout, N, E, dq); break;
    default: fusion_attn_bwd_kernel<4><<<grid, 256, 0, st>>>(q, qkv_bias, probs, dout, N, E, dq); break;
  }
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

int bmkg_abi_version(void) { return BMKG_ABI_VERSION; }

const char* bmkg_error_string(int code) {
  switch (code) {
    case BMKG_OK: return "ok";
    case BMKG_ERR_BAD_ARG: return "bad argument (null pointer, non-positive size or unsupported shape)";
    case BMKG_ERR_MISALIGNED: return "pointer not 16-byte aligned";
    case BMKG_ERR_WORKSPACE: return "workspace missing or too small";
    case BMKG_ERR_LAUNCH: return "CUDA launch failure";
    case BMKG_ERR_UNSUPPORTED: return "shape outside the compiled kernel range";
    case BMKG_ERR_DRIVER: return "CUDA driver entry point unavailable (cuTensorMapEncodeTiled)";
    default: return "unknown error";
  }
}

}  // extern "C"
