// Streaming elementwise / small-reduction kernels around the GCL step.
//   - mask_feature (model/gcl.py:40-41,75) fused with the fp32 -> bf16 cast of x
//   - ReLU+dropout backward (encoder.py:155-158) fused with the bias gradient
//   - modality mean (gcl_module.py:47-48)
//   - L2 row normalisation (PyGCL InfoNCE _similarity) + its backward
//   - deterministic column sums (two-stage, fixed order, no atomics)
// All HBM-bound, 128-bit vectorised, grid sized in multiples of the SM count.
#include "common.cuh"
#include "../../include/bmkg_b200.h"

namespace bmkg {

constexpr int kEwThreads = 256;
inline unsigned ew_grid(int64_t work_items) {
  int64_t g = ceil_div(work_items, kEwThreads);
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

// x fp32 [n] (n % 4 == 0) -> up to three bf16 copies: plain, masked by keep1, masked by keep2
__global__ void __launch_bounds__(kEwThreads) mask_cast_kernel(const float* __restrict__ x, const uint8_t* __restrict__ keep1,
                                                               const uint8_t* __restrict__ keep2, int64_t n4,
                                                               __nv_bfloat16* __restrict__ x0, __nv_bfloat16* __restrict__ x1,
                                                               __nv_bfloat16* __restrict__ x2) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    if (x0) reinterpret_cast<uint2*>(x0)[i] = make_uint2(pack2(v.x, v.y), pack2(v.z, v.w));
    if (x1) {
      const uint32_t m = reinterpret_cast<const uint32_t*>(keep1)[i];
      reinterpret_cast<uint2*>(x1)[i] = make_uint2(pack2((m & 0xffu) ? v.x : 0.f, (m & 0xff00u) ? v.y : 0.f),
                                                   pack2((m & 0xff0000u) ? v.z : 0.f, (m & 0xff000000u) ? v.w : 0.f));
    }
    if (x2) {
      const uint32_t m = reinterpret_cast<const uint32_t*>(keep2)[i];
      reinterpret_cast<uint2*>(x2)[i] = make_uint2(pack2((m & 0xffu) ? v.x : 0.f, (m & 0xff00u) ? v.y : 0.f),
                                                   pack2((m & 0xff0000u) ? v.z : 0.f, (m & 0xff000000u) ? v.w : 0.f));
    }
  }
}

// mean over the modality axis: x fp32 [N, M, F] -> fp32 / bf16 [N, F]
__global__ void __launch_bounds__(kEwThreads) modality_mean_kernel(const float* __restrict__ x, int64_t N, int M, int F4,
                                                                   float* __restrict__ out_f32,
                                                                   __nv_bfloat16* __restrict__ out_bf16) {
  const int64_t total = N * F4;
  const float inv = 1.0f / (float)M;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / F4, f = i % F4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int m = 0; m < M; ++m) {
      const float4 v = reinterpret_cast<const float4*>(x)[(n * M + m) * F4 + f];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
    if (out_f32) reinterpret_cast<float4*>(out_f32)[i] = a;
    if (out_bf16) reinterpret_cast<uint2*>(out_bf16)[i] = make_uint2(pack2(a.x, a.y), pack2(a.z, a.w));
  }
}

// g_pre = (y > 0) ? g_y * scale : 0   (y = dropout(relu(pre)) so y > 0 <=> kept and pre > 0)
// and per-CTA column partial sums of g_pre for the bias gradient.
// grid.x CTAs each own a contiguous row range; thread t owns column-octet t % (C/8).
__global__ void __launch_bounds__(kEwThreads) relu_dropout_bwd_kernel(const __nv_bfloat16* __restrict__ gy,
                                                                      const __nv_bfloat16* __restrict__ y, float scale,
                                                                      int64_t N, int C, int rows_per_cta,
                                                                      __nv_bfloat16* __restrict__ gpre,
                                                                      float* __restrict__ partial /*[grid, C]*/) {
  extern __shared__ float red[];  // [groups][C]
  const int oct = C / 8;
  const int groups = kEwThreads / oct;  // row groups per CTA iteration (host guarantees >= 1)
  const int g = threadIdx.x / oct, o = threadIdx.x % oct;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = min(N, r0 + rows_per_cta);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (g < groups) {
    for (int64_t r = r0 + g; r < r1; r += groups) {
      const uint4 ug = ldg_stream(gy + r * C + o * 8);
      const uint4 uy = ldg_stream(y + r * C + o * 8);
      float fg[8], fy[8], res[8];
      unpack8(ug, fg);
      unpack8(uy, fy);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        res[i] = fy[i] > 0.f ? fg[i] * scale : 0.f;
      }
      const uint4 packed = pack8(res);
      *reinterpret_cast<uint4*>(gpre + r * C + o * 8) = packed;
      float rb[8];
      unpack8(packed, rb);  // sum what downstream actually sees (bf16-rounded)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += rb[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[g * C + o * 8 + i] = acc[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kEwThreads) {
    float s = 0.f;
    for (int gg = 0; gg < groups; ++gg) s += red[gg * C + c];
    partial[(int64_t)blockIdx.x * C + c] = s;
  }
}

// out[c] = sum_b partial[b, c]: 32 columns x 32 interleaved row groups per CTA (1024 threads: the kernel is a chain of
// dependent L2 loads, so more groups = a shorter chain), fixed-order combine (deterministic)
constexpr int kFinishThreads = 1024;
__global__ void __launch_bounds__(kFinishThreads) colsum_finish_kernel(const float* __restrict__ partial, int nb, int stride, int C, float* __restrict__ out) {
  __shared__ float red[32][33];
  const int cl = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;  // 4 independent chains keep 4 loads in flight; combined in fixed order
  if (c < C) {
    int b = g;
    for (; b + 96 < nb; b += 128) {
      s0 += partial[(int64_t)b * stride + c];
      s1 += partial[(int64_t)(b + 32) * stride + c];
      s2 += partial[(int64_t)(b + 64) * stride + c];
      s3 += partial[(int64_t)(b + 96) * stride + c];
    }
    for (; b < nb; b += 32) s0 += partial[(int64_t)b * stride + c];
  }
  red[g][cl] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (g == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += red[k][cl];
    out[c] = t;
  }
}

// column partial sums of a fp32 [N, C] matrix (optionally row-weighted), same CTA/row layout.
// blockIdx.y selects a 1024-column slab so any C is covered.
__global__ void __launch_bounds__(kEwThreads) colsum_partial_kernel(const float* __restrict__ z, const float* __restrict__ roww,
                                                                    int64_t N, int C, int rows_per_cta,
                                                                    float* __restrict__ partial) {
  __shared__ float red[1024];  // groups * cw <= 256 * 4
  const int col0 = blockIdx.y * 1024;
  const int cw = min(1024, C - col0);
  const int quad = cw / 4;
  const int groups = kEwThreads / quad;
  const int g = threadIdx.x / quad, o = threadIdx.x % quad;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = min(N, r0 + rows_per_cta);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (g < groups) {
    for (int64_t r = r0 + g; r < r1; r += groups) {
      const float4 v = reinterpret_cast<const float4*>(z + r * C + col0)[o];
      const float w = roww ? roww[r] : 1.f;
      acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
    }
    reinterpret_cast<float4*>(red + g * cw)[o] = acc;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cw; c += kEwThreads) {
    float s = 0.f;
    for (int gg = 0; gg < groups; ++gg) s += red[gg * cw + c];
    partial[(int64_t)blockIdx.x * C + col0 + c] = s;
  }
}

// partial[b, 0:C] = sum_r w1[r,h(c)] x[r,c], partial[b, C:2C] = same with w2 (w == null -> 1); x bf16 [N, C], C = H*Ch
__global__ void __launch_bounds__(kEwThreads) colsum_bf16_partial_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld,
                                                                         const float* __restrict__ w1, const float* __restrict__ w2,
                                                                         int64_t N, int C, int H, int rows_per_cta,
                                                                         float* __restrict__ partial) {
  __shared__ float red[2][2048];  // groups * C <= 256 * 8
  const int oct = C / 8;
  const int groups = kEwThreads / oct;
  const int g = threadIdx.x / oct, o = threadIdx.x % oct;
  const int h = (o * 8) / (C / H);
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = min(N, r0 + rows_per_cta);
  float a1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, a2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (g < groups) {
    for (int64_t r = r0 + g; r < r1; r += groups) {
      float f[8];
      unpack8(ldg_stream(x + r * ld + o * 8), f);
      const float u1 = w1 ? w1[r * H + h] : 1.f;
      const float u2 = w2 ? w2[r * H + h] : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { a1[i] = fmaf(u1, f[i], a1[i]); a2[i] = fmaf(u2, f[i], a2[i]); }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { red[0][g * C + o * 8 + i] = a1[i]; red[1][g * C + o * 8 + i] = a2[i]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += kEwThreads) {
    const int which = c / C, cc = c % C;
    float t = 0.f;
    for (int gg = 0; gg < groups; ++gg) t += red[which][gg * C + cc];
    partial[(int64_t)blockIdx.x * 2 * C + c] = t;
  }
}

// dx = g0 + keep1 * g1 + keep2 * g2   (backward of mask_cast; any g may be null), bf16 grads -> fp32
__global__ void __launch_bounds__(kEwThreads) mask_cast_bwd_kernel(const __nv_bfloat16* __restrict__ g0, const __nv_bfloat16* __restrict__ g1,
                                                                   const __nv_bfloat16* __restrict__ g2, const uint8_t* __restrict__ keep1,
                                                                   const uint8_t* __restrict__ keep2, int64_t n4, float* __restrict__ dx) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    if (g0) {
      const uint2 u = reinterpret_cast<const uint2*>(g0)[i];
      a[0] += __uint_as_float(u.x << 16); a[1] += __uint_as_float(u.x & 0xffff0000u);
      a[2] += __uint_as_float(u.y << 16); a[3] += __uint_as_float(u.y & 0xffff0000u);
    }
    if (g1) {
      const uint2 u = reinterpret_cast<const uint2*>(g1)[i];
      const uint32_t m = reinterpret_cast<const uint32_t*>(keep1)[i];
      if (m & 0xffu) a[0] += __uint_as_float(u.x << 16);
      if (m & 0xff00u) a[1] += __uint_as_float(u.x & 0xffff0000u);
      if (m & 0xff0000u) a[2] += __uint_as_float(u.y << 16);
      if (m & 0xff000000u) a[3] += __uint_as_float(u.y & 0xffff0000u);
    }
    if (g2) {
      const uint2 u = reinterpret_cast<const uint2*>(g2)[i];
      const uint32_t m = reinterpret_cast<const uint32_t*>(keep2)[i];
      if (m & 0xffu) a[0] += __uint_as_float(u.x << 16);
      if (m & 0xff00u) a[1] += __uint_as_float(u.x & 0xffff0000u);
      if (m & 0xff0000u) a[2] += __uint_as_float(u.y << 16);
      if (m & 0xff000000u) a[3] += __uint_as_float(u.y & 0xffff0000u);
    }
    reinterpret_cast<float4*>(dx)[i] = make_float4(a[0], a[1], a[2], a[3]);
  }
}

// out = bf16(x - m): the deviation operand of a centred GEMM  x W^T = (x - 1 m^T) W^T + 1 (W m)^T  (projector, model/gcl.py:49-51)
__global__ void __launch_bounds__(kEwThreads) center_cast_kernel(const float* __restrict__ x, const float* __restrict__ m, int64_t n4,
                                                                 int C4, __nv_bfloat16* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const float4 c = reinterpret_cast<const float4*>(m)[i % C4];
    reinterpret_cast<uint2*>(out)[i] = make_uint2(pack2(v.x - c.x, v.y - c.y), pack2(v.z - c.z, v.w - c.w));
  }
}

// Centred InfoNCE operand (numerics: DESIGN.md "Centred bf16 operands").  At initialisation, and whenever the encoder
// over-smooths, the normalised rows z_u are nearly identical; the InfoNCE gradient then lives in deviations of relative
// size 1e-3 that a bf16 z (2^-9) cannot carry.  Z is therefore stored as a common fp32 vector mu plus bf16 deviations
// d_u = z_u - mu, and the kernels use  z_u . z_v = d_u . d_v + a_u + a_v + |mu|^2  with a_u = mu . d_u in fp32.
//
// pass 1: inv_norm[u] = 1 / max(|h_u|, eps) and per-CTA column partial sums of h_u * inv_norm[u]
__global__ void __launch_bounds__(256) l2norm_colsum_kernel(const float* __restrict__ h, int64_t N, int D, int rows_per_cta,
                                                            float* __restrict__ inv_norm, float* __restrict__ partial /*[grid, D]*/) {
  extern __shared__ float red[];  // [8][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta, r1 = min(N, r0 + rows_per_cta);
  float acc[8][4];  // D <= 1024: lane owns columns lane*4 + 128*k
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f;
  for (int64_t row = r0 + warp; row < r1; row += 8) {
    const float* hp = h + row * D;
    float4 v[8];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = lane * 4 + 128 * k;
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < D) v[k] = *reinterpret_cast<const float4*>(hp + c);
      ss += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
    }
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    if (lane == 0) inv_norm[row] = inv;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      acc[k][0] = fmaf(v[k].x, inv, acc[k][0]);
      acc[k][1] = fmaf(v[k].y, inv, acc[k][1]);
      acc[k][2] = fmaf(v[k].z, inv, acc[k][2]);
      acc[k][3] = fmaf(v[k].w, inv, acc[k][3]);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = lane * 4 + 128 * k;
    if (c < D) *reinterpret_cast<float4*>(red + warp * D + c) = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += 256) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w * D + c];   // fixed warp order: deterministic
    partial[(int64_t)blockIdx.x * D + c] = s;
  }
}

// pass 2: d_u = bf16(h_u * inv_norm[u] * scale - mu), a_u = mu . d_u (of the ROUNDED deviation, fp32)
__global__ void __launch_bounds__(256) center_scale_kernel(const float* __restrict__ h, const float* __restrict__ inv_norm,
                                                           const float* __restrict__ mu, int64_t N, int D, float scale,
                                                           __nv_bfloat16* __restrict__ z, float* __restrict__ a) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= N) return;
  const float s = inv_norm[row] * scale;
  const float* hp = h + row * D;
  float dot = 0.f;
  for (int c = lane * 4; c < D; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(hp + c);
    const float4 m = *reinterpret_cast<const float4*>(mu + c);
    const uint32_t lo = pack2(fmaf(v.x, s, -m.x), fmaf(v.y, s, -m.y));
    const uint32_t hi = pack2(fmaf(v.z, s, -m.z), fmaf(v.w, s, -m.w));
    *reinterpret_cast<uint2*>(z + row * D + c) = make_uint2(lo, hi);
    dot = fmaf(m.x, __uint_as_float(lo << 16), dot);
    dot = fmaf(m.y, __uint_as_float(lo & 0xffff0000u), dot);
    dot = fmaf(m.z, __uint_as_float(hi << 16), dot);
    dot = fmaf(m.w, __uint_as_float(hi & 0xffff0000u), dot);
  }
  dot = warp_sum(dot);
  if (lane == 0) a[row] = dot;
}

// dh = scale * inv * (dz - u (u . dz)),  u = h * inv   (straight-through the bf16 rounding)
__global__ void __launch_bounds__(256) l2norm_scale_bwd_kernel(const float* __restrict__ h, const float* __restrict__ inv_norm,
                                                               const float* __restrict__ dz, int64_t N, int D, float scale,
                                                               float* __restrict__ dh) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= N) return;
  const float inv = inv_norm[row];
  const float* hp = h + row * D;
  const float* gp = dz + row * D;
  float dot = 0.f;
  for (int c = lane * 4; c < D; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(hp + c);
    const float4 g = *reinterpret_cast<const float4*>(gp + c);
    dot += v.x * g.x + v.y * g.y + v.z * g.z + v.w * g.w;
  }
  dot = warp_sum(dot) * inv;  // u . dz
  const float s = scale * inv;
  for (int c = lane * 4; c < D; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(hp + c);
    const float4 g = *reinterpret_cast<const float4*>(gp + c);
    float4 o;
    o.x = s * (g.x - v.x * inv * dot);
    o.y = s * (g.y - v.y * inv * dot);
    o.z = s * (g.z - v.z * inv * dot);
    o.w = s * (g.w - v.w * inv * dot);
    *reinterpret_cast<float4*>(dh + row * D + c) = o;
  }
}

inline int colsum_ctas(int64_t N) {
  int64_t b = ceil_div(N, 64);
  const int64_t cap = kNumSMs * 4;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace bmkg

using namespace bmkg;

extern "C" {

int bmkg_bind_device(int device) {
  // The library links its own static cudart: bind that runtime (and the calling thread's driver context, which
  // cuTensorMapEncodeTiled needs) to the caller's device.  Call once per host thread before the first launch.
  return cudaSetDevice(device) == cudaSuccess ? BMKG_OK : BMKG_ERR_BAD_ARG;
}

int bmkg_mask_cast(const float* x, const uint8_t* keep1, const uint8_t* keep2, int64_t n, void* x0, void* x1, void* x2,
                   void* stream) {
  BMKG_REQUIRE(x && n >= 0 && n % 4 == 0, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE((!x1 || keep1) && (!x2 || keep2), BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(aligned16(x), BMKG_ERR_MISALIGNED);
  if (n == 0) return BMKG_OK;
  mask_cast_kernel<<<ew_grid(n / 4), kEwThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      x, keep1, keep2, n / 4, static_cast<__nv_bfloat16*>(x0), static_cast<__nv_bfloat16*>(x1),
      static_cast<__nv_bfloat16*>(x2));
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

int bmkg_mask_cast_bwd(const void* g0, const void* g1, const void* g2, const uint8_t* keep1, const uint8_t* keep2, int64_t n,
                       float* dx, void* stream) {
  BMKG_REQUIRE(dx && n >= 0 && n % 4 == 0, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE((!g1 || keep1) && (!g2 || keep2), BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(aligned16(dx), BMKG_ERR_MISALIGNED);
  if (n == 0) return BMKG_OK;
  mask_cast_bwd_kernel<<<ew_grid(n / 4), kEwThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(g0), static_cast<const __nv_bfloat16*>(g1), static_cast<const __nv_bfloat16*>(g2), keep1,
      keep2, n / 4, dx);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

int bmkg_colsum_bf16(const void* x_bf16, const float* w1, const float* w2, int64_t N, int C, int H, float* out1, float* out2,
                     void* ws, size_t ws_bytes, void* stream) {
  // out1[c] = sum_r w1[r, head(c)] x[r, c] (w1 null -> plain column sum); out2 likewise with w2 (null -> skipped)
  BMKG_REQUIRE(x_bf16 && out1 && N > 0 && H >= 1 && C % 8 == 0 && C >= 8 && C % H == 0 && (C / H) % 8 == 0, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(C <= 2048 || (H == 1 && !w2), BMKG_ERR_UNSUPPORTED);
  BMKG_REQUIRE(!w2 || out2, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(ws && ws_bytes >= 2 * bmkg_colsum_workspace_bytes(N, C), BMKG_ERR_WORKSPACE);
  BMKG_REQUIRE(aligned16(x_bf16), BMKG_ERR_MISALIGNED);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nb = colsum_ctas(N);
  const int rows_per_cta = (int)ceil_div(N, nb);
  float* partial = static_cast<float*>(ws);
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(x_bf16);
  for (int c0 = 0; c0 < C; c0 += 2048) {  // 2048-column slabs (one launch when C <= 2048)
    const int cw = (C - c0) < 2048 ? (C - c0) : 2048;
    colsum_bf16_partial_kernel<<<nb, kEwThreads, 0, st>>>(x + c0, C, w1, w2, N, cw, H, rows_per_cta, partial);
    colsum_finish_kernel<<<(unsigned)ceil_div(cw, 32), kFinishThreads, 0, st>>>(partial, nb, 2 * cw, cw, out1 + c0);
    if (out2) colsum_finish_kernel<<<(unsigned)ceil_div(cw, 32), kFinishThreads, 0, st>>>(partial + cw, nb, 2 * cw, cw, out2 + c0);
  }
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

int bmkg_modality_mean(const float* x, int64_t N, int M, int F, float* out_f32, void* out_bf16, void* stream) {
  BMKG_REQUIRE(x && N > 0 && M > 0 && F > 0 && F % 4 == 0 && (out_f32 || out_bf16), BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(aligned16(x), BMKG_ERR_MISALIGNED);
  modality_mean_kernel<<<ew_grid(N * (F / 4)), kEwThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      x, N, M, F / 4, out_f32, static_cast<__nv_bfloat16*>(out_bf16));
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

size_t bmkg_colsum_workspace_bytes(int64_t N, int C) { return (size_t)colsum_ctas(N) * C * sizeof(float); }

int bmkg_relu_dropout_bwd(const void* gy_bf16, const void* y_bf16, float scale, int64_t N, int C, void* gpre_bf16, float* dbias,
                          void* ws, size_t ws_bytes, void* stream) {
  BMKG_REQUIRE(gy_bf16 && y_bf16 && gpre_bf16 && dbias && N > 0, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(C % 8 == 0 && C >= 8 && C / 8 <= kEwThreads, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(ws && ws_bytes >= bmkg_colsum_workspace_bytes(N, C), BMKG_ERR_WORKSPACE);
  BMKG_REQUIRE(aligned16(gy_bf16) && aligned16(y_bf16) && aligned16(gpre_bf16), BMKG_ERR_MISALIGNED);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nb = colsum_ctas(N);
  const int rows_per_cta = (int)ceil_div(N, nb);
  const int groups = kEwThreads / (C / 8);
  relu_dropout_bwd_kernel<<<nb, kEwThreads, (size_t)groups * C * sizeof(float), st>>>(
      static_cast<const __nv_bfloat16*>(gy_bf16), static_cast<const __nv_bfloat16*>(y_bf16), scale, N, C, rows_per_cta,
      static_cast<__nv_bfloat16*>(gpre_bf16), static_cast<float*>(ws));
  colsum_finish_kernel<<<(unsigned)ceil_div(C, 32), kFinishThreads, 0, st>>>(static_cast<const float*>(ws), nb, C, C, dbias);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

int bmkg_colsum(const float* z, const float* row_weight, int64_t N, int C, float* out, void* ws, size_t ws_bytes, void* stream) {
  BMKG_REQUIRE(z && out && N > 0 && C % 4 == 0 && C >= 4, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(ws && ws_bytes >= bmkg_colsum_workspace_bytes(N, C), BMKG_ERR_WORKSPACE);
  BMKG_REQUIRE(aligned16(z), BMKG_ERR_MISALIGNED);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nb = colsum_ctas(N);
  const int rows_per_cta = (int)ceil_div(N, nb);
  colsum_partial_kernel<<<dim3(nb, (unsigned)ceil_div(C, 1024)), kEwThreads, 0, st>>>(z, row_weight, N, C, rows_per_cta,
                                                                                       static_cast<float*>(ws));
  colsum_finish_kernel<<<(unsigned)ceil_div(C, 32), kFinishThreads, 0, st>>>(static_cast<const float*>(ws), nb, C, C, out);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

int bmkg_center_cast(const float* x, const float* m, int64_t N, int C, void* out_bf16, void* stream) {
  BMKG_REQUIRE(x && m && out_bf16 && N > 0 && C > 0 && C % 4 == 0, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(aligned16(x) && aligned16(m) && aligned16(out_bf16), BMKG_ERR_MISALIGNED);
  center_cast_kernel<<<ew_grid(N * (C / 4)), kEwThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      x, m, N * (C / 4), C / 4, static_cast<__nv_bfloat16*>(out_bf16));
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

int bmkg_l2norm_colsum(const float* h, int64_t N, int D, float* inv_norm, float* colsum, void* ws, size_t ws_bytes, void* stream) {
  BMKG_REQUIRE(h && inv_norm && colsum && N > 0 && D > 0 && D % 4 == 0 && D <= 1024, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(ws && ws_bytes >= bmkg_colsum_workspace_bytes(N, D), BMKG_ERR_WORKSPACE);
  BMKG_REQUIRE(aligned16(h), BMKG_ERR_MISALIGNED);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nb = colsum_ctas(N);
  const int rows_per_cta = (int)ceil_div(N, nb);
  l2norm_colsum_kernel<<<nb, 256, 8 * D * sizeof(float), st>>>(h, N, D, rows_per_cta, inv_norm, static_cast<float*>(ws));
  colsum_finish_kernel<<<(unsigned)ceil_div(D, 32), kFinishThreads, 0, st>>>(static_cast<const float*>(ws), nb, D, D, colsum);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

int bmkg_center_scale(const float* h, const float* inv_norm, const float* mu, int64_t N, int D, float scale, void* z_bf16,
                      float* a, void* stream) {
  BMKG_REQUIRE(h && inv_norm && mu && z_bf16 && a && N > 0 && D > 0 && D % 4 == 0, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(aligned16(h) && aligned16(z_bf16) && aligned16(mu), BMKG_ERR_MISALIGNED);
  center_scale_kernel<<<(unsigned)ceil_div(N, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      h, inv_norm, mu, N, D, scale, static_cast<__nv_bfloat16*>(z_bf16), a);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

int bmkg_l2norm_scale_bwd(const float* h, const float* inv_norm, const float* dz, int64_t N, int D, float scale, float* dh,
                          void* stream) {
  BMKG_REQUIRE(h && inv_norm && dz && dh && N > 0 && D > 0 && D % 4 == 0, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(aligned16(h) && aligned16(dz) && aligned16(dh), BMKG_ERR_MISALIGNED);
  l2norm_scale_bwd_kernel<<<(unsigned)ceil_div(N, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(h, inv_norm, dz, N, D,
                                                                                                  scale, dh);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

}  // extern "C"
