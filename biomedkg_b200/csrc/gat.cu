// A3/A4 - GAT neighbour aggregation (extension).
//
// The reference's GCL path only uses GCNEncoder; BASELINE.json configs 2 and 5 name a GAT
// encoder, specified as PyG GATConv(in, out, heads=H, concat=True, negative_slope=0.2,
// add_self_loops=True) inside the GCNEncoder layer pattern (encoder.py:124-162).
//
//   e_ij  = leaky_relu(a_src[j] + a_dst[i])          a_src = <xh, att_src>, a_dst = <xh, att_dst>
//   alpha = softmax over the in-edges of i (CSR row)  (max-subtracted, / (sum + 1e-16) as PyG)
//   out_i = sum_j alpha_ij xh[j] + bias   (+ ReLU + dropout epilogue)
//
// One warp per destination row: a warp-level segmented softmax (lanes = edges for the
// logits, shuffle max / sum), then the same coalesced 128-bit row gather as the GCN kernel
// with the attention weight broadcast by shuffle.  Nothing per-edge is stored: the backward
// recomputes alpha from the node arrays (a_src, a_dst, row max, row sum), so no atomics and
// no edge permutation are needed:
//   csr pass (per dst i):  d a_dst[i], t_i = sum_k alpha_ik <g_i, xh_k>
//   csc pass (per src j):  d xh[j] = sum_i alpha_ij g_i + d a_src[j] att_src + d a_dst[j] att_dst,  d a_src[j]
// HBM-bound; algorithmic bytes per forward call: E'*(C*2 + 4 + 4H) + N*C*s_out + 4*(N+1).
#include "common.cuh"
#include "../../include/bmkg_b200.h"

namespace bmkg {

constexpr int kGatWarps = 8;

__device__ __forceinline__ float lrelu(float x, float slope) { return x > 0.f ? x : slope * x; }

// a_src[n,h], a_dst[n,h]: one warp per node, xh row read once
template <int H>
__global__ void __launch_bounds__(256) gat_scores_kernel(const __nv_bfloat16* __restrict__ xh, const float* __restrict__ att_src,
                                                         const float* __restrict__ att_dst, int64_t N, int C,
                                                         float* __restrict__ a_src, float* __restrict__ a_dst) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= N) return;
  const int HC = H * C;
  float s[H], d[H];
#pragma unroll
  for (int h = 0; h < H; ++h) { s[h] = 0.f; d[h] = 0.f; }
  for (int c0 = lane * 8; c0 < HC; c0 += 256) {
    float f[8];
    unpack8(ldg_stream(xh + n * HC + c0), f);
    const int h = c0 / C;
    float ps = 0.f, pd = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      ps = fmaf(f[i], att_src[c0 + i], ps);
      pd = fmaf(f[i], att_dst[c0 + i], pd);
    }
#pragma unroll
    for (int hh = 0; hh < H; ++hh)
      if (hh == h) { s[hh] += ps; d[hh] += pd; }
  }
#pragma unroll
  for (int h = 0; h < H; ++h) {
    s[h] = warp_sum(s[h]);
    d[h] = warp_sum(d[h]);
    if (lane == 0) { a_src[n * H + h] = s[h]; a_dst[n * H + h] = d[h]; }
  }
}

struct GatEpilogue {
  const float* bias;
  int relu;
  float drop_scale;
  uint32_t drop_threshold;
  uint64_t drop_seed;
  const uint8_t* drop_keep;
};

template <int H, int NV, bool OUT_F32>
__global__ void __launch_bounds__(kGatWarps * 32) gat_aggregate_kernel(const int32_t* __restrict__ rowptr,
                                                                       const int32_t* __restrict__ colind,
                                                                       const __nv_bfloat16* __restrict__ xh,
                                                                       const float* __restrict__ a_src,
                                                                       const float* __restrict__ a_dst, int64_t N, int C,
                                                                       float slope, GatEpilogue ep, void* __restrict__ out,
                                                                       float* __restrict__ rowmax, float* __restrict__ rowsum) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kGatWarps + (threadIdx.x >> 5);
  if (row >= N) return;
  const int HC = H * C;
  const int beg = rowptr[row], end = rowptr[row + 1];
  float ad[H], mx[H], sum[H];
#pragma unroll
  for (int h = 0; h < H; ++h) { ad[h] = a_dst[row * H + h]; mx[h] = -INFINITY; sum[h] = 0.f; }

  // sweep 1: segment max of the logits (lanes = edges)
  for (int k = beg + lane; k < end; k += 32) {
    const int c = colind[k];
#pragma unroll
    for (int h = 0; h < H; ++h) mx[h] = fmaxf(mx[h], lrelu(a_src[(int64_t)c * H + h] + ad[h], slope));
  }
#pragma unroll
  for (int h = 0; h < H; ++h) mx[h] = warp_max(mx[h]);

  int hl[NV];   // head of this lane's column octet
  bool act[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) { act[v] = (v * 256 + lane * 8) < HC; hl[v] = act[v] ? (v * 256 + lane * 8) / C : 0; }
  float acc[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[v][i] = 0.f;

  // sweep 2: p = exp(e - max), row sum, weighted gather
  for (int base = beg; base < end; base += 32) {
    const int k = base + lane;
    int c = 0;
    float p[H];
#pragma unroll
    for (int h = 0; h < H; ++h) p[h] = 0.f;
    if (k < end) {
      c = colind[k];
#pragma unroll
      for (int h = 0; h < H; ++h) {
        p[h] = __expf(lrelu(a_src[(int64_t)c * H + h] + ad[h], slope) - mx[h]);
        sum[h] += p[h];
      }
    }
    const int cnt = min(32, end - base);
    for (int j = 0; j < cnt; j += 4) {
      uint4 u[4][NV];
      float w[4][NV];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int sl = (j + t) & 31;
        const int cj = __shfl_sync(0xffffffffu, c, sl);
        float pj[H];
#pragma unroll
        for (int h = 0; h < H; ++h) pj[h] = __shfl_sync(0xffffffffu, p[h], sl);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          w[t][v] = pj[0];
#pragma unroll
          for (int h = 1; h < H; ++h)
            if (hl[v] == h) w[t][v] = pj[h];
        }
        if (j + t < cnt) {
          const __nv_bfloat16* rp = xh + (int64_t)cj * HC + lane * 8;
#pragma unroll
          for (int v = 0; v < NV; ++v)
            if (act[v]) u[t][v] = ldg_cached(rp + v * 256);
        }
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (j + t < cnt) {
#pragma unroll
          for (int v = 0; v < NV; ++v)
            if (act[v]) {
              float f[8];
              unpack8(u[t][v], f);
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[v][i] = fmaf(w[t][v], f[i], acc[v][i]);
            }
        }
      }
    }
  }
  float inv[H];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    sum[h] = warp_sum(sum[h]);
    inv[h] = 1.0f / (sum[h] + 1e-16f);
    if (lane == 0) { rowmax[row * H + h] = mx[h]; rowsum[row * H + h] = sum[h]; }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    if (!act[v]) continue;
    const int c0 = v * 256 + lane * 8;
    float iv = inv[0];
#pragma unroll
    for (int h = 1; h < H; ++h)
      if (hl[v] == h) iv = inv[h];
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = acc[v][i] * iv + (ep.bias ? ep.bias[c0 + i] : 0.f);
    if (ep.relu) {
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = fmaxf(r[i], 0.f);
    }
    if (ep.drop_keep) {
      const uint2 m = *reinterpret_cast<const uint2*>(ep.drop_keep + row * HC + c0);
      const uint32_t mm[2] = {m.x, m.y};
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = ((mm[i >> 2] >> (8 * (i & 3))) & 0xff) ? r[i] * ep.drop_scale : 0.f;
    } else if (ep.drop_threshold) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        r[i] = hash_keep(ep.drop_seed, (uint64_t)(row * HC + c0 + i), ep.drop_threshold) ? r[i] * ep.drop_scale : 0.f;
    }
    if (OUT_F32) {
      float* o = static_cast<float*>(out) + row * HC + c0;
      *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(r[4], r[5], r[6], r[7]);
    } else {
      *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(out) + row * HC + c0) = pack8(r);
    }
  }
}

// per-head dot of two rows held as octets across the warp
template <int H, int NV>
__device__ __forceinline__ void head_dots(const float (&a)[NV][8], const uint4 (&ub)[NV], const bool (&act)[NV], const int (&hl)[NV],
                                          float (&dot)[H]) {
#pragma unroll
  for (int h = 0; h < H; ++h) dot[h] = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    if (!act[v]) continue;
    float f[8];
    unpack8(ub[v], f);
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) part = fmaf(a[v][i], f[i], part);
#pragma unroll
    for (int h = 0; h < H; ++h)
      if (hl[v] == h) dot[h] += part;
  }
#pragma unroll
  for (int h = 0; h < H; ++h) dot[h] = warp_sum(dot[h]);
}

// csr pass: per destination i, d a_dst[i] and t_i
template <int H, int NV>
__global__ void __launch_bounds__(kGatWarps * 32) gat_bwd_dst_kernel(const int32_t* __restrict__ rowptr,
                                                                     const int32_t* __restrict__ colind,
                                                                     const __nv_bfloat16* __restrict__ xh,
                                                                     const __nv_bfloat16* __restrict__ g,
                                                                     const float* __restrict__ a_src, const float* __restrict__ a_dst,
                                                                     const float* __restrict__ rowmax, const float* __restrict__ rowsum,
                                                                     int64_t N, int C, float slope, float* __restrict__ d_adst,
                                                                     float* __restrict__ tsum) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kGatWarps + (threadIdx.x >> 5);
  if (row >= N) return;
  const int HC = H * C;
  int hl[NV];
  bool act[NV];
  float gi[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    act[v] = (v * 256 + lane * 8) < HC;
    hl[v] = act[v] ? (v * 256 + lane * 8) / C : 0;
    if (act[v]) unpack8(ldg_stream(g + row * HC + v * 256 + lane * 8), gi[v]);
  }
  float ad[H], mx[H], inv[H], A[H], B[H], T[H];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    ad[h] = a_dst[row * H + h];
    mx[h] = rowmax[row * H + h];
    inv[h] = 1.0f / (rowsum[row * H + h] + 1e-16f);
    A[h] = B[h] = T[h] = 0.f;
  }
  const int beg = rowptr[row], end = rowptr[row + 1];
  for (int base = beg; base < end; base += 32) {
    const int k = base + lane;
    const int c = (k < end) ? colind[k] : 0;
    const int cnt = min(32, end - base);
    for (int j = 0; j < cnt; ++j) {
      const int cj = __shfl_sync(0xffffffffu, c, j);
      uint4 u[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (act[v]) u[v] = ldg_cached(xh + (int64_t)cj * HC + v * 256 + lane * 8);
      float dot[H];
      head_dots<H, NV>(gi, u, act, hl, dot);
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float raw = a_src[(int64_t)cj * H + h] + ad[h];
        const float alpha = __expf(lrelu(raw, slope) - mx[h]) * inv[h];
        const float dl = raw > 0.f ? 1.f : slope;
        A[h] = fmaf(alpha * dot[h], dl, A[h]);
        B[h] = fmaf(alpha, dl, B[h]);
        T[h] = fmaf(alpha, dot[h], T[h]);
      }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int h = 0; h < H; ++h) {
      d_adst[row * H + h] = A[h] - T[h] * B[h];
      tsum[row * H + h] = T[h];
    }
  }
}

// csc pass: per source j, d xh[j] and d a_src[j]
template <int H, int NV>
__global__ void __launch_bounds__(kGatWarps * 32) gat_bwd_src_kernel(const int32_t* __restrict__ csc_rowptr,
                                                                     const int32_t* __restrict__ csc_colind,
                                                                     const __nv_bfloat16* __restrict__ xh,
                                                                     const __nv_bfloat16* __restrict__ g,
                                                                     const float* __restrict__ a_src, const float* __restrict__ a_dst,
                                                                     const float* __restrict__ rowmax, const float* __restrict__ rowsum,
                                                                     const float* __restrict__ tsum, const float* __restrict__ d_adst,
                                                                     const float* __restrict__ att_src, const float* __restrict__ att_dst,
                                                                     int64_t N, int C, float slope, __nv_bfloat16* __restrict__ dxh,
                                                                     float* __restrict__ d_asrc) {
  const int lane = threadIdx.x & 31;
  const int64_t src = (int64_t)blockIdx.x * kGatWarps + (threadIdx.x >> 5);
  if (src >= N) return;
  const int HC = H * C;
  int hl[NV];
  bool act[NV];
  float xj[NV][8], acc[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    act[v] = (v * 256 + lane * 8) < HC;
    hl[v] = act[v] ? (v * 256 + lane * 8) / C : 0;
    if (act[v]) unpack8(ldg_stream(xh + src * HC + v * 256 + lane * 8), xj[v]);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[v][i] = 0.f;
  }
  float as[H], das[H];
#pragma unroll
  for (int h = 0; h < H; ++h) { as[h] = a_src[src * H + h]; das[h] = 0.f; }
  const int beg = csc_rowptr[src], end = csc_rowptr[src + 1];
  for (int base = beg; base < end; base += 32) {
    const int k = base + lane;
    const int c = (k < end) ? csc_colind[k] : 0;
    const int cnt = min(32, end - base);
    for (int j = 0; j < cnt; ++j) {
      const int ci = __shfl_sync(0xffffffffu, c, j);  // destination i of edge src -> i
      uint4 u[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (act[v]) u[v] = ldg_cached(g + (int64_t)ci * HC + v * 256 + lane * 8);
      float dot[H];
      head_dots<H, NV>(xj, u, act, hl, dot);
      float alpha[H];
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float raw = as[h] + a_dst[(int64_t)ci * H + h];
        alpha[h] = __expf(lrelu(raw, slope) - rowmax[(int64_t)ci * H + h]) / (rowsum[(int64_t)ci * H + h] + 1e-16f);
        const float de = alpha[h] * (dot[h] - tsum[(int64_t)ci * H + h]);
        das[h] = fmaf(de, raw > 0.f ? 1.f : slope, das[h]);
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (!act[v]) continue;
        float al = alpha[0];
#pragma unroll
        for (int h = 1; h < H; ++h)
          if (hl[v] == h) al = alpha[h];
        float f[8];
        unpack8(u[v], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[v][i] = fmaf(al, f[i], acc[v][i]);
      }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int h = 0; h < H; ++h) d_asrc[src * H + h] = das[h];
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    if (!act[v]) continue;
    const int c0 = v * 256 + lane * 8;
    float ds = das[0], dd = d_adst[src * H + 0];
#pragma unroll
    for (int h = 1; h < H; ++h)
      if (hl[v] == h) { ds = das[h]; dd = d_adst[src * H + h]; }
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = acc[v][i] + ds * att_src[c0 + i] + dd * att_dst[c0 + i];
    *reinterpret_cast<uint4*>(dxh + src * HC + c0) = pack8(r);
  }
}

template <int H, int NV>
static int launch_gat_fwd(const int32_t* rowptr, const int32_t* colind, const __nv_bfloat16* xh, const float* as, const float* ad,
                          int64_t N, int C, float slope, const GatEpilogue& ep, void* out, int out_f32, float* rowmax,
                          float* rowsum, cudaStream_t st) {
  const unsigned grid = (unsigned)ceil_div(N, kGatWarps);
  if (out_f32)
    gat_aggregate_kernel<H, NV, true><<<grid, kGatWarps * 32, 0, st>>>(rowptr, colind, xh, as, ad, N, C, slope, ep, out, rowmax, rowsum);
  else
    gat_aggregate_kernel<H, NV, false><<<grid, kGatWarps * 32, 0, st>>>(rowptr, colind, xh, as, ad, N, C, slope, ep, out, rowmax, rowsum);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

template <int H, int NV>
static int launch_gat_bwd(const int32_t* rowptr, const int32_t* colind, const int32_t* crp, const int32_t* cci,
                          const __nv_bfloat16* xh, const __nv_bfloat16* g, const float* as, const float* ad, const float* rmax,
                          const float* rsum, const float* att_s, const float* att_d, int64_t N, int C, float slope,
                          __nv_bfloat16* dxh, float* das, float* dad, float* tsum, cudaStream_t st) {
  const unsigned grid = (unsigned)ceil_div(N, kGatWarps);
  gat_bwd_dst_kernel<H, NV><<<grid, kGatWarps * 32, 0, st>>>(rowptr, colind, xh, g, as, ad, rmax, rsum, N, C, slope, dad, tsum);
  gat_bwd_src_kernel<H, NV><<<grid, kGatWarps * 32, 0, st>>>(crp, cci, xh, g, as, ad, rmax, rsum, tsum, dad, att_s, att_d, N, C, slope,
                                                            dxh, das);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

}  // namespace bmkg

using namespace bmkg;

#define GAT_DISPATCH(H_, NV_, CALL)                                   \
  switch ((H_) * 10 + (NV_)) {                                        \
    case 11: { constexpr int H = 1, NV = 1; CALL; } break;            \
    case 12: { constexpr int H = 1, NV = 2; CALL; } break;            \
    case 14: { constexpr int H = 1, NV = 4; CALL; } break;            \
    case 21: { constexpr int H = 2, NV = 1; CALL; } break;            \
    case 22: { constexpr int H = 2, NV = 2; CALL; } break;            \
    case 24: { constexpr int H = 2, NV = 4; CALL; } break;            \
    case 41: { constexpr int H = 4, NV = 1; CALL; } break;            \
    case 42: { constexpr int H = 4, NV = 2; CALL; } break;            \
    case 44: { constexpr int H = 4, NV = 4; CALL; } break;            \
    default: return BMKG_ERR_UNSUPPORTED;                             \
  }

static int gat_shape_ok(int H, int C) {
  if (!(H == 1 || H == 2 || H == 4)) return 0;
  if (C <= 0 || C % 8 != 0 || H * C > 1024) return 0;
  return 1;
}
static int gat_nv(int HC) { return HC <= 256 ? 1 : (HC <= 512 ? 2 : 4); }

extern "C" {

int bmkg_gat_scores(const void* xh_bf16, const float* att_src, const float* att_dst, int64_t N, int H, int C, float* a_src,
                    float* a_dst, void* stream) {
  BMKG_REQUIRE(xh_bf16 && att_src && att_dst && a_src && a_dst && N > 0, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(gat_shape_ok(H, C), BMKG_ERR_UNSUPPORTED);
  BMKG_REQUIRE(aligned16(xh_bf16), BMKG_ERR_MISALIGNED);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(xh_bf16);
  const unsigned grid = (unsigned)ceil_div(N, 8);
  if (H == 1) gat_scores_kernel<1><<<grid, 256, 0, st>>>(x, att_src, att_dst, N, C, a_src, a_dst);
  else if (H == 2) gat_scores_kernel<2><<<grid, 256, 0, st>>>(x, att_src, att_dst, N, C, a_src, a_dst);
  else gat_scores_kernel<4><<<grid, 256, 0, st>>>(x, att_src, att_dst, N, C, a_src, a_dst);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

int bmkg_gat_aggregate(const int32_t* rowptr, const int32_t* colind, const void* xh_bf16, const float* a_src, const float* a_dst,
                       int64_t N, int H, int C, float negative_slope, const float* bias, int relu, float drop_p,
                       uint64_t drop_seed, const uint8_t* drop_keep, void* out, int out_is_fp32, float* rowmax, float* rowsum,
                       void* stream) {
  BMKG_REQUIRE(rowptr && colind && xh_bf16 && a_src && a_dst && out && rowmax && rowsum && N > 0, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(gat_shape_ok(H, C), BMKG_ERR_UNSUPPORTED);
  BMKG_REQUIRE(drop_p >= 0.f && drop_p < 1.f, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(aligned16(xh_bf16) && aligned16(out) && (!drop_keep || aligned16(drop_keep)), BMKG_ERR_MISALIGNED);
  GatEpilogue ep;
  ep.bias = bias;
  ep.relu = relu;
  ep.drop_keep = (drop_p > 0.f) ? drop_keep : nullptr;
  ep.drop_scale = (drop_p > 0.f) ? 1.0f / (1.0f - drop_p) : 1.0f;
  ep.drop_threshold = (drop_p > 0.f && !drop_keep) ? (uint32_t)((double)drop_p * 4294967296.0) : 0u;
  ep.drop_seed = drop_seed;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(xh_bf16);
  GAT_DISPATCH(H, gat_nv(H * C), return (launch_gat_fwd<H, NV>(rowptr, colind, x, a_src, a_dst, N, C, negative_slope, ep, out,
                                                               out_is_fp32, rowmax, rowsum, st)));
  return BMKG_OK;
}

int bmkg_gat_aggregate_bwd(const int32_t* rowptr, const int32_t* colind, const int32_t* csc_rowptr, const int32_t* csc_colind,
                           const void* xh_bf16, const void* g_bf16, const float* a_src, const float* a_dst, const float* rowmax,
                           const float* rowsum, const float* att_src, const float* att_dst, int64_t N, int H, int C,
                           float negative_slope, void* dxh_bf16, float* d_asrc, float* d_adst, float* tsum_ws, void* stream) {
  BMKG_REQUIRE(rowptr && colind && csc_rowptr && csc_colind && xh_bf16 && g_bf16 && a_src && a_dst && rowmax && rowsum && att_src &&
                   att_dst && dxh_bf16 && d_asrc && d_adst && tsum_ws && N > 0,
               BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(gat_shape_ok(H, C), BMKG_ERR_UNSUPPORTED);
  BMKG_REQUIRE(aligned16(xh_bf16) && aligned16(g_bf16) && aligned16(dxh_bf16), BMKG_ERR_MISALIGNED);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GAT_DISPATCH(H, gat_nv(H * C),
               return (launch_gat_bwd<H, NV>(rowptr, colind, csc_rowptr, csc_colind, static_cast<const __nv_bfloat16*>(xh_bf16),
                                             static_cast<const __nv_bfloat16*>(g_bf16), a_src, a_dst, rowmax, rowsum, att_src, att_dst,
                                             N, C, negative_slope, static_cast<__nv_bfloat16*>(dxh_bf16), d_asrc, d_adst, tsum_ws, st)));
  return BMKG_OK;
}

}  // extern "C"
