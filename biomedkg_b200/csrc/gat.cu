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

// Hub rows (power-law graphs, BASELINE cfg 5) use the same fixed chunking of the CSR edge array as the GCN kernel
// (common.cuh: kHubThreshold / kHubSeg): one CTA per chunk pre-reduces the hub rows it meets - for the softmax that is a
// (max, sum, weighted row) triple per (chunk, slot), merged in chunk order with the usual exp(max_c - max) rescale.
// softmax statistics and weighted row sum of the edges [beg, end) of one destination row (whole warp)
template <int H, int NV>
__device__ __forceinline__ void gat_fwd_segment(const int32_t* __restrict__ colind, const __nv_bfloat16* __restrict__ xh,
                                                const float* __restrict__ a_src, const float (&ad)[H], int HC, float slope, int beg,
                                                int end, int lane, const bool (&act)[NV], const int (&hl)[NV], float (&mx)[H],
                                                float (&sum)[H], float (&acc)[NV][8]) {
#pragma unroll
  for (int h = 0; h < H; ++h) { mx[h] = -INFINITY; sum[h] = 0.f; }
  // sweep 1: segment max of the logits (lanes = edges)
  for (int k = beg + lane; k < end; k += 32) {
    const int c = colind[k];
#pragma unroll
    for (int h = 0; h < H; ++h) mx[h] = fmaxf(mx[h], lrelu(a_src[(int64_t)c * H + h] + ad[h], slope));
  }
#pragma unroll
  for (int h = 0; h < H; ++h) mx[h] = warp_max(mx[h]);
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
#pragma unroll
  for (int h = 0; h < H; ++h) sum[h] = warp_sum(sum[h]);
}

// forward pre-pass: per (chunk, slot) softmax partial of the hub rows.  pstat[(chunk*2+slot)*2H + h] = max, [.. + H + h] = sum
template <int H, int NV>
__global__ void __launch_bounds__(kGatWarps * 32) gat_hub_fwd_kernel(const int32_t* __restrict__ rowptr,
                                                                     const int32_t* __restrict__ colind,
                                                                     const __nv_bfloat16* __restrict__ xh,
                                                                     const float* __restrict__ a_src, const float* __restrict__ a_dst,
                                                                     int64_t N, int C, float slope, const int32_t* __restrict__ hub_rows,
                                                                     float* __restrict__ pacc, float* __restrict__ pstat) {
  __shared__ int s_rows[2];
  __shared__ float s_m[kGatWarps][H], s_s[kGatWarps][H];
  extern __shared__ float red[];  // [kGatWarps][HC]
  int cs, ce;
  if (!hub_chunk_rows(rowptr, N, hub_rows, s_rows, cs, ce)) return;
  const int HC = H * C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int hl[NV];
  bool act[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) { act[v] = (v * 256 + lane * 8) < HC; hl[v] = act[v] ? (v * 256 + lane * 8) / C : 0; }
  for (int slot = 0; slot < 2; ++slot) {
    const int r = s_rows[slot];
    if (r < 0) continue;
    const int sb = max(rowptr[r], cs), se = min(rowptr[r + 1], ce);
    const int per = (se - sb + kGatWarps - 1) / kGatWarps;
    const int wb = min(se, sb + warp * per), we = min(se, wb + per);
    float ad[H], mx[H], sum[H], acc[NV][8];
#pragma unroll
    for (int h = 0; h < H; ++h) ad[h] = a_dst[(int64_t)r * H + h];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[v][i] = 0.f;
    gat_fwd_segment<H, NV>(colind, xh, a_src, ad, HC, slope, wb, we, lane, act, hl, mx, sum, acc);
    if (lane == 0) {
#pragma unroll
      for (int h = 0; h < H; ++h) { s_m[warp][h] = mx[h]; s_s[warp][h] = sum[h]; }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v)
      if (act[v]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) red[warp * HC + v * 256 + lane * 8 + i] = acc[v][i];
      }
    __syncthreads();
    for (int c = threadIdx.x; c < HC; c += kGatWarps * 32) {
      const int h = c / C;
      float M = -INFINITY;
#pragma unroll
      for (int w = 0; w < kGatWarps; ++w) M = fmaxf(M, s_m[w][h]);
      float t = 0.f, ssum = 0.f;
#pragma unroll
      for (int w = 0; w < kGatWarps; ++w) {
        const float sc = (s_m[w][h] == -INFINITY) ? 0.f : __expf(s_m[w][h] - M);
        t = fmaf(red[w * HC + c], sc, t);
        ssum = fmaf(s_s[w][h], sc, ssum);
      }
      pacc[((int64_t)blockIdx.x * 2 + slot) * HC + c] = t;
      if (c % C == 0) {
        pstat[((int64_t)blockIdx.x * 2 + slot) * 2 * H + h] = M;
        pstat[((int64_t)blockIdx.x * 2 + slot) * 2 * H + H + h] = ssum;
      }
    }
    __syncthreads();
  }
}

template <int H, int NV, bool OUT_F32>
__global__ void __launch_bounds__(kGatWarps * 32) gat_aggregate_kernel(const int32_t* __restrict__ rowptr,
                                                                       const int32_t* __restrict__ colind,
                                                                       const __nv_bfloat16* __restrict__ xh,
                                                                       const float* __restrict__ a_src,
                                                                       const float* __restrict__ a_dst, int64_t N, int C,
                                                                       float slope, GatEpilogue ep, const float* __restrict__ pacc,
                                                                       const float* __restrict__ pstat, void* __restrict__ out,
                                                                       float* __restrict__ rowmax, float* __restrict__ rowsum) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kGatWarps + (threadIdx.x >> 5);
  if (row >= N) return;
  const int HC = H * C;
  const int beg = rowptr[row], end = rowptr[row + 1];
  float ad[H], mx[H], sum[H];
#pragma unroll
  for (int h = 0; h < H; ++h) ad[h] = a_dst[row * H + h];
  int hl[NV];   // head of this lane's column octet
  bool act[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) { act[v] = (v * 256 + lane * 8) < HC; hl[v] = act[v] ? (v * 256 + lane * 8) / C : 0; }
  float acc[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[v][i] = 0.f;

  if (pacc != nullptr && end - beg > kHubThreshold) {
    // hub row: merge the per-chunk softmax partials in chunk order
    const int c_first = beg / kHubSeg, c_last = (end - 1) / kHubSeg;
#pragma unroll
    for (int h = 0; h < H; ++h) { mx[h] = -INFINITY; sum[h] = 0.f; }
    for (int c = c_first; c <= c_last; ++c) {
      const float* st = pstat + ((int64_t)c * 2 + (c == c_first ? 1 : 0)) * 2 * H;
#pragma unroll
      for (int h = 0; h < H; ++h) mx[h] = fmaxf(mx[h], st[h]);
    }
    for (int c = c_first; c <= c_last; ++c) {
      const int64_t rec = (int64_t)c * 2 + (c == c_first ? 1 : 0);
      const float* st = pstat + rec * 2 * H;
      float sc[H];
#pragma unroll
      for (int h = 0; h < H; ++h) {
        sc[h] = __expf(st[h] - mx[h]);
        sum[h] = fmaf(st[H + h], sc[h], sum[h]);
      }
      const float* pp = pacc + rec * HC;
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (act[v]) {
          float s1 = sc[0];
#pragma unroll
          for (int h = 1; h < H; ++h)
            if (hl[v] == h) s1 = sc[h];
          const float4 p0 = *reinterpret_cast<const float4*>(pp + v * 256 + lane * 8);
          const float4 p1 = *reinterpret_cast<const float4*>(pp + v * 256 + lane * 8 + 4);
          acc[v][0] = fmaf(p0.x, s1, acc[v][0]); acc[v][1] = fmaf(p0.y, s1, acc[v][1]);
          acc[v][2] = fmaf(p0.z, s1, acc[v][2]); acc[v][3] = fmaf(p0.w, s1, acc[v][3]);
          acc[v][4] = fmaf(p1.x, s1, acc[v][4]); acc[v][5] = fmaf(p1.y, s1, acc[v][5]);
          acc[v][6] = fmaf(p1.z, s1, acc[v][6]); acc[v][7] = fmaf(p1.w, s1, acc[v][7]);
        }
    }
  } else {
    gat_fwd_segment<H, NV>(colind, xh, a_src, ad, HC, slope, beg, end, lane, act, hl, mx, sum, acc);
  }
  float inv[H];
#pragma unroll
  for (int h = 0; h < H; ++h) {
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

// csr pass over the in-edges [beg, end) of destination i: A += alpha dot dl, B += alpha dl, T += alpha dot
template <int H, int NV>
__device__ __forceinline__ void gat_bwd_dst_segment(const int32_t* __restrict__ colind, const __nv_bfloat16* __restrict__ xh,
                                                    const float* __restrict__ a_src, const float (&ad)[H], const float (&mx)[H],
                                                    const float (&inv)[H], const float (&gi)[NV][8], int HC, float slope, int beg,
                                                    int end, int lane, const bool (&act)[NV], const int (&hl)[NV], float (&A)[H],
                                                    float (&B)[H], float (&T)[H]) {
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
}

// backward csr pre-pass for hub destinations: pabt[(chunk*2+slot)*3H + {A[H] | B[H] | T[H]}]
template <int H, int NV>
__global__ void __launch_bounds__(kGatWarps * 32) gat_hub_bwd_dst_kernel(const int32_t* __restrict__ rowptr,
                                                                         const int32_t* __restrict__ colind,
                                                                         const __nv_bfloat16* __restrict__ xh,
                                                                         const __nv_bfloat16* __restrict__ g,
                                                                         const float* __restrict__ a_src, const float* __restrict__ a_dst,
                                                                         const float* __restrict__ rowmax, const float* __restrict__ rowsum,
                                                                         int64_t N, int C, float slope, const int32_t* __restrict__ hub_rows,
                                                                         float* __restrict__ pabt) {
  __shared__ int s_rows[2];
  __shared__ float s_abt[kGatWarps][3 * H];
  int cs, ce;
  if (!hub_chunk_rows(rowptr, N, hub_rows, s_rows, cs, ce)) return;
  const int HC = H * C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int hl[NV];
  bool act[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) { act[v] = (v * 256 + lane * 8) < HC; hl[v] = act[v] ? (v * 256 + lane * 8) / C : 0; }
  for (int slot = 0; slot < 2; ++slot) {
    const int r = s_rows[slot];
    if (r < 0) continue;
    const int sb = max(rowptr[r], cs), se = min(rowptr[r + 1], ce);
    const int per = (se - sb + kGatWarps - 1) / kGatWarps;
    const int wb = min(se, sb + warp * per), we = min(se, wb + per);
    float gi[NV][8], ad[H], mx[H], inv[H], A[H], B[H], T[H];
#pragma unroll
    for (int v = 0; v < NV; ++v)
      if (act[v]) unpack8(ldg_stream(g + (int64_t)r * HC + v * 256 + lane * 8), gi[v]);
#pragma unroll
    for (int h = 0; h < H; ++h) {
      ad[h] = a_dst[(int64_t)r * H + h];
      mx[h] = rowmax[(int64_t)r * H + h];
      inv[h] = 1.0f / (rowsum[(int64_t)r * H + h] + 1e-16f);
      A[h] = B[h] = T[h] = 0.f;
    }
    gat_bwd_dst_segment<H, NV>(colind, xh, a_src, ad, mx, inv, gi, HC, slope, wb, we, lane, act, hl, A, B, T);
    if (lane == 0) {
#pragma unroll
      for (int h = 0; h < H; ++h) { s_abt[warp][h] = A[h]; s_abt[warp][H + h] = B[h]; s_abt[warp][2 * H + h] = T[h]; }
    }
    __syncthreads();
    if (threadIdx.x < 3 * H) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kGatWarps; ++w) t += s_abt[w][threadIdx.x];
      pabt[((int64_t)blockIdx.x * 2 + slot) * 3 * H + threadIdx.x] = t;
    }
    __syncthreads();
  }
}

// csr pass: per destination i, d a_dst[i] and t_i
template <int H, int NV>
__global__ void __launch_bounds__(kGatWarps * 32) gat_bwd_dst_kernel(const int32_t* __restrict__ rowptr,
                                                                     const int32_t* __restrict__ colind,
                                                                     const __nv_bfloat16* __restrict__ xh,
                                                                     const __nv_bfloat16* __restrict__ g,
                                                                     const float* __restrict__ a_src, const float* __restrict__ a_dst,
                                                                     const float* __restrict__ rowmax, const float* __restrict__ rowsum,
                                                                     int64_t N, int C, float slope, const float* __restrict__ pabt,
                                                                     float* __restrict__ d_adst, float* __restrict__ tsum) {
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
  if (pabt != nullptr && end - beg > kHubThreshold) {
    const int c_first = beg / kHubSeg, c_last = (end - 1) / kHubSeg;
    for (int c = c_first; c <= c_last; ++c) {
      const float* pp = pabt + ((int64_t)c * 2 + (c == c_first ? 1 : 0)) * 3 * H;
#pragma unroll
      for (int h = 0; h < H; ++h) { A[h] += pp[h]; B[h] += pp[H + h]; T[h] += pp[2 * H + h]; }
    }
  } else {
    gat_bwd_dst_segment<H, NV>(colind, xh, a_src, ad, mx, inv, gi, HC, slope, beg, end, lane, act, hl, A, B, T);
  }
  if (lane == 0) {
#pragma unroll
    for (int h = 0; h < H; ++h) {
      d_adst[row * H + h] = A[h] - T[h] * B[h];
      tsum[row * H + h] = T[h];
    }
  }
}

// csc pass over the out-edges [beg, end) of source j: das += de dl, acc += alpha g_i
template <int H, int NV>
__device__ __forceinline__ void gat_bwd_src_segment(const int32_t* __restrict__ csc_colind, const __nv_bfloat16* __restrict__ g,
                                                    const float* __restrict__ a_dst, const float* __restrict__ rowmax,
                                                    const float* __restrict__ rowsum, const float* __restrict__ tsum,
                                                    const float (&as)[H], const float (&xj)[NV][8], int HC, float slope, int beg,
                                                    int end, int lane, const bool (&act)[NV], const int (&hl)[NV], float (&das)[H],
                                                    float (&acc)[NV][8]) {
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
}

// backward csc pre-pass for hub sources: pacc[(chunk*2+slot)*HC + c], pdas[(chunk*2+slot)*H + h]
template <int H, int NV>
__global__ void __launch_bounds__(kGatWarps * 32) gat_hub_bwd_src_kernel(const int32_t* __restrict__ csc_rowptr,
                                                                         const int32_t* __restrict__ csc_colind,
                                                                         const __nv_bfloat16* __restrict__ xh,
                                                                         const __nv_bfloat16* __restrict__ g,
                                                                         const float* __restrict__ a_src, const float* __restrict__ a_dst,
                                                                         const float* __restrict__ rowmax, const float* __restrict__ rowsum,
                                                                         const float* __restrict__ tsum, int64_t N, int C, float slope,
                                                                         const int32_t* __restrict__ hub_rows, float* __restrict__ pacc,
                                                                         float* __restrict__ pdas) {
  __shared__ int s_rows[2];
  __shared__ float s_das[kGatWarps][H];
  extern __shared__ float red[];  // [kGatWarps][HC]
  int cs, ce;
  if (!hub_chunk_rows(csc_rowptr, N, hub_rows, s_rows, cs, ce)) return;
  const int HC = H * C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int hl[NV];
  bool act[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) { act[v] = (v * 256 + lane * 8) < HC; hl[v] = act[v] ? (v * 256 + lane * 8) / C : 0; }
  for (int slot = 0; slot < 2; ++slot) {
    const int r = s_rows[slot];
    if (r < 0) continue;
    const int sb = max(csc_rowptr[r], cs), se = min(csc_rowptr[r + 1], ce);
    const int per = (se - sb + kGatWarps - 1) / kGatWarps;
    const int wb = min(se, sb + warp * per), we = min(se, wb + per);
    float xj[NV][8], acc[NV][8], as[H], das[H];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      if (act[v]) unpack8(ldg_stream(xh + (int64_t)r * HC + v * 256 + lane * 8), xj[v]);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[v][i] = 0.f;
    }
#pragma unroll
    for (int h = 0; h < H; ++h) { as[h] = a_src[(int64_t)r * H + h]; das[h] = 0.f; }
    gat_bwd_src_segment<H, NV>(csc_colind, g, a_dst, rowmax, rowsum, tsum, as, xj, HC, slope, wb, we, lane, act, hl, das, acc);
    if (lane == 0) {
#pragma unroll
      for (int h = 0; h < H; ++h) s_das[warp][h] = das[h];
    }
#pragma unroll
    for (int v = 0; v < NV; ++v)
      if (act[v]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) red[warp * HC + v * 256 + lane * 8 + i] = acc[v][i];
      }
    __syncthreads();
    for (int c = threadIdx.x; c < HC; c += kGatWarps * 32) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kGatWarps; ++w) t += red[w * HC + c];
      pacc[((int64_t)blockIdx.x * 2 + slot) * HC + c] = t;
    }
    if (threadIdx.x < H) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kGatWarps; ++w) t += s_das[w][threadIdx.x];
      pdas[((int64_t)blockIdx.x * 2 + slot) * H + threadIdx.x] = t;
    }
    __syncthreads();
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
                                                                     int64_t N, int C, float slope, const float* __restrict__ pacc,
                                                                     const float* __restrict__ pdas, __nv_bfloat16* __restrict__ dxh,
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
  if (pacc != nullptr && end - beg > kHubThreshold) {
    const int c_first = beg / kHubSeg, c_last = (end - 1) / kHubSeg;
    for (int c = c_first; c <= c_last; ++c) {
      const int64_t rec = (int64_t)c * 2 + (c == c_first ? 1 : 0);
#pragma unroll
      for (int h = 0; h < H; ++h) das[h] += pdas[rec * H + h];
      const float* pp = pacc + rec * HC;
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (act[v]) {
          const float4 p0 = *reinterpret_cast<const float4*>(pp + v * 256 + lane * 8);
          const float4 p1 = *reinterpret_cast<const float4*>(pp + v * 256 + lane * 8 + 4);
          acc[v][0] += p0.x; acc[v][1] += p0.y; acc[v][2] += p0.z; acc[v][3] += p0.w;
          acc[v][4] += p1.x; acc[v][5] += p1.y; acc[v][6] += p1.z; acc[v][7] += p1.w;
        }
    }
  } else {
    gat_bwd_src_segment<H, NV>(csc_colind, g, a_dst, rowmax, rowsum, tsum, as, xj, HC, slope, beg, end, lane, act, hl, das, acc);
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

// workspace carve for the hub partials: fwd [acc HC | stat 2H], bwd [abt 3H | acc HC | das H] per (chunk, slot)
struct GatHubWs {
  float *acc, *stat, *abt, *das;
};
static size_t gat_hub_ws_floats(int64_t nnz_capacity, int H, int C) {
  const size_t recs = (size_t)ceil_div(nnz_capacity > 0 ? nnz_capacity : 1, kHubSeg) * 2;
  return recs * ((size_t)H * C + 3 * (size_t)H + (size_t)H + 2 * (size_t)H);
}
static GatHubWs gat_hub_carve(void* ws, int64_t nnz_capacity, int H, int C) {
  const size_t recs = (size_t)ceil_div(nnz_capacity > 0 ? nnz_capacity : 1, kHubSeg) * 2;
  GatHubWs w;
  w.acc = static_cast<float*>(ws);
  w.stat = w.acc + recs * H * C;
  w.abt = w.stat + recs * 2 * H;
  w.das = w.abt + recs * 3 * H;
  return w;
}

template <int H, int NV>
static int launch_gat_fwd(const int32_t* rowptr, const int32_t* colind, const __nv_bfloat16* xh, const float* as, const float* ad,
                          int64_t N, int C, float slope, const GatEpilogue& ep, void* out, int out_f32, float* rowmax,
                          float* rowsum, int64_t nnz_capacity, const int32_t* hub_rows, void* hub_ws, cudaStream_t st) {
  const unsigned grid = (unsigned)ceil_div(N, kGatWarps);
  const float *pacc = nullptr, *pstat = nullptr;
  if (hub_ws) {
    GatHubWs w = gat_hub_carve(hub_ws, nnz_capacity, H, C);
    const unsigned chunks = (unsigned)ceil_div(nnz_capacity, kHubSeg);
    gat_hub_fwd_kernel<H, NV><<<chunks, kGatWarps * 32, (size_t)kGatWarps * H * C * sizeof(float), st>>>(
        rowptr, colind, xh, as, ad, N, C, slope, hub_rows, w.acc, w.stat);
    pacc = w.acc;
    pstat = w.stat;
  }
  if (out_f32)
    gat_aggregate_kernel<H, NV, true><<<grid, kGatWarps * 32, 0, st>>>(rowptr, colind, xh, as, ad, N, C, slope, ep, pacc, pstat, out,
                                                                       rowmax, rowsum);
  else
    gat_aggregate_kernel<H, NV, false><<<grid, kGatWarps * 32, 0, st>>>(rowptr, colind, xh, as, ad, N, C, slope, ep, pacc, pstat, out,
                                                                        rowmax, rowsum);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

template <int H, int NV>
static int launch_gat_bwd(const int32_t* rowptr, const int32_t* colind, const int32_t* crp, const int32_t* cci,
                          const __nv_bfloat16* xh, const __nv_bfloat16* g, const float* as, const float* ad, const float* rmax,
                          const float* rsum, const float* att_s, const float* att_d, int64_t N, int C, float slope,
                          __nv_bfloat16* dxh, float* das, float* dad, float* tsum, int64_t nnz_capacity,
                          const int32_t* hub_rows_csr, const int32_t* hub_rows_csc, void* hub_ws, cudaStream_t st) {
  const unsigned grid = (unsigned)ceil_div(N, kGatWarps);
  const float *pabt = nullptr, *pacc = nullptr, *pdas = nullptr;
  GatHubWs w{};
  const unsigned chunks = (unsigned)ceil_div(nnz_capacity > 0 ? nnz_capacity : 1, kHubSeg);
  if (hub_ws) {
    w = gat_hub_carve(hub_ws, nnz_capacity, H, C);
    gat_hub_bwd_dst_kernel<H, NV><<<chunks, kGatWarps * 32, 0, st>>>(rowptr, colind, xh, g, as, ad, rmax, rsum, N, C, slope,
                                                                  hub_rows_csr, w.abt);
    pabt = w.abt;
  }
  gat_bwd_dst_kernel<H, NV><<<grid, kGatWarps * 32, 0, st>>>(rowptr, colind, xh, g, as, ad, rmax, rsum, N, C, slope, pabt, dad, tsum);
  if (hub_ws) {
    gat_hub_bwd_src_kernel<H, NV><<<chunks, kGatWarps * 32, (size_t)kGatWarps * H * C * sizeof(float), st>>>(
        crp, cci, xh, g, as, ad, rmax, rsum, tsum, N, C, slope, hub_rows_csc, w.acc, w.das);
    pacc = w.acc;
    pdas = w.das;
  }
  gat_bwd_src_kernel<H, NV><<<grid, kGatWarps * 32, 0, st>>>(crp, cci, xh, g, as, ad, rmax, rsum, tsum, dad, att_s, att_d, N, C, slope,
                                                            pacc, pdas, dxh, das);
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

size_t bmkg_gat_workspace_bytes(int64_t nnz_capacity, int H, int C) { return gat_hub_ws_floats(nnz_capacity, H, C) * sizeof(float); }

int bmkg_gat_aggregate(const int32_t* rowptr, const int32_t* colind, const void* xh_bf16, const float* a_src, const float* a_dst,
                       int64_t N, int H, int C, float negative_slope, const float* bias, int relu, float drop_p,
                       uint64_t drop_seed, const uint8_t* drop_keep, void* out, int out_is_fp32, float* rowmax, float* rowsum,
                       int64_t nnz_capacity, const int32_t* hub_rows, void* hub_ws, size_t hub_ws_bytes, void* stream) {
  BMKG_REQUIRE(!hub_ws || (nnz_capacity > 0 && hub_ws_bytes >= bmkg_gat_workspace_bytes(nnz_capacity, H, C)), BMKG_ERR_WORKSPACE);
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
                                                               out_is_fp32, rowmax, rowsum, nnz_capacity, hub_rows, hub_ws, st)));
  return BMKG_OK;
}

int bmkg_gat_aggregate_bwd(const int32_t* rowptr, const int32_t* colind, const int32_t* csc_rowptr, const int32_t* csc_colind,
                           const void* xh_bf16, const void* g_bf16, const float* a_src, const float* a_dst, const float* rowmax,
                           const float* rowsum, const float* att_src, const float* att_dst, int64_t N, int H, int C,
                           float negative_slope, void* dxh_bf16, float* d_asrc, float* d_adst, float* tsum_ws,
                           int64_t nnz_capacity, const int32_t* hub_rows_csr, const int32_t* hub_rows_csc, void* hub_ws,
                           size_t hub_ws_bytes, void* stream) {
  BMKG_REQUIRE(!hub_ws || (nnz_capacity > 0 && hub_ws_bytes >= bmkg_gat_workspace_bytes(nnz_capacity, H, C)), BMKG_ERR_WORKSPACE);
  BMKG_REQUIRE(rowptr && colind && csc_rowptr && csc_colind && xh_bf16 && g_bf16 && a_src && a_dst && rowmax && rowsum && att_src &&
                   att_dst && dxh_bf16 && d_asrc && d_adst && tsum_ws && N > 0,
               BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(gat_shape_ok(H, C), BMKG_ERR_UNSUPPORTED);
  BMKG_REQUIRE(aligned16(xh_bf16) && aligned16(g_bf16) && aligned16(dxh_bf16), BMKG_ERR_MISALIGNED);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GAT_DISPATCH(H, gat_nv(H * C),
               return (launch_gat_bwd<H, NV>(rowptr, colind, csc_rowptr, csc_colind, static_cast<const __nv_bfloat16*>(xh_bf16),
                                             static_cast<const __nv_bfloat16*>(g_bf16), a_src, a_dst, rowmax, rowsum, att_src, att_dst,
                                             N, C, negative_slope, static_cast<__nv_bfloat16*>(dxh_bf16), d_asrc, d_adst, tsum_ws,
                                             nnz_capacity, hub_rows_csr, hub_rows_csc, hub_ws, st)));
  return BMKG_OK;
}

}  // extern "C"
