// A1/A2 - GCN neighbour aggregation over CSR (forward) / CSC (transposed backward).
//
// Replaces PyG GCNConv.propagate as reached from biomedkg/model/encoder.py:155,160:
// the reference gathers a [E',C] fp32 message tensor and scatter_add_s it with
// atomics.  Here one warp owns one destination row: each lane
// holds 8 of the C feature columns (one 128-bit bf16 load per neighbour row,
// fully coalesced: 32 lanes x 16 B = one 512 B row at C=256), accumulates in
// fp32 registers in CSR order (deterministic, no atomics), and the epilogue
// fuses the symmetric norm dis[i], bias, ReLU and dropout (encoder.py:155-158).
// Edge weights are never stored: w_ij = dis[i]*dis[j] is rebuilt from dis[N].
//
// HBM-bound.  Algorithmic bytes per call: E'*(C*2+4) + N*C*s_out + 4*(N+1).
#include "common.cuh"
#include "../../include/bmkg_b200.h"

namespace bmkg {

constexpr int kAggWarps = 8;
constexpr int kAggUnroll = 8;

struct AggEpilogue {
  const float* bias;        // [C] or null
  int relu;
  float drop_scale;         // 1/(1-p) or 1
  uint32_t drop_threshold;  // p * 2^32, 0 = no hashed dropout
  uint64_t drop_seed;
  const uint8_t* drop_keep;  // explicit [N,C] keep mask (tests / replay) or null
};

template <int NV, bool OUT_F32>
__global__ void __launch_bounds__(kAggWarps * 32) gcn_aggregate_kernel(const int32_t* __restrict__ rowptr,
                                                                       const int32_t* __restrict__ colind,
                                                                       const float* __restrict__ dis,
                                                                       const __nv_bfloat16* __restrict__ X, int64_t N, int C,
                                                                       AggEpilogue ep, void* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kAggWarps + (threadIdx.x >> 5);
  if (row >= N) return;
  const int beg = rowptr[row], end = rowptr[row + 1];

  float acc[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[v][i] = 0.f;

  bool act[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) act[v] = (v * 256 + lane * 8) < C;

  for (int base = beg; base < end; base += 32) {
    const int k = base + lane;
    int c = 0;
    float w = 0.f;
    if (k < end) {
      c = colind[k];
      w = dis[c];
    }
    const int cnt = min(32, end - base);
    for (int j = 0; j < cnt; j += kAggUnroll) {
      uint4 u[kAggUnroll][NV];
      float wj[kAggUnroll];
#pragma unroll
      for (int t = 0; t < kAggUnroll; ++t) {
        const int src_lane = (j + t) & 31;
        const int cj = __shfl_sync(0xffffffffu, c, src_lane);
        wj[t] = __shfl_sync(0xffffffffu, w, src_lane);
        if (j + t < cnt) {
          const __nv_bfloat16* rp = X + (int64_t)cj * C + lane * 8;
#pragma unroll
          for (int v = 0; v < NV; ++v)
            if (act[v]) u[t][v] = ldg_cached(rp + v * 256);
        }
      }
#pragma unroll
      for (int t = 0; t < kAggUnroll; ++t) {
        if (j + t < cnt) {
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            if (act[v]) {
              float f[8];
              unpack8(u[t][v], f);
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[v][i] = fmaf(wj[t], f[i], acc[v][i]);
            }
          }
        }
      }
    }
  }

  const float di = dis[row];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    if (!act[v]) continue;
    const int c0 = v * 256 + lane * 8;
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = acc[v][i] * di;
    if (ep.bias) {
      const float4 b0 = *reinterpret_cast<const float4*>(ep.bias + c0);
      const float4 b1 = *reinterpret_cast<const float4*>(ep.bias + c0 + 4);
      r[0] += b0.x; r[1] += b0.y; r[2] += b0.z; r[3] += b0.w;
      r[4] += b1.x; r[5] += b1.y; r[6] += b1.z; r[7] += b1.w;
    }
    if (ep.relu) {
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = fmaxf(r[i], 0.f);
    }
    if (ep.drop_keep) {
      const uint2 m = *reinterpret_cast<const uint2*>(ep.drop_keep + row * C + c0);
      const uint32_t mm[2] = {m.x, m.y};
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = ((mm[i >> 2] >> (8 * (i & 3))) & 0xff) ? r[i] * ep.drop_scale : 0.f;
    } else if (ep.drop_threshold) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        r[i] = hash_keep(ep.drop_seed, (uint64_t)(row * C + c0 + i), ep.drop_threshold) ? r[i] * ep.drop_scale : 0.f;
    }
    if (OUT_F32) {
      float* o = static_cast<float*>(out) + row * C + c0;
      *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(r[4], r[5], r[6], r[7]);
    } else {
      __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out) + row * C + c0;
      *reinterpret_cast<uint4*>(o) = pack8(r);
    }
  }
}

template <int NV>
static int launch_agg(const int32_t* rowptr, const int32_t* colind, const float* dis, const __nv_bfloat16* X, int64_t N, int C,
                      const AggEpilogue& ep, void* out, int out_f32, cudaStream_t st) {
  const unsigned grid = (unsigned)ceil_div(N, kAggWarps);
  if (out_f32)
    gcn_aggregate_kernel<NV, true><<<grid, kAggWarps * 32, 0, st>>>(rowptr, colind, dis, X, N, C, ep, out);
  else
    gcn_aggregate_kernel<NV, false><<<grid, kAggWarps * 32, 0, st>>>(rowptr, colind, dis, X, N, C, ep, out);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

}  // namespace bmkg

using namespace bmkg;

extern "C" int bmkg_gcn_aggregate(const int32_t* rowptr, const int32_t* colind, const float* dis, const void* x_bf16, int64_t N,
                                  int C, const float* bias, int relu, float drop_p, uint64_t drop_seed,
                                  const uint8_t* drop_keep, void* out, int out_is_fp32, void* stream) {
  BMKG_REQUIRE(rowptr && colind && dis && x_bf16 && out, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(N > 0 && C > 0 && C % 8 == 0 && C <= 1024, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(drop_p >= 0.f && drop_p < 1.f, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(aligned16(x_bf16) && aligned16(out) && (!bias || aligned16(bias)) && (!drop_keep || aligned16(drop_keep)),
               BMKG_ERR_MISALIGNED);
  AggEpilogue ep;
  ep.bias = bias;
  ep.relu = relu;
  ep.drop_keep = (drop_p > 0.f) ? drop_keep : nullptr;
  ep.drop_scale = (drop_p > 0.f) ? 1.0f / (1.0f - drop_p) : 1.0f;
  ep.drop_threshold = (drop_p > 0.f && !drop_keep) ? (uint32_t)((double)drop_p * 4294967296.0) : 0u;
  ep.drop_seed = drop_seed;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* X = static_cast<const __nv_bfloat16*>(x_bf16);
  const int nv = (C + 255) / 256;
  switch (nv) {
    case 1: return launch_agg<1>(rowptr, colind, dis, X, N, C, ep, out, out_is_fp32, st);
    case 2: return launch_agg<2>(rowptr, colind, dis, X, N, C, ep, out, out_is_fp32, st);
    case 3: return launch_agg<3>(rowptr, colind, dis, X, N, C, ep, out, out_is_fp32, st);
    default: return launch_agg<4>(rowptr, colind, dis, X, N, C, ep, out, out_is_fp32, st);
  }
}
