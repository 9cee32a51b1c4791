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
// Power-law graphs (BASELINE cfg 5): rows longer than kHubThreshold are cut at fixed kHubSeg-edge chunk boundaries of the CSR
// edge array, each chunk reduced by one CTA into a partial row (8 warps x contiguous sub-ranges, combined in warp order), and
// the row kernel sums the partials in chunk order.  The split depends only on rowptr, so the result is deterministic; no atomics.
// (kHubThreshold / kHubSeg live in common.cuh; a chunk intersects at most two hub rows - slot 0: continues, slot 1: starts)

struct AggEpilogue {
  const float* bias;        // [C] or null
  int relu;
  float drop_scale;         // 1/(1-p) or 1
  uint32_t drop_threshold;  // p * 2^32, 0 = no hashed dropout
  uint64_t drop_seed;
  const uint8_t* drop_keep;  // explicit [N,C] keep mask (tests / replay) or null
};

// acc[v][i] += sum_{k in [beg,end)} dis[colind[k]] * X[colind[k]][v*256 + lane*8 + i]   (CSR order, fp32)
// STAR (1-hop export, see gcn_star_aggregate_kernel): neighbours are leaves of a star - weight 1, rows of X - and the row's
// own self-loop entry takes weight dis[self_row] and its row from Xself instead.
template <int NV, bool STAR = false>
__device__ __forceinline__ void gather_accumulate(const int32_t* __restrict__ colind, const float* __restrict__ dis,
                                                  const __nv_bfloat16* __restrict__ X, int C, int beg, int end, int lane,
                                                  const bool (&act)[NV], float (&acc)[NV][8], int self_row = -1,
                                                  const __nv_bfloat16* __restrict__ Xself = nullptr) {
  for (int base = beg; base < end; base += 32) {
    const int k = base + lane;
    int c = 0;
    float w = 0.f;
    if (k < end) {
      c = colind[k];
      w = STAR ? (c == self_row ? dis[c] : 1.f) : dis[c];
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
          const __nv_bfloat16* rp = ((STAR && cj == self_row) ? Xself : X) + (int64_t)cj * C + lane * 8;
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
}

// one CTA per kHubSeg-edge chunk of the CSR edge array: partial[(chunk*2 + slot)*C + col] for the hub rows it intersects
template <int NV>
__global__ void __launch_bounds__(kAggWarps * 32) gcn_hub_partial_kernel(const int32_t* __restrict__ rowptr,
                                                                         const int32_t* __restrict__ colind,
                                                                         const float* __restrict__ dis,
                                                                         const __nv_bfloat16* __restrict__ X, int64_t N, int C,
                                                                         const int32_t* __restrict__ hub_rows /*hub_info*/,
                                                                         int64_t row_begin, int64_t row_end, float* __restrict__ partial) {
  __shared__ int s_rows[2];         // hub row containing the chunk start (or -1), hub row starting inside (or -1)
  extern __shared__ float red[];    // [kAggWarps][C]
  int cs, ce;
  if (!hub_chunk_rows(rowptr, N, hub_rows, s_rows, cs, ce)) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bool act[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) act[v] = (v * 256 + lane * 8) < C;
  for (int slot = 0; slot < 2; ++slot) {
    const int r = s_rows[slot];
    if (r < 0 || r < row_begin || r >= row_end) continue;  // uniform across the CTA; rows of other ranks are not ours to reduce
    const int sb = max(rowptr[r], cs), se = min(rowptr[r + 1], ce);
    const int per = (se - sb + kAggWarps - 1) / kAggWarps;
    const int wb = min(se, sb + warp * per), we = min(se, wb + per);
    float acc[NV][8];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[v][i] = 0.f;
    gather_accumulate<NV>(colind, dis, X, C, wb, we, lane, act, acc);
#pragma unroll
    for (int v = 0; v < NV; ++v)
      if (act[v]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) red[warp * C + v * 256 + lane * 8 + i] = acc[v][i];
      }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kAggWarps * 32) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kAggWarps; ++w) t += red[w * C + c];
      partial[((int64_t)blockIdx.x * 2 + slot) * C + c] = t;
    }
    __syncthreads();
  }
}

template <int NV, bool OUT_F32>
__global__ void __launch_bounds__(kAggWarps * 32) gcn_aggregate_kernel(const int32_t* __restrict__ rowptr,
                                                                       const int32_t* __restrict__ colind,
                                                                       const float* __restrict__ dis,
                                                                       const __nv_bfloat16* __restrict__ X, int64_t N, int C,
                                                                       AggEpilogue ep, const float* __restrict__ hub_partial,
                                                                       int64_t row_begin, void* __restrict__ out) {
  // rows [row_begin, row_begin + N) of the graph; `out` (and an explicit dropout mask) hold those N rows only: the row-sharded
  // multi-GPU encoder aggregates its own destination rows from the all-gathered X
  const int lane = threadIdx.x & 31;
  const int64_t lrow = (int64_t)blockIdx.x * kAggWarps + (threadIdx.x >> 5);
  if (lrow >= N) return;
  const int64_t row = row_begin + lrow;
  const int beg = rowptr[row], end = rowptr[row + 1];

  float acc[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[v][i] = 0.f;

  bool act[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) act[v] = (v * 256 + lane * 8) < C;

  if (hub_partial != nullptr && end - beg > kHubThreshold) {
    // hub row: sum the per-chunk partial rows in chunk order (slot 1 in the chunk where the row starts, slot 0 after)
    const int c_first = beg / kHubSeg, c_last = (end - 1) / kHubSeg;
    for (int c = c_first; c <= c_last; ++c) {
      const float* pp = hub_partial + ((int64_t)c * 2 + (c == c_first ? 1 : 0)) * C;
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
    gather_accumulate<NV>(colind, dis, X, C, beg, end, lane, act, acc);
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
      const uint2 m = *reinterpret_cast<const uint2*>(ep.drop_keep + lrow * C + c0);
      const uint32_t mm[2] = {m.x, m.y};
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = ((mm[i >> 2] >> (8 * (i & 3))) & 0xff) ? r[i] * ep.drop_scale : 0.f;
    } else if (ep.drop_threshold) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        r[i] = hash_keep(ep.drop_seed, (uint64_t)(row * C + c0 + i), ep.drop_threshold) ? r[i] * ep.drop_scale : 0.f;
    }
    if (OUT_F32) {
      float* o = static_cast<float*>(out) + lrow * C + c0;
      *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(r[4], r[5], r[6], r[7]);
    } else {
      __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out) + lrow * C + c0;
      *reinterpret_cast<uint4*>(o) = pack8(r);
    }
  }
}

// 1-hop "star" aggregation for the embedding-export path (biomedkg/data/node.py:193-241 drives BaseGCL.forward,
// gcl_module.py:55-58, over NeighborLoader(num_neighbors=[-1]) batches of ONE seed, data_module.py:71-79).  In such a batch
// the only edges are neighbour -> seed, so after gcn_norm every neighbour has degree 1 (its own self-loop) and the seed has
// the same in-degree d as in the full graph.  All N star graphs therefore share one "leaf" chain per node and differ only in
// the seed row:  out[s] = dis[s] * (sum_{j -> s, j != s} Tleaf[j] + dis[s] * Tseed[s]) + b,  dis = d^-1/2 - one pass over
// the full-graph CSR per layer instead of N tiny forward passes.  Rows longer than kHubThreshold are simply walked by their
// warp (export is a one-off pass).
template <int NV, bool OUT_F32>
__global__ void __launch_bounds__(kAggWarps * 32) gcn_star_aggregate_kernel(const int32_t* __restrict__ rowptr,
                                                                            const int32_t* __restrict__ colind,
                                                                            const float* __restrict__ dis,
                                                                            const __nv_bfloat16* __restrict__ Xleaf,
                                                                            const __nv_bfloat16* __restrict__ Xseed, int64_t N, int C,
                                                                            const float* __restrict__ bias, int relu,
                                                                            void* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kAggWarps + (threadIdx.x >> 5);
  if (row >= N) return;
  float acc[NV][8];
  bool act[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    act[v] = (v * 256 + lane * 8) < C;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[v][i] = 0.f;
  }
  gather_accumulate<NV, true>(colind, dis, Xleaf, C, rowptr[row], rowptr[row + 1], lane, act, acc, (int)row, Xseed);
  const float di = dis[row];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    if (!act[v]) continue;
    const int c0 = v * 256 + lane * 8;
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      r[i] = acc[v][i] * di + (bias ? bias[c0 + i] : 0.f);
      if (relu) r[i] = fmaxf(r[i], 0.f);
    }
    if (OUT_F32) {
      float* o = static_cast<float*>(out) + row * C + c0;
      *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(r[4], r[5], r[6], r[7]);
    } else {
      *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(out) + row * C + c0) = pack8(r);
    }
  }
}

template <int NV>
static int launch_star(const int32_t* rowptr, const int32_t* colind, const float* dis, const __nv_bfloat16* xl,
                       const __nv_bfloat16* xs, int64_t N, int C, const float* bias, int relu, void* out, int out_f32,
                       cudaStream_t st) {
  const unsigned grid = (unsigned)ceil_div(N, kAggWarps);
  if (out_f32)
    gcn_star_aggregate_kernel<NV, true><<<grid, kAggWarps * 32, 0, st>>>(rowptr, colind, dis, xl, xs, N, C, bias, relu, out);
  else
    gcn_star_aggregate_kernel<NV, false><<<grid, kAggWarps * 32, 0, st>>>(rowptr, colind, dis, xl, xs, N, C, bias, relu, out);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

template <int NV>
static int launch_agg(const int32_t* rowptr, const int32_t* colind, const float* dis, const __nv_bfloat16* X, int64_t N, int C,
                      const AggEpilogue& ep, void* out, int out_f32, float* hub_partial, const int32_t* hub_rows, int64_t nnz_capacity,
                      int64_t row_begin, int64_t total_rows, cudaStream_t st) {
  const unsigned grid = (unsigned)ceil_div(N, kAggWarps);
  if (hub_partial) {
    const unsigned chunks = (unsigned)ceil_div(nnz_capacity, kHubSeg);
    gcn_hub_partial_kernel<NV><<<chunks, kAggWarps * 32, (size_t)kAggWarps * C * sizeof(float), st>>>(
        rowptr, colind, dis, X, total_rows, C, hub_rows, row_begin, row_begin + N, hub_partial);
  }
  if (out_f32)
    gcn_aggregate_kernel<NV, true><<<grid, kAggWarps * 32, 0, st>>>(rowptr, colind, dis, X, N, C, ep, hub_partial, row_begin, out);
  else
    gcn_aggregate_kernel<NV, false><<<grid, kAggWarps * 32, 0, st>>>(rowptr, colind, dis, X, N, C, ep, hub_partial, row_begin, out);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

}  // namespace bmkg

using namespace bmkg;

extern "C" size_t bmkg_gcn_aggregate_workspace_bytes(int64_t nnz_capacity, int C) {
  return (size_t)ceil_div(nnz_capacity > 0 ? nnz_capacity : 1, kHubSeg) * 2 * (size_t)C * sizeof(float);
}

extern "C" int bmkg_gcn_aggregate(const int32_t* rowptr, const int32_t* colind, const float* dis, const void* x_bf16, int64_t N,
                                  int C, const float* bias, int relu, float drop_p, uint64_t drop_seed,
                                  const uint8_t* drop_keep, void* out, int out_is_fp32, int64_t nnz_capacity,
                                  const int32_t* hub_rows, void* hub_ws, size_t hub_ws_bytes, void* stream) {
  return bmkg_gcn_aggregate_rows(rowptr, colind, dis, x_bf16, N, 0, N, C, bias, relu, drop_p, drop_seed, drop_keep, out, out_is_fp32,
                                 nnz_capacity, hub_rows, hub_ws, hub_ws_bytes, stream);
}

extern "C" int bmkg_gcn_aggregate_rows(const int32_t* rowptr, const int32_t* colind, const float* dis, const void* x_bf16,
                                       int64_t total_rows, int64_t row_begin, int64_t N, int C, const float* bias, int relu,
                                       float drop_p, uint64_t drop_seed, const uint8_t* drop_keep, void* out, int out_is_fp32,
                                       int64_t nnz_capacity, const int32_t* hub_rows, void* hub_ws, size_t hub_ws_bytes,
                                       void* stream) {
  BMKG_REQUIRE(row_begin >= 0 && N > 0 && row_begin + N <= total_rows, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(!hub_ws || (nnz_capacity > 0 && hub_ws_bytes >= bmkg_gcn_aggregate_workspace_bytes(nnz_capacity, C)),
               BMKG_ERR_WORKSPACE);
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
  float* hub = static_cast<float*>(hub_ws);
  const int nv = (C + 255) / 256;
  switch (nv) {
    case 1: return launch_agg<1>(rowptr, colind, dis, X, N, C, ep, out, out_is_fp32, hub, hub_rows, nnz_capacity, row_begin, total_rows, st);
    case 2: return launch_agg<2>(rowptr, colind, dis, X, N, C, ep, out, out_is_fp32, hub, hub_rows, nnz_capacity, row_begin, total_rows, st);
    case 3: return launch_agg<3>(rowptr, colind, dis, X, N, C, ep, out, out_is_fp32, hub, hub_rows, nnz_capacity, row_begin, total_rows, st);
    default: return launch_agg<4>(rowptr, colind, dis, X, N, C, ep, out, out_is_fp32, hub, hub_rows, nnz_capacity, row_begin, total_rows, st);
  }
}


extern "C" int bmkg_gcn_star_aggregate(const int32_t* rowptr, const int32_t* colind, const float* dis, const void* leaf_bf16,
                                       const void* seed_bf16, int64_t N, int C, const float* bias, int relu, void* out,
                                       int out_is_fp32, void* stream) {
  BMKG_REQUIRE(rowptr && colind && dis && leaf_bf16 && seed_bf16 && out, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(N > 0 && N < (1ll << 31) && C > 0 && C % 8 == 0 && C <= 1024, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(aligned16(leaf_bf16) && aligned16(seed_bf16) && aligned16(out), BMKG_ERR_MISALIGNED);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* xl = static_cast<const __nv_bfloat16*>(leaf_bf16);
  const __nv_bfloat16* xs = static_cast<const __nv_bfloat16*>(seed_bf16);
  switch ((C + 255) / 256) {
    case 1: return launch_star<1>(rowptr, colind, dis, xl, xs, N, C, bias, relu, out, out_is_fp32, st);
    case 2: return launch_star<2>(rowptr, colind, dis, xl, xs, N, C, bias, relu, out, out_is_fp32, st);
    case 3: return launch_star<3>(rowptr, colind, dis, xl, xs, N, C, bias, relu, out, out_is_fp32, st);
    default: return launch_star<4>(rowptr, colind, dis, xl, xs, N, C, bias, relu, out, out_is_fp32, st);
  }
}
