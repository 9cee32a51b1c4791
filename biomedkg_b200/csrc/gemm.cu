// L1 - the Linear layers of the GCL step on tcgen05 / TMEM / TMA (sm_100a), replacing the cuBLAS calls of round 1.
//
// Reference call sites (all torch.nn.Linear / PyG Linear, y = x W^T + b):
//   biomedkg/utils/fusion.py:18-20   AttentionFusion q/k/v projections   [N*M,768] x [768,2304]   (one fused GEMM)
//   biomedkg/utils/fusion.py:70      ReDAF transform_layer               [N*M,768] x [768,768]
//   biomedkg/model/encoder.py:138-143 GCNConv.lin (PyG)                  [N,768] x [768,256], [N,256] x [256,256]
//   biomedkg/model/gcl.py:49-51      GRACE.project fc1 / fc2             [N,256] x [256,256]
//
// Two kernels, both persistent, warp-specialised (warp 0 TMA producer, warp 1 tcgen05.mma issuer, warps 2-5 epilogue), bf16
// operands, fp32 accumulation in TMEM, 128-byte-swizzled shared-memory tiles fed by cp.async.bulk.tensor:
//
//   gemm_nt   C[M,N] = A[M,K] B[N,K]^T (+ bias[N]) (ELU)      forward (B = W) and input gradient (B = W^T, prepared by the host)
//             tile 128 x BN (BN <= 256) x 64, 4-stage ring, TWO TMEM accumulators (2 x 256 columns) so the epilogue of one
//             tile overlaps the MMAs of the next; the epilogue adds the fp32 bias (which also carries the weight-residual
//             correction of ops._xw), applies ELU, optionally emits the GAT attention scores <row, att_src/att_dst> of the
//             bf16-rounded row (fusing bmkg_gat_scores), and stores bf16 or fp32 rows.
//   gemm_tn   C[N,K] = sum_m G[m,N] X[m,K]                    weight gradient: a reduction over the (huge) node dimension
//             both operands are read as MN-major tiles straight from their row-major storage (no transposed copies), the
//             node range is split over CTAs, every CTA writes its fp32 partial tile and a fixed-order pass adds them
//             (deterministic, no atomics).
//
// Tensor-bound for the fusion GEMM (6 N M 768^2 FLOP), bandwidth-bound for the 256-wide layers (one pass over A).
#include <cuda.h>

#include "common.cuh"
#include "umma.cuh"
#include "../../include/bmkg_b200.h"

namespace bmkg {
namespace gemm {

constexpr int kBM = 128;       // rows of A per tile (UMMA M)
constexpr int kBK = 64;        // K elements per stage: 64 bf16 = one 128-byte swizzle row
constexpr int kMaxBN = 256;    // columns of C per tile (UMMA N)
constexpr int kStages = 4;
constexpr int kThreads = 192;  // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue (TMEM lane quarters 2,3,0,1)
constexpr int kABytes = kBM * 128;           // 16 KB
constexpr int kBBytes = kMaxBN * 128;        // 32 KB
constexpr size_t kSmemBytes = 1024 + (size_t)kStages * (kABytes + kBBytes) + 256;

struct Epilogue {
  const float* bias;       // [N] fp32 or null
  int elu;                 // apply ELU(alpha = 1) after the bias
  int out_f32;             // C is fp32 (else bf16)
  const float* att_src;    // GAT: [N] fp32 (heads * channels), with att_dst, a_src, a_dst; null = off.  Needs N <= BN (one n tile)
  const float* att_dst;
  float* a_src;            // [M, heads]
  float* a_dst;
  int heads;
};

__global__ void __launch_bounds__(kThreads, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int M, int N, int K, int BN,
               Epilogue ep, void* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - ptx::smem_u32(smem_raw));
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)kStages * kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)kStages * kBBytes);
  uint64_t* full = bars;                 // [kStages]
  uint64_t* empty = full + kStages;      // [kStages]
  uint64_t* tfull = empty + kStages;     // [2] accumulator ready
  uint64_t* tempty = tfull + 2;          // [2] accumulator drained (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(&tfull[b], 1); ptx::mbar_init(&tempty[b], 4); }
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_b);
  }
  if (warp == 0) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_mt = (int)ceil_div(M, kBM), n_nt = (int)ceil_div(N, BN), n_kb = K / kBK;
  const int n_tiles = n_mt * n_nt;
  const uint32_t stage_bytes = (uint32_t)kABytes + (uint32_t)BN * 128u;

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer ----------------
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int mt = tile / n_nt, nt = tile % n_nt;   // n fastest: the A rows of one m tile stay L2-resident across its n tiles
        for (int kb = 0; kb < n_kb; ++kb) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&full[stage], stage_bytes);
          ptx::tma_load_2d(sA + (size_t)stage * kABytes, &tmap_a, &full[stage], kb * kBK, mt * kBM);
          ptx::tma_load_2d(sB + (size_t)stage * kBBytes, &tmap_b, &full[stage], kb * kBK, nt * BN);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {  // ---------------- MMA issuer ----------------
    const uint32_t idesc = ptx::idesc_bf16_f32(kBM, BN, 0, 0);
    int stage = 0, acc = 0;
    uint32_t phase = 0, accphase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      ptx::mbar_wait(&tempty[acc], accphase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * kMaxBN;
      for (int kb = 0; kb < n_kb; ++kb) {
        ptx::mbar_wait(&full[stage], phase);
        ptx::tc_fence_after();
        const uint64_t adesc = ptx::smem_desc_sw128(ptx::smem_u32(sA + (size_t)stage * kABytes), 16, 1024);
        const uint64_t bdesc = ptx::smem_desc_sw128(ptx::smem_u32(sB + (size_t)stage * kBBytes), 16, 1024);
        if (ptx::elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)   // K = 16 per instruction: +32 bytes inside the swizzle row (descriptor units of 16 B)
            ptx::umma_ss(d_tmem, adesc + (uint32_t)(kk * 2), bdesc + (uint32_t)(kk * 2), idesc, (kb > 0 || kk > 0) ? 1u : 0u);
          ptx::umma_commit(&empty[stage]);
          if (kb == n_kb - 1) ptx::umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; accphase ^= 1; }
    }
  } else {  // ---------------- epilogue warps ----------------
    const int quarter = warp & 3;
    const int lrow = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int acc = 0;
    uint32_t accphase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int mt = tile / n_nt, nt = tile % n_nt;
      const int64_t row = (int64_t)mt * kBM + lrow;
      const int col0 = nt * BN;
      const int ncols = min(BN, N - col0);
      ptx::mbar_wait(&tfull[acc], accphase);
      ptx::tc_fence_after();
      float as[4] = {0.f, 0.f, 0.f, 0.f}, ad[4] = {0.f, 0.f, 0.f, 0.f};   // GAT scores per head (heads <= 4)
      const int ch = ep.att_src ? N / ep.heads : 1;
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld32(lane_base + (uint32_t)acc * kMaxBN + (uint32_t)c0, r);
        ptx::tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (ep.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (c0 + j < ncols) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + c0 + j));
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          }
        }
        if (ep.elu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : expm1f(v[j]);
        }
        const int nv = min(32, ncols - c0);   // multiple of 16 (N % 16 == 0)
        if (ep.out_f32) {
          if (row < M) {
            float* o = static_cast<float*>(out) + row * N + col0 + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (j < nv) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        } else {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pack2(v[2 * j], v[2 * j + 1]);
          if (ep.att_src) {   // scores of the ROUNDED row (what the aggregation kernels read), as bmkg_gat_scores computes them
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (2 * j < nv) {
                const float lo = __uint_as_float(pk[j] << 16), hi = __uint_as_float(pk[j] & 0xffff0000u);
                const int c = col0 + c0 + 2 * j;
                const float2 s = __ldg(reinterpret_cast<const float2*>(ep.att_src + c));
                const float2 d = __ldg(reinterpret_cast<const float2*>(ep.att_dst + c));
                const int h = c / ch;
#pragma unroll
                for (int hh = 0; hh < 4; ++hh)
                  if (hh == h) {
                    as[hh] = fmaf(lo, s.x, fmaf(hi, s.y, as[hh]));
                    ad[hh] = fmaf(lo, d.x, fmaf(hi, d.y, ad[hh]));
                  }
              }
            }
          }
          if (row < M) {
            __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out) + row * N + col0 + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              if (j < nv) *reinterpret_cast<uint4*>(o + j) = make_uint4(pk[j / 2], pk[j / 2 + 1], pk[j / 2 + 2], pk[j / 2 + 3]);
          }
        }
      }
      // accumulator fully read: hand it back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
      if (ep.att_src && row < M) {
        for (int h = 0; h < ep.heads; ++h) {
          ep.a_src[row * ep.heads + h] = as[h];
          ep.a_dst[row * ep.heads + h] = ad[h];
        }
      }
      if (++acc == 2) { acc = 0; accphase ^= 1; }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<512>(tmem_base);
}

// ----------------------------------------------------------------------------
// weight gradient: C[N, K] = sum_m G[m, N] X[m, K]
// ----------------------------------------------------------------------------
// A operand = G tile [64 rows m x 128 columns n] (two 64-column panels), B operand = X tile [64 rows m x BK2 columns k]
// (BK2 / 64 panels), both MN-major: the contiguous dimension of the storage is the MMA's M / N dimension and the reduction
// dimension m runs over the rows.  K = 16 rows per instruction = 2048 bytes inside a panel.
constexpr int kTnRows = 64;                        // rows of m per stage
constexpr int kTnPanel = kTnRows * 128;            // one 64-column panel of a stage: 8 KB
constexpr int kTnABytes = 2 * kTnPanel;            // 128 columns of G
constexpr int kTnBBytes = 4 * kTnPanel;            // up to 256 columns of X
constexpr size_t kTnSmemBytes = 1024 + (size_t)kStages * (kTnABytes + kTnBBytes) + 256;

__global__ void __launch_bounds__(kThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_x, int M, int N, int K, int BK2,
               int rows_per_split, float* __restrict__ partial /*[splits][N][K]*/) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - ptx::smem_u32(smem_raw));
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)kStages * kTnABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)kStages * kTnBBytes);
  uint64_t* full = bars;
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    ptx::mbar_init(tfull, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tmap_g);
    ptx::prefetch_tensormap(&tmap_x);
  }
  if (warp == 0) ptx::tmem_alloc<256>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // blockIdx.x = output tile (n tile major, k tile minor), blockIdx.y = split of the node range
  const int n_kt = (int)ceil_div(K, BK2);
  const int nt = blockIdx.x / n_kt, kt = blockIdx.x % n_kt;
  const int m0 = blockIdx.y * rows_per_split, m1 = min(M, m0 + rows_per_split);
  const int n_it = (int)ceil_div(max(m1 - m0, 0), kTnRows);
  const int kcols = min(BK2, K - kt * BK2);          // multiple of 64
  const int kpanels = kcols / 64;
  const uint32_t stage_bytes = (uint32_t)kTnABytes + (uint32_t)kpanels * kTnPanel;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_it; ++it) {
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        ptx::mbar_arrive_expect_tx(&full[stage], stage_bytes);
        const int m = m0 + it * kTnRows;
        for (int p = 0; p < 2; ++p) ptx::tma_load_2d(sA + (size_t)stage * kTnABytes + p * kTnPanel, &tmap_g, &full[stage], nt * 128 + p * 64, m);
        for (int p = 0; p < kpanels; ++p) ptx::tma_load_2d(sB + (size_t)stage * kTnBBytes + p * kTnPanel, &tmap_x, &full[stage], kt * BK2 + p * 64, m);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = ptx::idesc_bf16_f32(128, kcols, 1, 1);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < n_it; ++it) {
      ptx::mbar_wait(&full[stage], phase);
      ptx::tc_fence_after();
      // MN-major: 64-element panels kTnPanel apart (LBO), 8-row groups 1024 B apart (SBO)
      const uint64_t adesc = ptx::smem_desc_sw128(ptx::smem_u32(sA + (size_t)stage * kTnABytes), kTnPanel, 1024);
      const uint64_t bdesc = ptx::smem_desc_sw128(ptx::smem_u32(sB + (size_t)stage * kTnBBytes), kTnPanel, 1024);
      if (ptx::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < kTnRows / 16; ++kk)
          ptx::umma_ss(tmem_base, adesc + (uint32_t)(kk * 2048 >> 4), bdesc + (uint32_t)(kk * 2048 >> 4), idesc, (it > 0 || kk > 0) ? 1u : 0u);
        ptx::umma_commit(&empty[stage]);
        if (it == n_it - 1) ptx::umma_commit(tfull);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else {
    const int quarter = warp & 3;
    const int lrow = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int n = nt * 128 + lrow;
    float* o = partial + ((size_t)blockIdx.y * N + n) * K + (size_t)kt * BK2;
    if (n_it > 0) {
      ptx::mbar_wait(tfull, 0);
      ptx::tc_fence_after();
    }
    for (int c0 = 0; c0 < kcols; c0 += 32) {
      uint32_t r[32];
      if (n_it > 0) {
        ptx::tmem_ld32(lane_base + (uint32_t)c0, r);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      if (n < N) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(o + c0 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<256>(tmem_base);
}

// out[i] = sum_s partial[s][i] (+ addend[i]), fixed order
__global__ void __launch_bounds__(256) reduce_splits_kernel(const float* __restrict__ partial, int splits, int64_t n4, const float* __restrict__ addend,
                                                            float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 a = addend ? reinterpret_cast<const float4*>(addend)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
      const float4 p = reinterpret_cast<const float4*>(partial)[(int64_t)s * n4 + i];
      a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
    }
    reinterpret_cast<float4*>(out)[i] = a;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 [rows, cols] row-major (leading dimension ld elements); box = 64 columns (128 B) x box_rows rows, 128-byte swizzle, OOB -> 0
static int make_tmap(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return BMKG_ERR_DRIVER;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? BMKG_OK : BMKG_ERR_DRIVER;
}

static int tn_splits(int64_t M, int tiles) {
  int s = (2 * kNumSMs) / (tiles > 0 ? tiles : 1);
  const int64_t max_by_rows = ceil_div(M, 16 * kTnRows);   // at least 1024 rows per split: the partial tiles cost N*K*4 bytes each
  if (s > max_by_rows) s = (int)max_by_rows;
  return s < 1 ? 1 : s;
}

}  // namespace gemm
}  // namespace bmkg

using namespace bmkg;
using namespace bmkg::gemm;

extern "C" {

int bmkg_linear_supported(int64_t M, int N, int K) { return (M > 0 && N >= 16 && N % 16 == 0 && K >= 64 && K % 64 == 0) ? 1 : 0; }

int bmkg_linear_nt(const void* a_bf16, const void* b_bf16, const float* bias, int64_t M, int N, int K, int elu, int out_f32, void* out,
                   const float* att_src, const float* att_dst, int heads, float* a_src, float* a_dst, void* stream) {
  BMKG_REQUIRE(a_bf16 && b_bf16 && out && bmkg_linear_supported(M, N, K) && M < (1ll << 31), BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(aligned16(a_bf16) && aligned16(b_bf16) && aligned16(out) && aligned16(bias), BMKG_ERR_MISALIGNED);
  const int BN = N >= kMaxBN ? kMaxBN : N;
  if (att_src) {
    BMKG_REQUIRE(att_dst && a_src && a_dst && heads >= 1 && heads <= 4 && N % heads == 0 && (N / heads) % 2 == 0, BMKG_ERR_BAD_ARG);
    BMKG_REQUIRE(N <= kMaxBN && !out_f32, BMKG_ERR_UNSUPPORTED);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap ta, tb;
  int rc = make_tmap(&ta, a_bf16, M, K, K, kBM);
  if (rc != BMKG_OK) return rc;
  rc = make_tmap(&tb, b_bf16, N, K, K, BN);
  if (rc != BMKG_OK) return rc;
  const int64_t tiles = ceil_div(M, kBM) * ceil_div(N, BN);
  const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  if (cudaFuncSetAttribute(gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess) return BMKG_ERR_LAUNCH;
  Epilogue ep{bias, elu, out_f32, att_src, att_dst, a_src, a_dst, heads};
  gemm_nt_kernel<<<grid, kThreads, kSmemBytes, st>>>(ta, tb, (int)M, N, K, BN, ep, out);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

size_t bmkg_linear_tn_workspace_bytes(int64_t M, int N, int K) {
  if (!(M > 0 && N >= 8 && N % 8 == 0 && K >= 64 && K % 64 == 0)) return 0;
  const int BK2 = K >= 256 ? 256 : K;
  const int tiles = (int)(ceil_div(N, 128) * ceil_div(K, BK2));
  return (size_t)tn_splits(M, tiles) * N * K * sizeof(float);
}

int bmkg_linear_tn(const void* g_bf16, const void* x_bf16, const float* addend, int64_t M, int N, int K, float* out, void* ws,
                   size_t ws_bytes, void* stream) {
  BMKG_REQUIRE(g_bf16 && x_bf16 && out && M > 0 && M < (1ll << 31), BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(N >= 8 && N % 8 == 0 && K >= 64 && K % 64 == 0, BMKG_ERR_UNSUPPORTED);
  BMKG_REQUIRE(aligned16(g_bf16) && aligned16(x_bf16) && aligned16(out) && aligned16(addend), BMKG_ERR_MISALIGNED);
  BMKG_REQUIRE(ws && ws_bytes >= bmkg_linear_tn_workspace_bytes(M, N, K), BMKG_ERR_WORKSPACE);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int BK2 = K >= 256 ? 256 : K;
  const int tiles = (int)(ceil_div(N, 128) * ceil_div(K, BK2));
  const int splits = tn_splits(M, tiles);
  const int rows_per_split = (int)(ceil_div(ceil_div(M, splits), kTnRows) * kTnRows);
  CUtensorMap tg, tx;
  int rc = make_tmap(&tg, g_bf16, M, N, N, kTnRows);
  if (rc != BMKG_OK) return rc;
  rc = make_tmap(&tx, x_bf16, M, K, K, kTnRows);
  if (rc != BMKG_OK) return rc;
  if (cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTnSmemBytes) != cudaSuccess) return BMKG_ERR_LAUNCH;
  gemm_tn_kernel<<<dim3(tiles, splits), kThreads, kTnSmemBytes, st>>>(tg, tx, (int)M, N, K, BK2, rows_per_split, static_cast<float*>(ws));
  const int64_t n4 = (int64_t)N * K / 4;
  const int rgrid = (int)(ceil_div(n4, 256) < 4 * kNumSMs ? ceil_div(n4, 256) : 4 * kNumSMs);
  reduce_splits_kernel<<<rgrid, 256, 0, st>>>(static_cast<const float*>(ws), splits, n4, addend, out);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

}  // extern "C"
