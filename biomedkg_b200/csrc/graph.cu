// G1/G2 - graph indexing for the GCL step.
//
// The reference has no CSR: PyG's GCNConv runs on the COO edge_index built at
// biomedkg/data/dataset/_base.py:80-86 and re-derives gcn_norm on every layer
// call (encoder.py:155,160).  Here the edge list is sorted ONCE per edge_index
// (bmkg_edge_sort: stable LSD radix sort on the key major<<b | minor) and each
// augmented view's canonical CSR/CSC - dropout_edge mask applied, existing
// self-loops removed, one self-loop per node inserted (PyG gcn_norm /
// add_remaining_self_loops) - is derived from the sorted parent by a stable
// stream compaction (bmkg_csr_filter).  A stable sort of a sub-sequence equals
// the sub-sequence of the stable sort, so the result is bit-identical to
// sorting the view's own edge list (oracle/pyg.py:canonical_csr).
//
// Integer work, HBM-bound; algorithmic bytes: 24*E + 4*(N+1) (+E for a mask).
#include "common.cuh"
#include "scan.cuh"
#include "../../include/bmkg_b200.h"

namespace bmkg {

// flag of sorted position k: edge kept by the view mask and not a self-loop
struct SortedFlag {
  const int32_t* major;
  const int32_t* minor;
  const int32_t* perm;
  const uint8_t* keep;  // may be null = keep all
  __device__ int operator()(int64_t k) const {
    if (major[k] == minor[k]) return 0;
    return keep ? (keep[perm[k]] != 0) : 1;
  }
};
// same flag, recorded as a byte while the scan's first pass evaluates it: the keep[perm[k]] lookup is a random 1-byte gather
// (a 32-byte sector per edge), so it is done once and the down-sweep and the compaction read the byte / the scanned positions
struct SortedFlagStore {
  SortedFlag f;
  uint8_t* out;
  __device__ int operator()(int64_t k) const {
    const int v = f(k);
    out[k] = (uint8_t)v;
    return v;
  }
};
// flag of original edge e
struct OrigFlag {
  const int64_t* src;
  const int64_t* dst;
  const uint8_t* keep;
  __device__ int operator()(int64_t e) const {
    if (src[e] == dst[e]) return 0;
    return keep ? (keep[e] != 0) : 1;
  }
};

// ---------------------------------------------------------------------------
// stable LSD radix sort of (key64, idx32), 8 bits per pass
// ---------------------------------------------------------------------------
constexpr int kRadixThreads = 256;
constexpr int kRadixItems = 16;
constexpr int kRadixTile = kRadixThreads * kRadixItems;  // 4096 keys per CTA
constexpr int kRadixBins = 256;

__global__ void __launch_bounds__(256) make_keys_kernel(const int64_t* __restrict__ ei, int64_t E, int by_src, int bits,
                                                        uint64_t* __restrict__ keys, uint32_t* __restrict__ idx) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  uint64_t s = (uint64_t)ei[e], d = (uint64_t)ei[E + e];
  uint64_t major = by_src ? s : d, minor = by_src ? d : s;
  keys[e] = (major << bits) | minor;
  idx[e] = (uint32_t)e;
}

__global__ void __launch_bounds__(kRadixThreads) radix_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift,
                                                                   int* __restrict__ tile_hist, int num_tiles) {
  __shared__ int h[kRadixBins];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kRadixTile;
#pragma unroll
  for (int j = 0; j < kRadixItems; ++j) {
    int64_t i = base + j * kRadixThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255], 1);  // integer: order-independent result
  }
  __syncthreads();
  tile_hist[(int64_t)threadIdx.x * num_tiles + blockIdx.x] = h[threadIdx.x];
}

// One pass of the stable scatter.  Keys are first ranked inside the tile (per-warp match_any ranking, warps and 32-key
// groups in input order => stable) and staged in shared memory in (digit, input order); the write-out then walks the staged
// tile linearly, so consecutive threads write consecutive global addresses inside each digit's run (16 keys = 128 B on
// average) instead of one 8-byte sector write per key.
constexpr size_t kRadixScatterSmem = (size_t)kRadixTile * (sizeof(uint64_t) + sizeof(uint32_t));

__global__ void __launch_bounds__(kRadixThreads, 3) radix_scatter_kernel(const uint64_t* __restrict__ keys_in,
                                                                      const uint32_t* __restrict__ idx_in,
                                                                      uint64_t* __restrict__ keys_out,
                                                                      uint32_t* __restrict__ idx_out, int64_t n, int shift,
                                                                      const int* __restrict__ tile_off, int num_tiles) {
  // cnt[w][d]: tile-local output cursor of digit d for warp w; bin 256 collects out-of-range lanes
  __shared__ int cnt[kRadixThreads / 32][kRadixBins + 1];
  __shared__ int gofs[kRadixBins];   // global position of staged slot i with digit d = gofs[d] + i
  __shared__ int sw[33];
  extern __shared__ __align__(16) unsigned char stage_raw[];
  uint64_t* skey = reinterpret_cast<uint64_t*>(stage_raw);
  uint32_t* sidx = reinterpret_cast<uint32_t*>(stage_raw + (size_t)kRadixTile * sizeof(uint64_t));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (kRadixThreads / 32) * (kRadixBins + 1); i += kRadixThreads) (&cnt[0][0])[i] = 0;
  __syncthreads();

  // each warp owns 512 consecutive keys and walks them 32 at a time, so ranks follow input order
  const int64_t tbase = (int64_t)blockIdx.x * kRadixTile;
  const int64_t wbase = tbase + (int64_t)warp * (kRadixItems * 32);
  uint64_t k[kRadixItems];
  int info[kRadixItems];  // digit (9 bits) | rank among the warp's equal digits << 9 | group size << 15 | leader << 22
  const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int it = 0; it < kRadixItems; ++it) {
    int64_t i = wbase + it * 32 + lane;
    bool valid = i < n;
    k[it] = valid ? keys_in[i] : 0;
    info[it] = valid ? (int)((k[it] >> shift) & 255) : kRadixBins;
  }
#pragma unroll
  for (int it = 0; it < kRadixItems; ++it) {
    // lanes holding the same digit, from 9 ballots (match.any is far slower than the vote path)
    const int dg = info[it];
    unsigned peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 9; ++b) {
      const bool bit = (dg >> b) & 1;
      const unsigned vote = __ballot_sync(0xffffffffu, bit);
      peers &= bit ? vote : ~vote;
    }
    const int leader = (lane == __ffs(peers) - 1) ? 1 : 0;
    const int group = __popc(peers);
    info[it] = dg | (__popc(peers & lt_mask) << 9) | (group << 15) | (leader << 22);
    if (leader) cnt[warp][dg] += group;
    __syncwarp();
  }
  __syncthreads();
  {
    const int dg = threadIdx.x;  // one thread per digit: exclusive scan over warps, then over digits
    int run = 0;
#pragma unroll
    for (int w = 0; w < kRadixThreads / 32; ++w) {
      int c = cnt[w][dg];
      cnt[w][dg] = run;
      run += c;
    }
    int tile_total;
    const int start = block_exclusive_scan(run, sw, tile_total);  // first staged slot of digit dg
    gofs[dg] = tile_off[(int64_t)dg * num_tiles + blockIdx.x] - start;
#pragma unroll
    for (int w = 0; w < kRadixThreads / 32; ++w) cnt[w][dg] += start;
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < kRadixItems; ++it) {
    const int dg = info[it] & 511;
    const int base = cnt[warp][dg];
    __syncwarp();
    if ((info[it] >> 22) & 1) cnt[warp][dg] = base + ((info[it] >> 15) & 127);
    __syncwarp();
    if (dg != kRadixBins) {
      const int pos = base + ((info[it] >> 9) & 63);
      skey[pos] = k[it];
      sidx[pos] = idx_in[wbase + it * 32 + lane];   // payload fetched late: 16 fewer live registers through the ranking
    }
  }
  __syncthreads();
  const int tile_n = (int)((n - tbase) < (int64_t)kRadixTile ? (n - tbase) : (int64_t)kRadixTile);
  for (int i = threadIdx.x; i < tile_n; i += kRadixThreads) {
    const uint64_t key = skey[i];
    const int pos = gofs[(int)((key >> shift) & 255)] + i;
    keys_out[pos] = key;
    idx_out[pos] = sidx[i];
  }
}

__global__ void __launch_bounds__(256) unpack_sorted_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ idx,
                                                            int64_t E, int bits, int32_t* __restrict__ major,
                                                            int32_t* __restrict__ minor, int32_t* __restrict__ perm) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= E) return;
  const uint64_t mmask = (1ull << bits) - 1ull;
  uint64_t key = keys[k];
  major[k] = (int32_t)(key >> bits);
  minor[k] = (int32_t)(key & mmask);
  perm[k] = (int32_t)idx[k];
}

// rowptr_raw[r] = first sorted position whose major >= r (r in [0, N])
__global__ void __launch_bounds__(256) rowptr_kernel(const int32_t* __restrict__ major, int64_t E, int64_t N,
                                                     int32_t* __restrict__ rowptr) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > N) return;
  int64_t lo = 0, hi = E;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if ((int64_t)major[mid] < r) lo = mid + 1; else hi = mid;
  }
  rowptr[r] = (int32_t)lo;
}

// selfsplit[i] = first sorted position in row i whose minor >= i (row end if none)
__global__ void __launch_bounds__(256) selfsplit_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ minor,
                                                        int64_t N, int32_t* __restrict__ split) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int lo = rowptr[i], hi = rowptr[i + 1];
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (minor[mid] < (int32_t)i) lo = mid + 1; else hi = mid;
  }
  split[i] = lo;
}

// ---------------------------------------------------------------------------
// per-view compaction
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) filter_rows_kernel(const int32_t* __restrict__ rowptr_raw, const int32_t* __restrict__ split,
                                                          const int* __restrict__ pos, int64_t N, int64_t E,
                                                          int32_t* __restrict__ rowptr, int32_t* __restrict__ colind,
                                                          int32_t* __restrict__ perm, float* __restrict__ dis,
                                                          int32_t* __restrict__ nnz_out, int32_t* __restrict__ hub_rows) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int start = pos[rowptr_raw[i]] + (int)i;
  const int end = pos[rowptr_raw[i + 1]] + (int)i + 1;
  const int sp = pos[split[i]] + (int)i;
  rowptr[i] = start;
  colind[sp] = (int32_t)i;
  if (perm) perm[sp] = pos[E] + (int32_t)i;
  if (dis) dis[i] = 1.0f / sqrtf((float)(end - start));
  if (hub_rows && end - start > kHubThreshold) atomicAdd(hub_rows, 1);  // integer count: order-independent
  if (i == N - 1) {
    rowptr[N] = end;
    if (nnz_out) *nnz_out = end;
  }
}

__global__ void __launch_bounds__(256) filter_edges_kernel(SortedFlag flag, const int* __restrict__ pos,
                                                           const int* __restrict__ rank_orig, int64_t E,
                                                           int32_t* __restrict__ colind, int32_t* __restrict__ perm) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= E) return;
  if (pos[k + 1] == pos[k]) return;   // dropped (pos is the exclusive scan of the flags, E + 1 entries)
  const int row = flag.major[k], c = flag.minor[k];
  const int o = pos[k] + row + (c > row ? 1 : 0);
  colind[o] = c;
  if (perm) perm[o] = rank_orig[flag.perm[k]];
}

// hub_info[1 + c] = row that contains edge c * kHubSeg of the view's CSR (or -1 past the end): lets the split-row pre-pass of the
// aggregation kernels find its rows without a serial binary search per CTA
__global__ void __launch_bounds__(256) chunk_rows_kernel(const int32_t* __restrict__ rowptr, int64_t N, int nchunks,
                                                         int32_t* __restrict__ hub_info) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchunks) return;
  const int x = c * kHubSeg;
  if (x >= rowptr[N]) { hub_info[1 + c] = -1; return; }
  int lo = 0, hi = (int)N;  // last row r with rowptr[r] <= x
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (rowptr[mid] <= x) lo = mid; else hi = mid;
  }
  hub_info[1 + c] = lo;
}

__global__ void empty_graph_kernel(int64_t N, int32_t* rowptr) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= N) rowptr[i] = 0;
}

static int key_bits(int64_t N) {
  int b = 1;
  while ((1ll << b) < N) ++b;
  return b;
}

}  // namespace bmkg

using namespace bmkg;

extern "C" {

size_t bmkg_edge_sort_workspace_bytes(int64_t N, int64_t E) {
  const int64_t tiles = ceil_div(E > 0 ? E : 1, kRadixTile);
  WsCarver c(nullptr);
  c.take<uint64_t>(E);
  c.take<uint64_t>(E);
  c.take<uint32_t>(E);
  c.take<uint32_t>(E);
  c.take<int>((size_t)tiles * kRadixBins + 1);
  c.take<int>((size_t)tiles * kRadixBins + 1);
  c.take<int>(scan_ws_ints((int64_t)tiles * kRadixBins));
  (void)N;
  return c.used();
}

int bmkg_edge_sort(const int64_t* edge_index, int64_t E, int64_t N, int by_src, int32_t* major_sorted, int32_t* minor_sorted,
                   int32_t* perm_sorted, int32_t* rowptr_raw, int32_t* selfsplit, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BMKG_REQUIRE(N > 0 && E >= 0 && E + N < (1ll << 31) - 1, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(rowptr_raw && selfsplit, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(E == 0 || (edge_index && major_sorted && minor_sorted && perm_sorted), BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(ws_bytes >= bmkg_edge_sort_workspace_bytes(N, E), BMKG_ERR_WORKSPACE);
  BMKG_REQUIRE(E == 0 || ws, BMKG_ERR_WORKSPACE);
  if (E == 0) {
    empty_graph_kernel<<<(unsigned)ceil_div(N + 1, 256), 256, 0, st>>>(N, rowptr_raw);
    empty_graph_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(N - 1, selfsplit);
    BMKG_CHECK_LAUNCH();
    return BMKG_OK;
  }
  const int tiles = (int)ceil_div(E, kRadixTile);
  WsCarver c(ws);
  uint64_t* k0 = c.take<uint64_t>(E);
  uint64_t* k1 = c.take<uint64_t>(E);
  uint32_t* i0 = c.take<uint32_t>(E);
  uint32_t* i1 = c.take<uint32_t>(E);
  int* hist = c.take<int>((size_t)tiles * kRadixBins + 1);
  int* offs = c.take<int>((size_t)tiles * kRadixBins + 1);
  int* sws = c.take<int>(scan_ws_ints((int64_t)tiles * kRadixBins));

  if (cudaFuncSetAttribute(radix_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRadixScatterSmem) != cudaSuccess)
    return BMKG_ERR_LAUNCH;
  const int bits = key_bits(N);
  make_keys_kernel<<<(unsigned)ceil_div(E, 256), 256, 0, st>>>(edge_index, E, by_src, bits, k0, i0);
  const int passes = (2 * bits + 7) / 8;
  for (int p = 0; p < passes; ++p) {
    radix_hist_kernel<<<tiles, kRadixThreads, 0, st>>>(k0, E, 8 * p, hist, tiles);
    int rc = exclusive_scan(LoadI32{hist}, (int64_t)tiles * kRadixBins, offs, sws, st);
    if (rc != BMKG_OK) return rc;
    radix_scatter_kernel<<<tiles, kRadixThreads, kRadixScatterSmem, st>>>(k0, i0, k1, i1, E, 8 * p, offs, tiles);
    uint64_t* tk = k0; k0 = k1; k1 = tk;
    uint32_t* ti = i0; i0 = i1; i1 = ti;
  }
  unpack_sorted_kernel<<<(unsigned)ceil_div(E, 256), 256, 0, st>>>(k0, i0, E, bits, major_sorted, minor_sorted, perm_sorted);
  rowptr_kernel<<<(unsigned)ceil_div(N + 1, 256), 256, 0, st>>>(major_sorted, E, N, rowptr_raw);
  selfsplit_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(rowptr_raw, minor_sorted, N, selfsplit);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

size_t bmkg_csr_filter_workspace_bytes(int64_t N, int64_t E) {
  WsCarver c(nullptr);
  c.take<int>(E + 1);
  c.take<int>(E + 1);
  c.take<int>(scan_ws_ints(E));
  c.take<uint8_t>(E + 1);
  (void)N;
  return c.used();
}

int bmkg_csr_filter(const int32_t* major_sorted, const int32_t* minor_sorted, const int32_t* perm_sorted,
                    const int32_t* rowptr_raw, const int32_t* selfsplit, const uint8_t* keep, const int64_t* edge_index,
                    int64_t E, int64_t N, int32_t* rowptr, int32_t* colind, int32_t* perm, float* dis, int32_t* nnz_out,
                    int32_t* hub_rows_out, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BMKG_REQUIRE(N > 0 && E >= 0 && E + N < (1ll << 31) - 1, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(rowptr_raw && selfsplit && rowptr && colind, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(E == 0 || (major_sorted && minor_sorted && perm_sorted), BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(!perm || E == 0 || edge_index, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(ws && ws_bytes >= bmkg_csr_filter_workspace_bytes(N, E), BMKG_ERR_WORKSPACE);
  WsCarver c(ws);
  int* pos = c.take<int>(E + 1);
  int* rank = c.take<int>(E + 1);
  int* sws = c.take<int>(scan_ws_ints(E));
  uint8_t* fbytes = c.take<uint8_t>(E + 1);
  SortedFlag sf{major_sorted, minor_sorted, perm_sorted, keep};
  int rc = exclusive_scan2(SortedFlagStore{sf, fbytes}, LoadU8{fbytes}, E, pos, sws, st);
  if (rc != BMKG_OK) return rc;
  if (perm && E > 0) {
    rc = exclusive_scan(OrigFlag{edge_index, edge_index + E, keep}, E, rank, sws, st);
    if (rc != BMKG_OK) return rc;
  }
  if (hub_rows_out) cudaMemsetAsync(hub_rows_out, 0, sizeof(int32_t), st);
  filter_rows_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(rowptr_raw, selfsplit, pos, N, E, rowptr, colind, perm, dis,
                                                                  nnz_out, hub_rows_out);
  if (E > 0)
    filter_edges_kernel<<<(unsigned)ceil_div(E, 256), 256, 0, st>>>(sf, pos, rank, E, colind, perm);
  if (hub_rows_out) {
    const int nchunks = (int)ceil_div(E + N, kHubSeg);
    chunk_rows_kernel<<<(unsigned)ceil_div(nchunks, 256), 256, 0, st>>>(rowptr, N, nchunks, hub_rows_out);
  }
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

int64_t bmkg_hub_info_len(int64_t N, int64_t E) { return 1 + ceil_div(E + N, kHubSeg); }

}  // extern "C"
