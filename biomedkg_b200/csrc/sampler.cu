// S1 - neighbour sampling for the reference's mini-batch regime.
//
// The reference trains on torch_geometric NeighborLoader batches built on the CPU (biomedkg/data_module.py:81-99:
// num_neighbors=[30]*3, batch_size seeds, shuffle=True; :71-79 uses [-1]).  Per hop, every node added in the previous hop
// draws min(in-degree, fanout) of its in-edges without replacement; the sampled edges are the batch's only edges; newly
// reached sources join the node list in order of first appearance; edges are relabelled to batch-local ids
// (row = source, col = target).  Here one hop is four device passes over the parent graph's raw CSR (the sorted edge list
// bmkg_edge_sort already produced): count -> scan -> pick -> relabel.
//
//  * pick: one warp per frontier node, fanout <= 32.  Lane i draws position hash(seed, hop, node, i, attempt) * deg >> 32;
//    a lane whose draw equals a lower lane's redraws (attempt + 1) until all are distinct - the process treats all
//    positions alike, so the result is a uniform random subset.  When deg <= 2*fanout the warp draws the deg - fanout
//    positions to LEAVE OUT instead, so every draw succeeds with probability >= 1/2.  Selected edges are emitted in CSR
//    order.  Counter-based: the same (seed, hop, node) always samples the same edges, whatever the batch or the launch.
//  * relabel: first appearance = atomicMin of the entry index per source node (order-independent result), flags ->
//    exclusive scan -> new local ids; a persistent local_id[N] map (all -1 between batches) is restored by bmkg_sample_reset.
//
// Integer work, bit-exact against the numpy restatement (oracle/sampler.py).  HBM-bound on random 4-byte gathers.
// Algorithmic bytes per hop: F*(8 + 4) + T*(4 + 4 + 8 + 8 + 4) with F frontier nodes and T sampled edges.
#include "common.cuh"
#include "scan.cuh"
#include "../../include/bmkg_b200.h"

namespace bmkg {

constexpr int kSampleWarps = 8;
constexpr int kNoPos = 0x7fffffff;

__device__ __forceinline__ int fanout_count(int deg, int fanout) { return (fanout < 0 || deg < fanout) ? deg : fanout; }

struct DegCount {
  const int32_t* rowptr;
  const int32_t* frontier;
  int fanout;
  __device__ int operator()(int64_t i) const {
    const int v = frontier[i];
    return fanout_count(rowptr[v + 1] - rowptr[v], fanout);
  }
};

__device__ __forceinline__ uint32_t sample_draw(uint64_t seed, int hop, int node, int lane, int attempt, int deg) {
  const uint64_t idx = ((uint64_t)(uint32_t)node << 32) | ((uint64_t)(uint32_t)attempt << 8) | (uint64_t)(uint32_t)lane;
  const uint32_t r = hash_u32(seed + 0x9E3779B97F4A7C15ull * (uint64_t)(hop + 1), idx);
  return (uint32_t)(((uint64_t)r * (uint64_t)(uint32_t)deg) >> 32);   // [0, deg)
}

__global__ void __launch_bounds__(kSampleWarps * 32) sample_pick_kernel(const int32_t* __restrict__ rowptr,
                                                                        const int32_t* __restrict__ colind,
                                                                        const int32_t* __restrict__ eperm,
                                                                        const int32_t* __restrict__ frontier, int64_t F, int fanout,
                                                                        const int32_t* __restrict__ off, uint64_t seed, int hop,
                                                                        int64_t frontier_base, int32_t* __restrict__ out_src,
                                                                        int64_t* __restrict__ out_col, int64_t* __restrict__ out_eid) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * kSampleWarps + (threadIdx.x >> 5);
  if (i >= F) return;
  const int v = frontier[i];
  const int beg = rowptr[v], deg = rowptr[v + 1] - beg;
  const int k = fanout_count(deg, fanout);
  const int64_t o = off[i];
  const int64_t col = frontier_base + i;
  if (k == deg) {  // take every in-edge, CSR order
    for (int p = lane; p < deg; p += 32) {
      out_src[o + p] = colind[beg + p];
      out_col[o + p] = col;
      if (out_eid) out_eid[o + p] = eperm[beg + p];
    }
    return;
  }
  // deg > k = fanout (<= 32): draw m distinct positions - the ones to take, or the ones to leave out when deg <= 2k
  const bool exclude = deg <= 2 * k;
  const int m = exclude ? deg - k : k;
  const bool active = lane < m;
  int attempt = 0;
  uint32_t cand = active ? sample_draw(seed, hop, v, lane, 0, deg) : 0xffffffffu - (uint32_t)lane;  // idle lanes: distinct sentinels
  const unsigned lower = (1u << lane) - 1u;
  for (;;) {
    const unsigned same = __match_any_sync(0xffffffffu, cand);
    const bool dup = active && (same & lower) != 0u;
    if (__ballot_sync(0xffffffffu, dup) == 0u) break;
    if (dup) cand = sample_draw(seed, hop, v, lane, ++attempt, deg);
  }
  if (exclude) {  // deg <= 64: bit p of (lo, hi) = position p is left out
    const unsigned lo = __reduce_or_sync(0xffffffffu, (active && cand < 32u) ? (1u << cand) : 0u);
    const unsigned hi = __reduce_or_sync(0xffffffffu, (active && cand >= 32u) ? (1u << (cand - 32u)) : 0u);
    const unsigned take_lo = ~lo & (deg >= 32 ? 0xffffffffu : ((1u << deg) - 1u));
    const unsigned take_hi = deg > 32 ? (~hi & (deg >= 64 ? 0xffffffffu : ((1u << (deg - 32)) - 1u))) : 0u;
    if ((take_lo >> lane) & 1u) {
      const int r = __popc(take_lo & lower);
      out_src[o + r] = colind[beg + lane];
      out_col[o + r] = col;
      if (out_eid) out_eid[o + r] = eperm[beg + lane];
    }
    if ((take_hi >> lane) & 1u) {
      const int r = __popc(take_lo) + __popc(take_hi & lower);
      out_src[o + r] = colind[beg + 32 + lane];
      out_col[o + r] = col;
      if (out_eid) out_eid[o + r] = eperm[beg + 32 + lane];
    }
  } else {  // rank of the lane's position among the m drawn ones
    int r = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const uint32_t cj = __shfl_sync(0xffffffffu, cand, j);
      r += (j < m && cj < cand) ? 1 : 0;
    }
    if (active) {
      out_src[o + r] = colind[beg + (int)cand];
      out_col[o + r] = col;
      if (out_eid) out_eid[o + r] = eperm[beg + (int)cand];
    }
  }
}

// first appearance of every not-yet-labelled source among the T sampled entries
__global__ void __launch_bounds__(256) sample_first_kernel(const int32_t* __restrict__ src, int64_t T,
                                                           const int32_t* __restrict__ local_id, int32_t* __restrict__ first_pos) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= T) return;
  const int s = src[e];
  if (local_id[s] < 0) atomicMin(&first_pos[s], (int)e);  // integer min: order-independent
}

struct NewFlag {
  const int32_t* src;
  const int32_t* local_id;
  const int32_t* first_pos;
  __device__ int operator()(int64_t e) const {
    const int s = src[e];
    return (local_id[s] < 0 && first_pos[s] == (int)e) ? 1 : 0;
  }
};

// new nodes take local ids n_before + rank (rank = position among first appearances); rank[T] = how many there are
__global__ void __launch_bounds__(256) sample_assign_kernel(const int32_t* __restrict__ src, int64_t T,
                                                            const int32_t* __restrict__ rank, int64_t n_before,
                                                            int32_t* __restrict__ local_id, const int32_t* __restrict__ first_pos,
                                                            int32_t* __restrict__ new_nodes, int32_t* __restrict__ new_count) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e == 0) *new_count = rank[T];
  if (e >= T) return;
  const int s = src[e];
  if (rank[e + 1] != rank[e]) {  // this entry is the first appearance of s (flag recorded by the scan, before local_id changes)
    new_nodes[rank[e]] = s;
    local_id[s] = (int)(n_before + rank[e]);
  }
  (void)first_pos;
}

__global__ void __launch_bounds__(256) sample_relabel_kernel(const int32_t* __restrict__ src, int64_t T,
                                                             const int32_t* __restrict__ local_id, int32_t* __restrict__ first_pos,
                                                             int64_t* __restrict__ out_row) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= T) return;
  const int s = src[e];
  out_row[e] = local_id[s];
  first_pos[s] = kNoPos;  // every entry of s writes the same value: restore the scratch map for the next hop
}

__global__ void __launch_bounds__(256) sample_set_ids_kernel(const int32_t* __restrict__ nodes, int64_t n, int32_t* __restrict__ local_id,
                                                             int reset) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) local_id[nodes[i]] = reset ? -1 : (int)i;
}

}  // namespace bmkg

using namespace bmkg;

extern "C" size_t bmkg_sample_workspace_bytes(int64_t max_entries) {
  WsCarver c(nullptr);
  c.take<int>(max_entries + 1);
  c.take<int>(scan_ws_ints(max_entries));
  return c.used();
}

extern "C" int bmkg_sample_count(const int32_t* rowptr, const int32_t* frontier, int64_t F, int fanout, int32_t* off, void* ws,
                                 size_t ws_bytes, void* stream) {
  BMKG_REQUIRE(rowptr && frontier && off && F > 0 && F < (1ll << 31) - 1, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(fanout == -1 || (fanout >= 1 && fanout <= 32), BMKG_ERR_UNSUPPORTED);
  BMKG_REQUIRE(ws && ws_bytes >= bmkg_sample_workspace_bytes(F), BMKG_ERR_WORKSPACE);
  WsCarver c(ws);
  c.take<int>(F + 1);
  int* sws = c.take<int>(scan_ws_ints(F));
  return exclusive_scan(DegCount{rowptr, frontier, fanout}, F, off, sws, static_cast<cudaStream_t>(stream));
}

extern "C" int bmkg_sample_pick(const int32_t* rowptr, const int32_t* colind, const int32_t* eperm, const int32_t* frontier, int64_t F,
                                int fanout, const int32_t* off, uint64_t seed, int hop, int64_t frontier_base, int32_t* out_src,
                                int64_t* out_col, int64_t* out_eid, void* stream) {
  BMKG_REQUIRE(rowptr && colind && frontier && off && out_src && out_col && F > 0, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(!out_eid || eperm, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(fanout == -1 || (fanout >= 1 && fanout <= 32), BMKG_ERR_UNSUPPORTED);
  sample_pick_kernel<<<(unsigned)ceil_div(F, kSampleWarps), kSampleWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      rowptr, colind, eperm, frontier, F, fanout, off, seed, hop, frontier_base, out_src, out_col, out_eid);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

extern "C" int bmkg_sample_relabel(const int32_t* src, int64_t T, int64_t n_before, int32_t* local_id, int32_t* first_pos,
                                   int32_t* new_nodes, int32_t* new_count, int64_t* out_row, void* ws, size_t ws_bytes, void* stream) {
  BMKG_REQUIRE(src && local_id && first_pos && new_nodes && new_count && out_row && T > 0 && T < (1ll << 31) - 1, BMKG_ERR_BAD_ARG);
  BMKG_REQUIRE(ws && ws_bytes >= bmkg_sample_workspace_bytes(T), BMKG_ERR_WORKSPACE);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  WsCarver c(ws);
  int* rank = c.take<int>(T + 1);
  int* sws = c.take<int>(scan_ws_ints(T));
  const unsigned grid = (unsigned)ceil_div(T, 256);
  sample_first_kernel<<<grid, 256, 0, st>>>(src, T, local_id, first_pos);
  int rc = exclusive_scan(NewFlag{src, local_id, first_pos}, T, rank, sws, st);
  if (rc != BMKG_OK) return rc;
  sample_assign_kernel<<<grid, 256, 0, st>>>(src, T, rank, n_before, local_id, first_pos, new_nodes, new_count);
  sample_relabel_kernel<<<grid, 256, 0, st>>>(src, T, local_id, first_pos, out_row);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}

extern "C" int bmkg_sample_set_ids(const int32_t* nodes, int64_t n, int32_t* local_id, int reset, void* stream) {
  BMKG_REQUIRE(nodes && local_id && n > 0, BMKG_ERR_BAD_ARG);
  sample_set_ids_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(nodes, n, local_id, reset);
  BMKG_CHECK_LAUNCH();
  return BMKG_OK;
}
