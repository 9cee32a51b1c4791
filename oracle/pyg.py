"""Restatement of the torch_geometric 2.5.3 functions the GCL path calls.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pure torch, CPU, any float
dtype (tests use fp64 / fp32).  Every function names the reference call site
it serves; the PyG semantics are those written out in SURVEY.md Appendix A.
Randomness is always injectable (``rand=``) so the CUDA path and the oracle
can be fed identical draws.
"""
from __future__ import annotations

import math

import numpy as np
import torch


# --------------------------------------------------------------------------
# initialisers (torch_geometric.nn.inits)
# --------------------------------------------------------------------------
def glorot_(t: torch.Tensor) -> torch.Tensor:
    """PyG ``inits.glorot``: U(-a, a), a = sqrt(6 / (size(-2) + size(-1))).
    Used by GCNConv.lin / GATConv.lin, att_src, att_dst (encoder.py:138-143)."""
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        return t.uniform_(-a, a)


def uniform_(size: int, t: torch.Tensor) -> torch.Tensor:
    """PyG ``inits.uniform(size, value)``: U(-1/sqrt(size), 1/sqrt(size)).
    Called on DGI.project.weight at model/gcl.py:13."""
    b = 1.0 / math.sqrt(size)
    with torch.no_grad():
        return t.uniform_(-b, b)


# --------------------------------------------------------------------------
# augmentations (torch_geometric.utils) - model/gcl.py:40-43, 75-76
# --------------------------------------------------------------------------
def dropout_edge(edge_index: torch.Tensor, p: float = 0.5, rand: torch.Tensor | None = None):
    """``dropout_edge(edge_index, p)`` with PyG defaults force_undirected=False,
    training=True (the reference never passes ``training``).  Keeps edge e iff
    rand[e] >= p, order preserved.  Returns (edge_index[:, mask], mask)."""
    if p < 0.0 or p > 1.0:
        raise ValueError(f"Dropout probability has to be between 0 and 1 (got {p})")
    E = edge_index.size(1)
    if rand is None:
        rand = torch.rand(E, device=edge_index.device)
    mask = rand >= p
    return edge_index[:, mask], mask


def mask_feature(x: torch.Tensor, p: float = 0.5, mode: str = "col", rand: torch.Tensor | None = None):
    """``mask_feature(x, p, mode)`` with fill_value=0, training=True.  The GCL
    path only uses mode="all" (model/gcl.py:40-41,75): element-wise mask
    rand_like(x) >= p, no 1/(1-p) rescale.  'col'/'row' kept for completeness."""
    if p < 0.0 or p > 1.0:
        raise ValueError(f"Masking ratio has to be between 0 and 1 (got {p})")
    assert x.dim() == 2, "mask_feature requires a 2-D feature matrix"
    if mode == "all":
        if rand is None:
            rand = torch.rand_like(x)
        mask = rand >= p
    elif mode == "col":
        if rand is None:
            rand = torch.rand(1, x.size(1), device=x.device)
        mask = rand.view(1, -1) >= p
    elif mode == "row":
        if rand is None:
            rand = torch.rand(x.size(0), 1, device=x.device)
        mask = rand.view(-1, 1) >= p
    else:
        raise ValueError(mode)
    return x.masked_fill(~mask, 0.0), mask


# --------------------------------------------------------------------------
# gcn_norm / GCNConv  (encoder.py:138-143,155,160 -> PyG GCNConv.forward)
# --------------------------------------------------------------------------
def add_remaining_self_loops(edge_index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """Existing self-loops are removed and exactly one self-loop per node is
    appended at the end (unweighted case of PyG add_remaining_self_loops);
    duplicate edges are kept.  Appendix A.1 step 1."""
    row, col = edge_index[0], edge_index[1]
    keep = row != col
    loop = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index[:, keep], torch.stack([loop, loop])], dim=1)


def gcn_norm(edge_index: torch.Tensor, num_nodes: int, dtype=torch.float32):
    """Returns (edge_index', weight) with weight = dis[row] * dis[col],
    dis = in-degree(incl. self-loop, multi-edges counted) ** -0.5, inf -> 0."""
    ei = add_remaining_self_loops(edge_index, num_nodes)
    row, col = ei[0], ei[1]
    w = torch.ones(ei.size(1), dtype=dtype, device=ei.device)
    deg = torch.zeros(num_nodes, dtype=dtype, device=ei.device).scatter_add_(0, col, w)
    dis = deg.pow(-0.5)
    dis = dis.masked_fill(dis == float("inf"), 0.0)
    return ei, dis[row] * w * dis[col]


def gcn_conv(x: torch.Tensor, edge_index: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None):
    """GCNConv.forward exactly as PyG executes it on a COO edge_index:
    gcn_norm -> lin (no bias) -> gather [E',C] -> scatter_add over targets -> + bias.
    This is the 'as-written' form that bench.py times as the CPU reference."""
    N = x.size(0)
    ei, w = gcn_norm(edge_index, N, dtype=x.dtype)
    xw = x @ weight.t()
    msg = w.unsqueeze(-1) * xw.index_select(0, ei[0])
    out = torch.zeros(N, weight.size(0), dtype=x.dtype, device=x.device)
    out.scatter_add_(0, ei[1].unsqueeze(-1).expand_as(msg), msg)
    if bias is not None:
        out = out + bias
    return out


def gcn_dense_adj(edge_index: torch.Tensor, num_nodes: int, dtype=torch.float64) -> torch.Tensor:
    """Dense A_hat = D^-1/2 (A' + I) D^-1/2 (row = target), an independent
    formulation used by the tests to cross-check gcn_conv on tiny graphs."""
    A = torch.zeros(num_nodes, num_nodes, dtype=dtype)
    for s, d in edge_index.t().tolist():
        if s != d:
            A[d, s] += 1.0
    A += torch.eye(num_nodes, dtype=dtype)
    deg = A.sum(1)
    dis = deg.pow(-0.5)
    return dis[:, None] * A * dis[None, :]


# --------------------------------------------------------------------------
# GATConv (extension - SURVEY.md Appendix A.6; no reference call site on the
# GCL path, nearest is RGAT at encoder.py:62-121)
# --------------------------------------------------------------------------
def gat_conv(x, edge_index, lin_weight, att_src, att_dst, bias, heads: int = 1, negative_slope: float = 0.2):
    """PyG GATConv(in, out, heads=H, concat=True, dropout=0, add_self_loops=True)."""
    N = x.size(0)
    H = heads
    C = lin_weight.size(0) // H
    xh = (x @ lin_weight.t()).view(N, H, C)
    a_s = (xh * att_src.view(1, H, C)).sum(-1)
    a_d = (xh * att_dst.view(1, H, C)).sum(-1)
    ei = add_remaining_self_loops(edge_index, N)
    row, col = ei[0], ei[1]
    e = torch.nn.functional.leaky_relu(a_s[row] + a_d[col], negative_slope)
    emax = torch.full((N, H), float("-inf"), dtype=x.dtype).scatter_reduce_(
        0, col.unsqueeze(-1).expand_as(e), e, reduce="amax", include_self=True)
    e = (e - emax[col]).exp()
    esum = torch.zeros(N, H, dtype=x.dtype).scatter_add_(0, col.unsqueeze(-1).expand_as(e), e)
    alpha = e / (esum[col] + 1e-16)
    msg = alpha.unsqueeze(-1) * xh[row]
    out = torch.zeros(N, H, C, dtype=x.dtype)
    out.scatter_add_(0, col.view(-1, 1, 1).expand_as(msg), msg)
    out = out.reshape(N, H * C)
    if bias is not None:
        out = out + bias
    return out


# --------------------------------------------------------------------------
# canonical CSR / CSC  (SURVEY.md Appendix A.8 - defines "bit-exact")
# --------------------------------------------------------------------------
def canonical_csr(edge_index_prime: torch.Tensor, num_nodes: int, by: str = "dst"):
    """Canonical compressed form of an edge list (already containing its
    self-loops): key = major * N + minor, perm = stable argsort(key),
    rowptr = [0, cumsum(bincount(major))], colind = minor[perm].
    by="dst": CSR over targets (forward aggregation); by="src": CSC."""
    src = edge_index_prime[0].to(torch.int64)
    dst = edge_index_prime[1].to(torch.int64)
    major, minor = (dst, src) if by == "dst" else (src, dst)
    key = major * num_nodes + minor
    perm = torch.sort(key, stable=True).indices
    counts = torch.bincount(major, minlength=num_nodes)
    rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(counts, 0)
    return rowptr.to(torch.int32), minor[perm].to(torch.int32), perm.to(torch.int32)


def canonical_csr_numpy(edge_index_prime: np.ndarray, num_nodes: int, by: str = "dst"):
    """numpy twin of canonical_csr for full-size (tens of millions of edges) checks."""
    src = edge_index_prime[0].astype(np.int64)
    dst = edge_index_prime[1].astype(np.int64)
    major, minor = (dst, src) if by == "dst" else (src, dst)
    key = major * num_nodes + minor
    perm = np.argsort(key, kind="stable")
    rowptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(np.bincount(major, minlength=num_nodes), out=rowptr[1:])
    return rowptr.astype(np.int32), minor[perm].astype(np.int32), perm.astype(np.int32)


def view_graph(edge_index: torch.Tensor, num_nodes: int, keep: torch.Tensor | None = None):
    """edge list a GCNConv layer sees for one augmented view: optional
    dropout_edge mask (A.2), then remove/append self-loops (A.1)."""
    ei = edge_index if keep is None else edge_index[:, keep]
    return add_remaining_self_loops(ei, num_nodes)
