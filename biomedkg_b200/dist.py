"""Row-sharded GRACE step across the GPUs of one node ("partition nodes by destination row").

Partition.  Nodes are cut into ``world`` contiguous blocks of B rows, B = ceil(N / world) rounded up to 128 (``shard_layout``).
Rank p owns the nodes [pB, (p+1)B) in BOTH views, so

  * encoder: fusion, feature masks, every layer's GEMMs and the projector run on the rank's rows only; each conv layer
    all-gathers its transformed rows (bf16 [N,C]) and aggregates its own destination rows; the backward all-gathers the
    pre-activation gradient and aggregates its own SOURCE rows over the CSC (gather form both ways: no scatter, no atomics);
  * InfoNCE: the stacked operand uses the block-interleaved layout of include/bmkg_b200.h with view block B - one
    all-gather of every rank's [2, B, D] block assembles it, and a rank's rows of both views are one contiguous 128-aligned
    range [2pB, 2(p+1)B).  forward: all-reduce of the [D] column sums (common vector mu), all-gather of the bf16
    deviations Z and of a = mu . d, the row-range tcgen05 kernel, all-gather of the per-row state (4 floats per row), all-reduce of the
    scalar loss share.  backward: the row-range kernel writes dZ of exactly the rows whose h this rank holds - NO collective.
  * parameter gradients are partial sums over the rank's rows -> one flat all-reduce (``allreduce_grads``).

The result is the single-GPU full-graph loss (unlike the reference's DDP, which contrasts per rank mini-batch).  No
collective touches the N x N work.

The compute is injected (``impl``): the default calls the CUDA kernels; the gloo CPU tests inject a torch restatement of
the same row-range math so the partitioning / collective / autograd plumbing is exercised with world_size 2 on CPU.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

ROW_ALIGN = 128
LOG2E = 1.4426950408889634


def shard_layout(num_nodes: int, world: int, align: int = ROW_ALIGN):
    """-> (B, [(n0, n1)] per rank): equal node blocks of B rows, B a multiple of ``align``; trailing ranks may be short / empty."""
    B = (num_nodes + world - 1) // world
    B = max(align, (B + align - 1) // align * align)
    return B, [(min(p * B, num_nodes), min((p + 1) * B, num_nodes)) for p in range(world)]


def node_partition(num_nodes: int, world: int):
    """Node blocks of the row-sharded step (same as ``shard_layout``; kept under the name loaders use)."""
    return shard_layout(num_nodes, world)


class CudaImpl:
    """Kernels from libbmkg_b200.so on this rank's rows."""

    def stats(self, h):
        """h fp32 [n, D] -> (inv_norm [n], column sums of the normalised rows [D])"""
        from .ops import _p, _stream, _ws, call, lib

        n, D = h.shape
        inv = torch.empty(max(n, 1), dtype=torch.float32, device=h.device)
        cs = torch.zeros(D, dtype=torch.float32, device=h.device)
        if n > 0:
            ws = _ws(lib.bmkg_colsum_workspace_bytes(n, D), h.device)
            call("bmkg_l2norm_colsum", _p(h), n, D, _p(inv), _p(cs), _p(ws), ws.numel(), _stream())
        return inv, cs

    def center(self, hs, invs, mu, B, scale):
        """-> (bf16 [2, B, D] deviations, fp32 [2, B] a = mu . d); rows beyond the rank's nodes stay zero"""
        from .ops import _p, _stream, call

        D = hs[0].size(1)
        z = torch.zeros(2, B, D, dtype=torch.bfloat16, device=mu.device)
        a = torch.zeros(2, B, dtype=torch.float32, device=mu.device)
        for v, (h, inv) in enumerate(zip(hs, invs)):
            if h.size(0) > 0:
                call("bmkg_center_scale", _p(h), _p(inv), _p(mu), h.size(0), D, scale, _p(z[v]), _p(a[v]), _stream())
        return z, a

    def fwd_rows(self, Z, A, N, B, r0, r1):
        """Z bf16 [R_all, D], A fp32 [R_all] (gathered) -> (loss share, forward->backward state [R_all, 4] with rows [r0, r1) filled)"""
        from .ops import _p, _stream, _ws, alloc_e_store, call, lib

        D = Z.size(1)
        rp = int(lib.bmkg_infonce_padded_rows(N, B))      # the kernels read a / write t in whole 128-row tiles
        if A.numel() < rp:
            raise ValueError(f"a must hold bmkg_infonce_padded_rows = {rp} entries (zero beyond the stacked rows), got {A.numel()}")
        loss = torch.zeros((), dtype=torch.float32, device=Z.device)
        t = torch.zeros(max(Z.size(0), rp), 4, dtype=torch.float32, device=Z.device)
        self.e_store = None
        if r1 > r0:
            xab = torch.empty(rp, 32, dtype=torch.bfloat16, device=Z.device)       # ext K columns of EVERY row, from the gathered a
            call("bmkg_infonce_ext", _p(A), N, B, _p(xab), _stream())
            ws = _ws(lib.bmkg_infonce_workspace_bytes_rows(N, B, D, r0, r1), Z.device)
            self.e_store = alloc_e_store(N, B, r0, r1, Z.device)      # E = 2^S'' of this rank's rows, kept for the backward if enabled
            call("bmkg_infonce_fwd_rows", _p(Z), _p(A), _p(xab), N, B, D, r0, r1, _p(loss), _p(t), _p(self.e_store), _p(ws), ws.numel(),
                 _stream())
        return loss, t

    def bwd_rows(self, Z, T, mu, g, N, B, r0, r1):
        """-> dZ fp32 [r1 - r0, D] of this range (the kernel addresses dz by global row: pass the buffer shifted by -r0 rows)"""
        from .ops import _p, _stream, _ws_optional, call, lib, release_e_store

        D = Z.size(1)
        if T.size(0) < int(lib.bmkg_infonce_padded_rows(N, B)):
            raise ValueError("the state must hold bmkg_infonce_padded_rows rows (zeros for padding rows)")
        dz = torch.zeros(max(r1 - r0, 1), D, dtype=torch.float32, device=Z.device)
        if r1 > r0:
            e_store = getattr(self, "e_store", None)
            ws = _ws_optional(lib.bmkg_infonce_bwd_workspace_bytes(N, B, D, r0, r1), Z.device) if e_store is None else None
            call("bmkg_infonce_bwd_rows", _p(Z), _p(T), _p(mu), _p(g), _p(e_store), N, B, D, r0, r1,
                 dz.data_ptr() - r0 * D * 4, _p(ws), 0 if ws is None else ws.numel(), _stream())
            release_e_store(getattr(self, "e_store", None))
            self.e_store = None
        return dz

    def norm_bwd(self, h, inv, dz, scale):
        from .ops import _p, _stream, call

        dh = torch.empty_like(h)
        if h.size(0) > 0:
            call("bmkg_l2norm_scale_bwd", _p(h), _p(inv), _p(dz), h.size(0), h.size(1), scale, _p(dh), _stream())
        return dh


class _ShardedInfoNCEFn(torch.autograd.Function):
    """h1_loc, h2_loc: this rank's node rows [n0, n1) of the two projected views; N: total nodes."""

    @staticmethod
    def forward(ctx, h1, h2, N, tau, group, impl):
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        h1, h2 = h1.contiguous(), h2.contiguous()           # fp32 for the CUDA kernels (the CPU test impl runs fp64)
        B, parts = shard_layout(N, world)
        n0, n1 = parts[rank]
        if h1.size(0) != n1 - n0 or h2.size(0) != n1 - n0:
            raise ValueError(f"rank {rank} must pass its node block [{n0}, {n1}) of both views")
        D = h1.size(1)
        dev = h1.device
        scale = math.sqrt(LOG2E / tau)
        inv1, cs1 = impl.stats(h1)
        inv2, cs2 = impl.stats(h2)
        cs = cs1 + cs2
        dist.all_reduce(cs, group=group)                       # every rank must centre with the same common vector
        mu = (cs * (scale / (2.0 * N))).contiguous()
        zb, ab = impl.center((h1, h2), (inv1, inv2), mu, B, scale)
        R_all = world * 2 * B                                  # >= bmkg_infonce_padded_rows(N, B); trailing blocks all zero
        Z = torch.empty(R_all, D, dtype=zb.dtype, device=dev)
        A = torch.empty(R_all, dtype=ab.dtype, device=dev)
        dist.all_gather_into_tensor(Z, zb.view(2 * B, D), group=group)
        dist.all_gather_into_tensor(A, ab.view(2 * B), group=group)
        nblk = (N + B - 1) // B
        r0, r1 = (rank * 2 * B, (rank + 1) * 2 * B) if rank < nblk else (0, 0)
        loss, t = impl.fwd_rows(Z, A, N, B, r0, r1)
        QW = torch.empty((R_all,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)      # the per-row forward -> backward state of every rank
        dist.all_gather_into_tensor(QW, t[rank * 2 * B: (rank + 1) * 2 * B].contiguous(), group=group)
        dist.all_reduce(loss, group=group)
        ctx.save_for_backward(h1, h2, inv1, inv2, Z, QW, mu)
        ctx.meta = (scale, impl, N, B, r0, r1)
        return loss

    @staticmethod
    def backward(ctx, g):
        h1, h2, inv1, inv2, Z, QW, mu = ctx.saved_tensors
        scale, impl, N, B, r0, r1 = ctx.meta
        n = h1.size(0)
        dz = impl.bwd_rows(Z, QW, mu, g.contiguous().to(h1.dtype), N, B, r0, r1)     # rows of THIS rank's nodes: no collective
        dh1 = impl.norm_bwd(h1, inv1, dz[:n], scale)
        dh2 = impl.norm_bwd(h2, inv2, dz[B: B + n] if n > 0 else dz[:0], scale)
        return dh1, dh2, None, None, None, None


def sharded_infonce_local(h1_loc, h2_loc, num_nodes, tau=0.2, group=None, impl=None):
    """InfoNCE of the full graph from this rank's node rows of the two views (see the module docstring)."""
    from . import ops

    if impl is None:
        h1_loc, h2_loc = ops.pad_infonce_dim(h1_loc.float(), h2_loc.float())
        impl = CudaImpl()
    return _ShardedInfoNCEFn.apply(h1_loc, h2_loc, int(num_nodes), float(tau), group, impl)


class _LocalRowsFn(torch.autograd.Function):
    """Replicated [N, D] -> this rank's rows.  Downstream every rank computes ITS rows' share of one global loss, so the
    gradient of the replicated tensor is the concatenation of the ranks' row gradients: backward = all-gather."""

    @staticmethod
    def forward(ctx, full, n0, n1, B, group):
        ctx.meta = (n0, n1, B, group, full.size(0))
        return full[n0:n1].contiguous()

    @staticmethod
    def backward(ctx, g):
        n0, n1, B, group, N = ctx.meta
        world = dist.get_world_size(group)
        blk = torch.zeros(B, g.size(1), dtype=g.dtype, device=g.device)
        blk[: n1 - n0] = g
        out = torch.empty(world * B, g.size(1), dtype=g.dtype, device=g.device)
        dist.all_gather_into_tensor(out, blk, group=group)
        return out[:N], None, None, None, None


def sharded_infonce_loss(h1, h2, tau=0.2, group=None, impl=None):
    """DualBranchContrast(InfoNCE(tau), "L2L", intraview_negs=True)(h1, h2) with the 2N x 2N work split by rows over ``group``.
    h1, h2 are REPLICATED (identical on every rank - e.g. a replicated GAT encoder) and so is the returned loss / gradient."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        from . import ops

        return ops.infonce_loss(h1, h2, tau)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    N = h1.size(0)
    B, parts = shard_layout(N, world)
    n0, n1 = parts[rank]
    return sharded_infonce_local(_LocalRowsFn.apply(h1, n0, n1, B, group), _LocalRowsFn.apply(h2, n0, n1, B, group), N, tau, group, impl)


class ShardedDualBranchContrast(torch.nn.Module):
    """Drop-in for losses.DualBranchContrast that row-shards the InfoNCE over the default process group."""

    def __init__(self, tau: float = 0.2, group=None, impl=None):
        super().__init__()
        self.tau, self.group, self.impl = tau, group, impl

    def forward(self, h1, h2):
        return sharded_infonce_loss(h1, h2, self.tau, self.group, self.impl)


# ---------------------------------------------------------------------------------------------------------------------------
# Row-sharded GCN encoder + full GRACE step
# ---------------------------------------------------------------------------------------------------------------------------
def _all_gather_rows(local: torch.Tensor, num_rows: int, block: int, group) -> torch.Tensor:
    """[n_local, C] blocks of every rank -> [num_rows, C] (NCCL all-gather of equal, zero-padded blocks)."""
    world = dist.get_world_size(group)
    padded = local
    if local.size(0) != block:
        padded = torch.zeros(block, local.size(1), dtype=local.dtype, device=local.device)
        padded[: local.size(0)] = local
    full = torch.empty(world * block, local.size(1), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(full, padded.contiguous(), group=group)
    return full[:num_rows]


class _GatherRowsFn(torch.autograd.Function):
    """All-gather node rows; downstream every rank computes the SAME function of the gathered tensor, so the gradient of the
    gathered tensor is identical on all ranks and the backward is just the local slice (no reduction)."""

    @staticmethod
    def forward(ctx, local, num_rows, block, r0, group):
        ctx.rng = (r0, r0 + local.size(0))
        return _all_gather_rows(local, num_rows, block, group)

    @staticmethod
    def backward(ctx, g):
        return g[ctx.rng[0]: ctx.rng[1]].contiguous(), None, None, None, None


class _ShardedGCNLayerFn(torch.autograd.Function):
    """One GCNConv(+ReLU+dropout) on this rank's destination rows.  forward: X_loc W^T -> all-gather -> CSR aggregation of the
    local rows; backward: ReLU/dropout grad of the local rows -> all-gather -> CSC aggregation of the local (source) rows ->
    local dW (summed over ranks by the caller), local dX.  Gather form in both directions: no scatter, no atomics."""

    @staticmethod
    def forward(ctx, x_loc, weight, bias, view, r0, r1, num_nodes, block, group, relu, drop_p, drop_seed, drop_keep, out_fp32, correct):
        from . import ops

        w16t = ops.weight_forms(weight)[1]
        # the weight-residual correction uses the mean of the LOCAL rows: any common vector restores the rounded-away part
        xw_loc = ops._xw(x_loc, weight, correct and x_loc.size(0) > 0)
        xw = _all_gather_rows(xw_loc, num_nodes, block, group).contiguous()
        hub = view.hub_csr if view.hub_possible else None
        y = ops.gcn_aggregate(view.rowptr, view.colind, view.dis, xw, bias.contiguous(), relu, drop_p, drop_seed, drop_keep, out_fp32,
                              hub_rows=hub, row_range=(r0, r1))
        ctx.meta = (view, r0, r1, num_nodes, block, group, relu, drop_p)
        ctx.save_for_backward(x_loc, w16t, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        from . import ops
        from .ops import _p, _stream, _ws, call, lib

        x_loc, w16t, y = ctx.saved_tensors
        view, r0, r1, num_nodes, block, group, relu, drop_p = ctx.meta
        gy = gy.contiguous()
        n, C = gy.shape
        if relu:
            gy16 = gy if gy.dtype == torch.bfloat16 else gy.to(torch.bfloat16)
            gpre = torch.empty(n, C, dtype=torch.bfloat16, device=gy.device)
            dbias = torch.zeros(C, dtype=torch.float32, device=gy.device)
            if n > 0:
                ws = _ws(lib.bmkg_colsum_workspace_bytes(n, C), gy.device)
                scale = 1.0 / (1.0 - drop_p) if drop_p > 0 else 1.0
                call("bmkg_relu_dropout_bwd", _p(gy16), _p(y), float(scale), n, C, _p(gpre), _p(dbias), _p(ws), ws.numel(), _stream())
        else:
            dbias = ops.colsum(gy.float()) if n > 0 else torch.zeros(C, dtype=torch.float32, device=gy.device)
            gpre = gy if gy.dtype == torch.bfloat16 else gy.to(torch.bfloat16)
        g_full = _all_gather_rows(gpre, num_nodes, block, group).contiguous()
        hub = view.hub_csc if view.hub_possible else None
        dxw = ops.gcn_aggregate(view.csc_rowptr, view.csc_colind, view.dis, g_full, hub_rows=hub, row_range=(r0, r1))
        dw = ops.gemm_tn(dxw, x_loc) if ctx.needs_input_grad[1] else None
        dx = ops.gemm_nt(dxw, w16t) if ctx.needs_input_grad[0] else None
        return (dx, dw, dbias) + (None,) * 12


def sharded_gcn_encoder(encoder, x_loc, view, r0, r1, num_nodes, block, group):
    """GCNEncoder.forward (encoder.py:153-162) on the node rows [r0, r1) of this rank; x_loc bf16 [r1-r0, in]."""
    x = x_loc
    layers = encoder.graph_layers
    for i, layer in enumerate(layers):
        last = i == len(layers) - 1
        p, seed, keep = 0.0, 0, None
        if not last and encoder.drop_out and encoder.training:
            p = 0.2
            # full-size draw in the reference's order (identical masks whatever the world size); an explicit mask (replayed
            # or graph-safe draws) is sliced to this rank's rows, the counter-hash stream is indexed by the global row
            seed, keep = encoder.draws.dropout((num_nodes, layer.out_channels), p, x.device)
            if keep is not None:
                keep = keep[r0:r1].contiguous()
        x = _ShardedGCNLayerFn.apply(x, layer.lin.weight, layer.bias, view, r0, r1, num_nodes, block, group, not last, p, seed, keep,
                                     last, i > 0)
    return x


def sharded_grace_loss(module, x, edge_index, group=None, num_nodes=None):
    """GRACEModule.training_step's loss with the node rows split over the ranks of ``group``.

    Every rank receives the same edge_index and the same parameters / RNG state.  Fusion, feature masks, the GCN layers'
    GEMMs, the projector and the normalisation run on this rank's node block only; each GCN layer all-gathers its transformed
    rows (bf16 [N,256] over NVLink) before aggregating its own destination rows; the InfoNCE rows are the rank's own nodes in
    both views (``sharded_infonce_local``).  The loss is the single-GPU full-graph loss; parameter gradients are partial sums
    that the caller all-reduces (``allreduce_grads``).

    ``x`` is either the full feature tensor (every rank slices its rows) or, with ``num_nodes`` given, only this rank's block
    ``shard_layout(num_nodes, world)[1][rank]`` - a sharded loader then moves 1/world of the features per rank."""
    from . import ops

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    model = module.model
    N = int(num_nodes) if num_nodes is not None else x.size(0)
    block, parts = shard_layout(N, world)
    r0, r1 = parts[rank]
    x_loc = x if num_nodes is not None else x[r0:r1]
    if x_loc.size(0) != r1 - r0:
        raise ValueError("x must hold this rank's node block when num_nodes is given")
    fused = module.fusion_fn(x_loc)                                      # row-local (attention fusion / mean)
    draws = model.draws
    m1 = draws.feature_mask(x.new_empty(N, fused.size(1)), 0.4)          # full-size draws in the reference's order, then sliced:
    m2 = draws.feature_mask(x.new_empty(N, fused.size(1)), 0.4)          # identical masks whatever the world size
    k1 = draws.edge_mask(edge_index, 0.4)
    k2 = draws.edge_mask(edge_index, 0.4)
    sg = ops.sorted_graph(edge_index, N)
    x0, x1, x2 = ops.mask_cast(fused.float().contiguous(), m1[r0:r1].contiguous(), m2[r0:r1].contiguous(),
                               want_plain=not model.skip_unused_view)
    enc = model.encoder
    views = [(x1, sg.view(k1)), (x2, sg.view(k2))]
    if not model.skip_unused_view:
        views.insert(0, (x0, sg.view(None)))       # the reference's unused un-augmented view comes first (dropout draw order)
    # (Tried in round 2: one CUDA stream per view so that a view's all-gathers overlap another view's compute - 32.5 -> 32.4 ms
    # at 4 GPUs, 19.6 -> 19.0 ms at 8, not worth the cross-stream allocator hazards; the passes run back to back.)
    outs = [sharded_gcn_encoder(enc, xv, view, r0, r1, N, block, group) for xv, view in views]
    z1, z2 = outs[-2], outs[-1]
    tau = module.contrast_model.loss.tau if hasattr(module.contrast_model, "loss") else 0.2
    return sharded_infonce_local(model.project(z1), model.project(z2), N, tau, group)


def edge_chunk(num_edges: int, world: int, rank: int):
    """Column range of edge_index a sharded loader reads on ``rank`` (equal chunks; the last may be short)."""
    per = (num_edges + world - 1) // world
    return per, min(rank * per, num_edges), min((rank + 1) * per, num_edges)


def gather_edge_index(chunk: torch.Tensor, num_edges: int, group=None) -> torch.Tensor:
    """Sharded loader, graph side: every rank brings ``per`` columns of the int64 [2, E] edge_index (its ``edge_chunk``,
    zero-padded to ``per``) and the full edge list is assembled over NVLink - E x 16 bytes cross the host's PCIe links ONCE per
    step instead of once per rank (cfg4 on 8 GPUs: 128 MB instead of 1 GB)."""
    world = dist.get_world_size(group)
    per = chunk.size(1)
    out = torch.empty(2, world * per, dtype=chunk.dtype, device=chunk.device)
    dist.all_gather_into_tensor(out[0], chunk[0].contiguous(), group=group)
    dist.all_gather_into_tensor(out[1], chunk[1].contiguous(), group=group)
    return out if world * per == num_edges else out[:, :num_edges].contiguous()


def allreduce_grads(params, group=None):
    """Sum the per-rank partial parameter gradients (one flat NCCL all-reduce)."""
    ps = [p for p in params if p.grad is not None]
    if not ps:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    dist.all_reduce(flat, group=group)
    off = 0
    for p in ps:
        p.grad.copy_(flat[off: off + p.numel()].view_as(p.grad))
        off += p.numel()
