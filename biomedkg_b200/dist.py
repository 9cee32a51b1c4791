"""Row-sharded GRACE InfoNCE across the GPUs of one node.

The stacked views Z = [a; b] (2N x D) are available on every rank (in round 1 the encoder is data-replicated; a row-sharded
encoder would all-gather them - the InfoNCE side is the same).  Rank p owns a contiguous, 128-aligned block of Z's rows and

    forward :  R_u for its rows against ALL columns (tcgen05 kernel on a row range)
               all_gather(1/R)            - the backward needs 1/R_v of every column
               all_reduce(loss share)     - scalar
    backward:  dZ rows of its block (needs the gathered 1/R), then all_gather(dZ) so every rank continues the (replicated)
               backward with the full gradient.

The result equals the single-GPU full-graph loss (unlike the reference's DDP, which contrasts per rank mini-batch).  No
collective touches the N x N work itself; only 2N floats and the 2N x D gradient cross NVLink.

The compute is injected (``impl``): the default calls the CUDA kernels; the gloo CPU tests inject a torch restatement so
the partitioning / collective / autograd plumbing is exercised with world_size 2 on CPU.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

ROW_ALIGN = 128
LOG2E = 1.4426950408889634


def row_partition(num_rows: int, world: int, align: int = ROW_ALIGN):
    """Contiguous [begin, end) row ranges, begins aligned to ``align``, sizes within one aligned block of each other.
    Ranks beyond the number of aligned blocks get an empty range."""
    blocks = (num_rows + align - 1) // align
    out = []
    for r in range(world):
        b0 = (blocks * r) // world
        b1 = (blocks * (r + 1)) // world
        out.append((min(b0 * align, num_rows), min(b1 * align, num_rows)))
    return out


class CudaImpl:
    """Kernels from libbmkg_b200.so on a row range."""

    def prep(self, h1, h2, tau):
        from . import ops
        from .ops import _p, _stream, call

        N, D = h1.shape
        scale = math.sqrt(LOG2E / tau)
        z = torch.empty(2 * N, D, dtype=torch.bfloat16, device=h1.device)
        inv_norm = torch.empty(2 * N, dtype=torch.float32, device=h1.device)
        call("bmkg_l2norm_scale", _p(h1), N, D, scale, _p(z), _p(inv_norm), _stream())
        call("bmkg_l2norm_scale", _p(h2), N, D, scale, z.data_ptr() + N * D * 2, inv_norm.data_ptr() + N * 4, _stream())
        return z, inv_norm, scale

    def fwd_rows(self, z, N, r0, r1):
        from .ops import _p, _stream, _ws, call, lib

        D = z.size(1)
        loss = torch.zeros((), dtype=torch.float32, device=z.device)
        inv_r = torch.zeros(lib.bmkg_infonce_padded_rows(N), dtype=torch.float32, device=z.device)
        if r1 > r0:
            ws = _ws(lib.bmkg_infonce_workspace_bytes_rows(N, D, r0, r1), z.device)
            call("bmkg_infonce_fwd_rows", _p(z), N, D, r0, r1, _p(loss), _p(inv_r), _p(ws), ws.numel(), _stream())
        return loss, inv_r

    def bwd_rows(self, z, inv_r, g, N, r0, r1):
        from .ops import _p, _stream, call

        D = z.size(1)
        dz = torch.zeros(2 * N, D, dtype=torch.float32, device=z.device)
        if r1 > r0:
            call("bmkg_infonce_bwd_rows", _p(z), _p(inv_r), _p(g), N, D, r0, r1, _p(dz), _stream())
        return dz

    def norm_bwd(self, h1, h2, inv_norm, dz, scale):
        from .ops import _p, _stream, call

        N, D = h1.shape
        dh1, dh2 = torch.empty_like(h1), torch.empty_like(h2)
        call("bmkg_l2norm_scale_bwd", _p(h1), _p(inv_norm), _p(dz), N, D, scale, _p(dh1), _stream())
        call("bmkg_l2norm_scale_bwd", _p(h2), inv_norm.data_ptr() + N * 4, dz.data_ptr() + N * D * 4, N, D, scale, _p(dh2), _stream())
        return dh1, dh2


class _ShardedInfoNCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h1, h2, tau, group, impl):
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        h1, h2 = h1.contiguous().float(), h2.contiguous().float()
        N = h1.size(0)
        parts = row_partition(2 * N, world)
        r0, r1 = parts[rank]
        z, inv_norm, scale = impl.prep(h1, h2, tau)
        loss, inv_r = impl.fwd_rows(z, N, r0, r1)
        # every rank filled only its own rows of inv_r (zeros elsewhere): a sum all-reduce is the all-gather of ragged blocks
        dist.all_reduce(inv_r, group=group)
        dist.all_reduce(loss, group=group)
        ctx.save_for_backward(h1, h2, z, inv_norm, inv_r)
        ctx.meta = (scale, group, impl, N, r0, r1)
        return loss

    @staticmethod
    def backward(ctx, g):
        h1, h2, z, inv_norm, inv_r = ctx.saved_tensors
        scale, group, impl, N, r0, r1 = ctx.meta
        dz = impl.bwd_rows(z, inv_r, g.contiguous().float(), N, r0, r1)
        dist.all_reduce(dz, group=group)          # rows are disjoint across ranks: sum == all-gather
        dh1, dh2 = impl.norm_bwd(h1, h2, inv_norm, dz, scale)
        return dh1, dh2, None, None, None


def sharded_infonce_loss(h1, h2, tau=0.2, group=None, impl=None):
    """DualBranchContrast(InfoNCE(tau), "L2L", intraview_negs=True)(h1, h2) with the 2N x 2N work split by rows over ``group``.
    h1, h2 must be identical on every rank (replicated encoder) and so is the returned loss / gradient."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        from . import ops

        return ops.infonce_loss(h1, h2, tau)
    return _ShardedInfoNCEFn.apply(h1, h2, float(tau), group, impl or CudaImpl())


class ShardedDualBranchContrast(torch.nn.Module):
    """Drop-in for losses.DualBranchContrast that row-shards the InfoNCE over the default process group."""

    def __init__(self, tau: float = 0.2, group=None, impl=None):
        super().__init__()
        self.tau, self.group, self.impl = tau, group, impl

    def forward(self, h1, h2):
        return sharded_infonce_loss(h1, h2, self.tau, self.group, self.impl)


# ---------------------------------------------------------------------------------------------------------------------------
# Row-sharded GCN encoder + full GRACE step ("partition nodes by destination row")
# ---------------------------------------------------------------------------------------------------------------------------
def node_partition(num_nodes: int, world: int):
    """Equal contiguous node blocks (the last ones may be short / empty): block p = [p*B, min((p+1)*B, N)), B = ceil(N / world)."""
    B = (num_nodes + world - 1) // world
    return B, [(min(p * B, num_nodes), min((p + 1) * B, num_nodes)) for p in range(world)]


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
    def forward(ctx, x_loc, weight, bias, view, r0, r1, num_nodes, block, group, relu, drop_p, drop_seed, out_fp32):
        from . import ops

        w16 = weight.to(torch.bfloat16)
        xw_loc = torch.mm(x_loc, w16.t())
        xw = _all_gather_rows(xw_loc, num_nodes, block, group).contiguous()
        hub = view.hub_csr if view.hub_possible else None
        y = ops.gcn_aggregate(view.rowptr, view.colind, view.dis, xw, bias.contiguous(), relu, drop_p, drop_seed, None, out_fp32,
                              hub_rows=hub, row_range=(r0, r1))
        ctx.meta = (view, r0, r1, num_nodes, block, group, relu, drop_p)
        ctx.save_for_backward(x_loc, w16, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        from . import ops
        from .ops import _p, _stream, _ws, call, lib

        x_loc, w16, y = ctx.saved_tensors
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
        dw = ops._mm_f32(dxw.t(), x_loc) if ctx.needs_input_grad[1] else None
        dx = torch.mm(dxw, w16) if ctx.needs_input_grad[0] else None
        return (dx, dw, dbias) + (None,) * 10


def sharded_gcn_encoder(encoder, x_loc, view, r0, r1, num_nodes, block, group):
    """GCNEncoder.forward (encoder.py:153-162) on the node rows [r0, r1) of this rank; x_loc bf16 [r1-r0, in]."""
    x = x_loc
    layers = encoder.graph_layers
    for i, layer in enumerate(layers):
        last = i == len(layers) - 1
        p, seed = 0.0, 0
        if not last and encoder.drop_out and encoder.training:
            p = 0.2
            seed, _ = encoder.draws.dropout((num_nodes, layer.out_channels), p, x.device)   # global-row hash: sharding-invariant
        x = _ShardedGCNLayerFn.apply(x, layer.lin.weight, layer.bias, view, r0, r1, num_nodes, block, group, not last, p, seed, last)
    return x


def sharded_grace_loss(module, x, edge_index, group=None, num_nodes=None):
    """GRACEModule.training_step's loss with the node rows split over the ranks of ``group``.

    Every rank receives the same (x, edge_index) and the same parameters / RNG state.  Fusion, feature masks, the GCN layers'
    GEMMs, the projector and the normalisation run on this rank's node block only; each GCN layer all-gathers its transformed
    rows (bf16 [N,256] over NVLink) before aggregating its own destination rows; the projected views are all-gathered once and
    the InfoNCE is split by rows of Z (sharded_infonce_loss).  The loss is the single-GPU full-graph loss; parameter gradients
    are partial sums that the caller all-reduces (``allreduce_grads``).

    ``x`` is either the full feature tensor (every rank slices its rows) or, with ``num_nodes`` given, only this rank's block
    ``node_partition(num_nodes, world)[1][rank]`` - a sharded loader then moves 1/world of the features per rank."""
    from . import ops

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    model = module.model
    N = int(num_nodes) if num_nodes is not None else x.size(0)
    block, parts = node_partition(N, world)
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
    if not model.skip_unused_view:
        sharded_gcn_encoder(enc, x0, sg.view(None), r0, r1, N, block, group)   # the reference's unused un-augmented view
    z1 = sharded_gcn_encoder(enc, x1, sg.view(k1), r0, r1, N, block, group)
    z2 = sharded_gcn_encoder(enc, x2, sg.view(k2), r0, r1, N, block, group)
    h1 = _GatherRowsFn.apply(model.project(z1), N, block, r0, group)
    h2 = _GatherRowsFn.apply(model.project(z2), N, block, r0, group)
    return sharded_infonce_loss(h1, h2, module.contrast_model.loss.tau if hasattr(module.contrast_model, "loss") else 0.2, group)


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
