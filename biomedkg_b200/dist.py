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
