"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row partitioning, the collectives and the autograd plumbing
of the row-sharded InfoNCE.  The CUDA kernels are replaced by an injected torch restatement of the same row-range math
(the product has no CPU path; this is the checker standing in for the kernels)."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pygcl


def test_shard_layout_properties():
    from biomedkg_b200.dist import shard_layout

    for n in (10, 128, 129, 1000, 28_000, 130_000, 1_000_000):
        for world in (1, 2, 3, 4, 8):
            B, parts = shard_layout(n, world)
            assert B % 128 == 0 and B * world >= n
            assert parts[0][0] == 0 and parts[-1][1] == n
            for p, (b0, e0) in enumerate(parts):
                assert b0 == min(p * B, n) and b0 <= e0 <= b0 + B
            for (b0, e0), (b1, e1) in zip(parts, parts[1:]):
                assert e0 == b1


def _node_of_row(u, B):
    blk = u // B
    return (blk // 2) * B + (u - blk * B), blk % 2


class TorchRowImpl:
    """fp64 torch restatement of what the row-range kernels compute on the block-interleaved stacked layout of
    include/bmkg_b200.h (centred operand z_u = mu + d_u, q_u = 1/R'_u, w_u = 2^a_u; exact exp2, no bf16 rounding)."""

    def stats(self, h):
        inv = 1 / h.double().norm(dim=1).clamp_min(1e-12)
        return inv, (h.double() * inv[:, None]).sum(0)

    def center(self, hs, invs, mu, B, scale):
        D = hs[0].size(1)
        z = torch.zeros(2, B, D, dtype=torch.float64)
        a = torch.zeros(2, B, dtype=torch.float64)
        for v, (h, inv) in enumerate(zip(hs, invs)):
            n = h.size(0)
            z[v, :n] = h.double() * inv[:, None] * scale - mu
            a[v, :n] = z[v, :n] @ mu
        return z, a

    def _valid(self, R_all, N, B):
        u = torch.arange(R_all)
        node, view = _node_of_row(u, B)
        return node < N, view

    def fwd_rows(self, Z, A, N, B, r0, r1):
        R_all = Z.size(0)
        qw = torch.zeros(R_all, 2, dtype=torch.float64)
        loss = torch.zeros((), dtype=torch.float64)
        if r1 > r0:
            valid, view = self._valid(R_all, N, B)
            E = torch.exp2(Z[r0:r1] @ Z.t() + A[None, :]) * valid[None, :]
            E[torch.arange(r1 - r0), torch.arange(r0, r1)] = 0
            Rp = E.sum(1)
            rows = torch.arange(r0, r1)
            ok = valid[r0:r1]
            qw[rows[ok], 0] = 1 / Rp[ok]
            qw[rows[ok], 1] = torch.exp2(A[rows[ok]])
            first = rows[ok & (view[r0:r1] == 0)]
            dots = (Z[first] * Z[first + B]).sum(1)
            loss = (Rp[ok].log().sum() - math.log(2.0) * A[rows[ok]].sum() - 2 * math.log(2.0) * dots.sum()) / (2 * N)
        return loss, qw

    def bwd_rows(self, Z, QW, mu, g, N, B, r0, r1):
        dz = torch.zeros(max(r1 - r0, 1), Z.size(1), dtype=torch.float64)
        if r1 > r0:
            q, w = QW[:, 0], QW[:, 1]
            P = torch.exp2(Z[r0:r1] @ Z.t()) * (q[r0:r1, None] * w[None, :] + q[None, :] * w[r0:r1, None])
            P[torch.arange(r1 - r0), torch.arange(r0, r1)] = 0
            rows = torch.arange(r0, r1)
            node, view = _node_of_row(rows, B)
            pair = torch.where(view == 0, rows + B, rows - B)
            out = float(g) * math.log(2.0) / (2 * N) * (P @ Z + (P.sum(1, keepdim=True) - 2.0) * mu[None, :] - 2 * Z[pair])
            dz[: r1 - r0] = out * (node < N)[:, None]
        return dz

    def norm_bwd(self, h, inv, dz, scale):
        u = h.double() * inv[:, None]
        d = dz.double()
        return (scale * inv[:, None] * (d - u * (u * d).sum(1, keepdim=True))).to(h.dtype)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, d, out, replicated):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from biomedkg_b200.dist import shard_layout, sharded_infonce_local, sharded_infonce_loss

        g = torch.Generator().manual_seed(5)           # identical (replicated) inputs on every rank
        h1 = torch.randn(n, d, generator=g, dtype=torch.float64)
        h2 = h1 + torch.randn(n, d, generator=g, dtype=torch.float64)
        if replicated:                                  # replicated encoder: full tensors in, full gradients out
            h1.requires_grad_(True), h2.requires_grad_(True)
            loss = sharded_infonce_loss(h1, h2, 0.2, None, TorchRowImpl())
            (loss * 3.0).backward()                    # non-trivial upstream gradient
            out[rank] = (float(loss), h1.grad.clone(), h2.grad.clone(), (0, n))
        else:                                           # row-sharded encoder: every rank holds only its node block
            B, parts = shard_layout(n, world)
            n0, n1 = parts[rank]
            a, b = h1[n0:n1].clone().requires_grad_(True), h2[n0:n1].clone().requires_grad_(True)
            loss = sharded_infonce_local(a, b, n, 0.2, None, TorchRowImpl())
            (loss * 3.0).backward()
            out[rank] = (float(loss), a.grad.clone(), b.grad.clone(), (n0, n1))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("replicated", [False, True])
@pytest.mark.parametrize("n,d,world", [(300, 64, 2), (129, 32, 2), (64, 32, 2)])
def test_sharded_infonce_matches_single_process_oracle(n, d, world, replicated):
    """(64, 32, 2): B = 128 > N, so rank 1 owns no rows at all - the collectives must still line up."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, d, out, replicated), nprocs=world, join=True)
    g = torch.Generator().manual_seed(5)
    h1 = torch.randn(n, d, generator=g, dtype=torch.float64)
    h2 = (h1 + torch.randn(n, d, generator=g, dtype=torch.float64)).requires_grad_(True)
    h1.requires_grad_(True)
    ref = pygcl.infonce_l2l_as_written(h1, h2, 0.2, True)
    (ref * 3.0).backward()
    for r in range(world):
        loss, g1, g2, (n0, n1) = out[r]
        assert abs(loss - float(ref)) < 1e-9 * abs(float(ref))
        assert torch.allclose(g1, h1.grad[n0:n1], rtol=1e-7, atol=1e-12)
        assert torch.allclose(g2, h2.grad[n0:n1], rtol=1e-7, atol=1e-12)
    assert out[0][0] == out[1][0]                      # every rank ends with the same loss


def _gather_worker(rank, world, port, n, c, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from biomedkg_b200.dist import _GatherRowsFn, node_partition

        block, parts = node_partition(n, world)
        r0, r1 = parts[rank]
        full = torch.arange(n * c, dtype=torch.float32).view(n, c)
        local = full[r0:r1].clone().requires_grad_(True)
        gathered = _GatherRowsFn.apply(local, n, block, r0, None)
        w = torch.arange(n, dtype=torch.float32).view(n, 1) + 1.0      # every rank applies the SAME function downstream
        (gathered * w).sum().backward()
        out[rank] = (torch.equal(gathered.detach(), full), torch.equal(local.grad, w[r0:r1].expand(-1, c)), (r0, r1))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 7, 1])
def test_node_partition_and_row_gather(n):
    from biomedkg_b200.dist import node_partition

    block, parts = node_partition(n, 2)
    assert parts[0][0] == 0 and parts[-1][1] == n and parts[0][1] == parts[1][0] and block * 2 >= n
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gather_worker, args=(2, _free_port(), n, 3, out), nprocs=2, join=True)
    assert all(out[r][0] and out[r][1] for r in range(2)), dict(out)
