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


def test_row_partition_properties():
    from biomedkg_b200.dist import row_partition

    for rows in (10, 128, 129, 1000, 56_000, 2_000_000):
        for world in (1, 2, 3, 4, 8):
            parts = row_partition(rows, world)
            assert parts[0][0] == 0 and parts[-1][1] == rows
            for (b0, e0), (b1, e1) in zip(parts, parts[1:]):
                assert e0 == b1 and b0 <= e0
            assert all(b % 128 == 0 for b, _ in parts)
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 128 + 127


class TorchRowImpl:
    """fp64 torch restatement of what the row-range kernels compute (same scaling convention, exact exp2)."""

    def prep(self, h1, h2, tau):
        scale = math.sqrt(1.4426950408889634 / tau)
        z = torch.cat([torch.nn.functional.normalize(h1.double()), torch.nn.functional.normalize(h2.double())]) * scale
        inv_norm = torch.cat([1 / h1.double().norm(dim=1), 1 / h2.double().norm(dim=1)])
        return z, inv_norm, scale

    def fwd_rows(self, z, N, r0, r1):
        inv_r = torch.zeros(((2 * N + 127) // 128) * 128, dtype=torch.float32)
        loss = torch.zeros((), dtype=torch.float32)
        if r1 > r0:
            S = torch.exp2(z[r0:r1] @ z.t())
            S[torch.arange(r1 - r0), torch.arange(r0, r1)] = 0
            R = S.sum(1)
            inv_r[r0:r1] = (1 / R).float()
            u = torch.arange(r0, r1)
            pos = u[u < N]
            dots = (z[pos] * z[pos + N]).sum(1)
            loss = ((R.log().sum() - 2 * math.log(2.0) * dots.sum()) / (2 * N)).float()
        return loss, inv_r

    def bwd_rows(self, z, inv_r, g, N, r0, r1):
        dz = torch.zeros(2 * N, z.size(1), dtype=torch.float32)
        if r1 > r0:
            c = inv_r[: 2 * N].double()
            P = torch.exp2(z[r0:r1] @ z.t()) * (c[r0:r1, None] + c[None, :])
            P[torch.arange(r1 - r0), torch.arange(r0, r1)] = 0
            pair = torch.cat([torch.arange(N, 2 * N), torch.arange(0, N)])[r0:r1]
            dz[r0:r1] = (float(g) * math.log(2.0) / (2 * N) * (P @ z - 2 * z[pair])).float()
        return dz

    def norm_bwd(self, h1, h2, inv_norm, dz, scale):
        out = []
        N = h1.size(0)
        for h, inv, d in ((h1, inv_norm[:N], dz[:N]), (h2, inv_norm[N:], dz[N:])):
            u = h.double() * inv[:, None]
            d = d.double()
            out.append((scale * inv[:, None] * (d - u * (u * d).sum(1, keepdim=True))).to(h.dtype))
        return out


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, d, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from biomedkg_b200.dist import _ShardedInfoNCEFn

        g = torch.Generator().manual_seed(5)           # identical (replicated) inputs on every rank
        h1 = torch.randn(n, d, generator=g, requires_grad=True)
        h2 = (h1.detach() + torch.randn(n, d, generator=g)).requires_grad_(True)
        loss = _ShardedInfoNCEFn.apply(h1, h2, 0.2, None, TorchRowImpl())
        (loss * 3.0).backward()                        # non-trivial upstream gradient
        out[rank] = (float(loss), h1.grad.clone(), h2.grad.clone())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,d,world", [(200, 64, 2), (129, 32, 2), (64, 32, 2)])
def test_sharded_infonce_matches_single_process_oracle(n, d, world):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, d, out), nprocs=world, join=True)
    g = torch.Generator().manual_seed(5)
    h1 = torch.randn(n, d, generator=g).double().requires_grad_(True)
    h2 = (h1.detach() + torch.randn(n, d, generator=g).double()).requires_grad_(True)
    ref = pygcl.infonce_l2l_as_written(h1, h2, 0.2, True)
    (ref * 3.0).backward()
    for r in range(world):
        loss, g1, g2 = out[r]
        assert abs(loss - float(ref)) < 1e-5 * abs(float(ref))
        assert torch.allclose(g1.double(), h1.grad, rtol=1e-4, atol=1e-7)
        assert torch.allclose(g2.double(), h2.grad, rtol=1e-4, atol=1e-7)
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
