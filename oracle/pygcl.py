"""Restatement of the PyGCL 0.1.2 objects the GCL path constructs.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Call sites: gcl_module.py:127 (SingleBranchContrast(JSD, "G2L")), :142 (its
call), :171-173 (DualBranchContrast(InfoNCE(0.2), "L2L", intraview_negs=True)),
:189 (its call).  Semantics: SURVEY.md Appendix A.4 / A.5.  Each loss exists
twice - the mask-materialising form PyGCL executes ("as written") and the
closed form the CUDA kernels implement - and tests/test_oracle.py checks
they agree.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# InfoNCE / DualBranchContrast (GRACE)
# --------------------------------------------------------------------------
def _similarity(h1: torch.Tensor, h2: torch.Tensor) -> torch.Tensor:
    # GCL.losses.infonce._similarity: F.normalize (p=2, dim=1, eps=1e-12) then h1 @ h2.T
    return F.normalize(h1) @ F.normalize(h2).t()


def _infonce_compute(anchor, sample, pos_mask, neg_mask, tau: float):
    # GCL.losses.InfoNCE.compute - no max-subtraction, masks multiply exp(sim)
    sim = _similarity(anchor, sample) / tau
    exp_sim = torch.exp(sim) * (pos_mask + neg_mask)
    log_prob = sim - torch.log(exp_sim.sum(dim=1, keepdim=True))
    loss = log_prob * pos_mask
    loss = loss.sum(dim=1) / pos_mask.sum(dim=1)
    return -loss.mean()


def _same_scale_sampler(anchor, sample, intraview_negs: bool):
    # GCL.models.samplers.SameScaleSampler + add_intraview_negs
    n = anchor.size(0)
    pos = torch.eye(n, dtype=anchor.dtype, device=anchor.device)
    neg = 1.0 - pos
    if intraview_negs:
        sample = torch.cat([sample, anchor], dim=0)
        pos = torch.cat([pos, torch.zeros_like(pos)], dim=1)
        neg = torch.cat([neg, 1.0 - torch.eye(n, dtype=anchor.dtype, device=anchor.device)], dim=1)
    return anchor, sample, pos, neg


def infonce_l2l_as_written(h1: torch.Tensor, h2: torch.Tensor, tau: float = 0.2, intraview_negs: bool = True):
    """DualBranchContrast.forward(h1, h2) in mode "L2L": materialises the
    [N, 2N] similarity / exp / mask tensors exactly as PyGCL does."""
    a1, s1, p1, n1 = _same_scale_sampler(h1, h2, intraview_negs)
    a2, s2, p2, n2 = _same_scale_sampler(h2, h1, intraview_negs)
    l1 = _infonce_compute(a1, s1, p1, n1, tau)
    l2 = _infonce_compute(a2, s2, p2, n2, tau)
    return (l1 + l2) * 0.5


def infonce_l2l_closed_form(h1: torch.Tensor, h2: torch.Tensor, tau: float = 0.2):
    """Closed form: with a = normalize(h1), b = normalize(h2),
    r1_i = sum_j exp(S12_ij) + sum_{j != i} exp(S11_ij), r2_i likewise with S12^T, S22;
    loss = -(1/2N) sum_i [ 2 S12_ii - log r1_i - log r2_i ]."""
    a, b = F.normalize(h1), F.normalize(h2)
    s12 = a @ b.t() / tau
    s11 = a @ a.t() / tau
    s22 = b @ b.t() / tau
    n = a.size(0)
    off = 1.0 - torch.eye(n, dtype=a.dtype)
    r1 = torch.exp(s12).sum(1) + (torch.exp(s11) * off).sum(1)
    r2 = torch.exp(s12).sum(0) + (torch.exp(s22) * off).sum(1)
    d = torch.diagonal(s12)
    return -(2.0 * d - torch.log(r1) - torch.log(r2)).sum() / (2.0 * n)


class _InfoNCEBlockwise(torch.autograd.Function):
    """The closed form above evaluated in row blocks of the stacked 2N x 2N similarity matrix, with a hand-written backward,
    so full-size configurations (N = 28k: the dense fp64 form needs ~40 GB) run in a few GB.  Optional hooks let
    oracle/emu.py round the operand / the probabilities exactly where the CUDA kernels do; with no hooks this is the
    plain fp64/fp32 closed form (tests/test_oracle.py checks it against infonce_l2l_closed_form and the as-written form)."""

    @staticmethod
    def forward(ctx, h1, h2, tau, block, z_hook, p_hook):
        n = h1.size(0)
        h = torch.cat([h1, h2], 0)
        inv = 1.0 / h.norm(dim=1).clamp_min(1e-12)          # F.normalize eps
        z = h * inv.unsqueeze(1)                              # unit rows
        mu = torch.zeros_like(z[0])
        if z_hook is not None:
            z, mu = z_hook(z)                                  # (deviations-or-rows used in products, common vector)
        zz = z + mu                                            # the represented rows
        R = torch.empty(2 * n, dtype=h.dtype)
        for b0 in range(0, 2 * n, block):
            e = torch.exp((zz[b0:b0 + block] @ zz.t()) / tau)
            idx = torch.arange(b0, min(b0 + block, 2 * n))
            e[idx - b0, idx] = 0.0
            R[b0:b0 + block] = e.sum(1)
        pos = (zz[:n] * zz[n:]).sum() / tau
        loss = (torch.log(R).sum() - 2.0 * pos) / (2.0 * n)
        ctx.save_for_backward(h, inv, z, mu, R)
        ctx.meta = (n, tau, block, p_hook)
        return loss

    @staticmethod
    def backward(ctx, g):
        h, inv, z, mu, R = ctx.saved_tensors
        n, tau, block, p_hook = ctx.meta
        zz = z + mu
        ir = 1.0 / R
        dzz = torch.empty_like(zz)
        for b0 in range(0, 2 * n, block):
            e = torch.exp((zz[b0:b0 + block] @ zz.t()) / tau)
            idx = torch.arange(b0, min(b0 + block, 2 * n))
            e[idx - b0, idx] = 0.0
            P = e * (ir[b0:b0 + block].unsqueeze(1) + ir.unsqueeze(0))
            Pr = p_hook(P) if p_hook is not None else P
            # rounded probabilities multiply the deviation part only; the common vector rides on the exact row sums
            dzz[b0:b0 + block] = Pr @ z + P.sum(1, keepdim=True) * mu
        pair = torch.cat([zz[n:], zz[:n]], 0)
        dzz = (g / (2.0 * n * tau)) * (dzz - 2.0 * pair)
        u = h * inv.unsqueeze(1)
        dh = inv.unsqueeze(1) * (dzz - u * (u * dzz).sum(1, keepdim=True))
        return dh[:n], dh[n:], None, None, None, None


def infonce_l2l_blockwise(h1, h2, tau: float = 0.2, block: int = 2048, z_hook=None, p_hook=None):
    """infonce_l2l_closed_form in O(block x 2N) memory (see _InfoNCEBlockwise)."""
    return _InfoNCEBlockwise.apply(h1, h2, tau, block, z_hook, p_hook)


class DualBranchContrast(torch.nn.Module):
    """Oracle stand-in for GCL.models.DualBranchContrast (L2L only, as used)."""

    def __init__(self, loss, mode: str = "L2L", intraview_negs: bool = False, **kwargs):
        super().__init__()
        assert mode == "L2L", "the GCL path only constructs mode='L2L' (gcl_module.py:171-173)"
        self.loss, self.mode, self.intraview_negs = loss, mode, intraview_negs

    def forward(self, h1=None, h2=None, **kw):
        return infonce_l2l_as_written(h1, h2, self.loss.tau, self.intraview_negs)


class InfoNCE:
    def __init__(self, tau):
        self.tau = tau


# --------------------------------------------------------------------------
# JSD / SingleBranchContrast (DGI)
# --------------------------------------------------------------------------
def jsd_g2l_as_written(h: torch.Tensor, g: torch.Tensor, hn: torch.Tensor):
    """SingleBranchContrast(JSD(), "G2L").forward(h=z, g=g, hn=zn), batch=None.
    CrossScaleSampler: anchor=g [1,D], sample=cat(h,hn) [2N,D], pos=[1_N|0_N]."""
    n = h.size(0)
    sample = torch.cat([h, hn], dim=0)
    pos = torch.cat([torch.ones(1, n, dtype=h.dtype), torch.zeros(1, n, dtype=h.dtype)], dim=1)
    neg = 1.0 - pos
    sim = g @ sample.t()
    # GCL.losses.JSD.compute
    num_pos = pos.int().sum()
    num_neg = neg.int().sum()
    e_pos = (math.log(2.0) - F.softplus(-sim * pos)).sum() / num_pos
    neg_sim = sim * neg
    e_neg = (F.softplus(-neg_sim) + neg_sim - math.log(2.0)).sum() / num_neg
    return e_neg - e_pos


def jsd_g2l_closed_form(h: torch.Tensor, g: torch.Tensor, hn: torch.Tensor):
    """mean softplus(-s+) + mean softplus(s-) - 2 ln2, s+ = h g^T, s- = hn g^T."""
    sp = (h @ g.t()).squeeze(-1)
    sn = (hn @ g.t()).squeeze(-1)
    return F.softplus(-sp).mean() + F.softplus(sn).mean() - 2.0 * math.log(2.0)


class SingleBranchContrast(torch.nn.Module):
    """Oracle stand-in for GCL.models.SingleBranchContrast (G2L only, as used)."""

    def __init__(self, loss, mode: str = "G2L", **kwargs):
        super().__init__()
        assert mode == "G2L", "the GCL path only constructs mode='G2L' (gcl_module.py:127)"
        self.loss, self.mode = loss, mode

    def forward(self, h=None, g=None, batch=None, hn=None, **kw):
        assert batch is None
        return jsd_g2l_as_written(h, g, hn)


class JSD:
    pass


# --------------------------------------------------------------------------
# GGD head + BCE (model/gcl.py:83-91, gcl_module.py:229-234)
# --------------------------------------------------------------------------
def ggd_loss_as_written(pos_z, neg_z, weight, bias):
    pos_h = (pos_z @ weight.t() + bias).sum(1)
    neg_h = (neg_z @ weight.t() + bias).sum(1)
    pred = torch.cat([pos_h, neg_h])
    gt = torch.cat([torch.ones_like(pos_h), torch.zeros_like(neg_h)])
    return F.binary_cross_entropy_with_logits(pred, gt)


def ggd_loss_closed_form(pos_z, neg_z, weight, bias):
    """(z W^T + b).sum(1) == z . (sum_rows W) + sum(b); BCE(1/0) == softplus sums."""
    wv = weight.sum(0)
    bs = bias.sum()
    sp = pos_z @ wv + bs
    sn = neg_z @ wv + bs
    return (F.softplus(-sp).sum() + F.softplus(sn).sum()) / (2 * pos_z.size(0))
