"""Loss objects with PyGCL's call surface (GCL.models / GCL.losses as used at
biomedkg/gcl_module.py:127,142,171-173,189), backed by the fused kernels."""
from __future__ import annotations

import math

import torch

from . import ops


class InfoNCE:
    def __init__(self, tau: float):
        self.tau = tau


class JSD:
    pass


class DualBranchContrast(torch.nn.Module):
    """contrast_model(h1, h2) -> scalar.  Only the configuration the reference builds is
    implemented: mode="L2L", intraview_negs=True, InfoNCE loss (one fused tcgen05 kernel)."""

    def __init__(self, loss, mode: str = "L2L", intraview_negs: bool = False, **kwargs):
        super().__init__()
        if mode != "L2L" or not intraview_negs or not isinstance(loss, InfoNCE):
            raise NotImplementedError("only DualBranchContrast(InfoNCE(tau), mode='L2L', intraview_negs=True) is on the GCL path")
        self.loss, self.mode, self.intraview_negs = loss, mode, intraview_negs

    def forward(self, h1=None, h2=None, g1=None, g2=None, batch=None, h3=None, h4=None, extra_pos_mask=None, extra_neg_mask=None):
        if any(a is not None for a in (g1, g2, batch, h3, h4, extra_pos_mask, extra_neg_mask)):
            raise NotImplementedError("only the L2L call contrast_model(h1, h2) is on the GCL path")
        return ops.infonce_loss(h1, h2, self.loss.tau)


class SingleBranchContrast(torch.nn.Module):
    """contrast_model(h=z, g=summary, hn=zn) -> scalar; mode="G2L" with JSD (DGI)."""

    def __init__(self, loss, mode: str = "G2L", **kwargs):
        super().__init__()
        if mode != "G2L" or not isinstance(loss, JSD):
            raise NotImplementedError("only SingleBranchContrast(JSD(), mode='G2L') is on the GCL path")
        self.loss, self.mode = loss, mode

    def forward(self, h=None, g=None, batch=None, hn=None, extra_pos_mask=None, extra_neg_mask=None):
        if batch is not None or extra_pos_mask is not None or extra_neg_mask is not None:
            raise NotImplementedError("only the single-graph call contrast_model(h=, g=, hn=) is on the GCL path")
        n = h.size(0)
        s_pos, s_neg = ops.rowdot(h, g), ops.rowdot(hn, g)
        return ops.softplus_pair_sum(s_pos, s_neg) / n - 2.0 * math.log(2.0)


def bce_with_logits_pos_neg(pos_h: torch.Tensor, neg_h: torch.Tensor) -> torch.Tensor:
    """F.binary_cross_entropy_with_logits(cat(pos,neg), cat(1,0)) of gcl_module.py:231-233."""
    return ops.softplus_pair_sum(pos_h, neg_h) / (pos_h.numel() + neg_h.numel())
