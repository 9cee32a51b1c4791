"""Modality fusion - same surface as biomedkg/utils/fusion.py:10-90."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..draws import DeviceDraws


class AttentionFusion(nn.Module):
    """utils/fusion.py:10-31.  q/k/v projections run as ONE [N*M,E]x[E,3E] bf16 GEMM; the per-node
    M x M softmax, PV product and the mean over M are one fused kernel (ops.fusion_attention)."""

    def __init__(self, embed_dim: int):
        super().__init__()
        self.q_proj = nn.Linear(embed_dim, embed_dim)
        self.k_proj = nn.Linear(embed_dim, embed_dim)
        self.v_proj = nn.Linear(embed_dim, embed_dim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 3:
            raise NotImplementedError("AttentionFusion expects stacked modality embeddings [N, M, E]")
        N, M, E = x.shape
        w = torch.cat([self.q_proj.weight, self.k_proj.weight, self.v_proj.weight], dim=0)
        b = torch.cat([self.q_proj.bias, self.k_proj.bias, self.v_proj.bias], dim=0)
        qkv = ops.linear(x.reshape(N * M, E), w, None, out_bf16=True)     # bias is added inside the fused kernel
        return ops.fusion_attention(qkv.contiguous(), b, N, M, E)


class ReDAF(nn.Module):
    """utils/fusion.py:34-90 (sub_type_ids=None, relational_context=0.2 as every caller uses it): the transform is one
    bf16 GEMM, everything after it (bias, ReLU, gate, dropout, ReLU, mean over modalities) one fused kernel."""

    def __init__(self, embed_dim: int, num_modalities: int = 2):
        super().__init__()
        self.embed_dim, self.num_modalities = embed_dim, num_modalities
        self.modal_weights = nn.Parameter(torch.ones(num_modalities, 1, embed_dim))
        self.sub_type_embeddings = nn.Embedding(num_modalities, embed_dim)
        self.transform_layer = nn.Linear(embed_dim, embed_dim)
        self.relational_context_layer = nn.Linear(embed_dim, embed_dim)
        self.dropout = nn.Dropout(0.1)
        self.activation = nn.ReLU()
        self.draws = DeviceDraws()

    def forward(self, x, relational_context=0.2, sub_type_ids=None):
        if sub_type_ids is not None:
            raise NotImplementedError("sub_type_ids is never passed on the GCL path")
        if x.dim() != 3 or x.size(1) != self.num_modalities:
            raise ValueError(f"ReDAF expects [N, {self.num_modalities}, E] stacked modality embeddings")  # fusion.py:58-60,75-79 broadcast
        N, M, E = x.shape
        ctx = torch.full((1, self.embed_dim), relational_context, device=x.device, dtype=torch.float32)
        zeta = torch.sigmoid(torch.nn.functional.linear(ctx, self.relational_context_layer.weight, self.relational_context_layer.bias))
        gate = self.modal_weights.squeeze(1) * zeta                               # [M, E] (fusion.py:82-84)
        t = ops.linear(x.reshape(N * M, E), self.transform_layer.weight, None, out_bf16=True).view(N, M, E)
        p, seed, keep = 0.0, 0, None
        if self.training:
            p = self.dropout.p
            seed, keep = self.draws.dropout((N, M, E), p, x.device)
        return ops.redaf_fuse(t.contiguous(), self.transform_layer.bias, gate, p, seed, keep)
