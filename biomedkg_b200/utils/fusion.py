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
    """utils/fusion.py:34-90 (sub_type_ids=None, relational_context=0.2 as every caller uses it)."""

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
        ctx = torch.full((1, self.embed_dim), relational_context, device=x.device, dtype=torch.float32)
        zeta = torch.sigmoid(torch.nn.functional.linear(ctx, self.relational_context_layer.weight, self.relational_context_layer.bias))
        t = self.activation(ops.linear(x, self.transform_layer.weight, self.transform_layer.bias))
        gate = self.modal_weights.transpose(0, 1) * zeta.unsqueeze(0)            # [1, M, E]
        h = self.activation(self.dropout(t * gate))
        if h.dim() == 3:
            h = ops.modality_mean(h)
        return h
