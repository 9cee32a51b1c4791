"""GCN encoder of the GCL path - same surface as biomedkg/model/encoder.py:124-162.

``GCNConv`` mirrors PyG 2.5.3's parameter layout (``lin.weight [out,in]`` glorot,
``bias [out]`` zeros) so reference checkpoints load, but the
forward is X W^T (bf16 tensor-core GEMM) followed by the fused CSR aggregation
kernel (norm + bias + ReLU + dropout in the epilogue).  ``edge_index`` may be the
reference's int64 [2,E] tensor or a prebuilt ``ops.GraphView``.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import ops
from ..draws import DeviceDraws


def _glorot_(t):
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        return t.uniform_(-a, a)


class _Lin(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))


class GCNConv(nn.Module):
    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        if out_channels % 8:
            raise ValueError("out_channels must be a multiple of 8 (128-bit bf16 rows)")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = _Lin(in_channels, out_channels)
        self.bias = nn.Parameter(torch.empty(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        _glorot_(self.lin.weight)
        with torch.no_grad():
            self.bias.zero_()

    def forward(self, x, edge_index, relu=False, drop_p=0.0, drop_seed=0, drop_keep=None, out_fp32=True):
        view = ops.as_view(edge_index, x.size(0))
        if x.dtype != ops.BF16:
            x = ops.mask_cast(x.float())[0]
        return ops.gcn_layer(x, self.lin.weight, self.bias, view, relu, drop_p, drop_seed, drop_keep, out_fp32)


class GCNEncoder(nn.Module):
    def __init__(self, in_dim: int, hidden_dim: int, out_dim: int, num_hidden_layers: int, drop_out: bool = True):
        super().__init__()
        self.drop_out = drop_out
        layers = [GCNConv(in_dim, hidden_dim)]
        for _ in range(num_hidden_layers):
            layers.append(GCNConv(hidden_dim, hidden_dim))
        layers.append(GCNConv(hidden_dim, out_dim))
        self.graph_layers = nn.ModuleList(layers)
        self.draws = DeviceDraws()
        self.reset_parameters()

    def reset_parameters(self):
        for layer in self.graph_layers:
            layer.reset_parameters()

    def forward(self, x, edge_index):
        """x: float [N,in] (fp32 or bf16); returns fp32 [N,out] like the reference."""
        view = ops.as_view(edge_index, x.size(0))
        if x.dtype != ops.BF16:
            x = ops.mask_cast(x.float())[0]
        for layer in self.graph_layers[:-1]:
            p, seed, keep = 0.0, 0, None
            if self.drop_out and self.training:
                p = 0.2
                seed, keep = self.draws.dropout((x.size(0), layer.out_channels), p, x.device)
            x = layer(x, view, relu=True, drop_p=p, drop_seed=seed, drop_keep=keep, out_fp32=False)
        return self.graph_layers[-1](x, view, relu=False, out_fp32=True)
