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

    def forward(self, x, edge_index, relu=False, drop_p=0.0, drop_seed=0, drop_keep=None, out_fp32=None, correct=False):
        """``out_fp32`` defaults to fp32 output (as PyG) unless ``relu`` is fused (the encoder's bf16 inter-layer format)."""
        view = ops.as_view(edge_index, x.size(0))
        if x.dtype != ops.BF16:
            x = ops.mask_cast(x.float())[0]
        out_fp32 = (not relu) if out_fp32 is None else out_fp32
        return ops.gcn_layer(x, self.lin.weight, self.bias, view, relu, drop_p, drop_seed, drop_keep, out_fp32, correct)


class GATConv(nn.Module):
    """Extension (BASELINE.json configs 2, 5): PyG GATConv(in, out, heads, concat=True, negative_slope=0.2,
    add_self_loops=True) parameter layout - lin.weight [H*out, in], att_src/att_dst [1,H,out] glorot, bias [H*out]."""

    def __init__(self, in_channels: int, out_channels: int, heads: int = 1, negative_slope: float = 0.2):
        super().__init__()
        if heads not in (1, 2, 4) or out_channels % 8 or heads * out_channels > 1024:
            raise ValueError("GATConv kernels support heads in {1,2,4}, out_channels % 8 == 0, heads*out_channels <= 1024")
        self.in_channels, self.out_channels, self.heads, self.negative_slope = in_channels, out_channels, heads, negative_slope
        self.lin = _Lin(in_channels, heads * out_channels)
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.empty(heads * out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        _glorot_(self.lin.weight)
        _glorot_(self.att_src)
        _glorot_(self.att_dst)
        with torch.no_grad():
            self.bias.zero_()

    def forward(self, x, edge_index, relu=False, drop_p=0.0, drop_seed=0, drop_keep=None, out_fp32=None, correct=False):
        view = ops.as_view(edge_index, x.size(0))
        if x.dtype != ops.BF16:
            x = ops.mask_cast(x.float())[0]
        out_fp32 = (not relu) if out_fp32 is None else out_fp32
        return ops.gat_layer(x, self.lin.weight, self.att_src, self.att_dst, self.bias, view, self.heads, self.negative_slope,
                             relu, drop_p, drop_seed, drop_keep, out_fp32, correct)


class GCNEncoder(nn.Module):
    conv_cls = GCNConv

    def __init__(self, in_dim: int, hidden_dim: int, out_dim: int, num_hidden_layers: int, drop_out: bool = True, **conv_kw):
        super().__init__()
        self.drop_out = drop_out
        layers = [self.conv_cls(in_dim, hidden_dim, **conv_kw)]
        for _ in range(num_hidden_layers):
            layers.append(self.conv_cls(hidden_dim, hidden_dim, **conv_kw))
        layers.append(self.conv_cls(hidden_dim, out_dim, **conv_kw))
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
        for i, layer in enumerate(self.graph_layers[:-1]):
            p, seed, keep = 0.0, 0, None
            if self.drop_out and self.training:
                p = 0.2
                seed, keep = self.draws.dropout((x.size(0), layer.heads * layer.out_channels if hasattr(layer, "heads") else layer.out_channels), p, x.device)
            # layers fed by an activation (i > 0) restore the common-mode part of the weight rounding in fp32 (ops._xw)
            x = layer(x, view, relu=True, drop_p=p, drop_seed=seed, drop_keep=keep, out_fp32=False, correct=i > 0)
        return self.graph_layers[-1](x, view, relu=False, out_fp32=True, correct=len(self.graph_layers) > 1)


class GATEncoder(GCNEncoder):
    """Extension: the GCNEncoder layer pattern (encoder.py:124-162) over GATConv; same forward(x, edge_index)."""

    conv_cls = GATConv
