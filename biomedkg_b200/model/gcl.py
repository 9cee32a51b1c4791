"""GRACE / DGI / GGD heads - same surface as biomedkg/model/gcl.py:8-93."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..draws import DeviceDraws


class GRACE(nn.Module):
    """model/gcl.py:31-51: two augmented views (feature mask p=0.4, edge drop p=0.4) + MLP projector."""

    def __init__(self, encoder, hidden_dim, proj_dim):
        super().__init__()
        self.encoder = encoder
        self.fc1 = nn.Linear(hidden_dim, proj_dim)
        self.fc2 = nn.Linear(proj_dim, hidden_dim)
        self.draws = DeviceDraws()
        #: the reference computes an un-augmented z that GRACEModule discards (gcl_module.py:187);
        #: False keeps that work (faithful), True skips the third encoder pass.
        self.skip_unused_view = False

    def forward(self, x, edge_index):
        N = x.size(0)
        m1 = self.draws.feature_mask(x, 0.4)
        m2 = self.draws.feature_mask(x, 0.4)
        k1 = self.draws.edge_mask(edge_index, 0.4)
        k2 = self.draws.edge_mask(edge_index, 0.4)
        sg = ops.sorted_graph(edge_index, N)
        x0, x1, x2 = ops.mask_cast(x.float(), m1, m2, want_plain=not self.skip_unused_view)
        z = None if self.skip_unused_view else self.encoder(x0, sg.view(None))
        z1 = self.encoder(x1, sg.view(k1))
        z2 = self.encoder(x2, sg.view(k2))
        return z, z1, z2

    def project(self, z: torch.Tensor) -> torch.Tensor:
        h = ops.centered_linear(z, self.fc1.weight, self.fc1.bias, elu=True)      # F.elu fused into the GEMM epilogue
        return ops.centered_linear(h, self.fc2.weight, self.fc2.bias)


class DGI(nn.Module):
    """model/gcl.py:8-27."""

    def __init__(self, encoder, hidden_dim):
        super().__init__()
        self.encoder = encoder
        self.project = nn.Linear(hidden_dim, hidden_dim)
        b = 1.0 / math.sqrt(hidden_dim)  # torch_geometric.nn.inits.uniform(hidden_dim, weight)
        with torch.no_grad():
            self.project.weight.uniform_(-b, b)
        self.draws = DeviceDraws()

    def corruption(self, x, edge_index):
        return x[self.draws.randperm(x.size(0)).to(x.device)], edge_index

    @staticmethod
    def summary(z: torch.Tensor):
        return ops.colmean_sigmoid(z)

    def forward(self, x, edge_index):
        view = ops.as_view(edge_index, x.size(0))
        x16 = ops.mask_cast(x.float())[0]            # one fp32 -> bf16 pass shared by both encoder calls
        z = self.encoder(x16, view)
        g = self.project(self.summary(z))
        zn = self.encoder(self.corruption(x16, edge_index)[0], view)
        return z, g, zn


class GGD(nn.Module):
    """model/gcl.py:54-93.  With n_proj=1 (gcl_module.py:215) the head (z W^T + b).sum(1) is the
    GEMV z . (sum_rows W) + sum(b), evaluated by the fused row-dot kernel."""

    def __init__(self, encoder, hidden_dim, n_proj, aug_p):
        super().__init__()
        self.encoder = encoder
        self.p = aug_p
        self.mlp = nn.ModuleList([nn.Linear(hidden_dim, hidden_dim) for _ in range(n_proj)])
        self.draws = DeviceDraws()

    def corruption(self, x, edge_index):
        return x[self.draws.randperm(x.size(0)).to(x.device)], edge_index

    def forward(self, x, edge_index):
        N = x.size(0)
        x = x.float()
        sg = ops.sorted_graph(edge_index, N)
        if self.draws.coin() < self.p:
            m = self.draws.feature_mask(x, 0.4)
            k = self.draws.edge_mask(edge_index, 0.4)
            x16 = ops.mask_cast(x, m, None, want_plain=False)[1]   # mask_feature fused with the bf16 cast
            view = sg.view(k)
        else:
            x16 = ops.mask_cast(x)[0]
            view = sg.view(None)
        pos_h = self.encoder(x16, view)
        neg_h = self.encoder(self.corruption(x16, edge_index)[0], view)
        for layer in self.mlp[:-1]:
            pos_h = F.relu(ops.linear(pos_h, layer.weight, layer.bias))
            neg_h = F.relu(ops.linear(neg_h, layer.weight, layer.bias))
        last = self.mlp[-1]
        wv, bs = last.weight.sum(0), last.bias.sum()
        return ops.rowdot(pos_h, wv) + bs, ops.rowdot(neg_h, wv) + bs
