"""CPU restatement of the reference's GCL model surface (no reference import).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows, line by line:
    biomedkg/model/encoder.py:124-162   GCNEncoder
    biomedkg/model/gcl.py:8-27          DGI
    biomedkg/model/gcl.py:31-51         GRACE
    biomedkg/model/gcl.py:54-93         GGD
    biomedkg/utils/fusion.py:10-31      AttentionFusion
    biomedkg/utils/fusion.py:34-90      ReDAF
    biomedkg/gcl_module.py:43-58        BaseGCL.fusion_fn / forward
    biomedkg/gcl_module.py:140-143      DGIModule.calculate_loss
    biomedkg/gcl_module.py:186-190      GRACEModule.calculate_loss
    biomedkg/gcl_module.py:229-234      GGDModule.calculate_loss

Every stochastic draw goes through a ``draws`` object (duck-typed, see
``TorchDraws`` / ``ReplayDraws``) in the reference's draw order (SURVEY.md
Appendix A.7), so the CUDA path and this oracle can consume identical masks.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pyg, pygcl


# --------------------------------------------------------------------------
# draw sources
# --------------------------------------------------------------------------
class TorchDraws:
    """Draws with torch's global generators exactly where the reference does:
    rand_like / rand / dropout on the tensor's device generator, randperm and
    the GGD coin on the CPU generator.  Records everything it hands out."""

    def __init__(self, record: bool = True):
        self.log = [] if record else None

    def _rec(self, kind, t):
        if self.log is not None:
            self.log.append((kind, t.clone() if torch.is_tensor(t) else t))
        return t

    def feature_mask(self, x, p):
        return self._rec("feature_mask", torch.rand_like(x) >= p)

    def edge_mask(self, edge_index, p):
        return self._rec("edge_mask", torch.rand(edge_index.size(1), device=edge_index.device) >= p)

    def dropout_mask(self, x, p):
        return self._rec("dropout_mask", torch.rand_like(x) >= p)

    def randperm(self, n):
        return self._rec("randperm", torch.randperm(n))

    def coin(self):
        return self._rec("coin", float(torch.rand(1).item()))


class ReplayDraws:
    """Replays a recorded list of (kind, value) draws, asserting the order."""

    def __init__(self, log):
        self.log = list(log)
        self.i = 0

    def _next(self, kind):
        k, v = self.log[self.i]
        assert k == kind, f"draw order mismatch: wanted {kind}, recorded {k} at {self.i}"
        self.i += 1
        return v

    def feature_mask(self, x, p):
        return self._next("feature_mask").to(x.device)

    def edge_mask(self, edge_index, p):
        return self._next("edge_mask").to(edge_index.device)

    def dropout_mask(self, x, p):
        return self._next("dropout_mask").to(x.device)

    def randperm(self, n):
        return self._next("randperm")

    def coin(self):
        return self._next("coin")


# --------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------
class _Lin(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(o, i))


class GCNConv(nn.Module):
    """PyG GCNConv(in, out) with defaults (Appendix A.1)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.lin = _Lin(in_channels, out_channels)
        self.bias = nn.Parameter(torch.empty(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        pyg.glorot_(self.lin.weight)
        with torch.no_grad():
            self.bias.zero_()

    def forward(self, x, edge_index):
        return pyg.gcn_conv(x, edge_index, self.lin.weight, self.bias)


class GATConv(nn.Module):
    """PyG GATConv(in, out, heads) (Appendix A.6) - extension."""

    def __init__(self, in_channels, out_channels, heads=1):
        super().__init__()
        self.heads, self.out_channels = heads, out_channels
        self.lin = _Lin(in_channels, heads * out_channels)
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.empty(heads * out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        pyg.glorot_(self.lin.weight)
        pyg.glorot_(self.att_src)
        pyg.glorot_(self.att_dst)
        with torch.no_grad():
            self.bias.zero_()

    def forward(self, x, edge_index):
        return pyg.gat_conv(x, edge_index, self.lin.weight, self.att_src, self.att_dst, self.bias, self.heads)


class GCNEncoder(nn.Module):
    """encoder.py:124-162."""

    conv_cls = GCNConv

    def __init__(self, in_dim, hidden_dim, out_dim, num_hidden_layers, drop_out=True, **conv_kw):
        super().__init__()
        self.drop_out = drop_out
        layers = [self.conv_cls(in_dim, hidden_dim, **conv_kw)]
        for _ in range(num_hidden_layers):
            layers.append(self.conv_cls(hidden_dim, hidden_dim, **conv_kw))
        layers.append(self.conv_cls(hidden_dim, out_dim, **conv_kw))
        self.graph_layers = nn.ModuleList(layers)
        self.reset_parameters()
        self.draws = None  # None -> torch's own F.dropout
        #: optional list of [N,C] bool masks consumed one per hidden layer: when set, the ReLU/dropout
        #: decisions are taken from it (x = pre * mask * scale) instead of from the fp64 pre-activation.
        #: Tests use it to compare gradients under the decisions a reduced-precision forward actually took.
        self.relu_override = None

    def reset_parameters(self):
        for layer in self.graph_layers:
            layer.reset_parameters()

    def _dropout(self, x):
        if not self.training:
            return x
        if self.draws is None:
            return F.dropout(x, p=0.2, training=True)
        keep = self.draws.dropout_mask(x, 0.2)
        return x * keep.to(x.dtype) / (1.0 - 0.2)

    def forward(self, x, edge_index):
        for layer in self.graph_layers[:-1]:
            if self.relu_override is not None:
                mask = self.relu_override.pop(0).to(x.dtype)
                scale = 1.0
                if self.drop_out and self.training:
                    scale = 1.0 / (1.0 - 0.2)
                    if self.draws is not None:
                        self.draws.dropout_mask(x, 0.2)      # keep the draw order; the decision is in `mask`
                x = layer(x, edge_index) * mask * scale
                continue
            x = F.relu(layer(x, edge_index))
            if self.drop_out:
                x = self._dropout(x)
        return self.graph_layers[-1](x, edge_index)


class GATEncoder(GCNEncoder):
    """Extension: the GCNEncoder layer pattern over GATConv (heads=1)."""

    conv_cls = GATConv


class GRACE(nn.Module):
    """model/gcl.py:31-51."""

    def __init__(self, encoder, hidden_dim, proj_dim, draws=None):
        super().__init__()
        self.encoder = encoder
        self.fc1 = nn.Linear(hidden_dim, proj_dim)
        self.fc2 = nn.Linear(proj_dim, hidden_dim)
        self.draws = draws or TorchDraws(record=False)

    def forward(self, x, edge_index):
        m1 = self.draws.feature_mask(x, 0.4)
        m2 = self.draws.feature_mask(x, 0.4)
        k1 = self.draws.edge_mask(edge_index, 0.4)
        k2 = self.draws.edge_mask(edge_index, 0.4)
        x1 = x.masked_fill(~m1, 0.0)
        x2 = x.masked_fill(~m2, 0.0)
        z = self.encoder(x, edge_index)
        z1 = self.encoder(x1, edge_index[:, k1])
        z2 = self.encoder(x2, edge_index[:, k2])
        return z, z1, z2

    def project(self, z):
        return self.fc2(F.elu(self.fc1(z)))


class DGI(nn.Module):
    """model/gcl.py:8-27."""

    def __init__(self, encoder, hidden_dim, draws=None):
        super().__init__()
        self.encoder = encoder
        self.project = nn.Linear(hidden_dim, hidden_dim)
        pyg.uniform_(hidden_dim, self.project.weight)
        self.draws = draws or TorchDraws(record=False)

    @staticmethod
    def summary(z):
        return torch.sigmoid(z.mean(dim=0, keepdim=True))

    def forward(self, x, edge_index):
        z = self.encoder(x, edge_index)
        g = self.project(self.summary(z))
        zn = self.encoder(x[self.draws.randperm(x.size(0))], edge_index)
        return z, g, zn


class GGD(nn.Module):
    """model/gcl.py:54-93."""

    def __init__(self, encoder, hidden_dim, n_proj, aug_p, draws=None):
        super().__init__()
        self.encoder = encoder
        self.p = aug_p
        self.mlp = nn.ModuleList([nn.Linear(hidden_dim, hidden_dim) for _ in range(n_proj)])
        self.draws = draws or TorchDraws(record=False)

    def forward(self, x, edge_index):
        if self.draws.coin() < self.p:
            x = x.masked_fill(~self.draws.feature_mask(x, 0.4), 0.0)
            edge_index = edge_index[:, self.draws.edge_mask(edge_index, 0.4)]
        pos_z = self.encoder(x, edge_index)
        neg_z = self.encoder(x[self.draws.randperm(x.size(0))], edge_index)
        pos_h, neg_h = pos_z, neg_z
        for layer in self.mlp[:-1]:
            pos_h = F.relu(layer(pos_h))
            neg_h = F.relu(layer(neg_h))
        pos_h = self.mlp[-1](pos_h).sum(1)
        neg_h = self.mlp[-1](neg_h).sum(1)
        return pos_h, neg_h


# --------------------------------------------------------------------------
# fusion
# --------------------------------------------------------------------------
class AttentionFusion(nn.Module):
    """utils/fusion.py:10-31; SDPA written out: softmax(q k^T / sqrt(E)) v."""

    def __init__(self, embed_dim):
        super().__init__()
        self.q_proj = nn.Linear(embed_dim, embed_dim)
        self.k_proj = nn.Linear(embed_dim, embed_dim)
        self.v_proj = nn.Linear(embed_dim, embed_dim)

    def forward(self, x):
        q, k, v = self.q_proj(x), self.k_proj(x), self.v_proj(x)
        att = torch.softmax(q @ k.transpose(-2, -1) / math.sqrt(q.size(-1)), dim=-1)
        x = att @ v
        if x.dim() == 3:
            x = x.mean(dim=1)
        return x


class ReDAF(nn.Module):
    """utils/fusion.py:34-90 with sub_type_ids=None, relational_context=0.2."""

    def __init__(self, embed_dim, num_modalities=2):
        super().__init__()
        self.embed_dim, self.num_modalities = embed_dim, num_modalities
        self.modal_weights = nn.Parameter(torch.ones(num_modalities, 1, embed_dim))
        self.sub_type_embeddings = nn.Embedding(num_modalities, embed_dim)
        self.transform_layer = nn.Linear(embed_dim, embed_dim)
        self.relational_context_layer = nn.Linear(embed_dim, embed_dim)
        self.draws = None

    def forward(self, x, relational_context=0.2):
        ctx = torch.full((1, self.embed_dim), relational_context, dtype=x.dtype)
        zeta = torch.sigmoid(self.relational_context_layer(ctx))        # [1, E]
        t = F.relu(self.transform_layer(x))                                # [N, M, E]
        w = t * self.modal_weights.transpose(0, 1) * zeta.unsqueeze(0)   # [N, M, E]
        if self.training:
            if self.draws is None:
                w = F.dropout(w, 0.1, True)
            else:
                w = w * self.draws.dropout_mask(w, 0.1).to(w.dtype) / 0.9
        h = F.relu(w)
        if h.dim() == 3:
            h = h.mean(dim=1)
        return h


def create_fuser(method, embed_dim):
    """factory.py:8-15."""
    if method == "attention":
        return AttentionFusion(embed_dim)
    if method == "redaf":
        return ReDAF(embed_dim)
    return None


# --------------------------------------------------------------------------
# task modules (gcl_module.py) - loss computation only
# --------------------------------------------------------------------------
class _BaseGCL(nn.Module):
    def __init__(self, model, embed_dim, fuse_method):
        super().__init__()
        self.model = model
        self.modality_transform = create_fuser(fuse_method, embed_dim)

    def fusion_fn(self, x):
        if self.modality_transform:
            x = self.modality_transform(x)
        elif x.dim() == 3:
            x = torch.mean(x, dim=1)
        return x

    def forward(self, x, edge_index):
        return self.model.encoder(self.fusion_fn(x), edge_index)

    def training_step(self, x, edge_index):
        return self.calculate_loss(self.fusion_fn(x), edge_index)


def _enc(encoder, in_dim, hidden_dim, out_dim, num_hidden_layers):
    cls = GATEncoder if encoder == "gat" else GCNEncoder
    return cls(in_dim, hidden_dim, out_dim, num_hidden_layers)


class GRACEModule(_BaseGCL):
    def __init__(self, in_dim, hidden_dim, out_dim, num_hidden_layers, fuse_method=None, encoder="gcn", closed_form=False):
        model = GRACE(_enc(encoder, in_dim, hidden_dim, out_dim, num_hidden_layers), hidden_dim, hidden_dim)
        super().__init__(model, in_dim, fuse_method)
        self.closed_form = closed_form

    def calculate_loss(self, x, edge_index):
        _, z1, z2 = self.model(x, edge_index)
        h1, h2 = [self.model.project(z) for z in (z1, z2)]
        if self.closed_form:
            return pygcl.infonce_l2l_closed_form(h1, h2, 0.2)
        return pygcl.infonce_l2l_as_written(h1, h2, 0.2, True)


class DGIModule(_BaseGCL):
    def __init__(self, in_dim, hidden_dim, out_dim, num_hidden_layers, fuse_method=None, encoder="gcn"):
        model = DGI(_enc(encoder, in_dim, hidden_dim, out_dim, num_hidden_layers), hidden_dim)
        super().__init__(model, in_dim, fuse_method)

    def calculate_loss(self, x, edge_index):
        pos_z, summary, neg_z = self.model(x, edge_index)
        return pygcl.jsd_g2l_as_written(pos_z, summary, neg_z)


class GGDModule(_BaseGCL):
    def __init__(self, in_dim, hidden_dim, out_dim, num_hidden_layers, fuse_method=None, encoder="gcn"):
        model = GGD(_enc(encoder, in_dim, hidden_dim, out_dim, num_hidden_layers), hidden_dim, 1, 0.5)
        super().__init__(model, in_dim, fuse_method)

    def calculate_loss(self, x, edge_index):
        pos_h, neg_h = self.model(x, edge_index)
        pred = torch.cat([pos_h, neg_h])
        gt = torch.cat([torch.ones_like(pos_h), torch.zeros_like(neg_h)])
        return F.binary_cross_entropy_with_logits(pred, gt)


def set_draws(module: nn.Module, draws):
    """Point every stochastic sub-module of an oracle model at one draw source."""
    for m in module.modules():
        if hasattr(m, "draws"):
            m.draws = draws
    return module
