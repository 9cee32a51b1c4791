"""bf16-emulating twin of oracle/models.py: the same GRACE training step in fp64, rounded to bf16 at exactly the points
where the CUDA path stores bf16.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Why it exists.  The north star allows "bf16-in / fp32-accumulate".  Against the pure fp64 oracle the device's gradients
then differ by the accumulated effect of those roundings, which says nothing about whether the KERNELS are right.  This
module separates the two questions:

    device  vs  emulation   -> kernel exactness      (only fp32 accumulation order / ex2.approx differ: <= 1e-3)
    emulation vs fp64 oracle -> cost of the bf16 storage format itself (measurable on CPU, per rounding point)

Rounding points (each can be switched off through ``points`` to attribute the error):

    "x"    encoder input after the feature mask (ops.mask_cast)                   fwd + bwd (dx of the first layer is bf16)
    "w"    parameters cast to bf16 per GEMM                                       fwd
    "xw"   X W^T output of every conv layer (GEMM epilogue)                       fwd; its gradient dxw (aggregation output) bwd
    "y"    inter-layer activations (aggregation epilogue)                         fwd; incoming gradient (GEMM output dx) bwd
    "gpre" gradient after the ReLU/dropout backward (bmkg_relu_dropout_bwd)       bwd
    "proj" projector GEMM operands (ops._LinearFn: z, h -> bf16; g -> bf16)       fwd + bwd
    "z"    InfoNCE operand Z = normalize(h) * sqrt(log2e / tau) as plain bf16 rows   fwd   (round 1)
    "zc"   ... as fp32 column mean mu + bf16 deviations (ops._InfoNCEFn now)       fwd
    "projc" centred projector GEMMs (ops._CenteredLinearFn)                        fwd + bwd
    "wc"   bf16 weights + fp32 rank-1 correction mean(x)(W - bf16 W)^T, layers>0   fwd   (ops._xw)
    "p"    InfoNCE backward probabilities P = 2^S (1/R_u + 1/R_v)                 bwd
    "qkv"  attention-fusion projection output and its gradient                    fwd + bwd

Follows the same reference lines as oracle/models.py (encoder.py:124-162, model/gcl.py:31-51, utils/fusion.py:10-31,
gcl_module.py:186-190) and the device data flow of biomedkg_b200/ops.py.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import pyg

ALL_POINTS = frozenset({"x", "w", "xw", "y", "gpre", "proj", "z", "p", "qkv"})      # round 1's data flow
#: the data flow the CUDA path implements now: centred InfoNCE operand, centred projector GEMMs, weight-residual correction
DEVICE_POINTS = frozenset({"x", "xw", "y", "gpre", "p", "qkv", "zc", "projc", "wc"})
LOG2E = 1.4426950408889634


def bf(t: torch.Tensor) -> torch.Tensor:
    """round-to-nearest-even to bf16, returned in the input dtype"""
    return t.to(torch.bfloat16).to(t.dtype)


class _Round(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fwd, bwd):
        ctx.bwd = bwd
        return bf(x) if fwd else x.clone()

    @staticmethod
    def backward(ctx, g):
        return (bf(g) if ctx.bwd else g), None, None


def rnd(x, fwd=False, bwd=False):
    if not fwd and not bwd:
        return x
    return _Round.apply(x, fwd, bwd)


def _infonce(h1, h2, tau, z_mode, round_p, block=2048):
    """ops._InfoNCEFn: l2norm_colsum + center_scale -> mu (fp32) + bf16 deviations; R'_u, P_uv as in csrc/infonce.cu.
    z_mode: "center" (the device), True (plain bf16 rows, mu = 0: round 1's format), False (no rounding)."""
    from . import pygcl

    scale = math.sqrt(LOG2E / tau)

    def z_hook(z):                                   # z: unit rows; the device scales by sqrt(log2e / tau) before rounding
        zs = z * scale
        if z_mode == "center":
            mu = zs.mean(0).float().to(z.dtype)
            return bf(zs - mu) / scale, mu / scale
        return (bf(zs) if z_mode else zs) / scale, torch.zeros_like(z[0])

    return pygcl.infonce_l2l_blockwise(h1, h2, tau, block, z_hook, bf if round_p else None)


def _agg(xw, adj):
    """A_hat xw with A_hat = D^-1/2 (A' + I) D^-1/2 as a sparse CSR matrix (no [E', C] message tensor: full-size graphs fit)."""
    return torch.sparse.mm(adj, xw)


def _view(edge_index, N, keep, dtype):
    ei = pyg.view_graph(edge_index, N, keep)
    deg = torch.zeros(N, dtype=dtype).index_add_(0, ei[1], torch.ones(ei.size(1), dtype=dtype))
    dis = deg.pow(-0.5)
    dis[torch.isinf(dis)] = 0.0
    w = dis[ei[0]] * dis[ei[1]]
    adj = torch.sparse_coo_tensor(torch.stack([ei[1], ei[0]]), w, (N, N)).coalesce().to_sparse_csr()   # duplicates add up
    return ei, adj


def _gat_agg(xh, ei, att_src, att_dst, heads, slope=0.2):
    N = xh.size(0)
    H = heads
    C = xh.size(1) // H
    x3 = xh.view(N, H, C)
    a_s = (x3 * att_src.view(1, H, C)).sum(-1)
    a_d = (x3 * att_dst.view(1, H, C)).sum(-1)
    row, col = ei[0], ei[1]
    e = F.leaky_relu(a_s[row] + a_d[col], slope)
    emax = torch.full((N, H), float("-inf"), dtype=xh.dtype).scatter_reduce_(0, col.unsqueeze(-1).expand_as(e), e.detach(),
                                                                             reduce="amax", include_self=True)
    e = (e - emax[col]).exp()
    esum = torch.zeros(N, H, dtype=xh.dtype).index_add_(0, col, e)
    alpha = e / (esum[col] + 1e-16)
    msg = alpha.unsqueeze(-1) * x3[row]
    return torch.zeros(N, H, C, dtype=xh.dtype).index_add_(0, col, msg).reshape(N, H * C)


def encoder_forward(enc, x, edge_index, keep, draws, points, training=True):
    """GCNEncoder / GATEncoder.forward (oracle.models layout) on one view with the device's rounding points."""
    N = x.size(0)
    ei, adj = _view(edge_index, N, keep, x.dtype)
    layers = list(enc.graph_layers)
    gat = hasattr(layers[0], "att_src")
    for i, layer in enumerate(layers):
        last = i == len(layers) - 1
        w16 = rnd(layer.lin.weight, fwd="w" in points or "wc" in points)
        xw = x @ w16.t()
        if "wc" in points and i > 0:          # rank-1 correction of the weight rounding: + mean(x) (W - bf16(W))^T, an fp32 bias vector
            xw = xw + (x.mean(0, keepdim=True).detach() @ (layer.lin.weight - w16).t())
        xw = rnd(xw, fwd="xw" in points, bwd="xw" in points)
        if gat:
            agg = _gat_agg(xw, ei, layer.att_src, layer.att_dst, layer.heads)
        else:
            agg = _agg(xw, adj)
        if last:
            return rnd(agg, bwd="gpre" in points) + layer.bias      # fp32 output; dbias = colsum of the unrounded gradient
        pre = rnd(agg + layer.bias, bwd="gpre" in points)           # gpre is rounded before dbias / the transposed aggregation
        y = F.relu(pre)
        if enc.drop_out and training:
            keepm = draws.dropout_mask(y, 0.2)
            y = y * keepm.to(y.dtype) / (1.0 - 0.2)
        x = rnd(y, fwd="y" in points, bwd="y" in points)
    raise AssertionError


def _linear(x, weight, bias, points):
    if "projc" in points:           # centred operand: (x - m) bf16 GEMM + fp32 rank-1 term m W^T
        m = x.mean(0, keepdim=True).detach()
        xc16 = rnd(x - m, fwd=True)                     # dx = g16 W16 leaves the GEMM in fp32 (no bf16 store)
        w16 = rnd(weight, fwd=True)
        return rnd(xc16 @ w16.t(), bwd=True) + (m @ weight.t() + bias)
    x16 = rnd(x, fwd="proj" in points, bwd="proj" in points)
    w16 = rnd(weight, fwd="w" in points)
    return rnd(x16 @ w16.t(), bwd="proj" in points) + bias


def fusion_forward(module, x, points):
    """BaseGCL.fusion_fn (gcl_module.py:43-50) with the device's data flow for AttentionFusion / mean."""
    mt = module.modality_transform
    if mt is None:
        return x.mean(dim=1) if x.dim() == 3 else x
    if not hasattr(mt, "q_proj"):
        raise NotImplementedError("emulation covers fuse_method none / mean / attention")
    N, M, E = x.shape
    w = torch.cat([mt.q_proj.weight, mt.k_proj.weight, mt.v_proj.weight], 0)
    b = torch.cat([mt.q_proj.bias, mt.k_proj.bias, mt.v_proj.bias], 0)
    x16 = rnd(x.reshape(N * M, E), fwd="qkv" in points)
    w16 = rnd(w, fwd="w" in points or "qkv" in points)        # ops._LinearFn casts both GEMM operands to bf16
    qkv = rnd(x16 @ w16.t(), fwd="qkv" in points)
    qkv = rnd(qkv + b, bwd="qkv" in points).view(N, M, 3, E)        # bias added in fp32 inside the kernel; dqkv is bf16
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    att = torch.softmax(q @ k.transpose(-2, -1) / math.sqrt(E), dim=-1)
    return (att @ v).mean(dim=1)


def grace_training_step(module, x, edge_index, draws, points=ALL_POINTS, tau=0.2):
    """oracle.models.GRACEModule.training_step with bf16 rounding at ``points``; ``module`` supplies the parameters
    (fp64), ``draws`` the recorded random draws (oracle.models.ReplayDraws)."""
    points = frozenset(points)
    model = module.model
    fused = fusion_forward(module, x, points)
    m1 = draws.feature_mask(fused, 0.4)
    m2 = draws.feature_mask(fused, 0.4)
    k1 = draws.edge_mask(edge_index, 0.4)
    k2 = draws.edge_mask(edge_index, 0.4)
    enc = model.encoder
    training = module.training
    x0 = rnd(fused, fwd="x" in points, bwd="x" in points)
    x1 = rnd(fused.masked_fill(~m1, 0.0), fwd="x" in points, bwd="x" in points)
    x2 = rnd(fused.masked_fill(~m2, 0.0), fwd="x" in points, bwd="x" in points)
    encoder_forward(enc, x0, edge_index, None, draws, points, training)          # the reference's unused view (draw order)
    z1 = encoder_forward(enc, x1, edge_index, k1, draws, points, training)
    z2 = encoder_forward(enc, x2, edge_index, k2, draws, points, training)

    def project(z):
        h = F.elu(_linear(z, model.fc1.weight, model.fc1.bias, points))
        return _linear(h, model.fc2.weight, model.fc2.bias, points)

    return _infonce(project(z1), project(z2), tau, "center" if "zc" in points else ("z" in points), "p" in points)
