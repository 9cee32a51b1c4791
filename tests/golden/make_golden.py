"""Generate golden vectors by running the REFERENCE'S OWN Python modules.

Run once in the dev container (needs /root/reference; the GPU box has no copy,
which is why the outputs are committed):

    python tests/golden/make_golden.py

What is real and what is restated: ``biomedkg/model/encoder.py``,
``biomedkg/model/gcl.py``, ``biomedkg/utils/fusion.py``, ``biomedkg/factory.py``
and ``biomedkg/gcl_module.py`` are imported unmodified from /root/reference and
executed; the three third-party packages they call and that are absent from
this image (torch_geometric 2.5.3, PyGCL 0.1.2, lightning) are bound to the
restatements in ``oracle/pyg.py`` / ``oracle/pygcl.py`` through ``sys.modules``
shims.  The vectors therefore pin the reference's own control flow, draw
order, parameter layout and loss plumbing; the third-party arithmetic itself
stays "parity unpinned" (oracle/__init__.py).

Each fixture ``<name>.pt`` holds: inputs (x, edge_index), the state_dict, the
recorded random draws in consumption order, and the reference outputs.
"""
from __future__ import annotations

import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import models as om  # noqa: E402
from oracle import pyg, pygcl  # noqa: E402

REF = "/root/reference"
DRAWS = om.TorchDraws(record=True)


def _install_shims():
    """Bind the absent third-party names to the oracle restatements."""
    tg = types.ModuleType("torch_geometric")
    tg_nn = types.ModuleType("torch_geometric.nn")
    tg_inits = types.ModuleType("torch_geometric.nn.inits")
    tg_utils = types.ModuleType("torch_geometric.utils")

    class _Unused(torch.nn.Module):  # RGCNConv / RGATConv / GAE: KGE path, never built here
        def __init__(self, *a, **k):
            raise NotImplementedError("KGE-path layer - out of scope")

    tg_nn.GCNConv = om.GCNConv
    tg_nn.RGCNConv = _Unused
    tg_nn.RGATConv = _Unused
    tg_nn.GAE = _Unused
    tg_inits.uniform = pyg.uniform_

    def dropout_edge(edge_index, p=0.5, force_undirected=False, training=True):
        mask = DRAWS.edge_mask(edge_index, p)
        return edge_index[:, mask], mask

    def mask_feature(x, p=0.5, mode="col", fill_value=0.0, training=True):
        assert mode == "all"
        mask = DRAWS.feature_mask(x, p)
        return x.masked_fill(~mask, fill_value), mask

    tg_utils.dropout_edge = dropout_edge
    tg_utils.mask_feature = mask_feature
    tg.nn, tg.utils = tg_nn, tg_utils
    tg_nn.inits = tg_inits

    gcl = types.ModuleType("GCL")
    gcl_losses = types.ModuleType("GCL.losses")
    gcl_models = types.ModuleType("GCL.models")
    gcl_losses.InfoNCE = pygcl.InfoNCE
    gcl_losses.JSD = pygcl.JSD
    gcl_models.DualBranchContrast = pygcl.DualBranchContrast
    gcl_models.SingleBranchContrast = pygcl.SingleBranchContrast
    gcl.losses, gcl.models = gcl_losses, gcl_models

    lightning = types.ModuleType("lightning")

    class LightningModule(torch.nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

    lightning.LightningModule = LightningModule
    omegaconf = types.ModuleType("omegaconf")
    omegaconf.DictConfig = dict

    for name, mod in {
        "torch_geometric": tg, "torch_geometric.nn": tg_nn, "torch_geometric.nn.inits": tg_inits,
        "torch_geometric.utils": tg_utils, "GCL": gcl, "GCL.losses": gcl_losses, "GCL.models": gcl_models,
        "lightning": lightning, "omegaconf": omegaconf,
    }.items():
        sys.modules[name] = mod


def _patch_draw_sites(gcl_model_mod):
    """The reference draws randperm / the GGD coin / dropout straight from torch.
    Record them (same generator, same order) so they can be replayed."""
    real_randperm, real_rand, real_dropout = torch.randperm, torch.rand, torch.nn.functional.dropout

    class _TorchProxy:
        def __getattr__(self, k):
            return getattr(torch, k)

        @staticmethod
        def randperm(n, *a, **k):
            return DRAWS._rec("randperm", real_randperm(n, *a, **k))

        @staticmethod
        def rand(*size, **k):
            r = real_rand(*size, **k)
            if tuple(size) == (1,):
                DRAWS._rec("coin", float(r.item()))
            return r

    gcl_model_mod.torch = _TorchProxy()
    return real_dropout


def graph(n, e, seed, self_loops=3, dups=5):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, e), generator=g, dtype=torch.int64)
    ei[1, :self_loops] = ei[0, :self_loops]          # pre-existing self-loops
    ei[:, e - dups:] = ei[:, :dups]                   # duplicate edges
    return ei


def redaf_train_fixture(ref_fusion, IN=32):
    """The reference's own ReDAF (utils/fusion.py:34-90) in TRAINING mode.  Its nn.Dropout draws from torch's generator
    and keeps no record, so the instance attribute `dropout` (not the source) is swapped for a module that draws the
    same Bernoulli(1-p) keep mask with torch.rand_like and remembers it; everything else is the reference's arithmetic."""

    class _RecordingDropout(torch.nn.Module):
        def __init__(self, p):
            super().__init__()
            self.p, self.mask = p, None

        def forward(self, x):
            self.mask = torch.rand_like(x) >= self.p
            return x * self.mask.to(x.dtype) / (1.0 - self.p)

    torch.manual_seed(31)
    red = ref_fusion.ReDAF(IN).train()
    with torch.no_grad():
        red.modal_weights.normal_(1.0, 0.5)               # some negative gates: the second ReLU must cut them
        red.transform_layer.bias.normal_(0.0, 0.2)
    red.dropout = _RecordingDropout(red.dropout.p)
    x = torch.randn(257, 2, IN)
    x = (x / x.norm(dim=1, keepdim=True)).to(torch.bfloat16).float().requires_grad_(True)   # representable on the device as is
    w = torch.randn(257, IN)
    out = red(x)
    (out * w).sum().backward()
    torch.save({"x": x.detach(), "w": w, "state_dict": {k: v.clone() for k, v in red.state_dict().items()},
                "mask": red.dropout.mask, "p": red.dropout.p, "out": out.detach(), "x_grad": x.grad.clone(),
                "grads": {k: p.grad.clone() for k, p in red.named_parameters() if p.grad is not None}},
               os.path.join(HERE, "fusion_redaf_m2_train.pt"))
    print("fusion_redaf_m2_train: out norm", float(out.norm()))


def main():
    _install_shims()
    sys.path.insert(0, REF)
    import biomedkg.gcl_module as ref_mod          # the reference, unmodified
    import biomedkg.model.encoder as ref_enc
    import biomedkg.model.gcl as ref_gcl
    import biomedkg.utils.fusion as ref_fusion

    only_new = "--only-redaf-train" in sys.argv      # added after the first batch of fixtures: leaves the others untouched
    if only_new:
        redaf_train_fixture(ref_fusion)

    _patch_draw_sites(ref_gcl)

    # dropout inside the reference encoder: record masks by replacing F.dropout
    def rec_dropout(x, p=0.5, training=True, inplace=False):
        if not training:
            return x
        keep = DRAWS.dropout_mask(x, p)
        return x * keep.to(x.dtype) / (1.0 - p)

    class _FProxy:
        def __getattr__(self, k):
            return getattr(torch.nn.functional, k)

        dropout = staticmethod(rec_dropout)

    ref_enc.F = _FProxy()

    IN, HID, L = 32, 64, 2
    out = {}

    class _LoggedDropout(torch.nn.Module):
        """stands in for the nn.Dropout INSTANCE of a reference ReDAF (the source is untouched): same Bernoulli(1-p) keep
        mask, drawn with torch.rand_like and logged like every other draw"""

        def __init__(self, p):
            super().__init__()
            self.p = p

        def forward(self, x):
            if not self.training:
                return x
            return x * DRAWS.dropout_mask(x, self.p).to(x.dtype) / (1.0 - self.p)

    def run(name, cls, n, e, seed, fuse, M, train=True, dtype=torch.float64, log_fuser_dropout=False):
        torch.manual_seed(seed)
        mod = cls(in_dim=IN, hidden_dim=HID, out_dim=HID, num_hidden_layers=L, fuse_method=fuse)
        if log_fuser_dropout:
            mod.modality_transform.dropout = _LoggedDropout(mod.modality_transform.dropout.p)
        mod.to(dtype)
        mod.train(train)
        g = torch.Generator().manual_seed(seed + 1)
        if M:
            x = torch.randn(n, M, IN, generator=g, dtype=dtype)
            x = x / x.norm(dim=1, keepdim=True)      # data/node.py:115-117
        else:
            x = torch.randn(n, IN, generator=g, dtype=dtype)
        ei = graph(n, e, seed + 2)
        DRAWS.log = []

        class Batch:
            pass

        b = Batch()
        b.x, b.edge_index = x, ei
        loss = mod.training_step(b)
        draws = list(DRAWS.log)
        loss.backward()
        grads = {k: p.grad.clone() for k, p in mod.named_parameters() if p.grad is not None}
        fx = {
            "x": x, "edge_index": ei, "state_dict": {k: v.clone() for k, v in mod.state_dict().items()},
            "draws": draws, "loss": loss.detach().clone(), "grads": grads,
            "cfg": dict(in_dim=IN, hidden_dim=HID, out_dim=HID, num_hidden_layers=L, fuse_method=fuse, M=M, cls=cls.__name__),
        }
        # embedding export path: BaseGCL.forward in eval mode (no dropout draw)
        mod.eval()
        with torch.no_grad():
            fx["embed_eval"] = mod(x, ei).clone()
        torch.save(fx, os.path.join(HERE, name + ".pt"))
        out[name] = float(loss)

    if only_new:
        run("grace_redaf_train", ref_mod.GRACEModule, 80, 300, 17, "redaf", 2, train=True, dtype=torch.float32, log_fuser_dropout=True)
        print({k: round(v, 6) for k, v in out.items()})
        return
    run("grace_redaf_train", ref_mod.GRACEModule, 80, 300, 17, "redaf", 2, train=True, dtype=torch.float32, log_fuser_dropout=True)
    run("grace_none", ref_mod.GRACEModule, 96, 400, 10, "none", 0)
    run("grace_mean2", ref_mod.GRACEModule, 80, 300, 11, None, 2)
    run("grace_attention", ref_mod.GRACEModule, 72, 260, 12, "attention", 2)
    run("dgi_none", ref_mod.DGIModule, 90, 350, 13, "none", 0)
    run("ggd_none_a", ref_mod.GGDModule, 90, 350, 14, "none", 0)     # coin decides the aug branch
    run("ggd_none_b", ref_mod.GGDModule, 90, 350, 19, "none", 0)
    # ReDAF hard-codes an fp32 torch.full (utils/fusion.py:54-56): fp32 fixture
    run("ggd_redaf", ref_mod.GGDModule, 64, 200, 16, "redaf", 2, train=False, dtype=torch.float32)

    # stand-alone fusion modules straight from the reference (no shim involved)
    torch.manual_seed(21)
    att = ref_fusion.AttentionFusion(IN).double()
    x3 = torch.randn(50, 3, IN, dtype=torch.float64)
    torch.save({"x": x3, "state_dict": att.state_dict(), "out": att(x3).detach()}, os.path.join(HERE, "fusion_attention_m3.pt"))
    red = ref_fusion.ReDAF(IN).eval()
    x2 = torch.randn(50, 2, IN)
    torch.save({"x": x2, "state_dict": red.state_dict(), "out": red(x2).detach()}, os.path.join(HERE, "fusion_redaf_m2.pt"))

    # the reference's own encoder class in eval mode on a graph with edge cases
    torch.manual_seed(22)
    enc = ref_enc.GCNEncoder(IN, HID, HID, L).double().eval()
    n = 40
    ei = graph(n, 150, 23)
    ei = ei[:, (ei[0] != 7) & (ei[1] != 7)]          # node 7 isolated
    xg = torch.randn(n, IN, dtype=torch.float64)
    with torch.no_grad():
        torch.save({"x": xg, "edge_index": ei, "state_dict": enc.state_dict(), "out": enc(xg, ei)}, os.path.join(HERE, "gcn_encoder_eval.pt"))

    redaf_train_fixture(ref_fusion)

    for k, v in out.items():
        print(f"{k}: loss={v:.9f}")
    print("ggd coins:", [d for d in DRAWS.log if d[0] == "coin"][:1])


if __name__ == "__main__":
    main()
