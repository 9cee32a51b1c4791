"""GCL task modules - same surface as biomedkg/gcl_module.py (BaseGCL, DGIModule, GRACEModule, GGDModule).

Lightning is an optional dependency: when it imports, the classes derive from
``LightningModule`` exactly like the reference; otherwise from a minimal stand-in with
the methods the reference uses (``save_hyperparameters``, ``log``), so the modules also
run under the dependency-free driver in ``biomedkg_b200/train_gcl.py``.
"""
from __future__ import annotations

import math

import torch

from . import losses as L
from .factory import FusionFactory
from .losses import DualBranchContrast, SingleBranchContrast
from .model import GATEncoder, GCNEncoder
from .model.gcl import DGI, GGD, GRACE
from . import ops

try:  # pragma: no cover - lightning is not installed in the build image
    from lightning import LightningModule
except Exception:  # noqa: BLE001

    class LightningModule(torch.nn.Module):
        """The slice of lightning.LightningModule the reference uses: hyper-parameter capture, ``log`` and the
        checkpoint layout of ``Trainer.save_checkpoint`` / ``load_from_checkpoint`` (train_gcl.py:78-83, node.py:204-209):
        ``{"state_dict": ..., "hyper_parameters": <kwargs of the outermost __init__>}``."""

        trainer = None

        def save_hyperparameters(self, *args, ignore=None, **kwargs):
            import inspect

            ignore = set([ignore] if isinstance(ignore, str) else (ignore or ()))
            frame, outer = inspect.currentframe().f_back, None
            while frame is not None and frame.f_code.co_name == "__init__" and frame.f_locals.get("self") is self:
                outer, frame = frame, frame.f_back           # BaseGCL.__init__ <- GRACEModule.__init__ <- caller
            hp = {}
            if outer is not None:
                info = inspect.getargvalues(outer)
                for name in info.args[1:]:
                    if name not in ignore:
                        hp[name] = info.locals[name]
                if info.keywords:
                    hp.update({k: v for k, v in info.locals[info.keywords].items() if k not in ignore})
            self.hparams = hp

        def log(self, name, value, **kwargs):
            self.logged = getattr(self, "logged", {})
            self.logged[name] = value.detach() if torch.is_tensor(value) else value

        @property
        def device(self):
            return next(self.parameters()).device

        def save_checkpoint(self, path):
            torch.save({"state_dict": self.state_dict(), "hyper_parameters": dict(self.hparams)}, path)

        @classmethod
        def load_from_checkpoint(cls, checkpoint_path, map_location=None, strict=True, **overrides):
            import inspect

            ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
            accepted = inspect.signature(cls.__init__).parameters
            hp = {k: v for k, v in dict(ckpt.get("hyper_parameters", {})).items() if k in accepted}
            hp.update(overrides)
            obj = cls(**hp)
            obj.load_state_dict(ckpt["state_dict"], strict=strict)
            return obj.to(map_location) if map_location is not None else obj


def _cosine_with_warmup(optimizer, num_warmup_steps, num_training_steps):
    """transformers.get_cosine_schedule_with_warmup (num_cycles=0.5)."""

    def f(step):
        if step < num_warmup_steps:
            return step / max(1, num_warmup_steps)
        prog = (step - num_warmup_steps) / max(1, num_training_steps - num_warmup_steps)
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * prog)))

    return torch.optim.lr_scheduler.LambdaLR(optimizer, f)


def _linear_with_warmup(optimizer, num_warmup_steps, num_training_steps):
    """transformers.get_linear_schedule_with_warmup."""

    def f(step):
        if step < num_warmup_steps:
            return step / max(1, num_warmup_steps)
        return max(0.0, (num_training_steps - step) / max(1, num_training_steps - num_warmup_steps))

    return torch.optim.lr_scheduler.LambdaLR(optimizer, f)


class BaseGCL(LightningModule):
    """gcl_module.py:17-100."""

    def __init__(self, model, embed_dim, scheduler_type, learning_rate, warm_up_ratio, feature_embedding_dim,
                 contrast_model=None, fuse_method=None):
        super().__init__()
        self.save_hyperparameters(ignore=["model"])
        self.model = model
        self.modality_transform = FusionFactory.create_fuser(method=fuse_method, embed_dim=embed_dim)
        self.contrast_model = contrast_model
        self.lr = learning_rate
        self.scheduler_type = scheduler_type
        self.warm_up_ratio = warm_up_ratio
        self.feature_embedding_dim = feature_embedding_dim

    def fusion_fn(self, x):
        if self.modality_transform:
            x = self.modality_transform(x)
        elif x.dim() == 3:
            x = ops.modality_mean(x)
        return x

    def calculate_loss(self, x, edge_index):
        raise NotImplementedError

    def forward(self, x, edge_index):
        x = self.fusion_fn(x=x)
        return self.model.encoder(x, edge_index)

    def training_step(self, batch, batch_idx=None):
        x = self.fusion_fn(x=batch.x)
        loss = self.calculate_loss(x, batch.edge_index)
        self.log("train_loss", loss, on_epoch=True, on_step=True, prog_bar=True)
        return loss

    def validation_step(self, batch, batch_idx=None):
        x = self.fusion_fn(x=batch.x)
        loss = self.calculate_loss(x, batch.edge_index)
        self.log("val_loss", loss, on_epoch=True, on_step=True, prog_bar=True)
        return loss

    def test_step(self, batch, batch_idx=None):
        x = self.fusion_fn(x=batch.x)
        loss = self.calculate_loss(x, batch.edge_index)
        self.log("test_loss", loss, on_epoch=True, on_step=True, prog_bar=True)
        return loss

    def configure_optimizers(self):
        # the reference optimises self.model only: fuser weights stay at init (gcl_module.py:81)
        optimizer = torch.optim.Adam(self.model.parameters(), lr=self.lr)
        return {"optimizer": optimizer, "lr_scheduler": self._get_scheduler(optimizer=optimizer)}

    def _get_scheduler(self, optimizer, num_training_steps=None):
        if num_training_steps is None:
            num_training_steps = int(self.trainer.estimated_stepping_batches)
        warm = int(num_training_steps * self.warm_up_ratio)
        if self.scheduler_type == "linear":
            return _linear_with_warmup(optimizer, warm, num_training_steps)
        if self.scheduler_type == "cosine":
            return _cosine_with_warmup(optimizer, warm, num_training_steps)


def _make_encoder(kind, in_dim, hidden_dim, out_dim, num_hidden_layers):
    """kind="gcn" is the reference's encoder (gcl_module.py:118,161,208); "gat" is the BASELINE.json extension."""
    if kind == "gcn":
        return GCNEncoder(in_dim=in_dim, hidden_dim=hidden_dim, out_dim=out_dim, num_hidden_layers=num_hidden_layers)
    if kind == "gat":
        return GATEncoder(in_dim=in_dim, hidden_dim=hidden_dim, out_dim=out_dim, num_hidden_layers=num_hidden_layers)
    raise ValueError(f"unknown encoder {kind!r}")


def _common(self_cls, model, in_dim, scheduler_type, learning_rate, warm_up_ratio, fuse_method, contrast_model=None):
    return dict(model=model, embed_dim=in_dim, scheduler_type=scheduler_type, learning_rate=learning_rate,
                warm_up_ratio=warm_up_ratio, feature_embedding_dim=in_dim, contrast_model=contrast_model,
                fuse_method=fuse_method)


class DGIModule(BaseGCL):
    """gcl_module.py:103-143."""

    def __init__(self, in_dim, hidden_dim, out_dim, num_hidden_layers, scheduler_type="cosine", learning_rate=2e-4,
                 warm_up_ratio=0.03, fuse_method=None, encoder="gcn"):
        model = DGI(encoder=_make_encoder(encoder, in_dim, hidden_dim, out_dim, num_hidden_layers),
                    hidden_dim=hidden_dim)
        contrast_model = SingleBranchContrast(loss=L.JSD(), mode="G2L")
        BaseGCL.__init__(self, **_common(self, model, in_dim, scheduler_type, learning_rate, warm_up_ratio, fuse_method, contrast_model))

    def calculate_loss(self, x, edge_index):
        pos_z, summary, neg_z = self.model(x, edge_index)
        return self.contrast_model(h=pos_z, g=summary, hn=neg_z)


class GRACEModule(BaseGCL):
    """gcl_module.py:146-190."""

    def __init__(self, in_dim, hidden_dim, out_dim, num_hidden_layers, scheduler_type="cosine", learning_rate=2e-4,
                 warm_up_ratio=0.03, fuse_method=None, encoder="gcn"):
        model = GRACE(encoder=_make_encoder(encoder, in_dim, hidden_dim, out_dim, num_hidden_layers),
                      hidden_dim=hidden_dim, proj_dim=hidden_dim)
        contrast_model = DualBranchContrast(loss=L.InfoNCE(tau=0.2), mode="L2L", intraview_negs=True)
        BaseGCL.__init__(self, **_common(self, model, in_dim, scheduler_type, learning_rate, warm_up_ratio, fuse_method, contrast_model))

    def calculate_loss(self, x, edge_index):
        _, z1, z2 = self.model(x, edge_index)
        h1, h2 = [self.model.project(z) for z in [z1, z2]]
        return self.contrast_model(h1, h2)


class GGDModule(BaseGCL):
    """gcl_module.py:193-234."""

    def __init__(self, in_dim, hidden_dim, out_dim, num_hidden_layers, scheduler_type="cosine", learning_rate=2e-4,
                 warm_up_ratio=0.03, fuse_method=None, encoder="gcn"):
        model = GGD(encoder=_make_encoder(encoder, in_dim, hidden_dim, out_dim, num_hidden_layers),
                    hidden_dim=hidden_dim, n_proj=1, aug_p=0.5)
        BaseGCL.__init__(self, **_common(self, model, in_dim, scheduler_type, learning_rate, warm_up_ratio, fuse_method))

    def calculate_loss(self, x, edge_index):
        pos_h, neg_h = self.model(x, edge_index)
        return L.bce_with_logits_pos_neg(pos_h, neg_h)
