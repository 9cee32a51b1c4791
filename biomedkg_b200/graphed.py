"""CUDA-graph capture of the GCL training step.

A full-graph GRACE step at BASELINE cfg 2 is ~250 kernel launches of 5-50 us each: issued one by one from Python the
GPU waits on the host between them (kernel time 6.8 ms, eager step 8.1-8.9 ms).  ``GraphedStep`` captures
``loss = module.training_step(batch); loss.backward()`` once into a CUDA graph over static input buffers and replays it:
one launch per step, no Python between kernels.  The optimiser tail (gradient all-reduce, clipping, Adam - a handful of
fused launches) stays outside so any optimiser / DDP wrapper keeps working.

What makes the step capturable: every kernel of this package launches on the current stream with caller-provided
workspaces and no host synchronisation; the random draws come from torch's device generator, whose Philox offset a
captured graph advances per replay (``GraphSafeDraws`` replaces the host-seeded dropout stream by explicit masks).
``resort=True`` also captures the radix sort of ``edge_index`` - use it when every step brings a new edge list
(mini-batches); with ``resort=False`` the sort is done once at capture time and only the features may change.
Shapes are static: a new (N, E) needs a new ``GraphedStep``.

DGI and GGD draw from the CPU generator inside the step (the corruption permutation, model/gcl.py:17,66, and GGD's
augmentation coin, :74).  ``GraphSafeDraws`` turns the permutation into a static device buffer that the host refreshes
before every replay, and ``graphed_step`` captures GGD once per coin branch and lets the host's coin pick the graph.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from . import ops
from .draws import GraphSafeDraws, set_draws


class GraphedStep:
    def __init__(self, module, x: torch.Tensor, edge_index: torch.Tensor, resort: bool = False, warmup: int = 3,
                 loss_fn=None, coin_value=None):
        from .model.gcl import GGD

        if loss_fn is None and isinstance(getattr(module, "model", None), GGD) and coin_value is None:
            raise ValueError("GGD's augmentation coin is a host-side branch: capture one graph per branch (graphed_step(module, ...))")
        ops._need_cuda(x, edge_index)
        self.module = module
        self.resort = bool(resort)
        self.x = x.clone()
        self.edge_index = edge_index.clone()
        self._batch = SimpleNamespace(x=self.x, edge_index=self.edge_index)
        self._loss_fn = loss_fn or (lambda m, b: m.training_step(b))
        self.draws = GraphSafeDraws(coin_value)
        set_draws(module, self.draws)
        params = [p for p in module.parameters() if p.requires_grad]
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):     # warm-up on a side stream (lazy initialisation, autograd threads, cuBLAS workspaces)
            for _ in range(max(1, warmup)):
                for p in params:
                    p.grad = None
                self._fwd_bwd()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for p in params:
            p.grad = None
        launches0 = _launch_counter()
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: helper threads of the process (NCCL watchdog, clock samplers) may make CUDA calls during the capture
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.loss = self._fwd_bwd()
        self.launches_per_replay = _launch_counter() - launches0
        self.params = params
        #: the gradient buffers the captured backward writes (graph-pool memory); they must be the parameters' ``.grad``
        #: when the optimiser runs - re-attach them if anything reset ``.grad`` in between (e.g. ``zero_grad(set_to_none=True)``)
        self.grads = [p.grad for p in params]

    def _fwd_bwd(self):
        prev = ops._CAPTURE_RESORT
        ops._CAPTURE_RESORT = self.resort
        try:
            loss = self._loss_fn(self.module, self._batch)
            loss.backward()
        finally:
            ops._CAPTURE_RESORT = prev
        return loss.detach()

    def __call__(self, x: torch.Tensor | None = None, edge_index: torch.Tensor | None = None) -> torch.Tensor:
        """Copy the batch into the static buffers (device-to-device or pinned-host-to-device, stream-ordered), replay.
        Returns the static loss tensor; parameter ``.grad`` s hold this step's gradients."""
        if x is not None and x.data_ptr() != self.x.data_ptr():
            if x.shape != self.x.shape:
                raise ValueError(f"captured for x of shape {tuple(self.x.shape)}, got {tuple(x.shape)}")
            self.x.copy_(x, non_blocking=True)
        if edge_index is not None and edge_index.data_ptr() != self.edge_index.data_ptr():
            if not self.resort:
                raise ValueError("this step was captured with resort=False: its edge list is fixed")
            if edge_index.shape != self.edge_index.shape:
                raise ValueError(f"captured for edge_index of shape {tuple(self.edge_index.shape)}, got {tuple(edge_index.shape)}")
            self.edge_index.copy_(edge_index, non_blocking=True)
        for p, g in zip(self.params, self.grads):
            if p.grad is not g:
                p.grad = g
        set_draws(self.module, self.draws)     # another captured branch of the same module may have installed its own
        self.draws.refresh()                   # host-side draws of this step (corruption permutations) into the static buffers
        self.graph.replay()
        return self.loss


def _launch_counter() -> int:
    from . import _cabi

    return _cabi.kernel_launches


class GraphedGGDStep:
    """GGD (model/gcl.py:54-93): the coin ``torch.rand(1) < p`` decides on the host whether the positive pass is augmented, so
    the step is captured twice - coin pinned below / above ``p`` - and every call flips the real coin (CPU generator, the
    reference's draw) and replays the matching graph."""

    def __init__(self, module, x, edge_index, **kw):
        self.module = module
        self.p = float(module.model.p)
        self.branches = {True: GraphedStep(module, x, edge_index, coin_value=0.0, **kw),
                         False: GraphedStep(module, x, edge_index, coin_value=1.0, **kw)}
        self.launches_per_replay = max(b.launches_per_replay for b in self.branches.values())

    def __call__(self, x=None, edge_index=None):
        aug = float(torch.rand(1).item()) < self.p
        return self.branches[aug](x, edge_index)


def graphed_step(module, x, edge_index, **kw):
    """The captured training step for any of the three GCL modules."""
    from .model.gcl import GGD

    if isinstance(getattr(module, "model", None), GGD):
        return GraphedGGDStep(module, x, edge_index, **kw)
    return GraphedStep(module, x, edge_index, **kw)
