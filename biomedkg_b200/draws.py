"""Random draws of the GCL step, in the reference's order.

``DeviceDraws`` is the product default: feature / edge masks are drawn by torch
on the tensor's device exactly as PyG's ``mask_feature`` / ``dropout_edge`` do
(``torch.rand_like(x) >= p`` / ``torch.rand(E) >= p``), ``randperm`` and the GGD
coin come from the CPU generator as in the reference, and encoder dropout is a
counter-based stream evaluated inside the aggregation kernel (seeded from
``torch.initial_seed()``).  ``ReplayDraws`` feeds recorded draws back (tests).
"""
from __future__ import annotations

import torch

_MASK64 = 0xFFFFFFFFFFFFFFFF


def _splitmix(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _MASK64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK64
    return z ^ (z >> 31)


def hash_keep_mask(seed: int, numel: int, p: float) -> torch.Tensor:
    """Host mirror of csrc/common.cuh:hash_keep - keep[i] = hash_u32(seed, i) >= p * 2^32."""
    import numpy as np

    idx = np.arange(numel, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed & _MASK64)
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    u = (z >> np.uint64(32)).astype(np.uint64)
    thr = np.uint64(int(p * 4294967296.0))
    return torch.from_numpy(u >= thr)


def hash_keep_mask16(seed: int, numel: int, p: float) -> torch.Tensor:
    """Host mirror of the ReDAF kernels' stream (csrc/redaf.cu): one hash_u32 per two consecutive elements,
    keep[i] = 16-bit half (i & 1) of hash_u32(seed, i >> 1) >= p * 2^16."""
    import numpy as np

    idx = np.arange((numel + 1) // 2, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed & _MASK64)
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    h = (z >> np.uint64(32)).astype(np.uint64)
    halves = np.stack([h & np.uint64(0xFFFF), h >> np.uint64(16)], axis=1).reshape(-1)[:numel]
    return torch.from_numpy(halves >= np.uint64(int(p * 65536.0)))


class DeviceDraws:
    """``stream`` separates the counter-hash dropout streams of different owners (encoder, ReDAF, ...: every instance gets its
    own id), ``salt`` lets data-parallel ranks that each train on their OWN graph decorrelate their masks (set it to the
    rank; the row-sharded path must keep it equal on all ranks - they evaluate one global mask), and ``counter`` can be
    restored from a checkpoint (``state`` / ``load_state``) so a resumed run does not replay the masks it already used."""

    _instances = 0

    def __init__(self, salt: int = 0):
        self._counter = 0
        DeviceDraws._instances += 1
        self.stream = DeviceDraws._instances
        self.salt = int(salt)

    def state(self):
        return {"counter": self._counter, "stream": self.stream, "salt": self.salt}

    def load_state(self, st):
        self._counter, self.stream, self.salt = int(st["counter"]), int(st["stream"]), int(st["salt"])

    def feature_mask(self, x, p):
        return torch.rand_like(x, dtype=torch.float32) >= p

    def edge_mask(self, edge_index, p):
        return torch.rand(edge_index.size(1), device=edge_index.device) >= p

    def dropout(self, shape, p, device):
        """-> (seed, explicit_keep_mask_or_None) for the fused aggregation epilogue."""
        self._counter += 1
        return _splitmix(torch.initial_seed() ^ _splitmix(self._counter) ^ _splitmix((self.stream << 32) ^ self.salt ^ 0x5bd1e995)), None

    def randperm(self, n):
        return torch.randperm(n)

    def coin(self):
        return float(torch.rand(1).item())


class GraphSafeDraws(DeviceDraws):
    """Draws for a CUDA-graph-captured step (biomedkg_b200/graphed.py).

    * A captured kernel's scalar arguments are frozen, so the per-call dropout seed of ``DeviceDraws`` would replay the same
      mask forever: the encoder dropout mask is drawn by torch's device generator (whose Philox offset a captured graph
      advances on every replay) and handed to the aggregation epilogue as an explicit keep mask.
    * ``randperm`` (DGI / GGD corruption, model/gcl.py:17,66) comes from the CPU generator in the reference.  Here it returns a
      STATIC device buffer; ``refresh()`` - called by ``GraphedStep`` before every replay - draws the permutation on the host
      exactly as the reference does (``torch.randperm(n)``) and copies it into the buffer from pinned memory.
    * ``coin`` (GGD's augmentation branch, model/gcl.py:74) is a host decision: one graph is captured per branch with the coin
      pinned to that side (``coin_value``) and the host flips the real coin to choose which graph to replay."""

    def __init__(self, coin_value=None):
        super().__init__()
        self._coin = coin_value
        self._perms = {}

    def dropout(self, shape, p, device):
        return 0, torch.rand(shape, device=device) >= p

    def randperm(self, n):
        ent = self._perms.get(n)
        if ent is None:
            host = torch.randperm(n).pin_memory()
            ent = self._perms[n] = (host.cuda(non_blocking=True), host)
        return ent[0]

    def refresh(self):
        for n, (dev, host) in self._perms.items():
            host.copy_(torch.randperm(n))
            dev.copy_(host, non_blocking=True)

    def coin(self):
        return super().coin() if self._coin is None else self._coin


class ReplayDraws:
    """Replays (kind, value) records, e.g. the ``draws`` list of a golden fixture."""

    def __init__(self, log, device):
        self.log, self.i, self.device = list(log), 0, device

    def _next(self, kind):
        k, v = self.log[self.i]
        if k != kind:
            raise AssertionError(f"draw order mismatch: wanted {kind}, recorded {k} at position {self.i}")
        self.i += 1
        return v

    def feature_mask(self, x, p):
        return self._next("feature_mask").to(self.device)

    def edge_mask(self, edge_index, p):
        return self._next("edge_mask").to(self.device)

    def dropout(self, shape, p, device):
        return 0, self._next("dropout_mask").to(self.device)

    def randperm(self, n):
        return self._next("randperm")

    def coin(self):
        return self._next("coin")


def set_draws(module: torch.nn.Module, draws) -> torch.nn.Module:
    for m in module.modules():
        if hasattr(m, "draws"):
            m.draws = draws
    return module
