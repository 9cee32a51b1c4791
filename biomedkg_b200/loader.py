"""Device-side neighbour loader - the reference's mini-batch regime (biomedkg/data_module.py:65-125).

``NeighborLoader`` mirrors the slice of ``torch_geometric.loader.NeighborLoader`` the reference uses
(``data, num_neighbors, batch_size, shuffle``; homogeneous ``Data``; ``num_workers=0``): it yields batches with ``.x``,
``.edge_index`` (batch-local, row = source, col = target), ``.n_id``, ``.e_id`` and ``.batch_size`` whose first ``batch_size``
nodes are the seeds, so ``GCLModule.training_step(batch)`` and ``model(batch.x, batch.edge_index)[: batch.batch_size]``
(node.py:229-234) work unchanged.  Sampling runs on the GPU over the sorted parent graph (``csrc/sampler.cu``); the CPU
never touches the edge list.  ``random_link_split`` is the ``T.RandomLinkSplit(num_val, num_test, neg_sampling_ratio=0.0)``
of data_module.py:65-69 restricted to what the GCL stage reads (the message-passing ``edge_index`` of each split).

Random streams are this package's own (PyG's C++ sampler RNG cannot be reproduced): draws are counter-based on
``(seed, hop, node)``, the seed advancing once per batch; ``oracle/sampler.py`` restates the same stream for bit-exact tests.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Iterator, Sequence

import torch

from . import ops
from ._cabi import call, lib
from .ops import _need_cuda, _p, _stream, _ws

_I32_MAX = 0x7FFFFFFF


class NeighborSampler:
    """k-hop in-neighbour sampling from one graph.  ``sample(seeds, seed)`` -> (n_id, edge_index, e_id), all int64 on the device."""

    def __init__(self, edge_index: torch.Tensor, num_nodes: int, num_neighbors: Sequence[int]):
        _need_cuda(edge_index)
        for f in num_neighbors:
            if not (f == -1 or 1 <= f <= 32):
                raise NotImplementedError("fan-outs of 1..32 or -1 (all neighbours) are supported (the reference uses 30 and -1)")
        self.N, self.num_neighbors = int(num_nodes), [int(f) for f in num_neighbors]
        s = ops.sorted_graph(edge_index, self.N).by[0]                    # raw CSR by destination
        self.rowptr, self.colind, self.eperm = s.rowptr_raw, s.minor, s.perm
        dev = edge_index.device
        self.local_id = torch.full((self.N,), -1, dtype=torch.int32, device=dev)
        self.first_pos = torch.full((self.N,), _I32_MAX, dtype=torch.int32, device=dev)
        self._count = torch.zeros(1, dtype=torch.int32, device=dev)

    def sample(self, seeds: torch.Tensor, seed: int):
        _need_cuda(seeds)
        dev = seeds.device
        frontier = seeds.to(torch.int32).contiguous()
        S = int(frontier.numel())
        if S == 0:
            raise ValueError("empty seed list")
        call("bmkg_sample_set_ids", _p(frontier), S, _p(self.local_id), 0, _stream())
        nodes, rows, cols, eids = [frontier], [], [], []
        n_nodes, base = S, 0
        for hop, fanout in enumerate(self.num_neighbors):
            F = int(frontier.numel())
            if F == 0:
                break
            off = torch.empty(F + 1, dtype=torch.int32, device=dev)
            ws = _ws(lib.bmkg_sample_workspace_bytes(F), dev)
            call("bmkg_sample_count", _p(self.rowptr), _p(frontier), F, fanout, _p(off), _p(ws), ws.numel(), _stream())
            T = int(off[F].item())                                          # host needs the size to allocate the hop's edges
            if T == 0:
                frontier = frontier[:0]
                continue
            src = torch.empty(T, dtype=torch.int32, device=dev)
            col = torch.empty(T, dtype=torch.int64, device=dev)
            eid = torch.empty(T, dtype=torch.int64, device=dev)
            call("bmkg_sample_pick", _p(self.rowptr), _p(self.colind), _p(self.eperm), _p(frontier), F, fanout, _p(off),
                 int(seed) & 0xFFFFFFFFFFFFFFFF, hop, base, _p(src), _p(col), _p(eid), _stream())
            row = torch.empty(T, dtype=torch.int64, device=dev)
            new_nodes = torch.empty(T, dtype=torch.int32, device=dev)
            ws = _ws(lib.bmkg_sample_workspace_bytes(T), dev)
            call("bmkg_sample_relabel", _p(src), T, n_nodes, _p(self.local_id), _p(self.first_pos), _p(new_nodes), _p(self._count),
                 _p(row), _p(ws), ws.numel(), _stream())
            n_new = int(self._count.item())
            rows.append(row)
            cols.append(col)
            eids.append(eid)
            base, frontier = n_nodes, new_nodes[:n_new]
            nodes.append(frontier)
            n_nodes += n_new
        n_id32 = torch.cat(nodes)
        call("bmkg_sample_set_ids", _p(n_id32), int(n_id32.numel()), _p(self.local_id), 1, _stream())   # map back to all -1
        if rows:
            edge_index = torch.stack([torch.cat(rows), torch.cat(cols)])
            e_id = torch.cat(eids)
        else:
            edge_index = torch.empty(2, 0, dtype=torch.int64, device=dev)
            e_id = torch.empty(0, dtype=torch.int64, device=dev)
        return n_id32.to(torch.int64), edge_index, e_id


class NeighborLoader:
    """``NeighborLoader(data, num_neighbors, batch_size=1, shuffle=False)`` as called at data_module.py:74-79,82-87,92-98.
    ``data`` needs ``.x`` and ``.edge_index`` (device tensors); ``input_nodes`` defaults to every node."""

    def __init__(self, data, num_neighbors: Sequence[int], batch_size: int = 1, shuffle: bool = False, input_nodes=None,
                 num_workers: int = 0, seed: int | None = None, **unused):
        self.data = data
        self.x, ei = data.x, data.edge_index
        _need_cuda(self.x, ei)
        self.N = int(self.x.size(0))
        self.batch_size, self.shuffle = int(batch_size), bool(shuffle)
        self.sampler = NeighborSampler(ei, self.N, num_neighbors)
        self.input_nodes = torch.arange(self.N, device=ei.device) if input_nodes is None else input_nodes.to(ei.device).long()
        self.base_seed = torch.initial_seed() if seed is None else int(seed)
        self._epoch = 0

    def __len__(self) -> int:
        return (int(self.input_nodes.numel()) + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[SimpleNamespace]:
        order = self.input_nodes
        if self.shuffle:   # torch's CPU generator, like the RandomSampler behind PyG's loader
            order = order[torch.randperm(order.numel()).to(order.device)]
        self._epoch += 1
        for b, start in enumerate(range(0, int(order.numel()), self.batch_size)):
            seeds = order[start : start + self.batch_size]
            n_id, edge_index, e_id = self.sampler.sample(seeds, self.base_seed + (self._epoch << 32) + b)
            yield SimpleNamespace(x=self.x.index_select(0, n_id), edge_index=edge_index, n_id=n_id, e_id=e_id,
                                  batch_size=int(seeds.numel()), input_id=seeds)


def random_link_split(data, num_val: float = 0.1, num_test: float = 0.2, generator: torch.Generator | None = None):
    """``T.RandomLinkSplit(num_val, num_test, neg_sampling_ratio=0.0)(data)`` (data_module.py:65-69), directed graph,
    restricted to the message-passing edges the GCL stage reads: train sees the training edges, validation sees the training
    edges, test sees training + validation edges.  The supervision labels (edge_label*) are KGE-stage inputs and out of scope."""
    ei = data.edge_index
    E = int(ei.size(1))
    n_val, n_test = int(num_val * E) if isinstance(num_val, float) else int(num_val), int(num_test * E) if isinstance(num_test, float) else int(num_test)
    n_train = E - n_val - n_test
    if n_train <= 0:
        raise ValueError("Insufficient number of edges for training")
    perm = torch.randperm(E, generator=generator).to(ei.device)
    train_e, val_e = perm[:n_train], perm[n_train : n_train + n_val]
    mk = lambda idx: SimpleNamespace(x=data.x, edge_index=ei.index_select(1, idx).contiguous())  # noqa: E731
    return mk(train_e), mk(train_e), mk(torch.cat([train_e, val_e]))
