"""Oracle for the embedding-export path (TEST INFRASTRUCTURE ONLY - see oracle/__init__.py).

Restates, as plain per-seed loops, what ``GCLEncode._get_embeddings`` does per node type
(biomedkg/data/node.py:193-241): iterate ``PrimeKGModule.subgraph_dataloader()``
(biomedkg/data_module.py:71-79: ``NeighborLoader(data, num_neighbors=[-1], shuffle=False)``, default batch size 1),
run ``model(batch.x, batch.edge_index)`` = ``BaseGCL.forward`` (biomedkg/gcl_module.py:55-58) under ``no_grad`` and keep
``out[: batch.batch_size]`` (node.py:229-234).

PARITY UNPINNED for the sampler: ``torch_geometric.loader.NeighborLoader`` (torch_geometric == 2.5.3, pyproject.toml:6)
is not installable here.  Its published behaviour for ``num_neighbors=[-1]``, one hop, one seed, homogeneous ``Data``,
restated in ``one_hop_batch``: the batch's nodes are the seed followed by its distinct in-neighbours (sources of edges
whose target is the seed) in order of first appearance; its edges are exactly the sampled edges neighbour -> seed
(all of them - duplicates kept, an existing seed -> seed loop included), relabelled to batch-local ids; no edges between
neighbours and no out-edges of the seed are included.
"""
from __future__ import annotations

import torch


def one_hop_batch(edge_index: torch.Tensor, seed: int):
    """(node ids [seed, neighbours...], local edge_index [2, deg]) of NeighborLoader(num_neighbors=[-1], batch_size=1)."""
    src, dst = edge_index[0], edge_index[1]
    sel = (dst == seed).nonzero(as_tuple=True)[0]
    nodes, local = [int(seed)], {int(seed): 0}
    rows = []
    for e in sel.tolist():
        j = int(src[e])
        if j not in local:
            local[j] = len(nodes)
            nodes.append(j)
        rows.append(local[j])
    ei = torch.tensor([rows, [0] * len(rows)], dtype=torch.int64).reshape(2, -1)
    return torch.tensor(nodes, dtype=torch.int64), ei


@torch.no_grad()
def export_loop(module, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    """node.py:224-236 for one node type: [N, out_dim], row s = module(batch_s.x, batch_s.edge_index)[:1].
    ``module`` is an oracle.models.*Module in eval mode (the reference leaves dropout on; see biomedkg_b200/export.py)."""
    outs = []
    for s in range(x.size(0)):
        nodes, ei = one_hop_batch(edge_index, s)
        outs.append(module(x[nodes], ei)[:1])
    return torch.cat(outs, dim=0)


@torch.no_grad()
def export_two_chains(module, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    """Closed form of ``export_loop`` for a GCNEncoder (cross-checked against it in tests/test_oracle.py): one leaf
    chain shared by all stars plus one seed chain, each layer one pass over the full edge list."""
    h = module.fusion_fn(x=x)
    N = h.size(0)
    src, dst = edge_index[0], edge_index[1]
    keep = src != dst                                  # gcn_norm drops existing self-loops and adds exactly one
    src, dst = src[keep], dst[keep]
    deg = torch.ones(N, dtype=h.dtype).index_add_(0, dst, torch.ones(dst.numel(), dtype=h.dtype))
    dis = deg.pow(-0.5)
    leaf = seed = h
    layers = list(module.model.encoder.graph_layers)
    for li, layer in enumerate(layers):
        t_leaf, t_seed = leaf @ layer.lin.weight.t(), seed @ layer.lin.weight.t()
        agg = torch.zeros_like(t_seed).index_add_(0, dst, t_leaf[src])
        seed = dis[:, None] * (agg + dis[:, None] * t_seed) + layer.bias
        leaf = t_leaf + layer.bias
        if li < len(layers) - 1:
            seed, leaf = torch.relu(seed), torch.relu(leaf)
    return seed
