"""Embedding export - the path the KGE stage consumes (biomedkg/data/node.py:155-241).

The reference's ``GCLEncode._get_embeddings`` (node.py:193-241) loads a trained GCL module and calls
``BaseGCL.forward`` (gcl_module.py:55-58) once per node over ``PrimeKGModule.subgraph_dataloader()`` =
``NeighborLoader(num_neighbors=[-1])`` with the default batch size 1 (data_module.py:71-79): every batch is ONE seed
node plus all its in-neighbours, and the only edges are neighbour -> seed.  It keeps row 0 (``[: batch.batch_size]``,
node.py:232-234) and pickles ``{node name: float32 [1, out_dim]}`` (node.py:236-241).

On such a star graph ``gcn_norm`` gives every neighbour degree 1 (its own self-loop) and the seed its full-graph
in-degree, so the N tiny forward passes collapse into two chains over the whole graph, evaluated here layer by layer:

    leaf chain   L(l+1)[j] = act(W_l L(l)[j] + b_l)                                    (one GEMM, shared by every star j is in)
    seed chain   S(l+1)[s] = act(dis_s * (sum_{j->s, j!=s} W_l L(l)[j] + dis_s * W_l S(l)[s]) + b_l)   (ops.gcn_star_aggregate)

with L(0) = S(0) = fusion_fn(x) and dis_s = indeg(s)^-1/2 (self-loop and duplicate edges counted as PyG does).  Result:
the same [N, out_dim] matrix the reference's loop produces row by row, from ``num_layers`` passes over the CSR.

The reference never calls ``.eval()`` on the loaded module, so its exported vectors carry dropout noise; this module
always evaluates the deterministic (eval-mode) forward.
"""
from __future__ import annotations

import pickle
from typing import Iterable, Sequence

import torch

from . import ops
from .model.encoder import GATEncoder, GCNEncoder


@torch.no_grad()
def star_embeddings(module, x: torch.Tensor, edge_index) -> torch.Tensor:
    """Row s = ``module(x_star_s, edge_index_star_s)[0]`` for the one-seed 1-hop star graph of every node s
    (node.py:224-234); fp32 [N, out_dim] on the device."""
    enc = module.model.encoder
    if isinstance(enc, GATEncoder) or not isinstance(enc, GCNEncoder):
        raise NotImplementedError("star export is defined for the reference's GCNEncoder (biomedkg/model/encoder.py:124-162)")
    was_training = module.training
    module.eval()
    try:
        h = module.fusion_fn(x=x)
        N = h.size(0)
        view = ops.as_view(edge_index, N)
        leaf = h if h.dtype == ops.BF16 else ops.mask_cast(h.float().contiguous())[0]
        seed = leaf
        layers = list(enc.graph_layers)
        for li, layer in enumerate(layers):
            last = li == len(layers) - 1
            w = layer.lin.weight
            if seed is leaf:
                t_leaf = t_seed = ops.linear(leaf, w, None, out_bf16=True)
            else:  # one GEMM for both chains
                t = ops.linear(torch.cat([leaf, seed], dim=0), w, None, out_bf16=True)
                t_leaf, t_seed = t[:N], t[N:]
            seed = ops.gcn_star_aggregate(view.rowptr, view.colind, view.dis, t_leaf.contiguous(), t_seed.contiguous(),
                                          layer.bias.contiguous(), relu=not last, out_fp32=last)
            if not last:
                leaf = torch.relu(t_leaf.float() + layer.bias).to(ops.BF16)
        return seed
    finally:
        module.train(was_training)


def export_embeddings(module, x: torch.Tensor, edge_index, node_names: Sequence, path: str | None = None,
                      mapping: dict | None = None) -> dict:
    """The ``node_mapping`` dict of node.py:193-241: ``{name: float32 ndarray [1, out_dim]}`` for one node type;
    pass the returned dict back in as ``mapping`` to accumulate gene / drug / disease as the reference does, and a
    ``path`` to pickle it with the reference's protocol (node.py:240-241)."""
    if len(node_names) != x.size(0):
        raise ValueError("one name per node expected")
    out = star_embeddings(module, x, edge_index).cpu().numpy()
    mapping = {} if mapping is None else mapping
    for i, name in enumerate(node_names):
        mapping[name] = out[i : i + 1]
    if path is not None:
        with open(path, "wb") as fh:
            pickle.dump(mapping, fh, protocol=pickle.HIGHEST_PROTOCOL)
    return mapping


class GCLEncode:
    """Lookup side of node.py:155-186: name list -> stacked [len, 1, embed_dim] tensor, Xavier-normal rows for names the
    export did not cover, ``random_init_ratio`` recorded."""

    def __init__(self, node_mapping: dict, embed_dim: int):
        self.node_mapping, self.embed_dim = node_mapping, embed_dim
        self.random_init_ratio = 0

    @classmethod
    def load(cls, artifact_path: str, embed_dim: int) -> "GCLEncode":
        with open(artifact_path, "rb") as fh:
            return cls(pickle.load(fh), embed_dim)

    def __call__(self, lst_node: Iterable[str]) -> torch.Tensor:
        rows, random_init = [], 0
        lst_node = list(lst_node)
        for name in lst_node:
            emb = self.node_mapping.get(name, None)
            if emb is None:
                emb = torch.nn.init.xavier_normal_(torch.empty(1, self.embed_dim))
                random_init += 1
            rows.append(torch.as_tensor(emb))
        self.random_init_ratio = random_init / len(lst_node)
        return torch.stack(rows, dim=0)
