"""Embedding-export path: biomedkg_b200.export evaluates every one-seed 1-hop star graph of
biomedkg/data/node.py:193-241 in one pass per layer; the oracle (oracle/export.py) runs the reference's per-seed loop.
Embeddings <= 1e-2 relative (bf16 operands, fp32 accumulation); the aggregation kernel itself <= 1e-5 on identical inputs."""
import pickle

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _graph(N, E, seed):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, N - 5, (2, E), generator=g)                 # last 5 nodes isolated
    ei[:, :6] = torch.tensor([[1, 1, 2, 3, 4, 7], [1, 1, 2, 0, 0, 7]])  # self-loops, one duplicated
    return torch.cat([ei, ei[:, 20:60]], dim=1)                        # duplicate edges


@pytest.mark.parametrize("cls,fuse,M", [("GRACEModule", "attention", 2), ("GRACEModule", "none", 0), ("DGIModule", None, 3),
                                        ("GGDModule", "redaf", 2)])
def test_star_export_matches_per_seed_loop(cls, fuse, M):
    import biomedkg_b200 as b
    from biomedkg_b200.export import star_embeddings
    from oracle import export as oe
    from oracle import models as om

    N, IN, HID = 300, 32, 64
    torch.manual_seed(5)
    orc = getattr(om, cls)(in_dim=IN, hidden_dim=HID, out_dim=HID, num_hidden_layers=2, fuse_method=fuse)
    for layer in orc.model.encoder.graph_layers:
        layer.bias.data.normal_(std=0.3)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(N, M, IN, generator=g) if M else torch.randn(N, IN, generator=g)
    ei = _graph(N, 2500, 7)
    mod = getattr(b, cls)(in_dim=IN, hidden_dim=HID, out_dim=HID, num_hidden_layers=2, fuse_method=fuse)
    mod.load_state_dict(orc.state_dict())
    mod = mod.to(DEV).train()                                          # export must not depend on the training flag
    got = star_embeddings(mod, x.to(DEV), ei.to(DEV))
    assert mod.training and got.dtype == torch.float32 and got.shape == (N, HID)
    ref = oe.export_loop(orc.double().eval(), x.double(), ei)
    assert rel_err(got, ref) < 1e-2
    # it is NOT the full-graph forward (neighbours of a star see nobody): guard against silently exporting that
    full = mod.eval()(x.to(DEV), ei.to(DEV))
    assert rel_err(full, ref) > 5e-2


def test_star_aggregate_kernel_exact_inputs():
    from biomedkg_b200 import ops

    N, C = 1000, 256
    g = torch.Generator().manual_seed(1)
    ei = _graph(N, 30000, 2).to(DEV)
    view = ops.as_view(ei, N)
    leaf = torch.randn(N, C, generator=g).to(DEV).to(torch.bfloat16)
    seed = torch.randn(N, C, generator=g).to(DEV).to(torch.bfloat16)
    bias = torch.randn(C, generator=g).to(DEV)
    for relu in (False, True):
        got = ops.gcn_star_aggregate(view.rowptr, view.colind, view.dis, leaf, seed, bias, relu=relu, out_fp32=True)
        src, dst = ei[0], ei[1]
        keep = src != dst
        src, dst = src[keep], dst[keep]
        deg = torch.ones(N, dtype=torch.float64, device=DEV).index_add_(0, dst, torch.ones(dst.numel(), dtype=torch.float64, device=DEV))
        dis = deg.pow(-0.5)[:, None]
        agg = torch.zeros(N, C, dtype=torch.float64, device=DEV).index_add_(0, dst, leaf.double()[src])
        ref = dis * (agg + dis * seed.double()) + bias.double()
        ref = torch.relu(ref) if relu else ref
        assert rel_err(got, ref) < 1e-5
        got16 = ops.gcn_star_aggregate(view.rowptr, view.colind, view.dis, leaf, seed, bias, relu=relu, out_fp32=False)
        assert torch.equal(got16, got.to(torch.bfloat16))
    # deterministic
    a = ops.gcn_star_aggregate(view.rowptr, view.colind, view.dis, leaf, seed, bias, relu=True, out_fp32=True)
    assert torch.equal(a, got)


def test_export_mapping_and_lookup(tmp_path):
    import biomedkg_b200 as b
    from biomedkg_b200.export import GCLEncode, export_embeddings, star_embeddings

    N, IN, HID = 64, 32, 64
    torch.manual_seed(0)
    mod = b.GRACEModule(in_dim=IN, hidden_dim=HID, out_dim=HID, num_hidden_layers=2, fuse_method="none").to(DEV)
    x = torch.randn(N, IN, device=DEV)
    ei = _graph(N, 400, 3).to(DEV)
    names = [f"node{i}" for i in range(N)]
    path = str(tmp_path / "grace_none.pickle")
    mapping = export_embeddings(mod, x, ei, names, path=path)
    ref = star_embeddings(mod, x, ei).cpu()
    with open(path, "rb") as fh:
        loaded = pickle.load(fh)
    assert list(loaded) == names
    for i in (0, 17, N - 1):
        assert loaded[names[i]].shape == (1, HID) and loaded[names[i]].dtype.name == "float32"      # node.py:232-236
        assert torch.equal(torch.from_numpy(loaded[names[i]]), ref[i : i + 1])
    enc = GCLEncode.load(path, HID)
    out = enc(["node3", "missing", "node5", "also missing"])
    assert out.shape == (4, 1, HID) and enc.random_init_ratio == 0.5                                 # node.py:173-186
    assert torch.equal(out[0], ref[3:4]) and torch.equal(out[2], ref[5:6])
    assert mapping is not loaded and set(mapping) == set(loaded)


def test_gat_export_is_refused():
    import biomedkg_b200 as b
    from biomedkg_b200.export import star_embeddings

    mod = b.GRACEModule(in_dim=32, hidden_dim=64, out_dim=64, num_hidden_layers=2, fuse_method="none", encoder="gat").to(DEV)
    with pytest.raises(NotImplementedError):
        star_embeddings(mod, torch.randn(16, 32, device=DEV), torch.randint(0, 16, (2, 40), device=DEV))
