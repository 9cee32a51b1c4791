"""S1 - device neighbour sampler / loader (biomedkg/data_module.py:65-99): bit-exact against the
restatement in oracle/sampler.py (integer work), plus the loader contract the reference's training loop relies on."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _graph(n, e, seed, hub=None):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, max(1, n - 3), (2, e), generator=g)              # last nodes isolated
    if e >= 8:
        ei[:, :4] = torch.tensor([[1, 1, 2, 0], [1, 1, 2, 2]])             # self-loops and duplicates are ordinary edges
    if hub is not None:
        ei[1, e // 2 :] = hub                                               # one very long in-edge list
    return ei


@pytest.mark.parametrize("n,e,fan,nseeds,hub", [(50, 600, [5, 5], 3, None), (300, 20000, [30, 30, 30], 16, None),
                                                 (300, 14000, [30, 30], 8, None),       # degrees ~47: the leave-out regime
                                                 (200, 3000, [-1], 1, None), (200, 3000, [-1, 2], 5, None),
                                                 (500, 6000, [32, 1, 7], 9, 7), (40, 0, [30, 30], 4, None),
                                                 (2000, 100000, [30, 30, 30], 128, 11)])
def test_sampler_bit_exact_vs_oracle(n, e, fan, nseeds, hub):
    from biomedkg_b200.loader import NeighborSampler
    from oracle import sampler as osamp

    ei = _graph(n, e, n + e, hub)
    g = torch.Generator().manual_seed(3)
    seeds = torch.randperm(n, generator=g)[:nseeds]
    if hub is not None:
        seeds[0] = hub if hub not in seeds.tolist() else seeds[0]
    smp = NeighborSampler(ei.to(DEV), n, fan)
    for seed in (0, 0xDEADBEEF12345):
        n_id, sub, eid = smp.sample(seeds.to(DEV), seed)
        rn, rs, re = osamp.sample(ei.numpy(), n, seeds.tolist(), fan, seed)
        assert n_id.dtype == torch.int64 and sub.dtype == torch.int64 and sub.shape[0] == 2
        assert np.array_equal(n_id.cpu().numpy(), rn)
        assert np.array_equal(sub.cpu().numpy(), rs)
        assert np.array_equal(eid.cpu().numpy(), re)
        assert int((smp.local_id != -1).sum()) == 0 and int((smp.first_pos != 0x7FFFFFFF).sum()) == 0   # scratch maps restored
    # a node draws the same in-edges whatever batch it is in (counter-based stream keyed by seed, hop, node)
    a = smp.sample(seeds[:1].to(DEV), 5)
    b = smp.sample(seeds.to(DEV), 5)
    first_hop_a = a[2][a[1][1] == 0]
    first_hop_b = b[2][b[1][1] == 0]
    assert torch.equal(first_hop_a[: first_hop_b.numel()], first_hop_b[: first_hop_a.numel()])


@pytest.mark.parametrize("n,e,fan,nseeds", [(400, 30000, [30, 30, 30], 64), (5000, 40000, [30, 30, 30], 128), (300, 9000, [-1], 1), (1000, 50000, [7, 3], 20)])
def test_sampler_structure_brute_force(n, e, fan, nseeds):
    """PyG NeighborLoader's structural rules (fan-out cap, no replacement, sampled edges are original edges, seeds first,
    first-appearance node order, hop/target grouping) checked by brute force on the GPU sampler's output - the checker
    (oracle.sampler.check_structure) shares no random stream or code with the kernels."""
    from biomedkg_b200.loader import NeighborSampler
    from oracle import sampler as osamp

    ei = _graph(n, e, 7 * n + e)
    seeds = torch.randperm(n, generator=torch.Generator().manual_seed(n))[:nseeds]
    smp = NeighborSampler(ei.to(DEV), n, fan)
    for seed in (1, 99, 2 ** 40 + 17):
        n_id, sub, eid = smp.sample(seeds.to(DEV), seed)
        assert osamp.check_structure(ei.numpy(), n, seeds.tolist(), fan, n_id.cpu().numpy(), sub.cpu().numpy(), eid.cpu().numpy()) == []


def test_loader_contract_and_training_step():
    import biomedkg_b200 as b
    from biomedkg_b200.loader import NeighborLoader, random_link_split
    from types import SimpleNamespace

    n, e, IN = 3000, 40000, 32
    g = torch.Generator().manual_seed(0)
    data = SimpleNamespace(x=torch.randn(n, IN, generator=g).to(DEV), edge_index=_graph(n, e, 1).to(DEV))
    train, val, test = random_link_split(data, num_val=0.2, num_test=0.2)      # data_module.py:65-69
    assert train.edge_index.size(1) == e - 2 * int(0.2 * e) and torch.equal(val.edge_index, train.edge_index)
    assert test.edge_index.size(1) == e - int(0.2 * e)
    loader = NeighborLoader(train, num_neighbors=[30] * 3, batch_size=128, shuffle=True, seed=7)   # data_module.py:92-98
    assert len(loader) == (n + 127) // 128
    torch.manual_seed(0)
    mod = b.GRACEModule(in_dim=IN, hidden_dim=64, out_dim=64, num_hidden_layers=2, fuse_method="none").to(DEV)
    opt = torch.optim.Adam(mod.model.parameters(), lr=1e-3)
    seen, losses = [], []
    for i, batch in enumerate(loader):
        assert batch.x.shape == (batch.n_id.numel(), IN) and torch.equal(batch.x, data.x[batch.n_id])
        assert torch.equal(batch.n_id[: batch.batch_size], batch.input_id)
        src, dst = train.edge_index[:, batch.e_id]
        assert torch.equal(batch.n_id[batch.edge_index[0]], src) and torch.equal(batch.n_id[batch.edge_index[1]], dst)
        seen.append(batch.input_id)
        if i < 6:
            loss = mod.training_step(batch)
            opt.zero_grad()
            loss.backward()
            opt.step()
            losses.append(float(loss))
    assert sorted(torch.cat(seen).tolist()) == list(range(n))                                   # every node is a seed exactly once
    assert all(np.isfinite(losses))
    # a second epoch reshuffles and redraws
    first = next(iter(loader))
    assert not torch.equal(first.input_id, seen[0])


def test_one_hop_loader_reproduces_the_export_path():
    """data_module.py:71-79 + node.py:224-234: NeighborLoader(num_neighbors=[-1]) batches of one seed, model(batch.x,
    batch.edge_index)[:1] - must equal the one-pass star export row by row."""
    import biomedkg_b200 as b
    from biomedkg_b200.export import star_embeddings
    from biomedkg_b200.loader import NeighborLoader
    from types import SimpleNamespace

    n, IN = 120, 32
    g = torch.Generator().manual_seed(2)
    data = SimpleNamespace(x=torch.randn(n, 2, IN, generator=g).to(DEV), edge_index=_graph(n, 900, 4).to(DEV))
    torch.manual_seed(1)
    mod = b.GRACEModule(in_dim=IN, hidden_dim=64, out_dim=64, num_hidden_layers=2, fuse_method="attention").to(DEV).eval()
    ref = star_embeddings(mod, data.x, data.edge_index)
    rows = []
    with torch.no_grad():
        for batch in NeighborLoader(data, num_neighbors=[-1], shuffle=False):
            assert batch.batch_size == 1
            rows.append(mod(batch.x, batch.edge_index)[: batch.batch_size])
    got = torch.cat(rows)
    assert got.shape == ref.shape and rel_err(got, ref) < 5e-3


def test_train_driver_minibatch_regime_runs_fit_validate_test():
    """biomedkg_b200.train_gcl --regime minibatch = the reference's train_gcl.py:108-122 loop (RandomLinkSplit -> NeighborLoader ->
    fit with validation every epoch -> test) on the GPU sampler: JSONL records carry train / val / test losses."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "biomedkg_b200.train_gcl", "--regime", "minibatch", "--model", "grace", "--nodes", "3000",
                        "--edges", "40000", "--in-dim", "64", "--hidden-dim", "64", "--epochs", "2", "--limit-batches", "4",
                        "--batch-size", "64"], cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    recs = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    epochs = [x for x in recs if "epoch" in x]
    assert len(epochs) == 2 and all(x["batches"] == 4 and x["train_loss"] == x["train_loss"] and x["val_loss"] == x["val_loss"] for x in epochs)
    assert "test_loss" in recs[-1] and recs[-1]["test_loss"] == recs[-1]["test_loss"]
