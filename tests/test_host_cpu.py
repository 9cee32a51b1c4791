"""Host-side logic that needs no GPU: the link split, the export lookup, the host mirrors of the device random streams."""
import pickle
from types import SimpleNamespace

import numpy as np
import torch


def test_random_link_split_message_passing_edges():
    """T.RandomLinkSplit(num_val, num_test, neg_sampling_ratio=0.0) on a directed graph (biomedkg/data_module.py:65-69):
    train and validation see the training edges, test sees training + validation edges, nothing sees the test edges."""
    from biomedkg_b200.loader import random_link_split

    g = torch.Generator().manual_seed(0)
    E = 1000
    data = SimpleNamespace(x=torch.zeros(50, 4), edge_index=torch.stack([torch.arange(E), torch.arange(E) + 7]))   # every edge unique
    tr, va, te = random_link_split(data, num_val=0.2, num_test=0.3, generator=g)
    assert tr.edge_index.shape == (2, 500) and torch.equal(va.edge_index, tr.edge_index) and te.edge_index.shape == (2, 700)
    ids = lambda d: set(d.edge_index[0].tolist())  # noqa: E731
    assert ids(tr) < ids(te) and len(ids(te)) == 700 and len(set(range(E)) - ids(te)) == 300
    assert torch.equal(te.edge_index[:, :500], tr.edge_index)              # cat([train, val]) order
    assert (tr.edge_index[1] - tr.edge_index[0] == 7).all()                # columns stay paired
    tr2, _, _ = random_link_split(data, num_val=0.2, num_test=0.3, generator=torch.Generator().manual_seed(0))
    assert torch.equal(tr2.edge_index, tr.edge_index)
    tr3, _, te3 = random_link_split(data, num_val=100, num_test=50)        # absolute counts, as PyG accepts
    assert tr3.edge_index.size(1) == 850 and te3.edge_index.size(1) == 950
    try:
        random_link_split(data, num_val=0.6, num_test=0.5)
        raise AssertionError("expected ValueError")
    except ValueError:
        pass


def test_gcl_encode_lookup(tmp_path):
    """biomedkg/data/node.py:173-186: stacked [len, 1, D] rows, Xavier rows for unknown names, random_init_ratio."""
    from biomedkg_b200.export import GCLEncode

    mapping = {f"n{i}": np.full((1, 8), float(i), dtype=np.float32) for i in range(5)}
    path = tmp_path / "grace_none.pickle"
    with open(path, "wb") as fh:
        pickle.dump(mapping, fh, protocol=pickle.HIGHEST_PROTOCOL)
    enc = GCLEncode.load(str(path), 8)
    out = enc(["n3", "zzz", "n0", "n4"])
    assert out.shape == (4, 1, 8) and enc.random_init_ratio == 0.25
    assert float(out[0].mean()) == 3.0 and float(out[2].abs().sum()) == 0.0 and float(out[3].mean()) == 4.0
    assert float(out[1].abs().sum()) > 0.0
    assert enc(["n1"]).shape == (1, 1, 8) and enc.random_init_ratio == 0.0


def test_host_mirrors_of_device_streams():
    """draws.hash_keep_mask / hash_keep_mask16 restate csrc/common.cuh hash_u32 (and the ReDAF two-per-hash split);
    the sampler oracle restates the same hash on Python ints."""
    from biomedkg_b200.draws import hash_keep_mask, hash_keep_mask16
    from oracle.sampler import hash_u32

    seed, n, p = 0xABCDEF0123, 4096, 0.2
    thr = int(p * 4294967296.0)
    ref = torch.tensor([hash_u32(seed, i) >= thr for i in range(n)])
    assert torch.equal(hash_keep_mask(seed, n, p), ref)
    thr16 = int(p * 65536.0)
    ref16 = torch.tensor([((hash_u32(seed, i >> 1) >> (16 * (i & 1))) & 0xFFFF) >= thr16 for i in range(n + 1)])
    assert torch.equal(hash_keep_mask16(seed, n + 1, p), ref16)
    for m in (hash_keep_mask(seed, 200_000, p), hash_keep_mask16(seed, 200_000, p)):
        assert abs(float(m.float().mean()) - (1 - p)) < 5e-3
    assert not torch.equal(hash_keep_mask(seed + 1, n, p), ref)


def test_lightning_stand_in_captures_outermost_init_kwargs():
    import biomedkg_b200 as b

    m = b.GGDModule(in_dim=32, hidden_dim=64, out_dim=64, num_hidden_layers=2, fuse_method="redaf", learning_rate=5e-4)
    assert m.hparams["learning_rate"] == 5e-4 and m.hparams["fuse_method"] == "redaf" and m.hparams["in_dim"] == 32
    assert "model" not in m.hparams and "embed_dim" not in m.hparams        # BaseGCL's own arguments are not the checkpoint's


def test_graphed_step_host_side_draws_contract():
    """GGD's augmentation coin is a host branch: GraphedStep wants one capture per branch (graphed_step does that); every
    module then fails loudly on CPU tensors - there is no CPU path to capture."""
    import pytest

    import biomedkg_b200 as b
    from biomedkg_b200.graphed import GraphedStep, graphed_step

    x, ei = torch.zeros(8, 32), torch.zeros(2, 4, dtype=torch.int64)
    with pytest.raises(ValueError):
        GraphedStep(b.GGDModule(in_dim=32, hidden_dim=64, out_dim=64, num_hidden_layers=2), x, ei)
    for cls in (b.DGIModule, b.GGDModule, b.GRACEModule):
        with pytest.raises(RuntimeError):
            graphed_step(cls(in_dim=32, hidden_dim=64, out_dim=64, num_hidden_layers=2), x, ei)


def test_sampler_structure_checker_accepts_oracle_and_rejects_corruptions():
    """oracle.sampler.check_structure is the stream-independent judge of a sampled batch (used on the GPU sampler's output in
    tests/test_gpu_sampler.py): it must accept the oracle's own batches and flag each kind of corruption."""
    import numpy as np

    from oracle import sampler as osamp

    rng = np.random.default_rng(0)
    n, e = 120, 3000
    ei = rng.integers(0, n - 5, size=(2, e))
    seeds = [3, 77, 10, 54]
    for fan in ([5, 5], [30, 30, 30], [-1], [-1, 2]):
        n_id, sub, eid = osamp.sample(ei, n, seeds, fan, 1234)
        assert osamp.check_structure(ei, n, seeds, fan, n_id, sub, eid) == []
        if sub.shape[1] > 4:
            dup = eid.copy(); dup[1] = dup[0]
            sub2 = sub.copy(); sub2[:, 1] = sub2[:, 0]
            assert osamp.check_structure(ei, n, seeds, fan, n_id, sub2, dup)                  # replacement
            assert osamp.check_structure(ei, n, seeds, fan, n_id, sub[:, :-1], eid[:-1])      # a missing edge breaks the fan-out count
            sw = n_id.copy(); sw[[len(seeds), len(n_id) - 1]] = sw[[len(n_id) - 1, len(seeds)]]
            assert osamp.check_structure(ei, n, seeds, fan, sw, sub, eid)                      # wrong node order


def test_hostmem_placement_hint_degrades_quietly():
    """biomedkg_b200.hostmem: without a GPU / NVML there is no placement hint - gpu_local_cpus answers None instead of raising
    (pinned_near then pins wherever the thread runs), and the calling thread's CPU affinity is left as it was."""
    import os

    from biomedkg_b200.hostmem import gpu_local_cpus

    before = os.sched_getaffinity(0)
    cpus = gpu_local_cpus(0)
    assert cpus is None or (isinstance(cpus, set) and cpus <= before)
    assert os.sched_getaffinity(0) == before
