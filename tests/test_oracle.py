"""CPU tests of the oracle itself (no GPU): it must reproduce the golden vectors generated from the
reference's own modules, its closed forms must equal the as-written PyG/PyGCL forms, and the
canonical CSR must have the properties SURVEY.md App. A.8 states."""
import glob
import math
import os

import numpy as np
import pytest
import torch

from oracle import models as om
from oracle import pyg, pygcl

MODULE_FIXTURES = ["grace_none", "grace_mean2", "grace_attention", "dgi_none", "ggd_none_a", "ggd_none_b", "ggd_redaf", "grace_redaf_train"]


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)


@pytest.mark.parametrize("name", MODULE_FIXTURES)
def test_oracle_reproduces_reference_golden(golden_dir, name):
    fx = _load(golden_dir, name)
    cfg = fx["cfg"]
    m = getattr(om, cfg["cls"])(cfg["in_dim"], cfg["hidden_dim"], cfg["out_dim"], cfg["num_hidden_layers"], fuse_method=cfg["fuse_method"])
    m = m.to(fx["x"].dtype)
    m.load_state_dict(fx["state_dict"])
    m.train(name != "ggd_redaf")
    om.set_draws(m, om.ReplayDraws(fx["draws"]))
    loss = m.training_step(fx["x"], fx["edge_index"])
    loss.backward()
    tol = 1e-5 if fx["x"].dtype == torch.float32 else 1e-10
    assert abs(float(loss) - float(fx["loss"])) <= tol * max(1.0, abs(float(fx["loss"])))
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert set(grads) == set(fx["grads"])
    for k, g in grads.items():
        assert torch.allclose(g, fx["grads"][k], rtol=tol * 10, atol=tol)
    m.eval()
    with torch.no_grad():
        assert torch.allclose(m(fx["x"], fx["edge_index"]), fx["embed_eval"], rtol=tol * 10, atol=tol)


def test_ggd_fixtures_cover_both_coin_branches(golden_dir):
    coins = [[d[1] for d in _load(golden_dir, n)["draws"] if d[0] == "coin"][0] for n in ("ggd_none_a", "ggd_none_b")]
    assert min(coins) < 0.5 <= max(coins)


def test_fusion_goldens(golden_dir):
    fx = _load(golden_dir, "fusion_attention_m3")
    att = om.AttentionFusion(fx["x"].size(-1)).double()
    att.load_state_dict(fx["state_dict"])
    assert torch.allclose(att(fx["x"]), fx["out"], atol=1e-12)
    fx = _load(golden_dir, "fusion_redaf_m2")
    red = om.ReDAF(fx["x"].size(-1)).eval()
    red.load_state_dict(fx["state_dict"])
    assert torch.allclose(red(fx["x"]), fx["out"], atol=1e-6)


def test_redaf_training_golden(golden_dir):
    """The reference's ReDAF in training mode (dropout mask recorded by the generator): forward and every gradient."""
    fx = _load(golden_dir, "fusion_redaf_m2_train")
    red = om.ReDAF(fx["x"].size(-1))
    red.load_state_dict(fx["state_dict"])
    red.train()
    red.draws = om.ReplayDraws([("dropout_mask", fx["mask"])])
    x = fx["x"].clone().requires_grad_(True)
    out = red(x)
    (out * fx["w"]).sum().backward()
    assert torch.allclose(out, fx["out"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(x.grad, fx["x_grad"], rtol=1e-4, atol=1e-6)
    for k, p_ in red.named_parameters():
        if k in fx["grads"]:
            assert torch.allclose(p_.grad, fx["grads"][k], rtol=1e-4, atol=1e-5), k
        else:
            assert p_.grad is None, k


def test_gcn_encoder_golden(golden_dir):
    fx = _load(golden_dir, "gcn_encoder_eval")
    enc = om.GCNEncoder(fx["x"].size(1), 64, 64, 2).double().eval()
    enc.load_state_dict(fx["state_dict"])
    with torch.no_grad():
        assert torch.allclose(enc(fx["x"], fx["edge_index"]), fx["out"], atol=1e-12)


# ---- hand-computable graphs: gcn_conv (gather/scatter form) vs dense A_hat --------------------------
GRAPHS = {
    "path": (4, [[0, 1, 1, 2, 2, 3], [1, 0, 2, 1, 3, 2]]),
    "star": (5, [[1, 2, 3, 4], [0, 0, 0, 0]]),
    "duplicates": (3, [[0, 0, 0, 1], [1, 1, 1, 2]]),
    "self_loops": (3, [[0, 1, 1, 2], [0, 1, 2, 2]]),
    "isolated": (4, [[0], [1]]),
    "asymmetric": (4, [[0, 1, 2, 3, 3], [1, 2, 3, 0, 1]]),
    "empty": (3, [[], []]),
}


@pytest.mark.parametrize("name", sorted(GRAPHS))
def test_gcn_conv_matches_dense_formula(name):
    n, e = GRAPHS[name]
    ei = torch.tensor(e, dtype=torch.int64).reshape(2, -1)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, 5, generator=g, dtype=torch.float64)
    w = torch.randn(3, 5, generator=g, dtype=torch.float64)
    b = torch.randn(3, generator=g, dtype=torch.float64)
    ref = pyg.gcn_dense_adj(ei, n) @ (x @ w.t()) + b
    assert torch.allclose(pyg.gcn_conv(x, ei, w, b), ref, atol=1e-12)


def test_star_values_by_hand():
    # node 0 has in-degree 4+1, leaves have in-degree 1 (self-loop only); leaf->hub weight = 1/sqrt(5*1)
    ei = torch.tensor(GRAPHS["star"][1])
    _, w = pyg.gcn_norm(ei, 5, torch.float64)
    assert torch.allclose(w[:4], torch.full((4,), 1 / math.sqrt(5.0), dtype=torch.float64))
    assert torch.allclose(w[4:], torch.tensor([1 / 5.0, 1, 1, 1, 1], dtype=torch.float64))


# ---- closed forms vs as-written ---------------------------------------------------------------------
@pytest.mark.parametrize("n,d", [(7, 8), (64, 32), (130, 16)])
def test_infonce_closed_form(n, d):
    g = torch.Generator().manual_seed(n)
    h1 = torch.randn(n, d, generator=g, dtype=torch.float64) * 3
    h2 = torch.randn(n, d, generator=g, dtype=torch.float64)
    a = pygcl.infonce_l2l_as_written(h1, h2, 0.2, True)
    b = pygcl.infonce_l2l_closed_form(h1, h2, 0.2)
    assert abs(float(a - b)) < 1e-11
    # invariant to positive row scaling of h (normalisation)
    c = pygcl.infonce_l2l_closed_form(h1 * torch.rand(n, 1, generator=g, dtype=torch.float64).add(0.1), h2, 0.2)
    assert abs(float(a - c)) < 1e-11


def test_jsd_and_ggd_closed_forms():
    g = torch.Generator().manual_seed(3)
    h, hn = torch.randn(50, 16, generator=g, dtype=torch.float64), torch.randn(50, 16, generator=g, dtype=torch.float64)
    s = torch.randn(1, 16, generator=g, dtype=torch.float64)
    assert abs(float(pygcl.jsd_g2l_as_written(h, s, hn) - pygcl.jsd_g2l_closed_form(h, s, hn))) < 1e-10
    w, b = torch.randn(16, 16, generator=g, dtype=torch.float64), torch.randn(16, generator=g, dtype=torch.float64)
    assert abs(float(pygcl.ggd_loss_as_written(h, hn, w, b) - pygcl.ggd_loss_closed_form(h, hn, w, b))) < 1e-10


# ---- canonical CSR properties -------------------------------------------------------------------------
def _rand_graph(n, e, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, n, (2, e), generator=g, dtype=torch.int64)


@pytest.mark.parametrize("seed", range(5))
def test_csr_invariant_to_edge_order_and_consistent(seed):
    n, e = 23, 200
    ei = pyg.view_graph(_rand_graph(n, e, seed), n)
    rp, ci, perm = pyg.canonical_csr(ei, n)
    g = torch.Generator().manual_seed(100 + seed)
    shuffle = torch.randperm(ei.size(1), generator=g)
    rp2, ci2, _ = pyg.canonical_csr(ei[:, shuffle], n)
    assert torch.equal(rp, rp2) and torch.equal(ci, ci2)           # (rowptr, colind) unique per edge multiset
    assert int(rp[-1]) == ei.size(1) and torch.equal(ei[0][perm.long()].int(), ci)
    dst_sorted = ei[1][perm.long()]
    assert bool((dst_sorted[1:] >= dst_sorted[:-1]).all())
    rn, cn, pn = pyg.canonical_csr_numpy(ei.numpy(), n)
    assert np.array_equal(rn, rp.numpy()) and np.array_equal(cn, ci.numpy()) and np.array_equal(pn, perm.numpy())
    # CSC is the CSR of the transposed graph
    rps, cis, _ = pyg.canonical_csr(ei, n, by="src")
    rpt, cit, _ = pyg.canonical_csr(ei.flip(0), n, by="dst")
    assert torch.equal(rps, rpt) and torch.equal(cis, cit)


def test_gat_conv_matches_dense_softmax():
    n = 6
    ei = torch.tensor([[0, 1, 2, 3, 4, 4, 2], [1, 2, 0, 0, 0, 5, 2]])
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, 4, generator=g, dtype=torch.float64)
    w = torch.randn(3, 4, generator=g, dtype=torch.float64)
    a_s, a_d = torch.randn(1, 1, 3, generator=g, dtype=torch.float64), torch.randn(1, 1, 3, generator=g, dtype=torch.float64)
    b = torch.randn(3, generator=g, dtype=torch.float64)
    out = pyg.gat_conv(x, ei, w, a_s, a_d, b)
    xh = x @ w.t()
    ei2 = pyg.add_remaining_self_loops(ei, n)
    ref = torch.zeros(n, 3, dtype=torch.float64)
    for i in range(n):
        src = ei2[0][ei2[1] == i]
        e = torch.nn.functional.leaky_relu((xh[src] * a_s.view(-1)).sum(-1) + (xh[i] * a_d.view(-1)).sum(), 0.2)
        ref[i] = torch.softmax(e, 0) @ xh[src]
    assert torch.allclose(out, ref + b, atol=1e-12)


# ---- embedding-export path (node.py:193-241) ----------------------------------------------------
def _export_case(seed=0, N=40, M=2, IN=32, fuse="attention"):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, M, IN, generator=g, dtype=torch.float64) if M else torch.randn(N, IN, generator=g, dtype=torch.float64)
    ei = torch.randint(0, N - 3, (2, 200), generator=g)            # the last 3 nodes stay isolated
    ei[:, :5] = torch.tensor([[1, 1, 2, 3, 4], [1, 1, 2, 0, 0]])    # existing (and duplicated) self-loops
    ei = torch.cat([ei, ei[:, 10:20]], dim=1)                       # duplicate edges
    torch.manual_seed(seed)
    m = om.GRACEModule(in_dim=IN, hidden_dim=64, out_dim=64, num_hidden_layers=2, fuse_method=fuse).double().eval()
    for layer in m.model.encoder.graph_layers:
        layer.bias.data.normal_(generator=g)
    return m, x, ei


def test_one_hop_batch_is_a_star():
    from oracle import export as oe

    ei = torch.tensor([[5, 2, 5, 3, 3, 0], [3, 3, 3, 3, 1, 5]])
    nodes, sub = oe.one_hop_batch(ei, 3)
    assert nodes.tolist() == [3, 5, 2]                                # seed first, neighbours in order of first appearance
    assert sub.tolist() == [[1, 2, 1, 0], [0, 0, 0, 0]]               # duplicates kept, the 3->3 loop kept, nothing else
    nodes, sub = oe.one_hop_batch(ei, 4)
    assert nodes.tolist() == [4] and sub.shape == (2, 0)


@pytest.mark.parametrize("fuse,M", [("attention", 2), (None, 2), ("none", 0)])
def test_export_loop_equals_two_chain_closed_form(fuse, M):
    from oracle import export as oe

    m, x, ei = _export_case(seed=3, M=M, fuse=fuse)
    loop, closed = oe.export_loop(m, x, ei), oe.export_two_chains(m, x, ei)
    assert loop.shape == (40, 64)
    assert float((loop - closed).abs().max()) < 1e-12 * max(1.0, float(loop.abs().max()))
    # an isolated node's star is the node alone: its embedding is the plain per-node MLP chain
    h = m.fusion_fn(x=x)[-1:]
    for i, layer in enumerate(m.model.encoder.graph_layers):
        h = h @ layer.lin.weight.t() + layer.bias
        h = torch.relu(h) if i < len(m.model.encoder.graph_layers) - 1 else h
    assert torch.allclose(loop[-1:], h, atol=1e-12)


# ---- hypothesis-driven properties -----------------------------------------------------------------
from hypothesis import given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 40), st.lists(st.tuples(st.integers(0, 39), st.integers(0, 39)), min_size=0, max_size=120), st.randoms())
def test_csr_properties_any_graph(n, edges, rnd):
    """Any edge multiset (empty, ragged, duplicates, self-loops): the canonical CSR of the gcn_norm'ed view is invariant to
    the order the edges arrive in, has exactly one self-loop per row, and its degree vector is what gcn_norm counts."""
    edges = [(s % n, d % n) for s, d in edges]
    ei = torch.tensor(edges, dtype=torch.int64).reshape(-1, 2).t().contiguous()
    view = pyg.view_graph(ei, n)
    rp, ci, _ = pyg.canonical_csr(view, n)
    shuffled = list(edges)
    rnd.shuffle(shuffled)
    ei2 = torch.tensor(shuffled, dtype=torch.int64).reshape(-1, 2).t().contiguous()
    rp2, ci2, _ = pyg.canonical_csr(pyg.view_graph(ei2, n), n)
    assert torch.equal(rp, rp2) and torch.equal(ci, ci2)
    non_self = sum(1 for s, d in edges if s != d)
    assert int(rp[-1]) == non_self + n
    for r in range(n):
        row = ci[int(rp[r]) : int(rp[r + 1])].tolist()
        assert row == sorted(row) and row.count(r) == 1
        assert len(row) == 1 + sum(1 for s, d in edges if d == r and s != r)


@settings(max_examples=25, deadline=None)
@given(st.integers(2, 24), st.integers(1, 12), st.integers(0, 2**31 - 1))
def test_infonce_properties_any_shape(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    h1 = torch.randn(n, d, generator=g, dtype=torch.float64)
    h2 = torch.randn(n, d, generator=g, dtype=torch.float64)
    a = pygcl.infonce_l2l_as_written(h1, h2, 0.2, True)
    assert abs(float(a - pygcl.infonce_l2l_closed_form(h1, h2, 0.2))) < 1e-10
    scale = torch.rand(n, 1, generator=g, dtype=torch.float64) + 0.05
    assert abs(float(a - pygcl.infonce_l2l_closed_form(h1 * scale, h2 / scale, 0.2))) < 1e-10     # row-scale invariant
    assert abs(float(a - pygcl.infonce_l2l_closed_form(h2, h1, 0.2))) < 1e-10                     # symmetric in the views
    p = torch.randperm(n, generator=g)
    assert abs(float(a - pygcl.infonce_l2l_closed_form(h1[p], h2[p], 0.2))) < 1e-10               # node order does not matter


# ---- neighbour sampler (data_module.py:71-99) ---------------------------------------------------------
def test_sampler_oracle_structure_and_uniformity():
    from collections import Counter

    from oracle import sampler as osamp

    rng = np.random.default_rng(1)
    n, e = 60, 900
    ei = rng.integers(0, n - 4, (2, e))                 # nodes n-4.. are isolated
    ei[:, :3] = [[5, 5, 9], [5, 5, 9]]                   # self-loops are ordinary edges for the sampler
    seeds = [3, 57, 11, 20]
    n_id, sub, eid = osamp.sample(ei, n, seeds, [5, 3, -1], 99)
    assert n_id[: len(seeds)].tolist() == seeds and len(set(n_id.tolist())) == len(n_id)
    assert (ei[0][eid] == n_id[sub[0]]).all() and (ei[1][eid] == n_id[sub[1]]).all()      # every edge is a real in-edge
    assert len(set(eid.tolist())) == len(eid)                                              # without replacement
    indeg = np.bincount(ei[1], minlength=n)
    per_target = Counter(sub[1].tolist())
    fan = {0: 5, 1: 3, 2: -1}
    # hop of a local node = the hop in which it was appended; every node of hop < 3 samples min(deg, fanout) in-edges
    first_seen, bounds, hop_of = {}, [len(seeds)], {}
    for pos, (r, c) in enumerate(zip(sub[0].tolist(), sub[1].tolist())):
        first_seen.setdefault(r, pos)
    for i in range(len(n_id)):
        hop_of[i] = 0 if i < len(seeds) else None
    order = sorted((p, r) for r, p in first_seen.items() if r >= len(seeds))
    assert [r for _, r in order] == list(range(len(seeds), len(n_id)))                    # appended in order of first appearance
    for i, v in enumerate(n_id.tolist()):
        if i in per_target:
            assert per_target[i] <= indeg[v]
    for i, v in enumerate(seeds):
        assert per_target.get(i, 0) == min(indeg[v], 5)
    assert per_target.get(1, 0) == 0 and indeg[57] == 0                                    # isolated seed stays alone
    # each in-edge position equally likely, in both the "take" (deg > 2k) and the "leave out" (deg <= 2k) regimes
    for deg, k in ((100, 30), (40, 30), (31, 30), (64, 32)):
        cnt = Counter()
        trials = 3000
        for node in range(trials):
            pos = osamp.pick_positions(5, 1, node, deg, k)
            assert len(pos) == k and len(set(pos)) == k and pos == sorted(pos) and 0 <= pos[0] and pos[-1] < deg
            cnt.update(pos)
        exp = trials * k / deg
        chi2 = sum((cnt[p] - exp) ** 2 / exp for p in range(deg)) / (1 - k / deg)
        assert chi2 < deg + 6 * math.sqrt(2 * deg), (deg, k, chi2)
    assert osamp.pick_positions(5, 0, 7, 12, 30) == list(range(12)) and osamp.pick_positions(5, 0, 7, 12, -1) == list(range(12))
