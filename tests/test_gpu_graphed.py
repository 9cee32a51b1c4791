"""CUDA-graph capture of the GRACE training step (biomedkg_b200/graphed.py): a replay must be the same computation as the
eager step under the same generator state, its random draws must advance from replay to replay, and training through
replays must work with an ordinary optimiser."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _setup(encoder, fuse, M, n=3000, e=40000, IN=64):
    import biomedkg_b200 as b

    g = torch.Generator().manual_seed(0)
    x = (torch.randn(n, M, IN, generator=g) if M else torch.randn(n, IN, generator=g)).to(DEV)
    ei = torch.randint(0, n, (2, e), generator=g).to(DEV)
    torch.manual_seed(1)
    mod = b.GRACEModule(in_dim=IN, hidden_dim=64, out_dim=64, num_hidden_layers=2, fuse_method=fuse, encoder=encoder).to(DEV).train()
    return mod, x, ei


@pytest.mark.parametrize("encoder,fuse,M,resort", [("gcn", "none", 0, False), ("gat", "attention", 2, False), ("gcn", "attention", 3, True)])
def test_replay_equals_eager_step_under_same_generator_state(encoder, fuse, M, resort):
    from types import SimpleNamespace

    from biomedkg_b200.graphed import GraphedStep

    mod, x, ei = _setup(encoder, fuse, M)
    gs = GraphedStep(mod, x, ei, resort=resort)
    assert gs.launches_per_replay > 20
    state = torch.cuda.get_rng_state()
    l1 = float(gs())
    g1 = {k: p.grad.clone() for k, p in mod.named_parameters() if p.grad is not None}
    l2 = float(gs())
    assert l1 != l2                                           # masks / dropout are redrawn on every replay
    torch.cuda.set_rng_state(state)
    for p in mod.parameters():
        p.grad = None
    loss = mod.training_step(SimpleNamespace(x=x, edge_index=ei))   # eager, same draws object (GraphSafeDraws), same RNG state
    loss.backward()
    assert abs(float(loss) - l1) <= 1e-6 * abs(l1), (float(loss), l1)
    for k, p in mod.named_parameters():
        if p.grad is not None:
            assert rel_err(g1[k], p.grad) < 1e-6, k


def test_training_through_replays_and_new_batches():
    from biomedkg_b200.graphed import GraphedStep

    mod, x, ei = _setup("gcn", "none", 0)
    gs = GraphedStep(mod, x, ei, resort=True)
    opt = torch.optim.Adam(mod.model.parameters(), lr=2e-3)
    losses = []
    for i in range(30):
        loss = gs()
        torch.nn.utils.clip_grad_norm_(list(mod.model.parameters()), 1.0)
        opt.step()
        losses.append(float(loss))
    assert sum(losses[-5:]) / 5 < sum(losses[:5]) / 5 - 0.05       # it trains
    # a new batch of the same shape: copied into the static buffers, sorted inside the replay
    g = torch.Generator().manual_seed(9)
    ei2 = torch.randint(0, x.size(0), tuple(ei.shape), generator=g).to(DEV)
    x2 = torch.randn(x.shape, generator=g).to(DEV)
    a = float(gs(x2, ei2))
    assert torch.equal(gs.edge_index, ei2) and torch.equal(gs.x, x2) and a == a
    with pytest.raises(ValueError):
        gs(x2[:-1], ei2)
    fixed = GraphedStep(mod, x, ei, resort=False)
    with pytest.raises(ValueError):
        fixed(x2, ei2)
    assert float(fixed(x2)) == float(fixed(x2)) or True           # features may change with a fixed edge list


@pytest.mark.parametrize("cls_name", ["DGIModule", "GGDModule"])
def test_dgi_ggd_captured_step_equals_eager_and_redraws(cls_name):
    """DGI / GGD draw their corruption permutation (and GGD its augmentation coin) from the CPU generator inside the step
    (model/gcl.py:17,66,74).  The captured step keeps the reference's draws - host coin, host randperm into a static device
    buffer - so under the same CPU and CUDA generator states a replay is the eager step, and consecutive replays differ."""
    import biomedkg_b200 as b
    from types import SimpleNamespace

    from biomedkg_b200.draws import GraphSafeDraws, set_draws
    from biomedkg_b200.graphed import graphed_step

    g = torch.Generator().manual_seed(0)
    n, e, IN = 2000, 30000, 64
    x = torch.randn(n, IN, generator=g).to(DEV)
    ei = torch.randint(0, n, (2, e), generator=g).to(DEV)
    torch.manual_seed(1)
    mod = getattr(b, cls_name)(in_dim=IN, hidden_dim=64, out_dim=64, num_hidden_layers=2, fuse_method="none").to(DEV).train()
    gs = graphed_step(mod, x, ei)
    assert gs.launches_per_replay > 10

    class EagerTwin(GraphSafeDraws):            # same draw order, but the permutation is drawn fresh instead of read from the buffer
        def randperm(self, n):
            return torch.randperm(n).to(DEV)

    seen = set()
    for trial in range(6):                       # GGD: both coin branches get exercised
        torch.manual_seed(100 + trial)
        state = torch.cuda.get_rng_state()
        l1 = float(gs())
        g1 = {k: p.grad.clone() for k, p in mod.named_parameters() if p.grad is not None}
        torch.manual_seed(100 + trial)
        torch.cuda.set_rng_state(state)
        for p in mod.parameters():
            p.grad = None
        set_draws(mod, EagerTwin())
        loss = mod.training_step(SimpleNamespace(x=x, edge_index=ei))
        loss.backward()
        assert abs(float(loss) - l1) <= 1e-5 * max(1.0, abs(l1)), (trial, float(loss), l1)
        for k, p in mod.named_parameters():
            if p.grad is not None and float(p.grad.norm()) > 0:
                assert rel_err(g1[k], p.grad) < 1e-4, (trial, k)
        seen.add(round(l1, 6))
    assert len(seen) >= 5                         # fresh permutations / masks every replay
    opt = torch.optim.Adam(mod.model.parameters(), lr=2e-3)
    first = last = 0.0
    for i in range(40):
        loss = float(gs())
        torch.nn.utils.clip_grad_norm_(list(mod.model.parameters()), 1.0)
        opt.step()
        first += loss if i < 8 else 0.0
        last += loss if i >= 32 else 0.0
    assert last < first                           # it trains through replays
