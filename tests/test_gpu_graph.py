"""G1/G2 parity (bit-exact): CUDA edge sort + per-view CSR/CSC vs oracle/pyg.py:canonical_csr."""
import numpy as np
import pytest
import torch

from oracle import pyg

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda", torch.cuda.current_device())


def _check_view(ei, n, keep, sg):
    from biomedkg_b200 import ops

    view = sg.view(None if keep is None else keep.to(_dev()), want_perm=True)
    eip = pyg.view_graph(ei, n, keep)
    for by, rp_d, ci_d, pm_d in (("dst", view.rowptr, view.colind, view.perm), ("src", view.csc_rowptr, view.csc_colind, view.csc_perm)):
        rp, ci, pm = pyg.canonical_csr(eip, n, by=by)
        nnz = int(rp[-1])
        assert torch.equal(rp_d.cpu(), rp), f"rowptr mismatch ({by})"
        assert torch.equal(ci_d.cpu()[:nnz], ci), f"colind mismatch ({by})"
        assert torch.equal(pm_d.cpu()[:nnz], pm), f"perm mismatch ({by})"
    assert int(view.nnz.item()) == eip.size(1)
    _, _ = pyg.gcn_norm(ei if keep is None else ei[:, keep], n)
    deg = torch.bincount(eip[1], minlength=n).float()
    assert torch.allclose(view.dis.cpu(), deg.pow(-0.5), rtol=1e-6, atol=0)


@pytest.mark.parametrize("n,e,seed", [(1, 0, 0), (3, 0, 1), (5, 7, 2), (40, 150, 3), (257, 4096, 4), (1000, 30000, 5), (4099, 70001, 6)])
def test_csr_bit_exact_random(n, e, seed):
    from biomedkg_b200 import ops

    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, e), generator=g, dtype=torch.int64)
    if e > 10:
        ei[1, :3] = ei[0, :3]       # self-loops
        ei[:, -4:] = ei[:, :4]      # duplicates
    sg = ops.SortedGraph(ei.to(_dev()), n)
    _check_view(ei, n, None, sg)
    if e:
        for s2 in range(2):
            keep = torch.rand(e, generator=g) >= 0.4
            _check_view(ei, n, keep, sg)
        _check_view(ei, n, torch.zeros(e, dtype=torch.bool), sg)   # everything dropped
        _check_view(ei, n, torch.ones(e, dtype=torch.bool), sg)


def test_csr_edge_cases_isolated_and_hub():
    from biomedkg_b200 import ops

    n = 300
    src = torch.cat([torch.arange(1, n), torch.tensor([5, 5, 5, 7])])
    dst = torch.cat([torch.zeros(n - 1, dtype=torch.int64), torch.tensor([9, 9, 5, 7])])   # star hub 0, dups, self-loops
    ei = torch.stack([src, dst])
    sg = ops.SortedGraph(ei.to(_dev()), n)
    _check_view(ei, n, None, sg)
    g = torch.Generator().manual_seed(9)
    _check_view(ei, n, torch.rand(ei.size(1), generator=g) >= 0.4, sg)


def test_csr_full_size_cfg2_matches_numpy():
    """BASELINE cfg 2 size (28k nodes, 650k edges): bit-exact vs the numpy oracle, plus sortedness."""
    from biomedkg_b200 import ops

    n, e = 28_000, 650_000
    g = torch.Generator().manual_seed(42)
    ei = torch.randint(0, n, (2, e), generator=g, dtype=torch.int64)
    keep = torch.rand(e, generator=g) >= 0.4
    sg = ops.SortedGraph(ei.to(_dev()), n)
    view = sg.view(keep.to(_dev()), want_perm=True)
    eip = pyg.view_graph(ei, n, keep).numpy()
    rp, ci, pm = pyg.canonical_csr_numpy(eip, n)
    nnz = int(rp[-1])
    assert np.array_equal(view.rowptr.cpu().numpy(), rp)
    assert np.array_equal(view.colind.cpu().numpy()[:nnz], ci)
    assert np.array_equal(view.perm.cpu().numpy()[:nnz], pm)


def test_csr_large_properties():
    """8M edges / 130k nodes (cfg 4 size): size-independent properties - rowptr monotone, rows sorted,
    multiset of edges preserved, perm is a permutation, idempotent under re-sorting."""
    from biomedkg_b200 import ops

    n, e = 130_000, 8_000_000
    g = torch.Generator().manual_seed(7)
    ei = torch.randint(0, n, (2, e), generator=g, dtype=torch.int64).to(_dev())
    keep = (torch.rand(e, device=_dev()) >= 0.4)
    sg = ops.SortedGraph(ei, n)
    view = sg.view(keep, want_perm=True)
    nnz = int(view.nnz.item())
    rp, ci, pm = view.rowptr.long(), view.colind[:nnz].long(), view.perm[:nnz].long()
    kept = ei[:, keep & (ei[0] != ei[1])]
    assert nnz == kept.size(1) + n and int(rp[-1]) == nnz and bool((rp[1:] >= rp[:-1]).all())
    row_of = torch.repeat_interleave(torch.arange(n, device=_dev()), rp[1:] - rp[:-1])
    key = row_of * n + ci
    assert bool((key[1:] >= key[:-1]).all())                                 # globally sorted by (dst, src)
    eip = torch.cat([kept, torch.arange(n, device=_dev()).expand(2, -1)], 1)
    assert torch.equal(torch.sort(eip[1] * n + eip[0]).values, key)           # same edge multiset
    assert torch.equal(torch.sort(pm).values, torch.arange(nnz, device=_dev()))
    assert torch.equal(eip[0][pm], ci) and torch.equal(eip[1][pm], row_of)   # perm maps ei' -> CSR
    same = pm[1:][key[1:] == key[:-1]] > pm[:-1][key[1:] == key[:-1]]
    assert bool(same.all())                                                   # stability among duplicates
