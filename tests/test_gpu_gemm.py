"""L1 parity: the hand-written tcgen05 Linear kernels (csrc/gemm.cu) against a plain PyTorch fp32 matmul of the SAME bf16
operands (so only the accumulation order differs): fp32 outputs <= 2e-5 relative, bf16 outputs = correctly rounded fp32
results up to 1 ulp, fused epilogues (fp32 bias, ELU, GATConv node scores) included."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ops():
    from biomedkg_b200 import ops

    return ops


@pytest.mark.parametrize("M,N,K", [(1000, 256, 768), (300, 2304, 768), (28_000, 256, 256), (130, 192, 64), (5, 64, 64), (4096, 768, 256),
                                   (129, 16, 128), (777, 320, 192)])
@pytest.mark.parametrize("bias,elu,out_f32", [(False, False, False), (True, False, True), (True, True, True), (True, False, False)])
def test_linear_nt_matches_torch(M, N, K, bias, elu, out_f32):
    ops = _ops()
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(DEV).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV).to(torch.bfloat16)
    b = torch.randn(N, generator=g).to(DEV) if bias else None
    before = ops.library_gemm_calls
    y = ops.gemm_nt(a, w, b, elu=elu, out_f32=out_f32)
    assert ops.library_gemm_calls == before                      # the tcgen05 kernel took it, not the library fallback
    ref = a.float() @ w.float().t()
    if bias:
        ref = ref + b
    if elu:
        ref = torch.nn.functional.elu(ref)
    assert y.shape == (M, N) and y.dtype == (torch.float32 if out_f32 else torch.bfloat16)
    if out_f32:
        assert rel_err(y, ref) < 2e-5
        assert float((y - ref).abs().max()) < 1e-3 * float(ref.abs().max())
    else:
        assert rel_err(y, ref) < 3e-3                              # bf16 rounding of the output
        assert float((y.float() - ref).abs().max()) <= 2.0 ** -7 * float(ref.abs().max())


@pytest.mark.parametrize("M,C,heads,K", [(1000, 256, 1, 768), (333, 64, 4, 256), (5000, 128, 2, 64)])
def test_linear_nt_gat_scores_epilogue(M, C, heads, K):
    """The GEMM epilogue's node scores equal bmkg_gat_scores' (<bf16-rounded row, att>) to fp32 accumulation order."""
    ops = _ops()
    from biomedkg_b200.ops import _p, _stream, call

    g = torch.Generator().manual_seed(M)
    N = heads * C
    a = torch.randn(M, K, generator=g).to(DEV).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV).to(torch.bfloat16)
    atts, attd = torch.randn(N, generator=g).to(DEV), torch.randn(N, generator=g).to(DEV)
    y, a_s, a_d = ops.gemm_nt(a, w, None, gat=(atts, attd, heads))
    y0 = ops.gemm_nt(a, w, None)
    assert torch.equal(y, y0)
    r_s = torch.empty(M, heads, device=DEV)
    r_d = torch.empty(M, heads, device=DEV)
    call("bmkg_gat_scores", _p(y0), _p(atts), _p(attd), M, heads, C, _p(r_s), _p(r_d), _stream())
    assert rel_err(a_s, r_s) < 1e-5 and rel_err(a_d, r_d) < 1e-5
    ref = (y0.float().view(M, heads, C) * atts.view(1, heads, C)).sum(-1)
    assert rel_err(a_s, ref) < 1e-5


@pytest.mark.parametrize("M,N,K", [(1000, 256, 768), (28_000, 256, 256), (50, 192, 64), (130_000, 256, 768), (70_000, 2304, 768), (63, 8, 64),
                                   (4097, 768, 256)])
def test_linear_tn_matches_torch(M, N, K):
    """dW = dY^T X reduced over the node dimension; with an fp32 addend; deterministic (no atomics)."""
    ops = _ops()
    g = torch.Generator().manual_seed(M + N)
    dy = torch.randn(M, N, generator=g).to(DEV).to(torch.bfloat16)
    x = torch.randn(M, K, generator=g).to(DEV).to(torch.bfloat16)
    add = torch.randn(N, K, generator=g).to(DEV)
    before = ops.library_gemm_calls
    out = ops.gemm_tn(dy, x, addend=add)
    assert ops.library_gemm_calls == before
    ref = (dy.double().t() @ x.double()).float() + add
    assert rel_err(out, ref) < 2e-5
    assert torch.equal(out, ops.gemm_tn(dy, x, addend=add))


def test_no_library_gemm_in_the_baseline_configurations():
    """cfg2's module (attention fusion + GAT + projector) and cfg1's (GCN) run every Linear on the tcgen05 kernels."""
    import biomedkg_b200 as b

    ops = _ops()
    torch.manual_seed(0)
    for enc, fuse, m in (("gat", "attention", 2), ("gcn", "none", 1)):
        n = 3000
        x = torch.randn(n, m, 768) if m > 1 else torch.randn(n, 768)
        mod = b.GRACEModule(768, 256, 256, 2, fuse_method=fuse, encoder=enc).to(DEV).train()

        class Batch:
            pass

        Batch.x, Batch.edge_index = x.to(DEV), torch.randint(0, n, (2, 30_000)).to(DEV)
        before = ops.library_gemm_calls
        mod.training_step(Batch).backward()
        assert ops.library_gemm_calls == before
