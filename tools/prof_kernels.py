"""Small driver for ncu captures: runs the hot kernels stand-alone at a BASELINE config size.
    ncu --set full --clock-control none --import-source on -k regex:infonce -s 2 -c 2 -o gpurun_out/prof python tools/prof_kernels.py infonce 28000
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from biomedkg_b200 import ops

what, n = sys.argv[1], int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = "cuda"
torch.manual_seed(0)
if what == "infonce":
    h1 = torch.randn(n, 256, device=dev, requires_grad=True)
    h2 = (h1.detach() + torch.randn(n, 256, device=dev)).requires_grad_(True)
    for _ in range(iters):
        loss = ops.infonce_loss(h1, h2, 0.2)
        loss.backward()
    torch.cuda.synchronize()
    e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e0.record()
    loss = ops.infonce_loss(h1, h2, 0.2)
    e1.record()
    loss.backward()
    e2.record()
    torch.cuda.synchronize()
    f, b = e0.elapsed_time(e1), e1.elapsed_time(e2)
    print(f"N={n} fwd {f:.3f} ms ({6*n*n*256/f/1e9:.0f} TF/s credited)  bwd {b:.3f} ms ({8*n*n*256/b/1e9:.0f} TF/s credited) loss {float(loss):.5f}")
elif what in ("gcn", "gat", "gcn_powerlaw"):
    e = int(sys.argv[4]) if len(sys.argv) > 4 else n * 23
    if what == "gcn_powerlaw":   # destinations ~ Pareto(alpha=2.1) degree sequence, sources uniform (cfg 5)
        w = (1.0 - torch.rand(n, device=dev, dtype=torch.float64)).pow(-1.0 / 1.1)
        dst = torch.multinomial((w / w.sum()).float(), e, replacement=True)
        ei = torch.stack([torch.randint(0, n, (e,), device=dev), dst])
    else:
        ei = torch.randint(0, n, (2, e), device=dev)
    view = ops.SortedGraph(ei, n).view(torch.rand(e, device=dev) >= 0.4)
    x = torch.randn(n, 256, device=dev).bfloat16()
    bias = torch.zeros(256, device=dev)
    nnz = int(view.nnz.item())
    deg = (view.rowptr[1:] - view.rowptr[:-1])
    print(f"max in-degree {int(deg.max())}, rows with degree > 1024: {int((deg > 1024).sum())}, edges in them {int(deg[deg > 1024].sum())}")
    for _ in range(iters):
        y = ops.gcn_aggregate(view.rowptr, view.colind, view.dis, x, bias, relu=True, drop_p=0.2, drop_seed=1, hub_rows=view.hub_csr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        y = ops.gcn_aggregate(view.rowptr, view.colind, view.dis, x, bias, relu=True, drop_p=0.2, drop_seed=1, hub_rows=view.hub_csr)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    byts = nnz * (256 * 2 + 4) + n * 256 * 2 + 4 * (n + 1)
    print(f"gcn_aggregate N={n} nnz={nnz}: {ms*1e3:.1f} us, {byts/ms/1e6:.0f} GB/s algorithmic")
if what == "redaf":   # python tools/prof_kernels.py redaf N [iters] [M]: fused ReDAF epilogue, forward and backward, E = 768
    M = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    E = 768
    t = torch.randn(n, M, E, device=dev).to(torch.bfloat16)
    bias = torch.randn(E, device=dev, requires_grad=True)
    gate = (torch.rand(M, E, device=dev) + 0.5).requires_grad_(True)
    tt = t.float().requires_grad_(True)

    def run():
        o = ops.redaf_fuse(tt.to(torch.bfloat16), bias, gate, 0.1, 1234, None)
        o.backward(torch.ones_like(o))

    from biomedkg_b200 import _cabi

    for _ in range(iters):
        run()
    torch.cuda.synchronize()
    _cabi.timed_entries.update({"bmkg_redaf_fwd", "bmkg_redaf_bwd"})
    _cabi.timings.clear()
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    tm = {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in _cabi.timings.items()}
    fb = n * M * E * 2 + n * E * 4
    bb = n * E * 4 + 2 * n * M * E * 2
    print(f"redaf N={n} M={M} E={E}: fwd {tm['bmkg_redaf_fwd']:.3f} ms = {fb / tm['bmkg_redaf_fwd'] / 1e6:.0f} GB/s algorithmic; "
          f"bwd {tm['bmkg_redaf_bwd']:.3f} ms = {bb / tm['bmkg_redaf_bwd'] / 1e6:.0f} GB/s algorithmic")
elif what == "star":  # python tools/prof_kernels.py star N [iters] [E]: export-path star aggregation, C = 256, rows > L2 when N >= 1M
    e = int(sys.argv[4]) if len(sys.argv) > 4 else n * 30
    ei = torch.randint(0, n, (2, e), device=dev)
    view = ops.SortedGraph(ei, n).view(None)
    leaf = torch.randn(n, 256, device=dev).to(torch.bfloat16)
    seed = torch.randn(n, 256, device=dev).to(torch.bfloat16)
    bias = torch.zeros(256, device=dev)
    for _ in range(iters):
        ops.gcn_star_aggregate(view.rowptr, view.colind, view.dis, leaf, seed, bias, relu=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.gcn_star_aggregate(view.rowptr, view.colind, view.dis, leaf, seed, bias, relu=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    nnz = int(view.nnz.item())
    byts = nnz * (256 * 2 + 4) + n * 256 * 2 * 2 + 4 * (n + 1)
    print(f"star N={n} nnz={nnz}: {ms:.3f} ms = {byts / ms / 1e6:.0f} GB/s algorithmic")
