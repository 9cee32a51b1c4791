"""A/B of the hand-written tcgen05 Linear kernels (csrc/gemm.cu) against the library GEMM (torch.mm -> cuBLAS) on the shapes of
the GCL step.  python tools/prof_gemm.py  -> a markdown table (committed as profiles/r2_gemm_ab.md)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from biomedkg_b200 import ops

DEV = "cuda"


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


rows = []
for name, M, N, K in [("cfg4 fusion q/k/v  [N*M,768]x[768,2304]", 390_000, 2304, 768), ("cfg2 fusion q/k/v", 56_000, 2304, 768),
                      ("cfg4 GCN layer 0   [N,768]x[768,256]", 130_000, 256, 768), ("cfg4 GCN layer 1-3 / projector [N,256]x[256,256]", 130_000, 256, 256),
                      ("cfg2 layer 1-3", 28_000, 256, 256), ("cfg5 layer 1-3", 1_000_000, 256, 256)]:
    a = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) / K ** 0.5).to(torch.bfloat16)
    g = torch.randn(M, N, device=DEV).to(torch.bfloat16)
    wt = w.t().contiguous()
    fl = 2.0 * M * N * K
    t_nt, t_nt_lib = timeit(lambda: ops.gemm_nt(a, w)), timeit(lambda: torch.mm(a, w.t()))
    t_dx, t_dx_lib = timeit(lambda: ops.gemm_nt(g, wt)), timeit(lambda: torch.mm(g, w))
    t_tn, t_tn_lib = timeit(lambda: ops.gemm_tn(g, a)), timeit(lambda: ops._mm_f32(g.t(), a))
    rows.append((name, M, N, K, t_nt, t_nt_lib, t_dx, t_dx_lib, t_tn, t_tn_lib, fl))
print("| shape | ours fwd ms (TF/s) | cuBLAS fwd ms | ours dX ms | cuBLAS dX ms | ours dW ms | cuBLAS dW ms |")
print("|---|---|---|---|---|---|---|")
for name, M, N, K, a, b, c, d, e, f, fl in rows:
    print(f"| {name} M={M} | {a:.3f} ({fl / a / 1e9:.0f}) | {b:.3f} ({fl / b / 1e9:.0f}) | {c:.3f} | {d:.3f} | {e:.3f} | {f:.3f} |")
