"""Time graph indexing: python tools/prof_graph.py N E"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from biomedkg_b200 import ops

n, e = int(sys.argv[1]), int(sys.argv[2])
ei = torch.randint(0, n, (2, e), device="cuda")
keep = torch.rand(e, device="cuda") >= 0.4
def timeit(f, k=5):
    f(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / k
t_sort = timeit(lambda: ops.SortedGraph(ei, n))
sg = ops.SortedGraph(ei, n)
t_view = timeit(lambda: sg.view(keep))
byts = 24 * e + 4 * (n + 1)
print(f"N={n} E={e}: edge_sort (both orientations) {t_sort:.3f} ms = {2*byts/t_sort/1e6:.0f} GB/s algorithmic; "
      f"view (CSR+CSC filter) {t_view:.3f} ms = {2*(byts+e)/t_view/1e6:.0f} GB/s algorithmic")
