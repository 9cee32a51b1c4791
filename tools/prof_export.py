"""Embedding-export throughput: all one-seed star graphs of node.py:193-241 in one pass per layer on the
GPU vs the reference's per-seed loop (oracle/export.py, CPU, bounded sample of seeds).  usage: prof_export.py [N E M]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import biomedkg_b200 as b  # noqa: E402
from biomedkg_b200 import _cabi  # noqa: E402
from biomedkg_b200.export import star_embeddings  # noqa: E402

N, E, M = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (130_000, 8_000_000, 3)
dev = "cuda"
g = torch.Generator().manual_seed(42)
x = torch.randn(N, M, 768, generator=g)
x = x / x.norm(dim=1, keepdim=True)
ei = torch.randint(0, N, (2, E), generator=g)
torch.manual_seed(42)
mod = b.GRACEModule(in_dim=768, hidden_dim=256, out_dim=256, num_hidden_layers=2, fuse_method="attention").to(dev)
xd, eid = x.to(dev), ei.to(dev)
for _ in range(3):
    out = star_embeddings(mod, xd, eid)
torch.cuda.synchronize()
_cabi.timed_entries.add("bmkg_gcn_star_aggregate")
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
t0.record()
for _ in range(reps):
    out = star_embeddings(mod, xd, eid)
t1.record()
torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / reps
tm = {k: sum(a.elapsed_time(b_) for a, b_ in v) / len(v) for k, v in _cabi.timings.items()}
res = {"N": N, "E": E, "M": M, "gpu_ms": ms, "gpu_nodes_per_s": N / (ms * 1e-3), "kernel_timings": tm}

# CPU: the reference's loop on a bounded sample of seeds
from oracle import export as oe  # noqa: E402
from oracle import models as om  # noqa: E402

orc = om.GRACEModule(in_dim=768, hidden_dim=256, out_dim=256, num_hidden_layers=2, fuse_method="attention")
orc.load_state_dict({k: v.cpu() for k, v in mod.state_dict().items()})
orc.eval()
S = int(os.environ.get("EXPORT_CPU_SEEDS", "200"))
order = torch.argsort(ei[1], stable=True)                      # pre-index in-edges so the sample timing is the model, not the scan
dst_sorted, src_sorted = ei[1][order], ei[0][order]
ptr = torch.searchsorted(dst_sorted, torch.arange(N + 1))
tc = time.perf_counter()
rows = []
with torch.no_grad():
    for s in range(S):
        nb = src_sorted[ptr[s] : ptr[s + 1]]
        sub = torch.stack([nb, torch.full_like(nb, s)])
        nodes, sei = oe.one_hop_batch(sub, s)
        rows.append(orc(x[nodes], sei)[:1])
cpu_s = time.perf_counter() - tc
ref = torch.cat(rows)
err = float((out[:S].cpu().double() - ref.double()).norm() / ref.double().norm())
res.update({"cpu_seeds": S, "cpu_nodes_per_s": S / cpu_s, "cpu_threads": torch.get_num_threads(), "rel_err_vs_loop": err})
print(json.dumps(res))
