"""Neighbour-sampler throughput: NeighborLoader([30]*3, batch_size=128) batches per second on the GPU vs the
plain-Python restatement on the host (bounded sample).  usage: prof_sampler.py [N E]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from biomedkg_b200 import _cabi  # noqa: E402
from biomedkg_b200.loader import NeighborSampler  # noqa: E402

N, E = (int(a) for a in sys.argv[1:3]) if len(sys.argv) >= 3 else (130_000, 8_000_000)
g = torch.Generator().manual_seed(42)
ei = torch.randint(0, N, (2, E), generator=g)
eid = ei.cuda()
smp = NeighborSampler(eid, N, [30, 30, 30])
order = torch.randperm(N, generator=g).cuda()
res = {"N": N, "E": E, "num_neighbors": [30, 30, 30]}
for bs in (128, 1024):
    batches = [order[i * bs : (i + 1) * bs] for i in range(12)]
    for s in batches[:2]:
        smp.sample(s, 1)
    torch.cuda.synchronize()
    for name in ("bmkg_sample_count", "bmkg_sample_pick", "bmkg_sample_relabel"):
        _cabi.timed_entries.add(name)
    _cabi.timings.clear()
    t0 = time.perf_counter()
    nodes = edges = 0
    for i, s in enumerate(batches[2:]):
        n_id, sub, _ = smp.sample(s, 100 + i)
        nodes += n_id.numel()
        edges += sub.size(1)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    kt = {k: sum(a.elapsed_time(b) for a, b in v) / 10 for k, v in _cabi.timings.items()}
    _cabi.timed_entries.clear()
    res[f"batch{bs}"] = {"ms_per_batch_wall": dt * 1e3, "nodes_per_batch": nodes / 10, "edges_per_batch": edges / 10,
                         "sampled_edges_per_s": edges / 10 / dt, "kernel_ms_per_batch": kt}
from oracle import sampler as osamp  # noqa: E402

t0 = time.perf_counter()
rn, rs, _ = osamp.sample(ei.numpy(), N, order[:8].tolist(), [30, 30, 30], 1)
dt = time.perf_counter() - t0
res["cpu_python_port"] = {"seeds": 8, "s": dt, "sampled_edges_per_s": rs.shape[1] / dt, "note": "includes the argsort of the edge list"}
print(json.dumps(res))
