"""torchrun --nproc-per-node N tools/prof_rowshard.py [cfg4] : kernel-level breakdown (torch.profiler, rank 0) of the row-sharded
GRACE step - which collectives / replicated kernels limit the strong scaling.  Output: profiles/r2_rowshard_breakdown_N*.md"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
from biomedkg_b200.dist import allreduce_grads, shard_layout, sharded_grace_loss

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
cfg = bench.CONFIGS[name]
x, ei = bench.synth(cfg, 42)
mod = bench._make_module(cfg, dev)
params = list(mod.model.parameters())
opt = torch.optim.Adam(params, lr=1e-3)
N = cfg["N"]
if world > 1:
    b0, b1 = shard_layout(N, world)[1][rank]
    xd = x[b0:b1].to(dev)
else:
    xd = x.to(dev)
eid = ei.to(dev)


class B:
    pass


B.x, B.edge_index = xd, eid


def step():
    opt.zero_grad(set_to_none=True)
    if world > 1:
        loss = sharded_grace_loss(mod, xd, eid, num_nodes=N)
        loss.backward()
        allreduce_grads(params)
    else:
        loss = mod.training_step(B)
        loss.backward()
    torch.nn.utils.clip_grad_norm_(params, 1.0)
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
if rank == 0:
    ev = [e for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA]
    tot = sum(e.self_device_time_total for e in ev) / 3
    print(f"# {name} world={world}: {ms:.2f} ms/step (CUDA events); kernel time {tot/1e3:.2f} ms/step on rank 0\n")
    print("| kernel | calls/step | ms/step | % |\n|---|---|---|---|")
    for e in sorted(ev, key=lambda e: -e.self_device_time_total)[:28]:
        print(f"| {e.key[:90]} | {e.count/3:.0f} | {e.self_device_time_total/3e3:.3f} | {100*e.self_device_time_total/3/tot:.1f} |")
if world > 1:
    dist.destroy_process_group()
