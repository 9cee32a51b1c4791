"""Dependency-free equivalent of the reference's train_gcl.py loop (hydra/Lightning/Comet are absent from the image).

    python -m biomedkg_b200.train_gcl --model grace --nodes 8000 --edges 200000 --epochs 5

Full-graph training on a synthetic graph of the requested shape: seed_everything(42) (configs/gcl.yaml:6), Adam over
``module.model`` only, cosine/linear warm-up schedule sized in steps, gradient clip 1.0 (train_gcl.py:99), JSONL log of
loss and nodes/s (the reference logs to Comet)."""
from __future__ import annotations

import argparse
import json
import time

import torch


def main():
    from . import DGIModule, GGDModule, GRACEModule

    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="grace", choices=["grace", "dgi", "ggd"])
    ap.add_argument("--encoder", default="gcn", choices=["gcn", "gat"])
    ap.add_argument("--fuse-method", default="none")
    ap.add_argument("--nodes", type=int, default=8000)
    ap.add_argument("--edges", type=int, default=200_000)
    ap.add_argument("--modalities", type=int, default=1)
    ap.add_argument("--in-dim", type=int, default=768)
    ap.add_argument("--hidden-dim", type=int, default=256)
    ap.add_argument("--num-hidden-layers", type=int, default=2)
    ap.add_argument("--learning-rate", type=float, default=1e-3)
    ap.add_argument("--warm-up-ratio", type=float, default=0.2)
    ap.add_argument("--scheduler-type", default="cosine")
    ap.add_argument("--epochs", type=int, default=10)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--log", default=None)
    ap.add_argument("--cuda-graph", action="store_true",
                    help="replay forward+backward from a CUDA graph (graphed.GraphedStep; GRACE only, fixed full graph)")
    a = ap.parse_args()

    torch.manual_seed(a.seed)
    dev = torch.device("cuda", torch.cuda.current_device())
    cls = {"grace": GRACEModule, "dgi": DGIModule, "ggd": GGDModule}[a.model]
    mod = cls(in_dim=a.in_dim, hidden_dim=a.hidden_dim, out_dim=a.hidden_dim, num_hidden_layers=a.num_hidden_layers,
              scheduler_type=a.scheduler_type, learning_rate=a.learning_rate, warm_up_ratio=a.warm_up_ratio,
              fuse_method=a.fuse_method, encoder=a.encoder).to(dev).train()
    g = torch.Generator().manual_seed(a.seed)
    if a.modalities > 1:
        x = torch.randn(a.nodes, a.modalities, a.in_dim, generator=g)
        x = x / x.norm(dim=1, keepdim=True)
    else:
        x = torch.nn.init.xavier_normal_(torch.empty(a.nodes, a.in_dim), generator=g)

    class Batch:
        pass

    Batch.x = x.to(dev)
    Batch.edge_index = torch.randint(0, a.nodes, (2, a.edges), generator=g, dtype=torch.int64).to(dev)
    opt = torch.optim.Adam(mod.model.parameters(), lr=a.learning_rate)
    sched = mod._get_scheduler(opt, num_training_steps=a.epochs)
    out = open(a.log, "a") if a.log else None
    graphed = None
    if a.cuda_graph:
        if a.model != "grace":
            raise SystemExit("--cuda-graph: DGI / GGD draw from the CPU generator inside the step; only GRACE is capturable")
        from .graphed import GraphedStep

        graphed = GraphedStep(mod, Batch.x, Batch.edge_index, resort=False)
    for epoch in range(a.epochs):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if graphed is not None:
            loss = graphed()
        else:
            opt.zero_grad(set_to_none=True)
            loss = mod.training_step(Batch)
            loss.backward()
        torch.nn.utils.clip_grad_norm_(mod.model.parameters(), 1.0)
        opt.step()
        sched.step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        rec = {"epoch": epoch, "train_loss": float(loss.detach()), "nodes_per_s": a.nodes / dt, "lr": sched.get_last_lr()[0]}
        print(json.dumps(rec))
        if out:
            out.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
