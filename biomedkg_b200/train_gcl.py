"""Dependency-free equivalent of the reference's train_gcl.py loop (hydra/Lightning/Comet are absent from the image).

    python -m biomedkg_b200.train_gcl --model grace --nodes 8000 --edges 200000 --epochs 5

``--regime fullgraph`` (default): full-graph training on a synthetic graph of the requested shape.  ``--regime minibatch``: the
reference's own loop - RandomLinkSplit(0.1, 0.2) -> NeighborLoader([30]*3, batch 64, shuffle) -> fit (train + validation every
epoch) -> test (train_gcl.py:108-122, data_module.py:65-99), with the GPU sampler of loader.py.  Both: seed_everything(42)
(configs/gcl.yaml:6), Adam over ``module.model`` only, cosine/linear warm-up schedule sized in steps, gradient clip 1.0
(train_gcl.py:99), JSONL log of the losses and nodes/s (the reference logs to Comet)."""
from __future__ import annotations

import argparse
import json
import time

import torch


def _evaluate(mod, loader, step_name):
    """validation / test epoch of the Lightning loop: module.validation_step / test_step over a loader, no gradients."""
    total, n = 0.0, 0
    mod.eval()
    with torch.no_grad():
        for i, batch in enumerate(loader):
            total += float(getattr(mod, step_name)(batch, i))
            n += 1
    mod.train()
    return total / max(n, 1)


def run_minibatch(a, mod, data, opt, log):
    """The reference's own regime (train_gcl.py:108-122 over GCLDataModule, data_module.py:65-99): RandomLinkSplit(0.1, 0.2) ->
    NeighborLoader([30]*3, batch_size, shuffle=True) for training, the same loader over the validation / test splits,
    ``trainer.fit`` (train + validate every epoch) then ``trainer.test``."""
    from .loader import NeighborLoader, random_link_split

    train_d, val_d, test_d = random_link_split(data, num_val=0.1, num_test=0.2)
    fan = [a.fanout] * a.hops
    train_l = NeighborLoader(train_d, fan, batch_size=a.batch_size, shuffle=True)
    val_l = NeighborLoader(val_d, fan, batch_size=a.val_batch_size, shuffle=False)
    test_l = NeighborLoader(test_d, fan, batch_size=a.val_batch_size, shuffle=False)
    steps = a.epochs * (min(len(train_l), a.limit_batches) if a.limit_batches else len(train_l))
    sched = mod._get_scheduler(opt, num_training_steps=steps)
    for epoch in range(a.epochs):
        torch.cuda.synchronize()
        t0, seen, tot, nb = time.perf_counter(), 0, 0.0, 0
        for i, batch in enumerate(train_l):
            if a.limit_batches and i >= a.limit_batches:
                break
            opt.zero_grad(set_to_none=True)
            loss = mod.training_step(batch, i)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(mod.model.parameters(), 1.0)
            opt.step()
            sched.step()
            tot += float(loss.detach())
            nb += 1
            seen += int(batch.x.size(0))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        val = _evaluate(mod, _limited(val_l, a.limit_batches), "validation_step")
        log({"epoch": epoch, "train_loss": tot / max(nb, 1), "val_loss": val, "batches": nb, "sampled_nodes_per_s": seen / dt,
             "lr": sched.get_last_lr()[0]})
    log({"test_loss": _evaluate(mod, _limited(test_l, a.limit_batches), "test_step")})


def _limited(loader, k):
    for i, b in enumerate(loader):
        if k and i >= k:
            return
        yield b


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
                    help="replay forward+backward from a CUDA graph (graphed.graphed_step; fixed full graph)")
    ap.add_argument("--regime", default="fullgraph", choices=["fullgraph", "minibatch"],
                    help="fullgraph: one full-graph step per epoch; minibatch: the reference's RandomLinkSplit + NeighborLoader loop with validation and test")
    ap.add_argument("--batch-size", type=int, default=64)
    ap.add_argument("--val-batch-size", type=int, default=128)
    ap.add_argument("--fanout", type=int, default=30)
    ap.add_argument("--hops", type=int, default=3)
    ap.add_argument("--limit-batches", type=int, default=0, help="cap the batches per epoch / evaluation (0 = all)")
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
    out = open(a.log, "a") if a.log else None

    def log(rec):
        print(json.dumps(rec))
        if out:
            out.write(json.dumps(rec) + "\n")

    if a.regime == "minibatch":
        run_minibatch(a, mod, Batch, opt, log)
        return
    sched = mod._get_scheduler(opt, num_training_steps=a.epochs)
    graphed = None
    if a.cuda_graph:
        from .graphed import graphed_step

        graphed = graphed_step(mod, Batch.x, Batch.edge_index, resort=False)
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
        log({"epoch": epoch, "train_loss": float(loss.detach()), "nodes_per_s": a.nodes / dt, "lr": sched.get_last_lr()[0]})


if __name__ == "__main__":
    main()
