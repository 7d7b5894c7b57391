#!/usr/bin/env python
"""BASELINE.json configs[2]: census wide & deep, async SGD, 4 workers (SURVEY.md section 8d item 3;
model_zoo/census_wide_deep_model/wide_deep_functional_api.py:164-212): 3 wide (dim 1) + 3 deep (dim 8) small
embedding tables, MLP 24 -> 16 -> 8 -> 4, batch 64 (scripts/client_test.sh:38), ids uniform, seed per worker
= 100 + rank, SGD lr 0.1 (and Adam 1e-3), lr-staleness modulation on / off.

The four workers are threads of one process, each with its own client view of the shared HBM shards and its own
stream (worker_ps_interaction_test.py:136-151 drives workers as threads too), and each runs the DROP-IN API:
ParameterServerTrainer.train_minibatch over six elasticdl Embedding layers.  At batch 64 a minibatch is a few
KB of rows: the step is bound by host latency (Python + ~40 small launches), not by the GPU -- the number says
what the drop-in API costs per minibatch, nothing about bandwidth.  Prints one JSON line per configuration."""
import argparse
import json
import os
import sys
import threading
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

OPTS = {"sgd": ("SGD", "learning_rate=0.1;momentum=0.0;nesterov=false;"),
        "adam": ("Adam", "learning_rate=0.001;beta_1=0.9;beta_2=0.999;epsilon=1e-07;amsgrad=false;")}
ROWS = (9, 16, 7, 1000, 1000, 1000)  # workclass / education / marital-status buckets, three hashed crosses


class CensusWideDeep(torch.nn.Module):
    def __init__(self, w, lr):
        super().__init__()
        from elasticdl_b200.layers import Embedding

        self.wide = torch.nn.ModuleList([Embedding(1, input_dim=ROWS[i], embeddings_initializer="zero",
                                                   name="wide_%d" % i) for i in range(3)])
        self.deep = torch.nn.ModuleList([Embedding(8, input_dim=ROWS[3 + i], embeddings_initializer="uniform",
                                                   name="deep_%d" % i) for i in range(3)])
        self.mlp = torch.nn.Sequential(torch.nn.Linear(24, 16), torch.nn.ReLU(), torch.nn.Linear(16, 8), torch.nn.ReLU(),
                                       torch.nn.Linear(8, 4), torch.nn.ReLU(), torch.nn.Linear(4, 1))
        self.optimizer = torch.optim.SGD(self.mlp.parameters(), lr=lr)
        bce = torch.nn.BCEWithLogitsLoss()
        self.loss = lambda labels, logits: bce(logits, labels)

    def forward(self, f):
        wide = sum(self.wide[i](f["w%d" % i]).squeeze(-1) for i in range(3))
        deep = torch.cat([self.deep[i](f["d%d" % i]) for i in range(3)], 1)
        return wide + self.mlp(deep).squeeze(1)


def run(opt, modulation, workers, steps, batch):
    from elasticdl_b200.ps import PSGroup
    from elasticdl_b200.worker.ps_client import PSClient
    from elasticdl_b200.worker.ps_trainer import ParameterServerTrainer

    torch.manual_seed(0)
    group = PSGroup(1, *OPTS[opt], device=0, lr_staleness_modulation=modulation)
    lr = 0.1 if opt == "sgd" else 1e-3
    dev = torch.device("cuda", 0)
    base = PSClient(group)
    base.dense_output = "torch"
    models = [CensusWideDeep(w, lr).to(dev) for w in range(workers)]
    for m in models[1:]:
        m.load_state_dict(models[0].state_dict())
    results, errors = [None] * workers, []
    # set-up is sequential (worker 0 registers the tables and initialises the dense parameters -- first writer
    # wins, server.go:209-221 -- the others attach as cloned views); only the timed loops run concurrently
    ctx = []
    for w in range(workers):
        view = group if w == 0 else group.clone_view()
        client = base if w == 0 else PSClient(view)
        client.dense_output = "torch"
        stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(stream):
            trainer = ParameterServerTrainer(models[w], client, args=types.SimpleNamespace(get_model_steps=1))
            gen = torch.Generator(device=dev).manual_seed(100 + w)
            data = []
            for s in range(8):
                f = {}
                for i in range(3):
                    f["w%d" % i] = torch.randint(0, ROWS[i], (batch,), generator=gen, device=dev)
                    f["d%d" % i] = torch.randint(0, ROWS[3 + i], (batch,), generator=gen, device=dev)
                data.append((f, (torch.rand(batch, generator=gen, device=dev) < 0.25).float()))
            for s in range(5):
                trainer.train_minibatch(*data[s % 8])
            stream.synchronize()
        ctx.append((view, stream, trainer, data))
    barrier = threading.Barrier(workers + 1)

    def worker(w):
        try:
            view, stream, trainer, data = ctx[w]
            with torch.cuda.stream(stream):
                barrier.wait()
                t0 = time.perf_counter()
                for s in range(steps):
                    accepted, version, loss = trainer.train_minibatch(*data[s % 8])
                stream.synchronize()
                results[w] = (time.perf_counter() - t0, version, float(loss))
                barrier.wait()
        except Exception as e:  # noqa: BLE001
            errors.append((w, repr(e)))
            barrier.abort()

    threads = [threading.Thread(target=worker, args=(w,)) for w in range(workers)]
    [t.start() for t in threads]
    try:
        barrier.wait()
        t0 = time.perf_counter()
        barrier.wait()
        wall = time.perf_counter() - t0
    except threading.BrokenBarrierError:
        wall = float("nan")
    [t.join() for t in threads]
    if errors:
        raise RuntimeError(errors)
    version = group.snapshot()[0][0]
    group.close()
    return {"config": "census wide&deep, async %s, %d workers (threads), batch %d, lr_staleness_modulation=%s"
                      % (opt, workers, batch, modulation),
            "samples_per_s": workers * steps * batch / wall, "minibatches_per_s": workers * steps / wall,
            "ms_per_minibatch_per_worker": 1e3 * wall / steps, "ps_version": version,
            "final_loss": results[0][2], "path": "ParameterServerTrainer.train_minibatch (drop-in API), eager"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=4)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    for opt in ("sgd", "adam"):
        for mod in (False, True):
            print(json.dumps(run(opt, mod, args.workers, args.steps, args.batch)), flush=True)


if __name__ == "__main__":
    main()
