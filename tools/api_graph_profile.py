"""Where the GPU time of the graphed drop-in API step goes: torch.profiler (CUPTI) over a few graph replays of
ParameterServerTrainer.train_minibatch(args.cuda_graph=True) on the bench workload; kernel time by name, per step.
Usage: python tools/api_graph_profile.py [--batch 32768] [--steps 5] > gpurun_out/api_graph_profile.json"""
import argparse
import collections
import json
import os
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from elasticdl_b200.ps import PSGroup  # noqa: E402
from elasticdl_b200.worker.ps_client import PSClient  # noqa: E402
from elasticdl_b200.worker.ps_trainer import ParameterServerTrainer  # noqa: E402
from elasticdl_b200.workloads.deepfm import DeepFMLayersModel, synthetic_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32768)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    batches = [synthetic_batch(args.batch, 100 + i, dev, "zipf") for i in range(4)]
    group = PSGroup(1, "Adam", bench.ADAM_ARGS, device=0)
    client = PSClient(group)
    client.dense_output = "torch"
    model = DeepFMLayersModel().to(dev)
    trainer = ParameterServerTrainer(model, client, args=types.SimpleNamespace(
        get_model_steps=1, batched_embedding_lookups=True, cuda_graph=True, cuda_graph_warmup=3))
    feats = [(DeepFMLayersModel.features_of(ids, dense), labels) for ids, dense, labels in batches]
    for i in range(6):
        trainer.train_minibatch(*feats[i % 4])
    torch.cuda.synchronize()
    assert isinstance(trainer._graph_state, dict), trainer.graph_fallback_reason
    t0 = time.perf_counter()
    for i in range(20):
        trainer.train_minibatch(*feats[i % 4])
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 20
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(args.steps):
            trainer.train_minibatch(*feats[i % 4])
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            a = agg[ev.name[:90]]
            a[0] += 1
            a[1] += ev.device_time
    rows = sorted(((n, c / args.steps, t / args.steps) for n, (c, t) in agg.items()), key=lambda r: -r[2])
    out = {"wall_ms_per_step": wall * 1e3, "gpu_busy_us_per_step": sum(r[2] for r in rows),
           "kernels_per_step": sum(r[1] for r in rows),
           "top": [{"name": n, "launches_per_step": c, "us_per_step": round(t, 2)} for n, c, t in rows[:45]]}
    print(json.dumps(out, indent=1))
    group.close()


if __name__ == "__main__":
    main()
