#!/usr/bin/env python
"""BASELINE.json configs[3]: ResNet-50-sized dense model (25.6 M fp32 in 214 tensors) under the allreduce
controller at 1 / 2 / 4 / 8 GPUs -- the time of one `synchronize() + step()` (gradient averaging + optimizer
update) with

  fused   DistributedOptimizer(fused=True): reduce-scatter + averaging + SGD-momentum update in ONE kernel
          reading the peers' gradient buckets over NVLink (b200ps_push_dense_reduce), all-gather by
          b200ps_pull_dense, device-side barriers -- no collective library on the data path;
  nccl    DistributedOptimizer(fused=False): one NCCL all-reduce of the flat bucket + the wrapped torch
          optimizer's step (the reference's structure, elasticai_api/pytorch/optimizer.py:141-207, with one
          collective instead of 214).

Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
            --master-port 29541 tools/bench_allreduce.py [--steps 20]
(N = 1: python tools/bench_allreduce.py).  Prints one JSON line (rank 0): ms per step (max over ranks, CUDA
events), NVLink bus GB/s = 2 (N-1)/N * bytes / time, and the update's HBM GB/s at N = 1."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from elasticdl_b200.elasticai_api.pytorch.optimizer import DistributedOptimizer  # noqa: E402
from mgpu_allreduce_check import resnet50_like_shapes  # noqa: E402


def timed(fn, steps, warmup, dev, world):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize(dev)
    ms = a.elapsed_time(b) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--momentum", type=float, default=0.9)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    shapes = resnet50_like_shapes()
    out = {"workload": "ResNet-50-sized dense model, %d tensors" % len(shapes), "n_gpus": world, "steps": args.steps,
           "optimizer": "SGD lr 0.1 momentum %g" % args.momentum}
    for mode in ("nccl", "fused"):
        if mode == "fused" and world == 1:
            continue
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.randn(s, device=dev) * 0.01) for s in shapes]
        opt = DistributedOptimizer(torch.optim.SGD(params, lr=0.1, momentum=args.momentum),
                                   named_parameters=[("p%d" % i, p) for i, p in enumerate(params)], fused=(mode == "fused"))
        flat = next(iter(opt._buckets.values()))
        gen = torch.Generator(device=dev).manual_seed(1 + rank)
        flat.normal_(0, 1e-3, generator=gen)  # synthetic gradients N(0, 1e-3), seed = rank
        n = sum(p.numel() for p in params)

        def step():
            opt.synchronize()
            with opt.skip_synchronize():
                opt.step()

        ms = timed(step, args.steps, args.warmup, dev, world)
        nbytes = n * 4
        out[mode] = {"ms_per_step": ms,
                     "bus_gbs": (2 * (world - 1) / world) * nbytes / (ms * 1e-3) / 1e9 if world > 1 else None,
                     "update_hbm_gbs": nbytes * 5 / (ms * 1e-3) / 1e9 if world == 1 else None}
        del opt, params
        torch.cuda.empty_cache()
    if "fused" in out:
        out["fused_speedup_vs_nccl"] = out["nccl"]["ms_per_step"] / out["fused"]["ms_per_step"]
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
