"""Dense allreduce controller on GPUs (BASELINE.json configs[3]: ResNet-50-sized gradients, 214
tensors, 25.6 M fp32), launched with torchrun, one rank per GPU, NCCL:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29521 tests/mgpu_allreduce_check.py

Checks: DistributedOptimizer's averaged gradient == mean over ranks (the invariant the reference's
mocked tests cannot pin), broadcast from rank 0, fixed-global-batch world-size invariance; then times
the one-bucket all-reduce (bus GB/s) and the fused momentum update kernel (HBM GB/s)."""
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from elasticdl_b200 import _lib  # noqa: E402
from elasticdl_b200.elasticai_api.pytorch.controller import create_elastic_controller  # noqa: E402
from elasticdl_b200.elasticai_api.pytorch.optimizer import DistributedOptimizer  # noqa: E402


def resnet50_like_shapes():
    """214 tensors totalling 25.6 M parameters (docs/benchmark/ftlib_benchmark.md:40-41,122-123)."""
    shapes = []
    for i in range(53):  # conv-like blocks: kernel + 3 per-channel vectors
        c = [64, 128, 256, 512][i % 4]
        shapes += [(c, c // 2, 3, 3), (c,), (c,), (c,)]
    shapes.append((1000,))
    total = sum(int(torch.Size(s).numel()) for s in shapes)
    assert len(shapes) == 213 and total < 25_600_000
    shapes.append((25_600_000 - total,))
    return shapes


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    # B200_SHARED_GPU=1: all ranks are processes on cuda:0, collectives over gloo (see tests/mgpu_check.py)
    shared = os.environ.get("B200_SHARED_GPU") == "1"
    local = 0 if shared else int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    backend = "gloo" if shared else "nccl"
    controller = create_elastic_controller(batch_size=32, num_epochs=1, dataset_size=3200, backend=backend)
    assert dist.get_backend() == backend and dist.get_world_size() == world
    # ---- correctness on a small model -------------------------------------------------
    torch.manual_seed(10 + rank)
    model = torch.nn.Sequential(torch.nn.Linear(32, 64), torch.nn.ReLU(), torch.nn.Linear(64, 1)).to(dev)
    opt = DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.1), named_parameters=model.named_parameters())
    controller.set_broadcast_model(model)
    controller.set_broadcast_optimizer(opt)
    controller.broadcast()
    ref = [p.detach().clone() for p in model.parameters()]
    for p in ref:
        q = p.clone()
        dist.broadcast(q, src=0)
        assert torch.equal(p, q), "broadcast left ranks different"
    g = torch.Generator(device=dev).manual_seed(500 + rank)
    x, y = torch.randn(16, 32, device=dev, generator=g), torch.randn(16, 1, device=dev, generator=g)
    opt.zero_grad()
    ((model(x) - y) ** 2).mean().backward()
    local_grads = [p.grad.detach().clone() for p in model.parameters()]
    opt.step()
    for p0, p1, gl in zip(ref, model.parameters(), local_grads):
        mean = gl.clone()
        dist.all_reduce(mean)
        mean /= world
        assert torch.allclose(p1.detach(), p0 - 0.1 * mean, atol=1e-6), "averaged gradient != mean over ranks"
    # ---- fused mode: reduce-scatter + averaging + momentum update in one kernel over peer memory ----------
    import numpy as np

    from oracle import ps_oracle as O

    torch.manual_seed(77)  # same initial model on every rank
    fmodel = torch.nn.Sequential(torch.nn.Linear(40, 33), torch.nn.Tanh(), torch.nn.Linear(33, 3)).to(dev)
    fopt = DistributedOptimizer(torch.optim.SGD(fmodel.parameters(), lr=0.05, momentum=0.9, nesterov=True),
                                named_parameters=fmodel.named_parameters(), fused=True)
    assert fopt._ps is not None, "fused backend was not selected"
    want_p = [p.detach().cpu().numpy().copy() for p in fmodel.parameters()]
    want_v = [np.zeros_like(w) for w in want_p]
    for step in range(3):
        gx = torch.Generator(device=dev).manual_seed(900 + 10 * step + rank)
        xb, yb = torch.randn(8, 40, device=dev, generator=gx), torch.randn(8, 3, device=dev, generator=gx)
        fopt.zero_grad()
        ((fmodel(xb) - yb) ** 2).mean().backward()
        mine = [p.grad.detach().clone() for p in fmodel.parameters()]
        fopt.step()
        for i, gl in enumerate(mine):
            allg = [torch.empty_like(gl) for _ in range(world)]
            dist.all_gather(allg, gl)
            acc = allg[0].cpu().numpy().astype(np.float32).copy()
            for r in range(1, world):  # the kernel sums the ranks in rank order
                acc = (acc + allg[r].cpu().numpy()).astype(np.float32)
            g_avg = (acc * np.float32(1.0 / world)).astype(np.float32)
            O.np_momentum(g_avg.reshape(-1), want_p[i].reshape(-1), want_v[i].reshape(-1), 0.9, True, 0.05)
        for w, p in zip(want_p, fmodel.parameters()):
            assert np.array_equal(w, p.detach().cpu().numpy()), "fused reduce+update differs from the oracle (step %d)" % step
    fopt._ps.group.check()
    # ---- config 4 timing: one-bucket allreduce of ResNet-50-sized gradients + fused momentum update ----
    shapes = resnet50_like_shapes()
    params = [torch.nn.Parameter(torch.zeros(s, device=dev)) for s in shapes]
    big = DistributedOptimizer(torch.optim.SGD(params, lr=0.1, momentum=0.9),
                               named_parameters=[("p%d" % i, p) for i, p in enumerate(params)])
    flat = next(iter(big._buckets.values()))
    flat.normal_(0, 1e-3)
    n = flat.numel()
    for _ in range(1 if shared else 3):
        big.synchronize()
    torch.cuda.synchronize()
    dist.barrier()
    evs = []
    for _ in range(2 if shared else 10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        big.synchronize()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2]
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_ar = float(t.item())
    # fused momentum update on the flat parameter (kernel_api.h Momentum on device arrays)
    lib = _lib.lib()
    p = torch.zeros(n, device=dev)
    vel = torch.zeros(n, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        _lib.check(lib.b200ps_kernel_momentum(flat.data_ptr(), p.data_ptr(), vel.data_ptr(), 0.9, 0, 0.1, n, st))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        _lib.check(lib.b200ps_kernel_momentum(flat.data_ptr(), p.data_ptr(), vel.data_ptr(), 0.9, 0, 0.1, n, st))
    b.record()
    torch.cuda.synchronize()
    ms_upd = a.elapsed_time(b) / 10
    dist.barrier()
    if rank == 0:
        nbytes = n * 4
        print(json.dumps({"check": "mgpu_allreduce ok", "world": world, "tensors": len(shapes), "numel": n,
                          "allreduce_ms": ms_ar,
                          "bus_gbs": (2 * (world - 1) / world) * nbytes / (ms_ar * 1e-3) / 1e9 if world > 1 else None,
                          "fused_momentum_update_ms": ms_upd, "update_hbm_gbs": n * 4 * 5 / (ms_upd * 1e-3) / 1e9}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
