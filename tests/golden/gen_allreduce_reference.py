"""Golden trajectories of the reference's allreduce optimizer, produced by EXECUTING
/root/reference/elasticai_api/pytorch/optimizer.py (class _DistributedOptimizer + DistributedOptimizer, unmodified) in
this container at world sizes 1 and 2:  python tests/golden/gen_allreduce_reference.py
Writes tests/golden/ref_allreduce_vectors.json (committed; replayed by tests/test_cpu_allreduce_reference_golden.py
against elasticdl_b200.elasticai_api.pytorch.DistributedOptimizer over gloo).

What is stubbed, and how: Horovod is not installed.  The file imports four names from it --
`horovod.torch.mpi_ops.{Average, allreduce_async_, size, synchronize}` and `horovod.torch.compression.Compression`.
Stand-ins with Horovod's documented semantics run over torch.distributed (gloo):
    allreduce_async_(t, name, op, prescale_factor, postscale_factor):  t <- post * sum_r(pre * t_r), where for
        op == Average the backend divides the postscale factor by size() (the reference's own comment says so,
        optimizer.py:143-157) -- in place, returns a handle;   synchronize(handle) -> the reduced tensor;
    size() = the world size;   Compression.none = identity.
Everything else -- hooks, accumulation over backward_passes_per_step, the fixed-global-batch logic of step() /
zero_grad(), the prescale / postscale composition -- is the reference's own code.  (The Horovod arithmetic itself stays
"unpinned", SURVEY 8c: no multi-rank numbers exist in the reference.)
"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))

SCENARIOS = [  # name, optimizer kwargs, wrapper kwargs as a function of world, micro-batches per rank
    ("sgd_momentum_average", {"lr": 0.1, "momentum": 0.9}, {}, 3),
    ("accumulate_2_passes", {"lr": 0.05, "momentum": 0.0}, {"backward_passes_per_step": 2}, 4),
    ("fixed_global_batch_4", {"lr": 0.5, "momentum": 0.0},
     {"fixed_global_batch_size": True, "global_batch_num_per_step": 4, "backward_passes_per_step": "4//world"}, "8//world"),
    ("predivide_2", {"lr": 0.1, "momentum": 0.0}, {"gradient_predivide_factor": 2.0, "global_batch_num_per_step": 1}, 2),
]

WORKER = r'''
import json, os, sys, types
import torch, torch.distributed as dist
mode, out_path, root = sys.argv[1], sys.argv[2], sys.argv[3]
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
if mode == "reference":
    Average, Sum = "Average", "Sum"
    class _Handle(object):
        def __init__(self, t): self.t = t
    def allreduce_async_(tensor, name=None, op=Average, prescale_factor=1.0, postscale_factor=1.0):
        tensor.mul_(prescale_factor)
        dist.all_reduce(tensor)
        post = postscale_factor / world if op == Average else postscale_factor
        tensor.mul_(post)
        return _Handle(tensor)
    hvd = types.ModuleType("horovod"); hvt = types.ModuleType("horovod.torch")
    comp = types.ModuleType("horovod.torch.compression"); ops = types.ModuleType("horovod.torch.mpi_ops")
    class Compression(object):
        class none(object):
            @staticmethod
            def compress(t): return t, None
            @staticmethod
            def decompress(t, ctx): return t
    comp.Compression = Compression
    ops.Average, ops.Sum, ops.allreduce_async_ = Average, Sum, allreduce_async_
    ops.size = lambda: world
    ops.synchronize = lambda h: h.t
    for n, m in (("horovod", hvd), ("horovod.torch", hvt), ("horovod.torch.compression", comp), ("horovod.torch.mpi_ops", ops)):
        sys.modules[n] = m
    sys.path.insert(0, "/root/reference")
    from elasticai_api.pytorch.optimizer import DistributedOptimizer
else:
    sys.path.insert(0, root)
    from elasticdl_b200.elasticai_api.pytorch.optimizer import DistributedOptimizer

scenarios = json.loads(sys.argv[4])
result = {}
for name, okw, wkw, n_micro in scenarios:
    wkw = {k: (eval(v, {"world": world}) if isinstance(v, str) else v) for k, v in wkw.items()}
    n_micro = eval(n_micro, {"world": world}) if isinstance(n_micro, str) else n_micro
    torch.manual_seed(11)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    if mode == "ours_q10":
        wkw = dict(wkw, reproduce_q10=True)
    opt = DistributedOptimizer(torch.optim.SGD(model.parameters(), **okw), named_parameters=model.named_parameters(), **wkw)
    bpps = wkw.get("backward_passes_per_step", 1)
    fixed = wkw.get("fixed_global_batch_size", False)
    traj = []
    for i in range(n_micro):
        g = torch.Generator().manual_seed(500 + 10 * i + rank)     # rank-dependent data
        x, y = torch.randn(4, 6, generator=g), torch.randn(4, 1, generator=g)
        if fixed or bpps == 1 or i % bpps == 0:
            opt.zero_grad()          # fixed mode: called every micro-batch, the optimizer decides (optimizer.py:254-263)
        ((model(x) - y) ** 2).mean().backward()
        if fixed or bpps == 1 or i % bpps == bpps - 1:
            opt.step()               # plain accumulation: step after the last pass (Horovod usage)
        traj.append([p.detach().reshape(-1).tolist() for p in model.parameters()])
    result[name] = traj
if rank == 0:
    json.dump(result, open(out_path, "w"))
dist.barrier()
'''


def run(mode, world, root):
    with tempfile.TemporaryDirectory() as tmp:
        script = os.path.join(tmp, "w.py")
        open(script, "w").write(WORKER)
        out = os.path.join(tmp, "out.json")
        port = str(32000 + os.getpid() % 2000 + world + {"reference": 7, "ours": 0, "ours_q10": 14}[mode])
        procs = []
        for r in range(world):
            env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=port,
                       CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
            procs.append(subprocess.Popen([sys.executable, script, mode, out, root, json.dumps(SCENARIOS)], env=env,
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        logs = [p.communicate(timeout=300)[0] for p in procs]
        if any(p.returncode for p in procs):
            raise RuntimeError("worker failed:\n" + "\n".join(l[-3000:] for l in logs))
        return json.load(open(out))


if __name__ == "__main__":
    root = os.path.dirname(os.path.dirname(HERE))
    out = {"scenarios": SCENARIOS, "world": {str(w): run("reference", w, root) for w in (1, 2)}}
    path = os.path.join(HERE, "ref_allreduce_vectors.json")
    json.dump(out, open(path, "w"))
    print("wrote", path, os.path.getsize(path), "bytes")
