"""Does concurrent H2D traffic slow the exchange's system-scope fences?  (torchrun, one rank per GPU)
Times the captured training step (a) quiet and (b) with a side stream copying pinned host batches
continuously, for the engine's exchange mode given by argv[1] (owner|direct)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "owner"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from elasticdl_b200.ps import PSGroup
    from elasticdl_b200.workloads.deepfm import DeepFMPSEngine, synthetic_batch

    group = PSGroup(world, "Adam", "learning_rate=0.001;beta_1=0.9;beta_2=0.999;epsilon=1e-07;amsgrad=False;",
                    device=local, local_shards=[rank])
    B = 32768
    eng = DeepFMPSEngine(group, B, exchange=mode)
    batches = [synthetic_batch(B, 1234 + p + 1000 * rank, dev, "zipf") for p in range(4)]
    for b in batches[:3]:
        eng.step(*b)
    eng.capture()
    host = torch.empty(12 << 20, dtype=torch.uint8).pin_memory()
    sink = torch.empty(12 << 20, dtype=torch.uint8, device=dev)
    side = torch.cuda.Stream(device=dev)
    out = {"mode": mode, "world": world}
    for label, dma in (("quiet", False), ("h2d", True), ("quiet2", False)):
        for i in range(5):
            eng.step_graph(*batches[i % 4])
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 100
        if dma:
            with torch.cuda.stream(side):
                for _ in range(400):  # ~12 MB each: far more than the timed steps need
                    sink.copy_(host, non_blocking=True)
        e0.record()
        for i in range(steps):
            eng.step_graph(*batches[i % 4])
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[label + "_us_per_step"] = round(float(t.item()), 1)
    group.check()
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    group.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
