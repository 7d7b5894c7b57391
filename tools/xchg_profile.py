"""Per-kernel timing of the owner-computes exchange on a DeepFM-shaped batch (torchrun)."""
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elasticdl_b200._lib import check  # noqa: E402
from elasticdl_b200.ps import PSGroup  # noqa: E402
from elasticdl_b200.workloads.deepfm import DeepFMPSEngine, synthetic_batch  # noqa: E402

ADAM = ("Adam", "learning_rate=0.001;beta_1=0.9;beta_2=0.999;epsilon=1e-07;amsgrad=false;")


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    group = PSGroup(world, *ADAM, device=local, local_shards=[rank])
    B = 32768
    eng = DeepFMPSEngine(group, B, exchange="owner")
    ids, dense, labels = synthetic_batch(B, 1234 + rank, dev, "zipf")
    for _ in range(3):
        eng.step(ids, dense, labels)
    torch.cuda.synchronize()
    dist.barrier()
    ms = (ctypes.c_float * 6)()
    acc = [0.0] * 6
    n = 10
    for _ in range(n):
        group.push_begin(0.001, [0] * world)
        torch.cuda.synchronize()
        dist.barrier()
        check(group.lib.b200ps_xchg_profile(group._h, eng.uniq.data_ptr(), eng.n_unique.data_ptr(), eng.bet_d.data_ptr(),
                                            eng.bet_w.data_ptr(), eng.gsum_d.data_ptr(), eng.gsum_w.data_ptr(), ms,
                                            group._stream()))
        group.push_end(sync=True)
        for i in range(6):
            acc[i] += ms[i] * 1e3 / n
    group.check()
    names = ["post", "serve", "unscatter", "send_upd", "apply", "wait_applied"]
    out = {k: round(v, 1) for k, v in zip(names, acc)}
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        print(json.dumps({"world": world, "unique_per_rank": int(eng.n_unique.sum().item()), "us_per_kernel_by_rank": gathered}))
    dist.barrier()
    group.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
