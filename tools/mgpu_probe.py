"""NVLink probe: scattered 32 B row READS vs WRITES against a peer GPU's shard (torchrun, 2+ ranks).
Decides whether the owner-computes exchange (only posted writes cross the link) is worth it."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elasticdl_b200.ps import PSGroup  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    group = PSGroup(world, "SGD", "learning_rate=0.1;momentum=0.0;nesterov=false;", device=local, local_shards=[rank])
    rows = 5_549_416
    group.register_table("t8", 8, "zero", rows)
    group.commit()
    dist.barrier()
    out = {}
    for U in (122_000, 1_000_000):
        peer = (rank + 1) % world
        slots = torch.randperm(rows // world - 1, device=dev)[:U]
        ids_remote = (slots * world + peer).contiguous()
        ids_local = (slots * world + rank).contiguous()
        buf = torch.randn((U, 8), device=dev)
        tid = group.tables["t8"][0]
        for name, ids in (("remote", ids_remote), ("local", ids_local)):
            arr, n = group.make_segs([(tid, U, ids, None, buf)])
            rd = timeit(lambda: group.lib.b200ps_pull_rows(group._h, arr, n, group._stream()))
            wr = timeit(lambda: group.lib.b200ps_set_rows(group._h, arr, n, group._stream()))
            out["%s_U%d" % (name, U)] = {"read_us": rd, "write_us": wr, "read_GBs": U * 32 / rd / 1e3, "write_GBs": U * 32 / wr / 1e3}
        torch.cuda.synchronize()
        dist.barrier()
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    group.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
