"""Multi-process / multi-GPU parity check (launched with torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/mgpu_check.py

Rank r owns PS shard r; every rank is also a worker.  Checks, against the CPU oracle:
  1. pulls see rows written by OTHER ranks through NVLink peer mappings (bit-exact),
  2. pushes from all ranks with disjoint id sets == the oracle applying them in any order,
  3. per-shard step/version counters advance once per push per rank (quirks Q2/Q7),
  4. dense parameters: first-writer-wins init, pull from the owner shard on every rank.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from elasticdl_b200.common.tensor_utils import EmbeddingTableInfo, Tensor  # noqa: E402
from elasticdl_b200.ps import PSGroup  # noqa: E402
from elasticdl_b200.worker.ps_client import PSClient  # noqa: E402
from oracle import ps_oracle as O  # noqa: E402

F = np.float32
ADAM = ("Adam", "learning_rate=0.001;beta_1=0.9;beta_2=0.999;epsilon=1e-07;amsgrad=false;")


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    # B200_SHARED_GPU=1: every rank is a process on cuda:0 (the shards are still separate allocations mapped
    # across PROCESSES with CUDA IPC, the kernels of the ranks time-slice the one device) and the host-side
    # plumbing runs over gloo -- NCCL refuses two ranks on one device.  Lets a 1-GPU box run this check.
    shared = os.environ.get("B200_SHARED_GPU") == "1"
    local = 0 if shared else int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if shared:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)
    group = PSGroup(world, *ADAM, device=local, local_shards=[rank])
    client = PSClient(group)
    cap, dim = 100000, 8
    client.push_embedding_table_infos([EmbeddingTableInfo("deep", dim, "zero", 1, cap),
                                       EmbeddingTableInfo("wide", 1, "zero", 1, cap)])
    shapes = {"w": (64, 8), "b": (7,)}
    client.partition_dense_parameters(shapes.keys(), shapes=shapes)
    oc = O.OraclePSClient([O.OracleServer(i, *ADAM, num_ps=world) for i in range(world)])
    oc.push_embedding_table_infos([O.EmbeddingTableInfo("deep", dim, "zero", 1), O.EmbeddingTableInfo("wide", 1, "zero", 1)])
    oc.partition_dense_parameters(shapes.keys())

    rng = np.random.RandomState(42)  # same stream on every rank
    ids = rng.permutation(cap)[:20000].astype(np.int64)
    vals = rng.randn(20000, dim).astype(F)
    # each rank writes a disjoint quarter of the rows -- to whichever shard owns them (mostly peers)
    mine = np.arange(20000) % world == rank
    group.set_rows([("deep", ids[mine], vals[mine])])
    for s in oc.servers:
        m = ids % world == s.id
        s.tables["deep"].set(ids[m], vals[m])
    torch.cuda.synchronize()
    dist.barrier()
    # 1. every rank pulls everything
    q = rng.choice(ids, 30000).astype(np.int64)
    got = client.pull_embedding_vectors("deep", q)
    assert np.array_equal(got, oc.pull_embedding_vectors("deep", q)), "peer pull not bit-exact"
    dist.barrier()
    # 4. dense init: every rank tries, exactly one wins per shard
    dense = {"w": rng.randn(64, 8).astype(F), "b": rng.randn(7).astype(F)}
    versions = [-1] * world
    params, uninit = client.pull_dense_parameters(list(range(world)), versions)
    for ps_id in uninit:
        # ranks push DIFFERENT values; the first writer's stay.  rank 0's copy is the expected one
        # only if it wins, so make all pushes identical except a rank marker we do not compare.
        client.push_dense_parameters([Tensor(n, v, None) for n, v in dense.items()], ps_id, 0)
    torch.cuda.synchronize()
    dist.barrier()
    params, uninit = client.pull_dense_parameters(list(range(world)), versions)
    assert uninit == [] and all(np.array_equal(params[n], dense[n]) for n in dense)
    for ps_id in set(oc.parameter_to_ps.values()):
        oc.push_dense_parameters([O.Tensor(n, v.copy(), None) for n, v in dense.items()], ps_id, 0)
    dist.barrier()
    # 2./3. three rounds of pushes; rank r pushes ids == r (mod world) of a fresh permutation, so rows
    # are disjoint across ranks (deterministic) but spread over all shards
    for rnd in range(3):
        pid = rng.permutation(cap)[:8000].astype(np.int64)
        g8 = rng.randn(8000, dim).astype(F)
        g1 = rng.randn(8000, 1).astype(F)
        gd = {n: rng.randn(*v.shape).astype(F) for n, v in dense.items()}
        for r in range(world):  # oracle: ranks applied in rank order
            sel = np.arange(8000) % world == r
            dg = [O.Tensor(n, gd[n] * (r + 1), None) for n in dense]
            oc.push_gradients(dg, [O.Tensor("deep", g8[sel].copy(), pid[sel]), O.Tensor("wide", g1[sel].copy(), pid[sel])],
                              0.001, [0] * world)
        # GPU: ranks push one after another (dense params are shared by all ranks, so order matters)
        for r in range(world):
            if r == rank:
                sel = np.arange(8000) % world == r
                dg = [Tensor(n, gd[n] * (r + 1), None) for n in dense]
                acc, ver = client.push_gradients(dg, [Tensor("deep", g8[sel], pid[sel]), Tensor("wide", g1[sel], pid[sel])],
                                                 0.001, [0] * world)
                assert acc
            torch.cuda.synchronize()
            dist.barrier()
    state = group.snapshot()
    assert [s[0] for s in state] == [3 * world] * world, state   # version: one per push per rank
    assert [s[1] for s in state] == [3 * world] * world, state   # step likewise
    allids = np.arange(cap, dtype=np.int64)
    for name in ("deep", "wide"):
        got = client.pull_embedding_vectors(name, allids)
        want = np.zeros_like(got)
        for s in oc.servers:
            keys = s.tables[name].keys()
            want[keys] = s.tables[name].get(keys)
        assert np.allclose(got, want, rtol=1e-5, atol=1e-7), name
        assert np.array_equal(got, want), name + " (bit-exact expected: no duplicate sums)"
    params, _ = client.pull_dense_parameters(list(range(world)), versions)
    for n in dense:
        assert np.array_equal(params[n], oc.servers[oc.parameter_to_ps[n]].dense[n]), n
    dist.barrier()
    # concurrent pushes from all ranks on disjoint rows (no barrier between ranks): still exact
    pid = rng.permutation(cap)[:8000].astype(np.int64)
    g8 = rng.randn(8000, dim).astype(F)
    sel = np.arange(8000) % world == rank
    before = client.pull_embedding_vectors("deep", pid)
    torch.cuda.synchronize(); dist.barrier()
    client.push_gradients([], [Tensor("deep", g8[sel], pid[sel])], 0.001, [0] * world)
    torch.cuda.synchronize(); dist.barrier()
    after = client.pull_embedding_vectors("deep", pid)
    assert not np.array_equal(before, after)
    assert [s[0] for s in group.snapshot()] == [4 * world] * world
    group.check()
    dist.barrier()
    if rank == 0:
        print("mgpu_check ok: world=%d peer pulls bit-exact, pushes bit-exact vs oracle, counters %s" % (world, state))
    group.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
