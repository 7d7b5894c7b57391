#!/usr/bin/env python
"""bench.py -- DeepFM (dac_ctr) training throughput over the HBM parameter server.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the Go PS path

One step = one pass of the PS hot path over one synthetic Criteo-shaped batch
(BASELINE.json configs[1]): pull dense -> unique -> pull rows of 76 tables -> tower
fwd/bwd -> dedup-sum -> push with fused Adam -> version++.  Prints ONE JSON line.

`value`   : whole-job samples/s with inputs resident in HBM (CUDA events, max over ranks)
`e2e`     : same metric with the step's ids/features/labels copied from pinned host
            memory and the loss copied back every step, inside the timed region
`roofline`: the dominant PS kernel's algorithmic bytes / CUDA-event duration vs the
            measured HBM copy bandwidth (MEASURED_PEAKS.json)
`cpu_baseline`: the oracle's C restatement of the Go PS path on the host cores
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ADAM_ARGS = "learning_rate=0.001;beta_1=0.9;beta_2=0.999;epsilon=1e-07;amsgrad=false;"  # dac_ctr/elasticdl_train.py:47-48
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback


def ncu_traffic(kernel, world):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the named kernel from the committed
    `ncu --set full` capture of this workload at N=1 (profiles/ncu_traffic.json); None when the capture
    does not cover the configuration (a number taken under ncu is never measured live by the bench)."""
    if world != 1:
        return None
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")) as f:
            k = json.load(f)["kernels"][kernel]
        return k["dram_read_bytes"] + k["dram_write_bytes"]
    except (OSError, KeyError, ValueError):
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def _cpu_reference_once(T, batch, steps, warmup, dist_kind, seed=1234):
    import ctypes

    import numpy as np

    from elasticdl_b200.workloads.deepfm import DEEP_DIM, GROUP_ROWS, synthetic_batch
    from oracle import ps_oracle as O

    G = len(GROUP_ROWS)
    n_shards = 1
    vp = ctypes.c_void_p
    ids = np.stack([synthetic_batch(batch, seed + t, "cpu", dist_kind)[0].numpy() for t in range(T)])  # [T, G, B]
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    rng = np.random.RandomState(0)
    total = 0.0
    for dim in (1, DEEP_DIM):
        grads = (rng.randn(T, G, batch, dim) * 1e-3).astype(np.float32)
        tabs = [[O.OracleTable(dim, "zero") for _ in range(G * n_shards)] for _ in range(3)]
        arrs = [(vp * (G * n_shards))(*[t._h for t in fam]) for fam in tabs]
        rows = (ctypes.c_double * 2)()

        def run(n):
            return O.lib.oracle_bench_ps(T, G, n_shards, arrs[0], arrs[1], arrs[2], O._i64(ids.reshape(-1)),
                                         O._f32(grads.reshape(-1)), batch, dim, 1e-3, 0.9, 0.999, 1e-7, n, rows)

        if warmup:
            run(warmup)
        total += run(steps)
        del tabs
    return T * batch * steps / total


def cpu_reference(batch, steps, warmup, dist_kind, threads=None):
    """Times the oracle's C restatement of the Go PS path (oracle/ps_oracle.c
    oracle_bench_ps): T host threads, each a worker with its own batch, doing unique ->
    pull -> dedup -> scatter -> SparseAdam on in-process hash-map shards (RWMutex per table,
    go/pkg/common/embedding_table.go:22-58) for both table families (dim 1 and dim 8).
    gRPC/protobuf and the TF tower are NOT included, which flatters the reference.  The
    thread count is swept (lock contention on Zipf-hot rows can make fewer threads faster)
    and the best is reported.  Returns (samples_per_sec, threads, description)."""
    ncpu = os.cpu_count() or 1
    cands = [threads] if threads else sorted({min(c, ncpu) for c in (8, 16, 32, 64, ncpu)})
    best = (0.0, cands[0])
    if len(cands) > 1:
        for T in cands:
            sps = _cpu_reference_once(T, batch, 1, 1, dist_kind)
            if sps > best[0]:
                best = (sps, T)
    T = best[1]
    sps = _cpu_reference_once(T, batch, steps, warmup, dist_kind)
    return sps, T, ("C restatement of the Go PS path (unique+pull+dedup+SparseAdam, 76 tables, no gRPC/protobuf, "
                    "no tower): best of thread counts %s = %d threads x batch %d x %d steps, %s ids, %d host cores"
                    % (cands, T, batch, steps, dist_kind, ncpu))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32768, help="samples per GPU per step")
    ap.add_argument("--dist", default="zipf", choices=["zipf", "uniform"])
    ap.add_argument("--pool", type=int, default=4, help="distinct pre-generated batches cycled through")
    ap.add_argument("--cpu-batch", type=int, default=4096)
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tower", default="fused", choices=["fused", "mma", "torch"])
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of the captured CUDA graph")
    ap.add_argument("--profile-step", action="store_true",
                    help="run ONE eager step between cudaProfilerStart/Stop after the warm-up and exit "
                         "(for `ncu --profile-from-start off`; prints no bench line)")
    ap.add_argument("--lookahead", default="on", choices=["on", "off"],
                    help="deduplicate the ids of batch i+1 on a second stream while batch i trains (graph mode)")
    ap.add_argument("--paired", default="off", choices=["on", "off"],
                    help="store each group's deep+wide rows as one record per id")
    ap.add_argument("--exchange", default="auto", choices=["auto", "owner", "direct"],
                    help="multi-GPU row exchange: owner-computes bulk exchange (default for N>1) or direct peer access")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "DeepFM dac_ctr synthetic Criteo: 38 id groups x (wide dim1 + deep dim8) = 76 PS tables, "
                          "5549416 rows/family, Adam 1e-3, DNN[16,4]+FM tower",
              "batch_per_gpu": args.batch, "global_batch": args.batch * max(world, 1), "id_distribution": args.dist,
              "ps_shards": max(world, 1), "sharding": "id % N over GPUs (NVLink P2P)" if world > 1 else "1 shard",
              "step": "pull_dense+unique+pull(76 tables)+tower fwd/bwd+dedup-sum+push(dense+76 tables, Adam)+version++",
              "l2_policy": "tables (0.6 GB/family) and per-step id sets exceed L2; pool of %d batches cycled" % args.pool}

    if args.impl == "reference":
        if rank != 0:
            return
        sps, T, desc = cpu_reference(args.cpu_batch, max(args.steps, 1), min(args.warmup, 2), args.dist)
        line = {"impl": "reference", "metric": "deepfm_train_samples_per_sec", "value": sps, "unit": "samples/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * args.cpu_batch * T / sps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": T, "kind": "port", "sample": desc},
                "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)

    from elasticdl_b200 import _lib
    from elasticdl_b200.ps import PSGroup
    from elasticdl_b200.workloads.deepfm import DeepFMPSEngine, N_DENSE, synthetic_batch

    group = PSGroup(world, "Adam", ADAM_ARGS, device=local_rank,
                    local_shards=[rank] if world > 1 else None)
    engine = DeepFMPSEngine(group, args.batch, tower=args.tower, paired=args.paired == "on",
                            exchange=None if args.exchange == "auto" else args.exchange)
    config["record_layout"] = "paired deep+wide record per id" if engine.paired else "one record slab per table"
    config["exchange"] = engine.exchange
    use_graph = args.tower != "torch" and not args.no_graph
    lookahead = use_graph and args.lookahead == "on"
    config["pipeline"] = ("CUDA graph per step; id dedup of batch i+1 overlapped with step i on a second stream"
                          if lookahead else ("CUDA graph per step" if use_graph else "eager launches"))
    if world > 1:
        dist.barrier()
    B = args.batch
    G = engine.G

    # pre-generated batches: pinned host copies (e2e) and device copies (kernel-resident timing)
    host, devb = [], []
    for p in range(args.pool):
        ids, dense, labels = synthetic_batch(B, 1234 + p + 1000 * rank, dev, args.dist)
        devb.append((ids, dense, labels))
        host.append(tuple(t.cpu().pin_memory() for t in (ids, dense, labels)))
    uniq_per_batch = []
    for ids, _, _ in devb:
        _, _, n = group.unique(ids.view(-1), G)
        uniq_per_batch.append(int(n.sum().item()))
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    loss_pin = torch.empty(args.steps + args.warmup + 8, dtype=torch.float32).pin_memory()

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def my_launches():
        return group.launch_count + _lib.lib().b200ps_launch_count(None) + _lib.lib().b200_deepfm_launch_count() + _lib.lib().b200_deepfm_mma_launch_count()

    # ---- eager pass: every kernel launched from the host, CUDA-event pairs around the PS kernels ----
    for i in range(args.warmup):
        engine.step(*devb[i % args.pool])
    sync_all()
    if args.profile_step:
        torch.cuda.cudart().cudaProfilerStart()
        engine.step(*devb[args.warmup % args.pool])
        torch.cuda.synchronize(dev)
        torch.cuda.cudart().cudaProfilerStop()
        group.check()
        return
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = my_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ev = {}
    for i in range(args.steps):
        engine.step(*devb[i % args.pool], ev=ev)
    e1.record()
    sync_all()
    ms_eager = max_over_ranks(e0.elapsed_time(e1))
    launches = my_launches() - launches0
    group.check()
    ms = ms_eager
    if use_graph:
        # ---- headline: the same step replayed from a CUDA graph (inputs resident in HBM) ----
        if lookahead:
            # batch i trains while the ids of batch i+1 are deduplicated on a second stream
            engine.prepare(devb[0][0])
            engine.capture_ahead()

            def graph_step(i):
                _, dense_i, labels_i = devb[i % args.pool]
                return engine.step_ahead_graph(dense_i, labels_i, devb[(i + 1) % args.pool][0])
        else:
            engine.capture()

            def graph_step(i):
                return engine.step_graph(*devb[i % args.pool])
        for i in range(args.warmup):
            graph_step(i)
        sync_all()
        e0.record()
        for i in range(args.warmup, args.warmup + args.steps):
            graph_step(i)
        e1.record()
        sync_all()
        ms = max_over_ranks(e0.elapsed_time(e1))
        group.check()
    clocks = sampler.stop() if rank == 0 else None
    step_fn = engine.step_graph if use_graph else engine.step

    # ---- per-kernel durations from the CUDA events recorded inside the timed region ----
    kern = engine.kernel_report(ev, [uniq_per_batch[i % args.pool] for i in range(args.steps)])

    # ---- end to end: host buffers in, loss out, every step ----------------------------
    sync_all()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ahead = 2 if lookahead else 1  # batches in flight ahead of the running step
    feeder = engine.host_feeder(ahead + 1, lookahead=lookahead) if use_graph else None
    if feeder is not None and lookahead:
        engine.prepare(devb[0][0])  # graphs exist already; the feeder re-prepares on its first batch
        sync_all()
    e2.record()
    if feeder is not None:
        for j in range(min(ahead, args.steps)):
            feeder.submit(*host[j % args.pool])
    for i in range(args.steps):
        if feeder is not None:
            # every step's inputs come from pinned host memory; the copies of the next batches (side
            # stream) overlap the kernels of batch i, like the reference's dataset.prefetch(1)
            if i + ahead < args.steps:
                feeder.submit(*host[(i + ahead) % args.pool])
            loss = feeder.run_next()
        else:
            hi, hd, hl = host[i % args.pool]
            loss = engine.step(hi.to(dev, non_blocking=True), hd.to(dev, non_blocking=True),
                               hl.to(dev, non_blocking=True))
        loss_pin[i].copy_(loss, non_blocking=True)
    e3.record()
    sync_all()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3))
    group.check()

    if rank != 0:
        if world > 1:
            dist.barrier()
        return

    samples = args.steps * B * world
    value = samples / (ms * 1e-3)
    e2e_value = samples / (ms_e2e * 1e-3)
    peak, peak_kind = measured_peak()
    line = {"metric": "deepfm_train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "gpu_launches_per_step": launches / max(args.steps, 1),
            "launch_mode": "cuda_graph" if use_graph else "eager", "eager_ms_per_step": ms_eager / args.steps,
            "tower": args.tower,
            "final_loss": float(loss_pin[args.steps - 1])}
    if kern:
        top = max((k for k in kern if "gbs" in kern[k] and k.startswith(("pull", "push"))), key=lambda k: kern[k]["ms"])
        line["kernels"] = kern
        k = kern[top]
        line["roofline"] = {"kernel": top, "bound": "hbm", "achieved": k["gbs"], "peak": peak, "unit": "GB/s",
                            "frac": k["gbs"] / peak, "traffic": ncu_traffic(top, world) if (args.batch == 32768 and args.dist == "zipf") else None, "peak_kind": peak_kind,
                            "algorithmic_bytes_per_launch": k["bytes"], "us_per_launch": k["ms"] * 1e3}
        if "pull_deep" in kern:
            line["pull_gbs"] = kern["pull_deep"]["gbs"]
            line["pull_frac_of_peak"] = kern["pull_deep"]["gbs"] / peak
    line["unique_ids_per_step"] = statistics.mean(uniq_per_batch)
    if not args.no_cpu_baseline and world == 1:
        sps, T, desc = cpu_reference(args.cpu_batch, args.cpu_steps, 1, args.dist)
        line["cpu_baseline"] = {"value": sps, "unit": "samples/s", "cores": T, "kind": "port", "sample": desc}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()


if __name__ == "__main__":
    main()
