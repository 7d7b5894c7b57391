"""Event logs of the reference's elastic allreduce controller, produced by EXECUTING
/root/reference/elasticai_api/common/base_controller.py and elasticai_api/pytorch/controller.py (unmodified) in this
container:  python tests/golden/gen_controller_reference.py  ->  tests/golden/ref_controller_vectors.json (committed;
replayed by tests/test_cpu_controller_reference_golden.py against elasticdl_b200.elasticai_api).

Stand-ins (Horovod, the ElasticDL master and its generated protos are not installed): `horovod.torch` (init / shutdown /
size / rank scripted by the scenario), `horovod.torch.functions.broadcast_*` (recorded), `HorovodInternalError`, a
scripted master client (`get_comm_rank`, `report_training_loop_status`), a counting data-shard service, and the modules'
`time` (a clock that advances 40 s per reading, so that the periodic rendezvous check runs on every call, and a
recording `sleep`) and `socket.gethostbyname`.  Everything the log records is decided by the reference's own code:
when the function is (re)run, when the group is rebuilt, what is broadcast, which optimizer methods are called on a
failure, how `global_completed_batch_num` and `backward_passes_per_step` evolve.
"""
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))

# scenario: world size / rank as Horovod reports them, rendezvous ids the master hands out per get_comm_rank() call,
# calls of the wrapped function that raise (counted over all invocations, 0-based), optimizer mode
SCENARIOS = [
    {"name": "fixed_batch_w2_rank0", "worker_num": 5, "size": 2, "rank": 0, "rdzv": [1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2],
     "fail_at": [3], "fixed": True, "calls": 5},
    {"name": "fixed_batch_w2_rank1", "worker_num": 5, "size": 2, "rank": 1, "rdzv": [1] * 12, "fail_at": [], "fixed": True,
     "calls": 4},
    {"name": "plain_w3", "worker_num": 3, "size": 3, "rank": 2, "rdzv": [7] * 12, "fail_at": [1, 2], "fixed": False, "calls": 3},
    {"name": "gives_up_after_5", "worker_num": 1, "size": 1, "rank": 0, "rdzv": [1] * 20, "fail_at": list(range(1, 30)),
     "fixed": False, "calls": 1},
]


class FakeOptimizer(object):
    """What the controller touches of a DistributedOptimizer (optimizer.py): the fixed-global-batch flags and the
    three methods restore() / reset_backward_passes_per_step() call.  update_gradients follows step()'s rule."""

    def __init__(self, log, fixed):
        self.log = log
        if fixed:
            self.fixed_global_batch_size = True
        self.backward_passes_per_step = 1
        self.update_gradients = True
        self._passes = 0

    def set_backward_passes_per_step(self, n):
        self.log.append(["optimizer.set_backward_passes_per_step", n])
        self.backward_passes_per_step = n

    def state_dict(self):
        return {"state": 1}

    def load_state_dict(self, sd):
        self.log.append(["optimizer.load_state_dict"])

    def zero_grad(self):
        self.log.append(["optimizer.zero_grad"])

    def step(self):  # optimizer.py:227-239
        self._passes += 1
        if getattr(self, "fixed_global_batch_size", False) and self._passes % self.backward_passes_per_step != 0:
            self.update_gradients = False
        else:
            self.update_gradients = True
            self._passes = 0


class FakeClock(object):
    def __init__(self, log):
        self.t, self.log = 1000.0, log

    def time(self):
        self.t += 40.0
        return self.t

    def sleep(self, secs):
        self.log.append(["sleep", secs])


def run_scenario(make_controller, sc):
    """Drives a controller (the reference's or this repo's) through a scenario; returns log + states."""
    log = []
    opt = FakeOptimizer(log, sc["fixed"])
    state = {"n_func": 0, "n_rank": 0}

    class Master(object):
        def get_comm_rank(self_):
            i = min(state["n_rank"], len(sc["rdzv"]) - 1)
            state["n_rank"] += 1
            log.append(["master.get_comm_rank"])
            return types.SimpleNamespace(rank_id=sc["rank"], world_size=sc["size"], rendezvous_id=sc["rdzv"][i],
                                         rendezvous_port=1234)

        def report_training_loop_status(self_, status):
            log.append(["master.report_training_loop_status", int(status)])

    class Shards(object):
        def get_minibatch_count_per_epoch(self_):
            return 10

        def report_batch_done(self_):
            log.append(["report_batch_done"])

    controller = make_controller(Master(), Shards(), log, sc)
    controller.set_broadcast_model(types.SimpleNamespace(state_dict=lambda: {"w": 0}))
    controller.set_broadcast_optimizer(opt)

    def train_one_batch(tag):
        i = state["n_func"]
        state["n_func"] += 1
        log.append(["func", tag, i])
        if i in sc["fail_at"]:
            raise RuntimeError("injected failure %d" % i)
        opt.step()
        return "loss-%s" % tag

    elastic = controller.elastic_run(train_one_batch)
    states, error = [], None
    try:
        with controller.scope():
            for c in range(sc["calls"]):
                r = elastic(c)
                states.append({"result": r, "global_completed_batch_num": controller.global_completed_batch_num,
                               "backward_passes_per_step": controller.backward_passes_per_step,
                               "epoch": controller.get_current_epoch()})
    except RuntimeError as err:
        error = str(err)
    return {"log": log, "states": states, "error": error}


def make_reference_controller(master, shards, log, sc):
    os.environ["USE_TORCH"] = "1"
    os.environ["WORKER_NUM"] = str(sc["worker_num"])
    hvd = types.ModuleType("horovod.torch")
    hvd.init = lambda: log.append(["group.init"])
    hvd.shutdown = lambda: log.append(["group.shutdown"])
    hvd.size = lambda: sc["size"]
    hvd.rank = lambda: sc["rank"]
    fn = types.ModuleType("horovod.torch.functions")
    fn.broadcast_parameters = lambda sd, root_rank=0: log.append(["broadcast_parameters", root_rank])
    fn.broadcast_optimizer_state = lambda o, root_rank=0: log.append(["broadcast_optimizer_state", root_rank])

    def broadcast_object(obj, root_rank=0, name=None):
        log.append(["broadcast_object", name])
        return obj

    fn.broadcast_object = broadcast_object
    exc = types.ModuleType("horovod.common.exceptions")
    exc.HorovodInternalError = type("HorovodInternalError", (Exception,), {})
    mods = {"horovod": types.ModuleType("horovod"), "horovod.torch": hvd, "horovod.torch.functions": fn,
            "horovod.common": types.ModuleType("horovod.common"), "horovod.common.exceptions": exc,
            "elasticai_api.common.data_shard_service": types.ModuleType("elasticai_api.common.data_shard_service"),
            "elasticai_api.common.master_client": types.ModuleType("elasticai_api.common.master_client")}
    mods["elasticai_api.common.data_shard_service"].RecordIndexService = object
    mods["elasticai_api.common.master_client"].build_master_client = lambda: None
    sys.modules.update(mods)
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    for m in ("elasticai_api.pytorch.controller", "elasticai_api.common.base_controller"):
        sys.modules.pop(m, None)
    import importlib

    base = importlib.import_module("elasticai_api.common.base_controller")
    ctl = importlib.import_module("elasticai_api.pytorch.controller")
    clock = FakeClock(log)
    base.time = clock
    ctl.time = clock
    base.socket = types.SimpleNamespace(gethostbyname=lambda h: "127.0.0.1", gethostname=lambda: "localhost")
    ctl.traceback = types.SimpleNamespace(print_exc=lambda: None)
    return ctl.PyTorchAllReduceController(master, shards)


if __name__ == "__main__":
    out = {"scenarios": SCENARIOS, "runs": {sc["name"]: run_scenario(make_reference_controller, sc) for sc in SCENARIOS}}
    path = os.path.join(HERE, "ref_controller_vectors.json")
    json.dump(out, open(path, "w"), indent=0)
    print("wrote", path, os.path.getsize(path), "bytes")
    for name, r in out["runs"].items():
        print(name, r["error"], [s["global_completed_batch_num"] for s in r["states"]], len(r["log"]))
