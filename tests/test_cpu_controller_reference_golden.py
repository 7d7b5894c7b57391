"""The elastic allreduce controller of this repo against event logs produced by EXECUTING the reference's
elasticai_api/common/base_controller.py + elasticai_api/pytorch/controller.py (tests/golden/gen_controller_reference.py).
Same scripted master / Horovod-size-rank / failing-function scenarios; the log of decisions must be identical: when the
wrapped function is (re)run, when the master is asked for the rank, when the group is rebuilt (a new rendezvous id),
what is broadcast and in which order, which optimizer methods a failed step triggers (sleep, load_state_dict, zero_grad),
when a batch is reported done, how global_completed_batch_num / backward_passes_per_step / the epoch evolve, and the
error after five failed attempts.

Normalisation (the two designs differ there on purpose): the reference rebuilds Horovod with shutdown() + init() and
also calls a bare hvd.init() before the very first function call; here the group is a torch.distributed process group
rebuilt inside RendevousManager._restart.  Both become one "group.restart" event; the bare first init() is dropped.
What _restart does inside is covered by the gloo world-2 test (tests/test_cpu_allreduce_controller.py)."""
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import gen_controller_reference as G  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "ref_controller_vectors.json")))


def _normalise_reference(log):
    out, i = [], 0
    while i < len(log):
        if log[i] == ["group.shutdown"] and i + 1 < len(log) and log[i + 1] == ["group.init"]:
            out.append(["group.restart"])
            i += 2
        elif log[i] == ["group.init"]:
            i += 1  # hvd.init() of _init_variables_before_first_calling / init_horovod_locally
        else:
            out.append(log[i])
            i += 1
    return out


def _make_our_controller(master, shards, log, sc, monkeypatch):
    from elasticdl_b200.elasticai_api.common import base_controller as bc
    from elasticdl_b200.elasticai_api.pytorch import controller as ctl

    monkeypatch.setenv("WORKER_NUM", str(sc["worker_num"]))
    clock = G.FakeClock(log)
    monkeypatch.setattr(bc, "time", clock)
    monkeypatch.setattr(ctl, "time", clock)
    monkeypatch.setattr(bc, "RETRY_ALLREDUCE_INTERVAL_SECS", 30)
    monkeypatch.setattr(ctl.traceback, "print_exc", lambda: None)
    monkeypatch.setattr(ctl, "comm_size", lambda: sc["size"])
    monkeypatch.setattr(ctl, "comm_rank", lambda: sc["rank"])
    monkeypatch.setattr(ctl, "broadcast_parameters", lambda sd, root_rank=0: log.append(["broadcast_parameters", root_rank]))
    monkeypatch.setattr(ctl, "broadcast_optimizer_state",
                        lambda o, root_rank=0: log.append(["broadcast_optimizer_state", root_rank]))

    def broadcast_object(obj, root_rank=0, name=None):
        log.append(["broadcast_object", name])
        return obj

    monkeypatch.setattr(ctl, "broadcast_object", broadcast_object)

    def restart(self, r):
        log.append(["group.restart"])
        self._rendezvous_id = r.rendezvous_id
        self.need_broadcast = True

    monkeypatch.setattr(bc.RendevousManager, "_restart", restart)
    return ctl.PyTorchAllReduceController(master, shards, backend="gloo")


def test_golden_file_matches_the_generator_scenarios():
    assert GOLD["scenarios"] == json.loads(json.dumps(G.SCENARIOS))


@pytest.mark.parametrize("sc", G.SCENARIOS, ids=[s["name"] for s in G.SCENARIOS])
def test_controller_makes_the_decisions_of_the_executed_reference(sc, monkeypatch):
    want = GOLD["runs"][sc["name"]]
    got = G.run_scenario(lambda m, s, log, sc_: _make_our_controller(m, s, log, sc_, monkeypatch), sc)
    assert got["error"] == want["error"]
    assert got["states"] == want["states"]
    assert got["log"] == _normalise_reference(want["log"]), "\n".join(
        "%-55s %s" % (a, b) for a, b in zip(got["log"] + [None] * 50, _normalise_reference(want["log"]) + [None] * 50)
        if a is not None or b is not None)
