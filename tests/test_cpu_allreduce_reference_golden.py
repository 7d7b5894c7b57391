"""elasticdl_b200's DistributedOptimizer against trajectories produced by EXECUTING the reference's
elasticai_api/pytorch/optimizer.py (tests/golden/gen_allreduce_reference.py: the reference class over a Horovod
stand-in backed by gloo, world sizes 1 and 2).  Same model, same rank-dependent data, same call pattern
(zero_grad / backward / step per micro-batch): the parameters after every micro-batch must agree -- plain averaged
SGD with momentum, local accumulation over backward_passes_per_step, the fixed-global-batch mode (step() and
zero_grad() are no-ops until the last local pass; the average is over global_batch_num_per_step micro-batches whatever
the world size) and gradient_predivide_factor.  gloo, CPU, world_size 1 and 2."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import gen_allreduce_reference as G  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "ref_allreduce_vectors.json")))


def test_golden_file_matches_the_generator_scenarios():
    assert [s[0] for s in GOLD["scenarios"]] == [s[0] for s in G.SCENARIOS]
    assert json.loads(json.dumps(G.SCENARIOS)) == GOLD["scenarios"]


def _same(a, b):
    return all(np.allclose(pa, pb, rtol=1e-6, atol=1e-7) for sa, sb in zip(a, b) for pa, pb in zip(sa, sb))


@pytest.mark.parametrize("world", [1, 2])
def test_distributed_optimizer_follows_the_executed_reference(world):
    """reproduce_q10=True: every scenario equals the executed reference.  Default: the fixed-global-batch mode (the
    only one the reference itself runs, allreduce_trainer_test.py:182, torch_optimizer_test.py:28) equals it too; the
    plain Average mode returns the MEAN over ranks where the reference's literal postscale (optimizer.py:154,157:
    predivide * size()) returns the SUM -- quirk Q10, visible from world size 2 on."""
    want = GOLD["world"][str(world)]
    q10 = G.run("ours_q10", world, os.path.dirname(HERE))
    assert set(q10) == set(want)
    for name in want:
        assert len(q10[name]) == len(want[name]) and _same(q10[name], want[name]), name
    got = G.run("ours", world, os.path.dirname(HERE))
    assert _same(got["fixed_global_batch_4"], want["fixed_global_batch_4"])
    for name in ("sgd_momentum_average", "accumulate_2_passes", "predivide_2"):
        assert _same(got[name], want[name]) == (world == 1), name
    # the fixed-global-batch update is world-size invariant (optimizer.py:141-157): same parameters after the same
    # number of GLOBAL micro-batches at world 1 and world 2 is NOT expected here (the data is rank-dependent) -- what
    # is checked is that both implementations skip and apply updates at the same calls:
    fixed = want["fixed_global_batch_4"]
    changed = [i for i in range(1, len(fixed)) if fixed[i] != fixed[i - 1]]
    assert changed == [i for i in range(1, len(fixed)) if (i + 1) % (4 // world) == 0]
