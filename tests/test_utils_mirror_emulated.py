"""The dict-shaped mirror API of the reference's lib/utils.js (node-cueball_b200/utils.py: planRebalance, shuffle,
currentMillis) on the CPU: the reference's own planRebalance known-answer tests (test/utils.test.js:13-285, recorded
under the reference engine) and utils.shuffle results go through the SAME Python as on the GPU, with the context's
kernels compiled for the host (tests/emu_context.py).  The GPU run of the same cases is tests/test_gpu_parity.py."""
import numpy as np
import pytest

import util
from emu_context import EmuContext
from oracle import oracle as O

PLAN = util.load_golden("plan_rebalance_kat.json")
SLOTS = util.load_golden("slots_golden.json")


@pytest.fixture(scope="module")
def ctx():
    return EmuContext()


@pytest.mark.parametrize("c", PLAN["kat"], ids=[c["test"] for c in PLAN["kat"]])
def test_plan_rebalance_reference_kat_through_the_mirror(ctx, c):
    import cueball_b200 as cb
    conns = dict((k, list(v)) for k, v in c["conns"])
    dead = dict((k, True) for k in c["dead"])
    plan = cb.planRebalance(conns, dead, c["target"], c["max"], c["singleton"], context=ctx)
    assert plan["add"] == c["add"] and plan["remove"] == c["remove"]


def test_plan_rebalance_random_reference_plans_through_the_mirror(ctx):
    import cueball_b200 as cb
    for c in PLAN["random"][::8]:
        conns = dict((k, list(v)) for k, v in c["conns"])
        plan = cb.planRebalance(conns, dict((k, True) for k in c["dead"]), c["target"], c["max"], c["singleton"], context=ctx)
        assert plan["add"] == c["add"] and plan["remove"] == c["remove"], c


def test_shuffle_through_the_mirror(ctx):
    """utils.shuffle (lib/utils.js:207-217) with injected draws: in place, returns the array, same permutation as
    the oracle's restatement (itself pinned to the reference engine's results, tests/test_oracle_golden.py)."""
    from cueball_b200 import utils as U
    rng = np.random.RandomState(11)
    for n in (0, 1, 2, 5, 64, 65, 200):
        arr = ["item%d" % i for i in range(n)]
        draws = rng.random_sample(max(n, 1)).tolist()
        want = O.shuffle(list(range(n)), draws) if n else []
        got = U.shuffle(arr, draws, context=ctx)
        assert got is arr and arr == ["item%d" % int(i) for i in want]


def test_current_millis_is_the_references_arithmetic(monkeypatch):
    """lib/utils.js:198-204: secs * 1000 + ns / 1e6, three roundings -- against the oracle's ns_to_ms."""
    import time

    from cueball_b200 import utils as U
    rng = np.random.RandomState(5)
    for t in [0, 1, 999_999_999, 10 ** 9, 86_400 * 100 * 10 ** 9 + 123_456_789] + [int(x) for x in rng.randint(0, 2 ** 62, 200)]:
        monkeypatch.setattr(time, "monotonic_ns", lambda t=t: t)
        assert np.float64(U.currentMillis()).tobytes() == np.float64(O.ns_to_ms(t)).tobytes()
