"""tools/proto_queue.py (the executable specification the cb_queue_* kernels follow: elementwise
enqueue, bisected expiry, closed-form idle event) against the oracle's event-driven pool model,
which is pinned to the reference's own JavaScript (tests/test_pool_sim_oracle.py)."""
import os
import sys

import numpy as np
import pytest

import util
from oracle import oracle as O

sys.path.insert(0, os.path.join(util.ROOT, "tools"))
import proto_queue  # noqa: E402

G = util.load_golden("pool_sim_golden.json")


def check(target, nslots, ty, a, b, idle0, state=None):
    cd = O.new_codel(target, state)
    res, tmo, out = O.pool_sim(cd, nslots, ty, a, b, idle0=idle0)
    q = proto_queue.DeviceQueueSpec(target, nslots, idle0=idle0, state=state)
    got = q.run(ty, a, b)
    assert got.tolist() == res.tolist()
    assert q.counters == out.counters()
    assert (q.idle, q.cursor, q.tail - q.head) == (out.idle, out.cursor, out.queued)
    util.assert_state_equal(q.state(), cd.as_dict())
    return res


@pytest.mark.parametrize("case", G["cases"], ids=[c["name"] for c in G["cases"]])
def test_spec_on_golden_scripts(case):
    check(case["target"], case["nslots"], case["ev_type"], case["ev_a"], case["ev_b"], case["idle0"])


@pytest.mark.parametrize("seed", range(24))
def test_spec_on_random_scripts(seed):
    rng = np.random.RandomState(7000 + seed)
    tgt = int(rng.choice([1, 2, 5, 10, 20]))
    ty, a, b = util.gen_pool_events(9000 + seed, 4000, tgt, float(rng.choice([1.3, 2.0, 5.0, 9.0, 12.0])),
                                    float(rng.choice([0.0, 0.1, 0.5])))
    res = check(tgt, int(rng.choice([1, 7, 64])), ty, a, b, int(rng.randint(0, 4)))
    assert len(res) > 0


def test_spec_carried_dropping_state_and_big_events():
    """Idle events that consume hundreds of handles at once, starting inside a dropping episode
    (drop_next in the past: every above-target handle of the event is dropped and counted)."""
    ms = 1_000_000
    t0 = 50 * 10 ** 9
    ty, a, b = [], [], []
    for i in range(2250):  # a claim every 20 us for 45 ms
        ty.append(0); a.append(t0 + i * 20_000); b.append(0)  # noqa: E702
    # sojourn 0..45 ms: the first 2000 handles are above target 5 and are all dropped (already
    # dropping, drop_next in the past), then 50 of the 250 fresh ones get the slots
    ty.append(1); a.append(t0 + 45 * ms); b.append(50)        # noqa: E702
    ty.append(2); a.append(2100); b.append(0)                 # noqa: E702
    for i in range(2250, 3000):
        ty.append(0); a.append(t0 + i * 20_000); b.append(0)  # noqa: E702
    ty.append(1); a.append(t0 + 61 * ms); b.append(700)       # noqa: E702
    ty.append(1); a.append(t0 + 400 * ms); b.append(5000)     # noqa: E702  whatever is left has timed out
    state = {"count": 7, "drop_next": (t0 + 10 * ms) / 1e6, "first_above_time": (t0 - 200 * ms) / 1e6,
             "dropping": True, "last_empty": None}
    res = check(5, 13, ty, a, b, 0, state)
    codes = np.bincount(res[:, 1], minlength=4)
    assert codes[1] > 1000 and codes[3] > 0
