"""The event-driven pool model (claim() enqueue side with getMaxIdle, idle-first
fast path, the slot-idle drain loop, hard timeouts, counters): the C restatement (oracle/, test
infrastructure) against golden vectors produced by the REFERENCE'S OWN lib/codel.js + lib/queue.js
under SpiderMonkey (tests/golden/gen_pool_golden.py), and against live reference runs where
oracle/_ref exists."""
import numpy as np
import pytest

import util
from oracle import oracle as O
from oracle import ref

G = util.load_golden("pool_sim_golden.json")


def run_c(case_or_target, nslots, ty, a, b, idle0):
    cd = O.new_codel(case_or_target)
    res, tmo, out = O.pool_sim(cd, nslots, ty, a, b, idle0=idle0)
    return cd, res, tmo, out


@pytest.mark.parametrize("case", G["cases"], ids=[c["name"] for c in G["cases"]])
def test_pool_sim_matches_reference(case):
    cd, res, tmo, out = run_c(case["target"], case["nslots"], case["ev_type"], case["ev_a"], case["ev_b"], case["idle0"])
    e = case["expect"]
    assert res.reshape(-1).tolist() == e["res"]
    assert tmo.tolist() == [float(x) for x in e["timeouts"]]
    assert out.counters() == e["counters"]
    assert (out.idle, out.cursor, out.queued) == (e["idle"], e["cursor"], e["queued"])
    util.assert_state_equal(cd.as_dict(), util.state_unhex(e["state"]), case["name"])


def test_golden_covers_every_outcome():
    """The fixtures exercise admits, CoDel drops, skipped (cancelled or timed-out) handles, hard
    timeouts, the idle-first fast path and the lowered getMaxIdle (3 x target)."""
    codes = np.zeros(4, int)
    low = fast = 0
    for c in G["cases"]:
        r = np.array(c["expect"]["res"], dtype=np.int64).reshape(-1, 3)
        codes += np.bincount(r[:, 1], minlength=4)
        low += sum(1 for t in c["expect"]["timeouts"] if t == 3 * c["target"])
        fast += c["expect"]["counters"]["claim"] - c["expect"]["counters"]["queued-claim"]
    assert all(codes > 100), codes
    assert low > 0 and fast > 0


@pytest.mark.ref
@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", [11, 12, 13, 14])
def test_live_reference_random_scripts(seed):
    rng = np.random.RandomState(seed)
    tgt = int(rng.choice([2, 5, 10, 20]))
    ty, a, b = util.gen_pool_events(5000 + seed, 1500, tgt, float(rng.choice([2.0, 5.0, 9.0, 12.0])),
                                    float(rng.choice([0.0, 0.1, 0.5])))
    r = ref.pool_sim(tgt, 11, ty, a, b, idle0=seed % 3)
    cd, res, tmo, out = run_c(tgt, 11, ty, a, b, seed % 3)
    assert res.reshape(-1).tolist() == r["res"]
    assert tmo.tolist() == [float(x) for x in r["timeouts"]]
    assert out.counters() == r["counters"]
    assert (out.idle, out.cursor, out.queued) == (r["idle"], r["cursor"], r["queued"])
    util.assert_state_equal(cd.as_dict(), r["state"], "seed %d" % seed)
