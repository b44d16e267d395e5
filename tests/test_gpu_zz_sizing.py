"""The pool's sizing loop on a live ConnectionPool (lib/pool.js:251-262 sampler, :541-592 _rebalance up to the
plan): the filter's input comes from the device queue's counters (cb_queue_info_get), the target from the
host mirror of the reference's arithmetic (pinned to the reference's own FIRFilter by
tests/test_lpf_golden.py), the plan from the device's planRebalance (cb_plan_rebalance).  Every step is
compared with the oracle (cbo_lpf_run + cbo_plan_rebalance) driven with the same occupancy.

(The file sorts last among the GPU tests: it was written after the round's last full GPU session; its scenarios ran
green on a B200 through tests/gpu_zz_quick.py -- profiles/r02_late_gpu_tests.jsonl -- and run on the CPU through the
emulated context, tests/test_pool_mirror_emulated.py.)"""
import pytest

from oracle import oracle as O

pytestmark = [pytest.mark.gpu]


def sizing_scenario(cb):
    clock = {"now": 10 ** 9}
    draws = iter([0.3, 0.8, 0.1])
    pool = cb.ConnectionPool({"targetClaimDelay": 5, "maximum": 8, "spares": 2, "clock": lambda: clock["now"],
                              "random": lambda: next(draws)})
    for k in ("b0", "b1", "b2"):
        pool.on_resolver_added(k, {})
    keys = list(pool.p_keys)
    have = [int((pool._slot_backend == pool._key_ids[k]).sum()) for k in keys]
    assert sum(have) == 8 and sorted(have) == [2, 3, 3]
    ev = {"ty": [], "a": [], "b": [], "c": [], "d": []}
    plans = []

    def occupancy():
        st = pool.getStats()
        return st["totalConnections"], st["idleConnections"], st["waiterCount"]

    def tick():
        total, idle, _ = occupancy()
        for key, v in zip("ty a b c d".split(), (0, total, idle, 0, 0)):
            ev[key].append(v)
        was_clamped = pool.p_lastRebalClamped
        if was_clamped:
            record_rebalance(None)
        out = pool.lpTick()
        assert (out is not None) == was_clamped          # lib/pool.js:260-261
        if out is not None:
            plans[-1] = out

    def record_rebalance(plan):
        total, idle, waiters = occupancy()
        for key, v in zip("ty a b c d".split(), (1, total, idle, 0, waiters)):
            ev[key].append(v)
        plans.append(plan)

    def rebalance():
        record_rebalance(None)
        plans[-1] = pool.rebalance()

    # idle pool: the filter fills with `spares`, the target is clamped to what the filter remembers
    for _ in range(3):
        tick()
    rebalance()
    # load arrives: 8 claims take every slot, 5 more wait
    handles = [pool.claim(lambda *a: None) for _ in range(13)]
    assert [h.ch_state for h in handles].count("claimed") == 8
    for _ in range(4):
        clock["now"] += 200_000_000 // 50
        tick()
    rebalance()
    # the load goes away: the clamp holds the size up while the filter decays
    clock["now"] += 1_000_000
    for h in handles:                  # a release hands the slot to a waiter, which is then released in its turn
        if h.isInState("claimed"):
            h.release()
    for _ in range(6):
        tick()
    rebalance()
    # the oracle on the same occupancy
    want = O.lpf_run(2, 8, ev["ty"], ev["a"], ev["b"], ev["c"], ev["d"])
    assert len(want) == len(plans) and len(plans) >= 3
    saw_clamped = saw_free = False
    for row, plan in zip(want, plans):
        assert plan is not None
        assert (plan["target"], plan["clamped"]) == (int(row[0]), bool(row[1]))
        owant, oadd, orem = O.plan_rebalance(have, [0] * len(keys), int(row[0]), 8)
        assert plan["add"] == [keys[int(i)] for i in oadd]
        assert plan["remove"] == [(keys[int(k)], int(o)) for k, o in orem]
        saw_clamped |= plan["clamped"]
        saw_free |= not plan["clamped"]
    assert saw_clamped and saw_free
    assert pool.p_lastRebalance == plans[-1]
    return plans


@pytest.fixture(scope="module")
def cb():
    import cueball_b200
    return cueball_b200


def test_sizing_loop_on_the_device_counters(cb):
    sizing_scenario(cb)
