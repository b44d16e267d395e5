"""ConnectionPool.claim() / handle.cancel() / release() (the host mirror of lib/pool.js:859-969 in
node-cueball_b200/pool.py, which keeps the pending-claim FIFO on the device: cb_queue_*) pinned to the
pool-model oracle: the 12 golden event scripts recorded from the reference's own lib/codel.js +
lib/queue.js (tests/golden/pool_sim_golden.json) and 24 random scripts must come out of the mirror with
exactly the oracle's (ticket, code, slot) resolution records, counters, idle count and queue length --
and every handle's callback must fire accordingly (ClaimTimeoutError for CoDel drops and hard timeouts,
lib/connection-fsm.js:626-642; cancellation test/pool.test.js:727)."""
import numpy as np
import pytest

import util
from oracle import oracle as O

pytestmark = [pytest.mark.gpu]

G = util.load_golden("pool_sim_golden.json")


def replay(cb, target, nslots, ty, a, b, idle0):
    clock = {"now": 0}
    draws = iter([0.1, 0.6, 0.3, 0.9])
    pool = cb.ConnectionPool({"targetClaimDelay": target, "maximum": nslots, "clock": lambda: clock["now"],
                              "random": lambda: next(draws), "initialIdleSlots": idle0, "queueCapacity": 1 << 16})
    for k in range(4):
        pool.on_resolver_added("b%d" % k, {})
    handles, got = [], {"admit": 0, "timeout": 0}

    def make_cb():
        def _cb(err, hdl, conn):
            if err is None:
                got["admit"] += 1
                assert hdl.isInState("claimed") and conn["slot"] == hdl.ch_slot
            else:
                assert isinstance(err, cb.ClaimTimeoutError)
                got["timeout"] += 1
        return _cb
    for t, x, y in zip(ty, a, b):
        if t == 0:
            clock["now"] = int(x)
            handles.append(pool.claim(make_cb()))
        elif t == 1:
            pool.slots_idle(int(y), now_ns=int(x))
        elif handles[int(x)].isInState("waiting"):
            # the script's CANCEL is handle.cancel() of a WAITING handle (the model ignores it otherwise); cancel() of
            # a claimed handle is a release in the reference and in the mirror: test_cancel_of_a_claimed_handle...
            handles[int(x)].cancel()
    return pool, handles, got


def check(cb, target, nslots, ty, a, b, idle0):
    cd = O.new_codel(target)
    want, tmo, out = O.pool_sim(cd, nslots, ty, a, b, idle0=idle0)
    pool, handles, got = replay(cb, target, nslots, ty, a, b, idle0)
    assert [list(r) for r in pool.records] == want.tolist()
    stats = pool.getStats()
    assert sorted(stats) == ["counters", "idleConnections", "pendingConnections", "totalConnections", "waiterCount"]
    assert stats["counters"] == {k: v for k, v in out.counters().items() if v}
    assert (stats["idleConnections"], stats["waiterCount"], stats["totalConnections"]) == (out.idle, out.queued, nslots)
    assert [h.ch_claimTimeout for h in handles] == tmo.tolist()
    # callbacks: an admit record resolves its handle to 'claimed'; a CoDel drop or a hard timeout fails a handle
    # that was still waiting with ClaimTimeoutError (a record for a cancelled handle fires nothing)
    codes = want[:, 1] if len(want) else np.zeros(0, int)
    assert got["admit"] == int((codes == 0).sum())
    assert got["timeout"] == sum(1 for h in handles if h.isInState("failed"))
    assert got["timeout"] <= int(((codes == 1) | (codes == 3)).sum())
    st = pool._ctx.stats()
    assert st.cursor == out.cursor
    util.assert_state_equal({"count": st.state.count, "drop_next": st.state.drop_next,
                             "first_above_time": st.state.first_above_time, "dropping": bool(st.state.dropping),
                             "last_empty": st.state.last_empty if st.state.has_last_empty else None}, cd.as_dict())


@pytest.fixture(scope="module")
def cb():
    import cueball_b200
    return cueball_b200


@pytest.mark.parametrize("case", G["cases"], ids=[c["name"] for c in G["cases"]])
def test_pool_mirror_on_golden_scripts(cb, case):
    check(cb, case["target"], case["nslots"], case["ev_type"], case["ev_a"], case["ev_b"], case["idle0"])


@pytest.mark.parametrize("seed", range(24))
def test_pool_mirror_on_random_scripts(cb, seed):
    rng = np.random.RandomState(7000 + seed)
    tgt = int(rng.choice([2, 5, 10, 20]))
    ty, a, b = util.gen_pool_events(9000 + seed, 600, tgt, float(rng.choice([2.0, 5.0, 9.0, 12.0])),
                                    float(rng.choice([0.0, 0.1, 0.5])))
    check(cb, tgt, int(rng.choice([3, 11, 37])), ty, a, b, seed % 3)


def test_cancel_of_a_claimed_handle_releases_it(cb):
    """lib/connection-fsm.js:580-587: cancel() on a claimed handle is a release()."""
    clock = {"now": 10 ** 9}
    pool = cb.ConnectionPool({"targetClaimDelay": 5, "maximum": 2, "clock": lambda: clock["now"], "random": lambda: 0.5})
    pool.on_resolver_added("b0", {})
    h1, h2, h3 = pool.claim(lambda *a: None), pool.claim(lambda *a: None), pool.claim(lambda *a: None)
    assert h1.isInState("claimed") and h2.isInState("claimed") and h3.isInState("waiting")
    clock["now"] += 1_000_000
    h1.cancel()
    assert h1.isInState("released") and h3.isInState("claimed")
    with pytest.raises(cb.ClaimHandleMisusedError):
        h1.release()


def test_waiters_time_out_without_any_release(cb):
    """The claim timer (getMaxIdle, lib/codel.js:109-118 -> lib/connection-fsm.js:626-642) fires from expire()
    and from the next claim(), not only when a slot is released."""
    clock = {"now": 10 ** 9}
    errs = []
    pool = cb.ConnectionPool({"targetClaimDelay": 5, "maximum": 1, "clock": lambda: clock["now"], "random": lambda: 0.5})
    pool.on_resolver_added("b0", {})
    pool.claim(lambda *a: None)                       # takes the only slot and never releases it
    w = pool.claim(lambda e, h, c: errs.append(e))
    assert w.isInState("waiting") and w.ch_claimTimeout == 50.0
    clock["now"] += 49_000_000
    pool.expire()
    assert w.isInState("waiting")
    clock["now"] += 2_000_000
    pool.expire()
    assert w.isInState("failed") and isinstance(errs[0], cb.ClaimTimeoutError)
    assert pool.getStats()["counters"]["claim-timeout"] == 1


def test_kang_snapshot_shape(cb):
    """What lib/pool-monitor.js:91-133 publishes for a pool: backends, per-backend connection counts by slot
    state, dead backends, state, counters, options."""
    clock = {"now": 10 ** 9}
    draws = iter([0.1, 0.9, 0.4])
    pool = cb.ConnectionPool({"targetClaimDelay": 5, "maximum": 4, "spares": 2, "clock": lambda: clock["now"],
                              "random": lambda: next(draws)})
    for k in ("b0", "b1"):
        pool.on_resolver_added(k, {"address": "10.0.0.%s" % k[1], "port": 80})
    hs = [pool.claim(lambda *a: None) for _ in range(5)]
    assert [h.ch_state for h in hs] == ["claimed"] * 4 + ["waiting"]
    snap = pool.toKangSnapshot()
    assert set(snap) == {"backends", "connections", "dead_backends", "state", "counters", "options"}
    assert snap["backends"] == {"b0": {"address": "10.0.0.0", "port": 80}, "b1": {"address": "10.0.0.1", "port": 80}}
    assert snap["connections"] == {"b0": {"busy": 2}, "b1": {"busy": 2}}
    assert snap["state"] == "running" and snap["dead_backends"] == []
    assert snap["options"] == {"spares": 2, "maximum": 4}
    assert snap["counters"] == {"claim": 5, "queued-claim": 1, "max-claim-queue": 1}
    clock["now"] += 1_000_000
    hs[0].release()                  # the waiter takes the slot at once: still four busy
    hs[1].release()
    snap = pool.toKangSnapshot()
    assert sorted(sum((list(v.items()) for v in snap["connections"].values()), [])) == [("busy", 1), ("busy", 2), ("idle", 1)]
    pool.stop()
    assert pool.toKangSnapshot()["state"] == "stopping"
