"""The reference's own CoDel tests (test/codel.test.js) restated on the mirror pool over a virtual clock:

  * 'implicit high timeout' (:114-180): with no connection up yet a claim fails with ClaimTimeoutError after
    the implicit getMaxIdle() timeout; once the slots are connected a claim succeeds;
  * the seven load runs 'delay target: 300 ... 5000 ms' (:186-297): 5 claims every 10 ms for 5 s against 2
    slots held 50 ms each; when every callback has fired, the mean claim delay (admits AND timeouts) must lie
    within 175 ms of the target, successes != 0, timeouts != 0, nothing pending.

The reference asserts these statistically on wall-clock timers; on a virtual clock the run is deterministic,
so on top of the reference's assertions the event script the run generated (claims, releases, timer sweeps) is
replayed through the pool-model oracle (cbo_pool_sim, pinned to the reference's lib/codel.js + lib/queue.js)
and every (ticket, code, slot) record and counter must match exactly.

(The file sorts last among the GPU tests: it was written after the round's last full GPU session; its scenarios ran
green on a B200 through tests/gpu_zz_quick.py -- profiles/r02_late_gpu_tests.jsonl -- and run on the CPU through the
emulated context, tests/test_pool_mirror_emulated.py.)"""
import heapq

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = [pytest.mark.gpu]
MS = 1_000_000


def load_run(cb, delay, claim_ms=5000, hold_ms=50, per_tick=5, tick_ms=10, setup=None, check_bound=True):
    clock = {"now": 10 ** 9}
    t_begin = clock["now"]
    pool = cb.ConnectionPool({"targetClaimDelay": delay, "maximum": 2, "spares": 2, "clock": lambda: clock["now"],
                              "random": lambda: 0.5})
    if setup:
        setup(pool)
    else:
        pool.on_resolver_added("b1", {})
    assert pool.getStats()["totalConnections"] == 2          # test/codel.test.js:211-214
    res = {"delays": [], "successes": 0, "timeouts": 0, "failures": 0, "count": 0}
    script = {"ty": [], "a": [], "b": []}
    timers = []   # (due ns, seq, handle): the setTimeout(release, 50) of :236-238
    seq = [0]

    def enqueue():
        res["count"] += 1
        start = clock["now"]

        def _cb(err, hdl, conn):
            res["delays"].append((clock["now"] - start) / MS)
            if isinstance(err, cb.ClaimTimeoutError):
                res["timeouts"] += 1
            elif err is not None:
                res["failures"] += 1
            else:
                res["successes"] += 1
                seq[0] += 1
                heapq.heappush(timers, (clock["now"] + hold_ms * MS, seq[0], hdl))
        script["ty"].append(0); script["a"].append(start); script["b"].append(0)   # noqa: E702
        pool.claim(_cb)

    next_tick = t_begin + tick_ms * MS
    t_stop = t_begin + claim_ms * MS
    while True:
        pending = res["count"] - res["successes"] - res["timeouts"] - res["failures"]
        cands = []
        if next_tick <= t_stop:
            cands.append((next_tick, 0))
        if timers:
            cands.append((timers[0][0], 1))
        dl = pool.nextDeadline()
        if dl is not None:
            cands.append((dl, 2))
        if not cands:
            break
        t, kind = min(cands)
        assert t >= clock["now"]
        clock["now"] = t
        if kind == 0:                      # the setInterval of :244-248
            for _ in range(per_tick):
                enqueue()
            next_tick += tick_ms * MS
        elif kind == 1:                    # a held connection is released: its slot turns idle
            _, _, hdl = heapq.heappop(timers)
            script["ty"].append(1); script["a"].append(t); script["b"].append(1)   # noqa: E702
            hdl.release()
        else:                              # a claim timer came due
            script["ty"].append(1); script["a"].append(t); script["b"].append(0)   # noqa: E702
            pool.expire()
        assert pending >= 0
    # the reference's assertions (test/codel.test.js:258-272)
    delavg = float(np.mean(res["delays"]))
    pstats = pool.getStats()
    if check_bound:
        assert delavg < delay + 175 and delavg > delay - 175, delavg
    assert res["successes"] + res["timeouts"] + res["failures"] == res["count"] == per_tick * (claim_ms // tick_ms)
    assert res["successes"] != 0 and res["failures"] == 0 and (res["timeouts"] != 0 or not check_bound)
    assert pstats["counters"]["claim-timeout"] == res["timeouts"]
    assert pstats["waiterCount"] == 0 or pool.nextDeadline() is None
    # and exactly what the reference's own ControlledDelay + Queue do with this event script
    cd = O.new_codel(delay)
    want, tmo, out = O.pool_sim(cd, 2, script["ty"], script["a"], script["b"], idle0=2)
    assert [list(r) for r in pool.records] == want.tolist()
    assert pstats["counters"] == {k: v for k, v in out.counters().items() if v}
    assert (pstats["idleConnections"], pstats["waiterCount"]) == (out.idle, out.queued)
    res["delavg"] = delavg
    res["pool"] = pool
    return res


def config_c1(cb):
    """BASELINE.json configs[0]: StaticResolver with 4 backends, ConnectionPool, 256 claim()s, CoDel target 5 ms -- the
    reference's own CPU-runnable case (2 slots held 50 ms each, 5 claims every 10 ms, as test/codel.test.js:196-243
    loads a pool).  The resolver's 'added' events reach the pool in list order with srvKey keys (lib/resolver.js:1387-1456,
    lib/pool.js:319-355); every callback fires; the run's event script replayed through the pool-model oracle gives
    exactly the same records."""
    import cueball_b200 as real
    backends = [{"address": "10.0.0.%d" % i, "port": 80} for i in (1, 2, 3, 4)]
    resolver = real.StaticIpResolver({"backends": backends})
    seen = []

    def setup(pool):
        resolver.on("added", lambda k, be: seen.append((k, be["name"])))
        resolver.attach(pool)
        resolver.start()
    per_tick, ticks = 4, 64                      # 256 claims
    res = load_run(cb, 5, claim_ms=ticks * 10, per_tick=per_tick, setup=setup, check_bound=False)
    pool = res["pool"]
    assert res["count"] == 256 and res["successes"] + res["timeouts"] == 256
    assert [n for _, n in seen] == ["10.0.0.%d:80" % i for i in (1, 2, 3, 4)]
    assert [k for k, _ in seen] == [real.srvKey({"name": n, "port": 80, "address": n.split(":")[0]}) for _, n in seen]
    assert resolver.isInState("running") and resolver.count() == 4 and sorted(pool.p_keys) == sorted(k for k, _ in seen)
    snap = pool.toKangSnapshot()
    assert len(snap["backends"]) == 4 and sum(sum(d.values()) for d in snap["connections"].values()) == 2
    assert pool.getStats()["counters"]["claim"] == 256
    return res


def implicit_high_timeout(cb):
    """test/codel.test.js:114-180 with delay = 100: the pool's two connections have not connected yet."""
    clock = {"now": 10 ** 9}
    delay = 100
    pool = cb.ConnectionPool({"targetClaimDelay": delay, "maximum": 2, "spares": 2, "clock": lambda: clock["now"],
                              "random": lambda: 0.5, "initialIdleSlots": 0})
    pool.on_resolver_added("b1", {})
    errs = []
    h = pool.claim(lambda err, hdl, conn: errs.append(err))
    assert h.isInState("waiting") and h.ch_claimTimeout == 10 * delay      # getMaxIdle(), lib/codel.js:109-118
    assert pool.nextDeadline() == clock["now"] + 10 * delay * MS
    clock["now"] = pool.nextDeadline() - 1
    pool.expire()
    assert errs == []
    clock["now"] += 1
    pool.expire()
    assert len(errs) == 1 and isinstance(errs[0], cb.ClaimTimeoutError) and "timed out" in str(errs[0]).lower()
    assert pool.nextDeadline() is None
    pool.slots_idle(2)                     # c.connect() on both connections (:158-161)
    got = []
    pool.claim(lambda err, hdl, conn: got.append((err, hdl, conn)))
    err, hdl, conn = got[0]
    assert err is None and hdl is not None and conn is not None
    hdl.release()
    pool.stop()
    stopped = []
    pool.claim(lambda e, hh, c: stopped.append(e))
    assert isinstance(stopped[0], cb.PoolStoppingError)


@pytest.fixture(scope="module")
def cb():
    import cueball_b200
    return cueball_b200


def test_implicit_high_timeout(cb):
    implicit_high_timeout(cb)


@pytest.mark.parametrize("delay", [300, 500, 1000, 1500, 2000, 2500, 5000])
def test_delay_target(cb, delay):
    load_run(cb, delay)


def test_config_c1_static_resolver_256_claims(cb):
    config_c1(cb)
