"""The host mirror `ConnectionPool` (node-cueball_b200/pool.py: claim() / cancel() / release() / expire() /
getStats() / toKangSnapshot(), lib/pool.js:859-969, lib/connection-fsm.js:580-642, lib/pool-monitor.js:91-133)
replayed in THIS container, without a GPU: the pool runs on tests/emu_context.EmuContext, which serves the
Context calls pool.py makes with the library's own queue and topology kernels compiled for the host
(tests/emu).  Same scripts, same checks as tests/test_gpu_pool_mirror.py (whose helpers are reused), so the
mirror's host logic -- records -> handle transitions and callbacks, counters, the snapshot shapes -- is pinned
to the pool-model oracle on the CPU as well.  An emulated launch costs ~20 ms, so the default run replays the
head of three golden scripts; CB_EMU_FULL=1 replays all twelve and six random ones."""
import os
import types

import numpy as np
import pytest

import test_gpu_pool_mirror as GP
import util
from emu_context import EmuContext

FULL = os.environ.get("CB_EMU_FULL") == "1"
QUICK = ["hard_timeouts_t5", "pauses_and_stalls_t10", "idle_slots_first_t5"]
CASES = GP.G["cases"] if FULL else [c for c in GP.G["cases"] if c["name"] in QUICK]
HEAD = None if FULL else 260


@pytest.fixture(scope="module")
def cb():
    """The package with ConnectionPool bound to an emulated context (everything else is the real thing)."""
    import cueball_b200 as real

    def pool(options):
        o = dict(options)
        o["context"] = EmuContext()
        return real.ConnectionPool(o)
    return types.SimpleNamespace(ConnectionPool=pool, ClaimTimeoutError=real.ClaimTimeoutError,
                                 ClaimHandleMisusedError=real.ClaimHandleMisusedError,
                                 PoolStoppingError=real.PoolStoppingError, NoBackendsError=real.NoBackendsError)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_pool_mirror_on_golden_scripts_emulated(cb, case):
    GP.check(cb, case["target"], case["nslots"], case["ev_type"][:HEAD], case["ev_a"][:HEAD], case["ev_b"][:HEAD],
             case["idle0"])


@pytest.mark.parametrize("seed", range(6) if FULL else [5])
def test_pool_mirror_on_random_scripts_emulated(cb, seed):
    rng = np.random.RandomState(7000 + seed)
    tgt = int(rng.choice([2, 5, 10, 20]))
    ty, a, b = util.gen_pool_events(9000 + seed, 600 if FULL else 220, tgt, float(rng.choice([2.0, 5.0, 9.0, 12.0])),
                                    float(rng.choice([0.0, 0.1, 0.5])))
    GP.check(cb, tgt, int(rng.choice([3, 11, 37])), ty, a, b, seed % 3)


def test_cancel_of_a_claimed_handle_releases_it_emulated(cb):
    GP.test_cancel_of_a_claimed_handle_releases_it(cb)


def test_waiters_time_out_without_any_release_emulated(cb):
    GP.test_waiters_time_out_without_any_release(cb)


def test_kang_snapshot_shape_emulated(cb):
    GP.test_kang_snapshot_shape(cb)


def test_topology_changes_move_idle_slots(cb):
    """_refresh_slots: the slots of a grown table turn idle as an event (they drain waiters, lib/pool.js:715);
    the slots of a table that shrank to nothing leave the idle queue (setUnwanted, lib/pool.js:601-631)."""
    clock = {"now": 10 ** 9}
    draws = iter([0.2, 0.7, 0.5])
    got = []
    pool = cb.ConnectionPool({"targetClaimDelay": 5, "maximum": 3, "clock": lambda: clock["now"],
                              "random": lambda: next(draws)})
    waiting = [pool.claim(lambda e, h, c: got.append((e, c))) for _ in range(2)]  # no backend yet: both wait
    assert [h.ch_state for h in waiting] == ["waiting", "waiting"]
    assert pool.getStats()["totalConnections"] == 0
    clock["now"] += 1_000_000
    pool.on_resolver_added("b0", {})          # 3 slots appear and drain the two waiters
    assert [h.ch_state for h in waiting] == ["claimed", "claimed"]
    assert [c["slot"] for _, c in got] == [0, 1] and all(e is None for e, _ in got)
    st = pool.getStats()
    assert (st["totalConnections"], st["idleConnections"], st["waiterCount"]) == (3, 1, 0)
    pool.on_resolver_added("b1", {})          # same table size: nothing moves
    assert pool.getStats()["idleConnections"] == 1
    assert sorted(pool.toKangSnapshot()["connections"]) == ["b0", "b1"]
    pool.on_resolver_removed("b0")
    pool.on_resolver_removed("b1")            # no backend, no slots: the idle one is gone
    st = pool.getStats()
    assert (st["totalConnections"], st["idleConnections"]) == (0, 0)
    with pytest.raises(AssertionError):
        pool.on_resolver_removed("b1")


def test_sizing_loop_emulated(cb):
    """lpTick() / rebalance(): the filter's input from the queue's counters, the plan from the planRebalance
    kernel, every step against the oracle (the scenario of tests/test_gpu_zz_sizing.py)."""
    import test_gpu_zz_sizing as Z
    plans = Z.sizing_scenario(cb)
    assert plans[0]["target"] == 2 and not plans[0]["clamped"]  # idle pool: busy 0 + spares 2; the filter still remembers less


def test_reference_codel_tests_emulated(cb):
    """test/codel.test.js restated (tests/test_gpu_zz_codel_load_runs.py): 'implicit high timeout', and one load run
    shortened to 1.5 s of claims (CB_EMU_FULL=1: the reference's 5 s, targets 300 and 1000 ms)."""
    import test_gpu_zz_codel_load_runs as Z
    Z.implicit_high_timeout(cb)
    if FULL:
        for delay in (300, 1000):
            Z.load_run(cb, delay)
    else:
        res = Z.load_run(cb, 300, claim_ms=1500)
        assert res["count"] == 750


# ---------------------------------------------------------------- test/pool.test.js, the claim-queue parts, restated
def _plain_pool(cb, clock, **kw):
    """A pool WITHOUT targetClaimDelay, as test/pool.test.js makes them (spares 2, maximum 2 unless given): no CoDel,
    per-claim options.timeout; the mirror runs it as a plain host FIFO."""
    opts = {"maximum": 2, "spares": 2, "clock": lambda: clock["now"], "random": lambda: 0.5}
    opts.update(kw)
    return cb.ConnectionPool(opts)


def test_empty_pool(cb):
    """test/pool.test.js:106-134: errorOnEmpty -> NoBackendsError at once; a plain claim times out."""
    import cueball_b200 as real
    clock = {"now": 10 ** 9}
    pool = _plain_pool(cb, clock, maximum=4)
    errs = []
    pool.claim({"errorOnEmpty": True}, lambda e, h, c: errs.append(e))
    assert isinstance(errs[0], real.NoBackendsError) and "no backend" in str(errs[0]).lower()
    pool.claim({"timeout": 100}, lambda e, h, c: errs.append(e))
    assert len(errs) == 1
    clock["now"] += 100 * 1_000_000
    pool.expire()
    assert isinstance(errs[1], cb.ClaimTimeoutError) and "timed out" in str(errs[1]).lower()


def test_pool_with_one_backend(cb):
    """test/pool.test.js:136-203: claims time out while the two connections are still connecting, succeed once
    they are up, and the third concurrent claim times out again."""
    clock = {"now": 10 ** 9}
    pool = _plain_pool(cb, clock, initialIdleSlots=0)
    pool.on_resolver_added("b1", {})
    assert pool.getStats()["totalConnections"] == 2
    log = []
    pool.claim({"timeout": 100}, lambda e, h, c: log.append(("first", e, h, c)))
    clock["now"] += 100 * 1_000_000
    pool.expire()
    assert isinstance(log[0][1], cb.ClaimTimeoutError)
    pool.slots_idle(2)                                   # c.connect() on both
    pool.claim({"timeout": 100}, lambda e, h, c: log.append(("again", e, h, c)))
    pool.claim({"timeout": 100}, lambda e, h, c: log.append(("once more", e, h, c)))
    assert [x[1] for x in log[1:]] == [None, None] and all(x[2] is not None and x[3] is not None for x in log[1:])
    assert sorted(x[3]["slot"] for x in log[1:]) == [0, 1]
    pool.claim({"timeout": 100}, lambda e, h, c: log.append(("empty", e, h, c)))
    assert len(log) == 3
    clock["now"] += 100 * 1_000_000
    pool.expire()
    assert isinstance(log[3][1], cb.ClaimTimeoutError)


def test_claim_cancellation(cb):
    """test/pool.test.js:727-780: a cancelled claim's callback never fires -- not when a slot turns up, not when
    its timeout would have come due."""
    clock = {"now": 10 ** 9}
    pool = _plain_pool(cb, clock, initialIdleSlots=0)
    pool.on_resolver_added("b1", {})
    fired = []
    claim = pool.claim({"timeout": 100}, lambda e, h, c: fired.append(e))
    claim.cancel()
    pool.slots_idle(2)
    clock["now"] += 150 * 1_000_000
    pool.expire()
    assert fired == [] and claim.isInState("cancelled")
    assert pool.getStats()["idleConnections"] == 2
    pool.stop()


def test_getstats_shape_issue_132(cb):
    """test/pool.test.js:902-966 (cueball#132): five keys; zeros before a backend exists, then 2 / 2 / 0 / 0."""
    clock = {"now": 10 ** 9}
    pool = _plain_pool(cb, clock)
    s = pool.getStats()
    assert isinstance(s, dict) and len(s) == 5 and isinstance(s["counters"], dict)
    assert (s["totalConnections"], s["idleConnections"], s["pendingConnections"], s["waiterCount"]) == (0, 0, 0, 0)
    pool.on_resolver_added("b1", {})
    s = pool.getStats()
    assert len(s) == 5 and isinstance(s["counters"], dict)
    assert (s["totalConnections"], s["idleConnections"], s["pendingConnections"], s["waiterCount"]) == (2, 2, 0, 0)
    assert pool.toKangSnapshot()["connections"] == {"b1": {"idle": 2}}
    pool.stop()


def test_config_c1_emulated(cb):
    """BASELINE.json configs[0] (StaticResolver, 4 backends, 256 claim()s, target 5 ms) through the mirror."""
    import test_gpu_zz_codel_load_runs as Z
    res = Z.config_c1(cb)
    assert res["successes"] > 0


def test_next_deadline_survives_heap_compaction(cb):
    """nextDeadline() keeps (deadline, ticket) entries lazily; claim() rebuilds the heap once resolved handles' entries
    outnumber the live ones: the answer must not change."""
    clock = {"now": 10 ** 9}
    pool = cb.ConnectionPool({"targetClaimDelay": 5, "maximum": 1, "clock": lambda: clock["now"], "random": lambda: 0.5})
    pool._DEADLINE_SLACK = 4
    pool.on_resolver_added("b0", {})
    held = pool.claim(lambda *a: None)
    assert held.isInState("claimed") and pool.nextDeadline() is None
    waiters = []
    for i in range(12):                       # admitted one after the other: their entries go stale
        clock["now"] += 100_000
        w = pool.claim(lambda *a: None)
        waiters.append(w)
        held.release()
        held = w
        assert held.isInState("claimed")
    assert len(pool._deadlines) <= 2 + pool._DEADLINE_SLACK + 1   # 13 claims so far, 12 of them resolved
    clock["now"] += 100_000
    last = pool.claim(lambda *a: None)
    assert last.isInState("waiting")
    assert pool.nextDeadline() == last.ch_started_ns + int(last.ch_claimTimeout * 1e6)
    last.cancel()
    assert pool.nextDeadline() is None
