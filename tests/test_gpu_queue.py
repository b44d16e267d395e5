"""GPU parity tests of the device-resident pending-claim queue (cb_queue_*) against
the pinned pool-model oracle.  The kernels follow tools/proto_queue.py, which passes the same
comparisons on CPU (tests/test_proto_queue.py); validated on a B200 at the start of round 2
(profiles/r02_validate_pending.txt)."""
import os

import numpy as np
import pytest

import util
from oracle import oracle as O

pytestmark = [pytest.mark.gpu]

G = util.load_golden("pool_sim_golden.json")


@pytest.fixture(scope="module")
def cb():
    import cueball_b200
    return cueball_b200


def run_device(cb, target, nslots, ty, a, b, idle0, state=None):
    ctx = cb.Context(0)
    ctx.codel_init(target)
    if state:
        ctx.set_state(state)
    ctx.backends_set(4, [0.1, 0.6, 0.3, 0.9], nslots)
    ctx.queue_create(1 << 16, 1 << 18, idle0)
    for t, x, y in zip(ty, a, b):
        if t == 0:
            ctx.queue_push([x])
        elif t == 1:
            ctx.queue_idle(x, y)
        else:
            ctx.queue_cancel([x])
    res = ctx.queue_results()
    return ctx, res


def check(cb, target, nslots, ty, a, b, idle0, state=None):
    cd = O.new_codel(target, state)
    want, tmo, out = O.pool_sim(cd, nslots, ty, a, b, idle0=idle0)
    ctx, got = run_device(cb, target, nslots, ty, a, b, idle0, state)
    assert got.tolist() == want.tolist()
    info = ctx.queue_info()
    assert info.counters() == out.counters()
    assert (info.idle, info.tail - info.head) == (out.idle, out.queued)
    st = ctx.stats()
    assert st.cursor == out.cursor
    util.assert_state_equal({"count": st.state.count, "drop_next": st.state.drop_next,
                             "first_above_time": st.state.first_above_time, "dropping": bool(st.state.dropping),
                             "last_empty": st.state.last_empty if st.state.has_last_empty else None}, cd.as_dict())


@pytest.mark.parametrize("case", G["cases"], ids=[c["name"] for c in G["cases"]])
def test_device_queue_on_golden_scripts(cb, case):
    check(cb, case["target"], case["nslots"], case["ev_type"], case["ev_a"], case["ev_b"], case["idle0"])


def test_device_queue_big_events_in_dropping_state(cb):
    ms = 1_000_000
    t0 = 50 * 10 ** 9
    ty, a, b = [], [], []
    for i in range(2250):
        ty.append(0); a.append(t0 + i * 20_000); b.append(0)  # noqa: E702
    ty.append(1); a.append(t0 + 45 * ms); b.append(50)        # noqa: E702
    ty.append(2); a.append(2100); b.append(0)                 # noqa: E702
    for i in range(2250, 3000):
        ty.append(0); a.append(t0 + i * 20_000); b.append(0)  # noqa: E702
    ty.append(1); a.append(t0 + 61 * ms); b.append(700)       # noqa: E702
    ty.append(1); a.append(t0 + 400 * ms); b.append(5000)     # noqa: E702
    state = {"count": 7, "drop_next": (t0 + 10 * ms) / 1e6, "first_above_time": (t0 - 200 * ms) / 1e6,
             "dropping": True, "last_empty": None}
    check(cb, 5, 13, ty, a, b, 0, state)


def test_batched_push_equals_single_pushes(cb):
    """cb_queue_push with many claims per call (timeouts elementwise, idle slots first)."""
    tgt, nslots = 5, 9
    ms = 1_000_000
    t0 = 10 ** 10
    starts = t0 + np.arange(5000, dtype=np.int64) * 3_000
    ctx = cb.Context(0)
    ctx.codel_init(tgt)
    ctx.backends_set(4, [0.1, 0.6, 0.3, 0.9], nslots)
    ctx.queue_create(1 << 14, 1 << 16, 7)
    first, tmo = ctx.queue_push(starts, want_timeouts=True)
    assert first == 0 and np.all(tmo == 50.0)
    ctx.queue_idle(int(starts[-1]) + 20 * ms, 100)
    got = ctx.queue_results()
    ty = [0] * 5000 + [1]
    a = starts.tolist() + [int(starts[-1]) + 20 * ms]
    b = [0] * 5000 + [100]
    cd = O.new_codel(tgt)
    want, _, out = O.pool_sim(cd, nslots, ty, a, b, idle0=7)
    assert got.tolist() == want.tolist()
