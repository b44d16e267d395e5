"""The `ControlledDelay` mirror (node-cueball_b200/codel.py: overloaded / overloaded_batch / empty / getMaxIdle and the
cd_* fields, lib/codel.js:24-118) on the CPU against the vectors recorded from the reference's own lib/codel.js
(tests/golden/codel_golden.json): the mirror runs on tests/emu_context.EmuContext, i.e. on the library's scalar kernels
and its scan -> resolve -> emit pipeline compiled for the host.  Claim by claim (the call shape of lib/pool.js:735-736)
and batch by batch.  The GPU run of the same vectors is tests/test_gpu_parity.py."""
import os

import numpy as np
import pytest

import util
from emu_context import EmuContext

G = util.load_golden("codel_golden.json")
# per-claim launches cost ~1 ms each under the emulation: the long generated traces go through the batch form only
FULL = os.environ.get("CB_EMU_FULL") == "1"
# the 20k-claim backwards-clock trace takes the exact one-warp path: three minutes under the emulation (the path itself
# is covered at small sizes by tests/test_pipeline_kernels_emulated.py and here by the short backwards-clock vectors)
BATCH = [c for c in G["cases"] if FULL or c["name"] != "gen_backstep"]
SHORT = [c for c in G["cases"] if c["trace"]["kind"] == "explicit" or c["trace"].get("n", 0) <= 3000]


def mirror(case):
    import cueball_b200 as cb
    ctx = EmuContext()
    cd = cb.ControlledDelay(case["target"], context=ctx)
    st = util.state_unhex(case["init_state"])
    if st:
        ctx.set_state(st)
    ctx.backends_set(0, [], 0)
    return cd, ctx


def state_of(cd):
    return {"count": float(cd.cd_count), "drop_next": cd.cd_drop_next, "first_above_time": cd.cd_first_above_time,
            "dropping": cd.cd_dropping, "last_empty": cd.cd_last_empty}


@pytest.mark.parametrize("case", SHORT, ids=[c["name"] for c in SHORT])
def test_claim_by_claim(case):
    """p_codel.overloaded(hdl.ch_started) per dequeued claim, p_codel.empty() where the queue ran dry."""
    start, now, flags = util.make_trace(case["trace"])
    cd, ctx = mirror(case)
    want = util.unrle_codes(case["expect"]["dec_rle"])
    lo = 0
    for b, wst in zip(case["batches"], case["expect"]["states"]):
        for i in range(lo, lo + b):
            drop = cd.overloaded(int(start[i]), int(now[i]))
            if want[i] != 2:                      # a skipped handle's decision is not recorded (lib/pool.js:738-740)
                assert drop == (want[i] == 1), (case["name"], i)
            if flags is not None and flags[i] & 2:
                cd.empty(int(now[i]))
        util.assert_state_equal(state_of(cd), util.state_unhex(wst), case["name"])
        lo += b


@pytest.mark.parametrize("case", BATCH, ids=[c["name"] for c in BATCH])
def test_batch_by_batch(case):
    """overloaded_batch(): the whole drain loop of lib/pool.js:733-753 per call."""
    start, now, flags = util.make_trace(case["trace"])
    cd, ctx = mirror(case)
    want = util.unrle_codes(case["expect"]["dec_rle"])
    lo = 0
    for b, wst in zip(case["batches"], case["expect"]["states"]):
        f = None if flags is None else flags[lo:lo + b]
        dec, slot, _, st = cd.overloaded_batch(start[lo:lo + b], now[lo:lo + b], f)
        assert np.array_equal(dec, want[lo:lo + b]), case["name"]
        util.assert_state_equal(state_of(cd), util.state_unhex(wst), case["name"])
        lo += b
    assert int((want == 1).sum()) == case["expect"]["drops"]


@pytest.mark.parametrize("case", G["max_idle"], ids=lambda c: "t%s_%s" % (c["target"], c["now_ns"]))
def test_get_max_idle(case):
    import cueball_b200 as cb
    ctx = EmuContext()
    cd = cb.ControlledDelay(case["target"], context=ctx)
    st = util.state_unhex(case["state"])
    if st:
        ctx.set_state(st)
    assert util.same_bits(cd.getMaxIdle(case["now_ns"]), util.unhx(case["expect"]))


def test_constructor_asserts_like_the_reference():
    import cueball_b200 as cb
    for bad in (0, -5, 2.5, float("nan"), "5"):
        with pytest.raises(AssertionError):
            cb.ControlledDelay(bad, context=EmuContext())
