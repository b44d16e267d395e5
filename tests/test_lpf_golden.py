"""The pool's low-pass filter and _rebalance's target sizing (lib/pool.js:44-100, 251-258, 556-587; SURVEY 8(f)
row 4, "LPF clamp input") against vectors recorded from the reference's OWN FIRFilter / genTaps run under the
reference tree's engine (tests/golden/lpf_golden.json, tests/golden/gen_lpf_golden.py):
  * the oracle's C restatement (cbo_lpf_taps, cbo_lpf_run), bit for bit; and against live reference runs on
    fresh seeds where oracle/_ref exists;
  * the product's host mirror (cueball_b200.pool: genTaps, FIRFilter, rebalanceTarget) -- host arithmetic by
    design (128 multiply-adds five times a second); its input on a live pool is the device queue's counters
    (tests/test_pool_mirror_emulated.py, tests/test_gpu_zz_sizing.py)."""
import numpy as np
import pytest

import util
from oracle import oracle as O
from oracle import ref

G = util.load_golden("lpf_golden.json")


def bits(xs):
    return np.asarray(xs, dtype=np.float64).view(np.uint64).tolist()


def test_taps_oracle_and_product():
    want = [util.unhx(t) for t in G["taps"]]
    assert len(want) == 128
    assert bits(O.lpf_taps()) == bits(want)
    from cueball_b200 import pool
    assert bits(pool.LP_TAPS) == bits(want)
    assert pool.LP_INT == G["lp_int_ms"] == 200 and pool.LP_RATE == 5


@pytest.mark.parametrize("case", G["cases"], ids=[c["name"] for c in G["cases"]])
def test_oracle_lpf_on_golden_scripts(case):
    got = O.lpf_run(case["spares"], case["max"], case["ty"], case["a"], case["b"], case["c"], case["d"])
    assert got[:, 0].astype(int).tolist() == case["target"]
    assert got[:, 1].astype(int).tolist() == case["clamped"]
    assert bits(got[:, 2]) == bits([util.unhx(x) for x in case["lpf"]])


@pytest.mark.parametrize("case", G["cases"], ids=[c["name"] for c in G["cases"]])
def test_product_filter_and_target_on_golden_scripts(case):
    from cueball_b200 import pool
    lpf = pool.FIRFilter(pool.LP_TAPS)
    targets, clamps, gets = [], [], []
    for ty, a, b, c, d in zip(case["ty"], case["a"], case["b"], case["c"], case["d"]):
        if ty == 0:
            lpf.put(a - b + case["spares"])      # lib/pool.js:256-258
        else:
            gets.append(lpf.get())
            t, cl = pool.rebalanceTarget(lpf, a, b, c, d, case["spares"], case["max"])
            targets.append(t)
            clamps.append(int(cl))
    assert targets == case["target"] and clamps == case["clamped"]
    assert bits(gets) == bits([util.unhx(x) for x in case["lpf"]])


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [4101, 4102])
def test_oracle_lpf_against_live_reference(seed):
    ty, a, b, c, d = util.gen_lpf_events(seed, 500, 40)
    r = ref.lpf_run(3, 50, ty, a, b, c, d)
    want = np.array(r["out"], dtype=np.float64).reshape(-1, 3)
    got = O.lpf_run(3, 50, ty, a, b, c, d)
    assert bits(got) == bits(want)
