#!/usr/bin/env python3
"""tests/golden/gen_lpf_golden.py -- golden vectors for the pool's low-pass filter and _rebalance's target
sizing ("LPF clamp input"), produced by running the REFERENCE'S OWN FIRFilter / genTaps
(lib/pool.js:44-100, verbatim) under the SpiderMonkey engine vendored in the reference tree (oracle/_ref)
through cb_lpf_run (oracle/js/driver.js: the sampler body lib/pool.js:251-258 and the target arithmetic
lib/pool.js:556-587 restated around it).

Needs /root/reference (this container only); writes tests/golden/lpf_golden.json, which holds the event
scripts themselves.  Doubles are stored as C99 hex."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref  # noqa: E402
from util import gen_lpf_events, hx  # noqa: E402  (tests/util.py)

CASES = [
    # name, seed, events, spares, maximum, peak connections
    ("ramp_up_and_down", 3001, 700, 2, 40, 30),
    ("steady_small_pool", 3003, 400, 1, 8, 6),
    ("bursty_load", 3007, 900, 4, 64, 60),
    ("always_clamped_by_max", 3011, 500, 3, 5, 25),
    ("no_spares", 3013, 500, 0, 32, 20),
    ("longer_than_the_ring", 3017, 1500, 2, 100, 90),
]


def main():
    out = []
    taps = None
    for name, seed, nev, spares, mx, peak in CASES:
        ty, a, b, c, d = gen_lpf_events(seed, nev, peak)
        r = ref.lpf_run(spares, mx, ty, a, b, c, d)
        taps = r["taps"]
        assert r["lp_int"] == 200
        o = np.array(r["out"], dtype=np.float64).reshape(-1, 3)
        out.append({"name": name, "seed": seed, "spares": spares, "max": mx, "ty": ty, "a": a, "b": b, "c": c, "d": d,
                    "target": [int(x) for x in o[:, 0]], "clamped": [int(x) for x in o[:, 1]],
                    "lpf": [hx(float(x)) for x in o[:, 2]]})
        print(name, len(o), "targets,", int(o[:, 1].sum()), "clamped")
    with open(os.path.join(HERE, "lpf_golden.json"), "w") as f:
        json.dump({"generator": "tests/golden/gen_lpf_golden.py",
                   "engine": "reference lib/pool.js:44-100 (FIRFilter, genTaps) under SpiderMonkey 1.7 (oracle/_ref); "
                             "Math.exp is the host libm's",
                   "lp_int_ms": 200, "taps": [hx(float(t)) for t in taps], "cases": out}, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
