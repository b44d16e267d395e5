#!/usr/bin/env python3
"""tests/fuzz_emulated_queue.py <seconds> [first_seed] -- random pool event scripts (claims, idle events
with 1..3 slots, cancels, arrival pauses, service stalls) through the emulated queue kernels
(tests/emu, csrc/cb_queue.cuh) against the oracle's pool model.  Prints every mismatch and a summary."""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import numpy as np  # noqa: E402

import util  # noqa: E402
import test_queue_kernels_emulated as T  # noqa: E402

L = T.load_emu()
t_end = time.time() + float(sys.argv[1])
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 500000
it = bad = 0
codes = np.zeros(4, int)
while time.time() < t_end:
    seed = seed0 + it
    it += 1
    rng = np.random.RandomState(seed)
    tgt = int(rng.choice([1, 2, 5, 10, 20]))
    ty, a, b = util.gen_pool_events(seed, int(rng.choice([150, 300, 600])), tgt,
                                    float(rng.choice([1.3, 2.0, 5.0, 9.0, 12.0])), float(rng.choice([0.0, 0.1, 0.5])))
    try:
        T.check(L, tgt, int(rng.choice([1, 7, 64])), ty, a, b, int(rng.randint(0, 4)))
    except AssertionError as e:
        bad += 1
        print("MISMATCH seed", seed, str(e)[:200], flush=True)
print("iterations", it, "bad", bad)
sys.exit(1 if bad else 0)
