#!/usr/bin/env python3
"""tests/fuzz_emulated_pipeline.py <seconds> [first_seed [cases]] -- random small traces (odd sizes, ragged batch
cuts, carried-in states, waiting flags and empty() markers, both commit variants) through the
emulated batch pipeline (tests/emu) against the oracle.  Prints every mismatch and a summary;
round 1 ran about 2 900 cases clean, round 2's final build 14 398 (tests/emu/README.md).  A stalled barrier (lanes that disagree about reaching a collective)
is reported with the stuck threads' stacks by the emulator instead of hanging."""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import numpy as np  # noqa: E402

import util
import test_pipeline_kernels_emulated as T
from cueball_b200 import tracegen
from oracle import oracle as O
L = T.load_emu()
t_end = time.time() + float(sys.argv[1])
max_cases = int(sys.argv[3]) if len(sys.argv) > 3 else None
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
it = 0
bad = 0
paths_seen = {0: 0, 1: 0}
while (time.time() < t_end) if max_cases is None else (it < max_cases):
    seed = seed0 + it; it += 1
    rng = np.random.RandomState(seed)
    n = int(rng.choice([1, 2, 63, 64, 65, 255, 256, 257, 2047, 2048, 2049, 4095, 4096, 4097, int(rng.randint(1, 9000))]))
    tgt = int(rng.choice([1, 2, 5, 20]))
    de = int(rng.choice([200, 3_000, 20_000, 100_000, 400_000]))
    unit = int(rng.choice([64, 512, 1 << 10, 1 << 12]))
    kw = {}
    r = rng.rand()
    if r < 0.25: kw["start_overloaded"] = True
    elif r < 0.5: kw["all_overloaded_from"] = int(rng.randint(0, max(1, n)))
    start, now = tracegen.gen_trace(n, seed=seed, de=de, unit=unit, target_ns=tgt * 1_000_000, **kw)
    if rng.rand() < float(os.environ.get("CB_FUZZ_BACKSTEP", "0.1")) and n > 1:
        # the clock steps back here and there: the batch takes the exact one-warp path (lib/codel.js:72-79's count - 2 /
        # sqrt(count) branch is only reachable this way)
        now = now - (rng.rand(n) < float(rng.choice([0.002, 0.02, 0.2]))).astype(np.int64) * int(rng.choice([1_000, 3_000_000, 150_000_000]))
    flags = None
    if rng.rand() < 0.4:
        flags = util.make_flags(n, seed, cancel_mod=int(rng.choice([0, 3, 9])), empty_mod=int(rng.choice([0, 97, 1501])))
    ncut = int(rng.randint(0, 4))
    cuts = sorted(set([0, n] + [int(x) for x in rng.randint(0, n + 1, size=ncut)]))
    state = None
    if rng.rand() < 0.3 and n > 0:
        t0 = float(now[0]) / 1e6
        state = {"count": int(rng.randint(0, 50)), "drop_next": t0 - float(rng.uniform(-50, 150)),
                 "first_above_time": float(rng.choice([0.0, t0 - float(rng.uniform(-50, 300))])),
                 "dropping": bool(rng.rand() < 0.5), "last_empty": None}
    fchain = int(rng.randint(0, 4))  # resolve-stage variant (mode bits of emu_pipe_create)
    if os.environ.get("CB_FUZZ_VERBOSE"):
        print("case", seed, dict(n=n, tgt=tgt, de=de, unit=unit, kw=kw, flags=flags is not None, cuts=cuts, fchain=fchain,
                                 state=state is not None), flush=True)
    cd = O.new_codel(tgt, state)
    od, os_, ost = O.drain(cd, start, now, flags, nslots=37)
    try:
        dec, slot, (st, cursor, td, ta), paths = T.run_batches(L, start, now, cuts, tgt, 37, state, fchain=fchain, flags=flags)
        ok = np.array_equal(dec, od) and np.array_equal(slot, os_) and cursor == ost.cursor and td == ost.drops
        util.assert_state_equal(st, cd.as_dict())
        for p in paths: paths_seen[p] = paths_seen.get(p, 0) + 1
    except AssertionError as e:
        ok = False
    if not ok:
        bad += 1
        print("MISMATCH seed", seed, dict(n=n, tgt=tgt, de=de, unit=unit, kw=kw, flags=flags is not None, cuts=cuts, state=state, fchain=fchain), flush=True)
print("iterations", it, "bad", bad, "paths", paths_seen)
sys.exit(1 if bad else 0)
