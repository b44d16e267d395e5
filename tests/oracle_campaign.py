#!/usr/bin/env python3
"""tests/oracle_campaign.py <seconds> [first_seed] -- pin the oracle's C restatement to the reference harder than the
test suite has time for: fresh random traces (ragged cuts, carried state, cancelled handles, empty() markers, backwards
clock steps), pool event scripts, planRebalance inputs, sizing scripts and Queue operation scripts through the REFERENCE'S
OWN JavaScript under the engine vendored in the reference tree (oracle/_ref; needs /root/reference to have been built
here) and through oracle/cueball_oracle.c; every output compared bit for bit.  Prints a JSON summary.  TEST INFRASTRUCTURE."""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import numpy as np  # noqa: E402

import util  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import ref, tracegen  # noqa: E402


def one_trace(seed):
    rng = np.random.RandomState(seed)
    n = int(rng.choice([1, 2, 300, 5000, 12000]))
    tgt = int(rng.choice([1, 2, 5, 20]))
    start, now = tracegen.gen_trace(n, seed=seed, de=int(rng.choice([2_000, 90_000, 600_000])),
                                    unit=int(rng.choice([64, 512, 4096])), target_ns=tgt * 1_000_000,
                                    start_overloaded=bool(rng.rand() < 0.3))
    flags = util.make_flags(n, seed, cancel_mod=int(rng.choice([0, 3, 13])), empty_mod=int(rng.choice([0, 97])))
    if rng.rand() < 0.3:
        now = now - (rng.rand(n) < 0.01).astype(np.int64) * 120_000_000
    cuts = sorted(set([0, n] + [int(x) for x in rng.randint(0, n + 1, size=int(rng.randint(0, 3)))]))
    cd = O.new_codel(tgt)
    st, cursor = None, 0
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        r = ref.run_trace(tgt, start[lo:hi], now[lo:hi], flags[lo:hi], state=st, nslots=37, cursor=cursor, want_slots=True)
        d, s, stats = O.drain(cd, start[lo:hi], now[lo:hi], flags[lo:hi], nslots=37, cursor=cursor)
        assert "".join(util.DEC_CHARS[x] for x in d) == r["dec"]
        assert s.tolist() == r["slots"]
        util.assert_state_equal(cd.as_dict(), r["state"], "seed %d" % seed)
        st, cursor = r["state"], r["cursor"]
        assert stats.cursor == cursor
    return n


def one_script(seed):
    rng = np.random.RandomState(seed)
    tgt = int(rng.choice([1, 2, 5, 10, 20]))
    ty, a, b = util.gen_pool_events(seed, int(rng.choice([300, 1500])), tgt, float(rng.choice([1.3, 2.0, 5.0, 9.0, 12.0])),
                                    float(rng.choice([0.0, 0.1, 0.5])))
    nslots, idle0 = int(rng.choice([1, 7, 11, 64])), int(rng.randint(0, 4))
    r = ref.pool_sim(tgt, nslots, ty, a, b, idle0=idle0)
    cd = O.new_codel(tgt)
    res, tmo, out = O.pool_sim(cd, nslots, ty, a, b, idle0=idle0)
    assert res.reshape(-1).tolist() == r["res"]
    assert tmo.tolist() == [float(x) for x in r["timeouts"]]
    assert out.counters() == r["counters"]
    assert (out.idle, out.cursor, out.queued) == (r["idle"], r["cursor"], r["queued"])
    util.assert_state_equal(cd.as_dict(), r["state"], "seed %d" % seed)
    return len(ty)


def one_plan(seed):
    rng = np.random.RandomState(seed)
    B = int(rng.randint(1, 12))
    names = ["b%d" % i for i in range(B)]
    conns = [[k, ["%s.c%d" % (k, j) for j in range(int(rng.randint(0, 5)))]] for k in names]
    dead = [k for k in names if rng.rand() < 0.25]
    mx = int(rng.randint(1, 25))
    target = int(rng.randint(0, mx + 1))
    single = bool(rng.rand() < 0.3)
    r = ref.plan(conns, dead, target, mx, single)
    have = [len(v) for _, v in conns]
    _, add, rem = O.plan_rebalance(have, [1 if k in dead else 0 for k in names], target, mx, single)
    cm = dict((k, v) for k, v in conns)
    assert [names[i] for i in add] == r["add"] and [cm[names[k]][o] for k, o in rem] == r["remove"]
    return 1


def one_lpf(seed):
    rng = np.random.RandomState(seed)
    ty, a, b, c, d = util.gen_lpf_events(seed, int(rng.choice([200, 900])), int(rng.choice([6, 40, 90])))
    sp, mx = int(rng.randint(0, 6)), int(rng.randint(4, 100))
    r = ref.lpf_run(sp, mx, ty, a, b, c, d)
    want = np.array(r["out"], dtype=np.float64).reshape(-1, 3)
    got = O.lpf_run(sp, mx, ty, a, b, c, d)
    assert got.view(np.uint64).tolist() == want.view(np.uint64).tolist()
    return len(ty)


def main():
    if not ref.available():
        raise SystemExit("oracle/_ref is not built (make -C oracle ref, where /root/reference exists)")
    t_end = time.time() + float(sys.argv[1])
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 700000
    kinds = (("traces", one_trace), ("pool_scripts", one_script), ("plans", one_plan), ("sizing_scripts", one_lpf))
    tot = {k: {"cases": 0, "units": 0} for k, _ in kinds}
    bad = []
    while time.time() < t_end:
        for name, fn in kinds:
            try:
                tot[name]["units"] += fn(seed)
                tot[name]["cases"] += 1
            except AssertionError as e:
                bad.append((name, seed, str(e)[:200]))
                print("MISMATCH", name, seed, str(e)[:200], flush=True)
        seed += 1
    print(json.dumps({"totals": tot, "mismatches": bad, "first_seed": int(sys.argv[2]) if len(sys.argv) > 2 else 700000,
                      "next_seed": seed}))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
