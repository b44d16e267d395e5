#!/usr/bin/env python3
"""tests/golden/gen_golden.py -- (re)generate the committed golden vectors by
running the REFERENCE'S OWN JavaScript (lib/codel.js, lib/queue.js,
lib/utils.js, test/utils.test.js -- verbatim) under the SpiderMonkey engine
vendored in the reference tree (built into oracle/_ref by `make -C oracle ref`).

Needs /root/reference (this container only).  The GPU box never runs this; it
reads the JSON files written here:

    codel_golden.json        drain-loop decisions, slots, CoDel state (bit-exact)
    plan_rebalance_kat.json  the reference's 20 planRebalance known-answer tests
                             (test/utils.test.js) + randomized cases
    slots_golden.json        preference list / slot table / shuffle vectors

Doubles are stored as C99 hex strings so comparisons are bit-for-bit.
"""
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref, tracegen  # noqa: E402
from util import hx, make_trace, rle, state_hex  # noqa: E402  (tests/util.py)


def run_case(name, target, spec, batches=None, nslots=0, state=None, note=""):
    start, now, flags = make_trace(spec)
    n = len(start)
    batches = batches or [n]
    assert sum(batches) == n
    dec, slots, states = "", [], []
    cursor, lo = 0, 0
    st = state
    totals = {"drops": 0, "admits": 0, "skipped": 0}
    for b in batches:
        r = ref.run_trace(target, start[lo:lo + b], now[lo:lo + b],
                          None if flags is None else flags[lo:lo + b],
                          state=st, nslots=nslots, cursor=cursor, want_slots=True)
        dec += r["dec"]
        slots += r["slots"]
        st = r["state"]
        cursor = r["cursor"]
        states.append(state_hex(st))
        for k in totals:
            totals[k] += r[k]
        lo += b
    sl = np.array(slots, dtype=np.int32)
    case = {"name": name, "note": note, "target": target, "trace": spec, "batches": batches,
            "nslots": nslots, "init_state": None if state is None else state_hex(state),
            "expect": {"dec_rle": rle(dec), "slots_crc32": zlib.crc32(sl.tobytes()) & 0xFFFFFFFF,
                       "slots_head": slots[:64], "states": states, "cursor": cursor, **totals}}
    print("%-28s n=%-7d drops=%-6d admits=%-6d skipped=%-5d final=%s" %
          (name, n, totals["drops"], totals["admits"], totals["skipped"], st))
    return case


def ms(x):
    return int(round(x * 1_000_000))


def codel_cases():
    cases = []
    # --- SURVEY.md Appendix B vectors (hand traces) ---
    s1 = [ms(1000 + i) for i in range(40)]
    n1 = [ms(1100 + 10 * i) for i in range(40)]
    cases.append(run_case("appB1_ramp", 5, {"kind": "explicit", "start_ns": s1, "now_ns": n1}, nslots=4))
    n2 = [ms(1000 + 10 * i) for i in range(45)] + [ms(1450)] + [ms(1451 + 10 * i) for i in range(30)]
    s2 = [t - ms(100) for t in n2]
    s2[45] = ms(1449)
    cases.append(run_case("appB2_reset_reenter", 5, {"kind": "explicit", "start_ns": s2, "now_ns": n2}, nslots=3))
    n3 = [ms(1000 + 10 * i) for i in range(31)] + [ms(1310), ms(1320), ms(1330), ms(1345), ms(1100.25), ms(1250.5)]
    s3 = [0] * len(n3)
    s3[34] = ms(1344)
    cases.append(run_case("appB3_backwards_sqrt3", 5, {"kind": "explicit", "start_ns": s3, "now_ns": n3}, nslots=2,
                          note="non-monotone clock: exercises count-2 / sqrt(count>1)"))
    # same traces split into batches with carried state
    cases.append(run_case("appB1_split", 5, {"kind": "explicit", "start_ns": s1, "now_ns": n1}, batches=[7, 14, 1, 18], nslots=4))
    cases.append(run_case("appB2_split", 5, {"kind": "explicit", "start_ns": s2, "now_ns": n2}, batches=[21, 25, 30], nslots=3))
    # equal timestamps (ties) around F and drop_next
    n4 = [ms(2000)] * 5 + [ms(2100)] * 5 + [ms(2200)] * 5 + [ms(2300)] * 6 + [ms(2300.000001)] * 3
    s4 = [t - ms(5) for t in n4]
    cases.append(run_case("ties_exact_boundaries", 5, {"kind": "explicit", "start_ns": s4, "now_ns": n4}, nslots=5,
                          note="sojourn == target exactly; now == F, now == drop_next exactly"))
    # sojourn just below / at target, sub-microsecond fractions
    n5 = [ms(5000) + 137 * i for i in range(300)]
    s5 = [t - (ms(5) - 1 + (i % 3)) for i, t in enumerate(n5)]
    cases.append(run_case("sojourn_at_target_edge", 5, {"kind": "explicit", "start_ns": s5, "now_ns": n5}, nslots=7))
    # sparse stream: every claim 120 ms apart, alternating above / below (many 1-claim runs)
    n6, s6 = [], []
    for i in range(200):
        t = ms(10_000 + 210 * i)
        n6.append(t)
        s6.append(t - (ms(50) if (i % 3) else ms(1)))
    cases.append(run_case("sparse_aab_pattern", 5, {"kind": "explicit", "start_ns": s6, "now_ns": n6}, nslots=2,
                          note="an ok-run every third claim: adversarial for the episode decomposition"))
    # monotone clock, yet the `count > 2 ? count - 2` re-entry branch is taken: drop_next + 100
    # rounds DOWN across the 1024 binade, so (now - drop_next) < 100 with now == first_above_time
    tB = 830_000_024
    tA, tD, tE = tB - ms(250), tB + ms(100), tB + ms(200)
    n8 = [tA, tB, tD, tD, tD, tD, tD, tE] + [tE + ms(8 * i) for i in range(1, 12)]
    s8 = [t - ms(50) for t in n8]
    s8[5] = n8[5] - ms(1)
    cases.append(run_case("monotone_rounding_count_minus2", 5, {"kind": "explicit", "start_ns": s8, "now_ns": n8}, nslots=3,
                          note="monotone clock reaching count-2 and sqrt(2) via a rounding coincidence"))
    # empty batch / single claims
    cases.append(run_case("empty_batch", 5, {"kind": "explicit", "start_ns": [], "now_ns": []}, nslots=4))
    cases.append(run_case("single_below", 5, {"kind": "explicit", "start_ns": [ms(10)], "now_ns": [ms(11)]}, nslots=4))
    cases.append(run_case("single_above", 5, {"kind": "explicit", "start_ns": [ms(10)], "now_ns": [ms(400)]}, nslots=4))
    # carried-in dropping state, first claim continues the episode
    carried = {"count": 16, "drop_next": 1300.0, "first_above_time": 1100.0, "dropping": True, "last_empty": None}
    n7 = [ms(1290 + 2 * i) for i in range(40)]
    s7 = [t - ms(100) for t in n7]
    s7[30] = n7[30] - ms(1)
    cases.append(run_case("carry_in_dropping", 5, {"kind": "explicit", "start_ns": s7, "now_ns": n7}, nslots=4, state=carried))
    # large-uptime clock (ulp(ms) = 2^-19 ms > 1 ns): 100 days.  (The JS harness carries total
    # ns in a double, so traces must stay below 2^53 ns = 104 days; process.hrtime() itself
    # has no such limit and neither have the C oracle or the CUDA path.)
    t0 = 100 * 86400 * 10**9
    cases.append(run_case("uptime_100d", 5, {"kind": "gen", "n": 6000, "seed": 77, "params":
                          {"t0": t0, "de": 50_000, "target_ns": 5_000_000, "unit": 512}}, nslots=16))
    # --- generated traces (tracegen spec) ---
    P = lambda **kw: dict({"t0": 10**9, "de": 1000, "target_ns": 5_000_000, "unit": 1 << 16}, **kw)  # noqa: E731
    cases.append(run_case("gen_dense_48k", 5, {"kind": "gen", "n": 48_000, "seed": tracegen.SEED,
                          "params": P(de=20_000, unit=2048)}, nslots=256))
    cases.append(run_case("gen_dense_48k_3batches", 5, {"kind": "gen", "n": 48_000, "seed": tracegen.SEED,
                          "params": P(de=20_000, unit=2048)}, batches=[10_001, 20_479, 17_520], nslots=256))
    cases.append(run_case("gen_seed2_flags", 5, {"kind": "gen", "n": 40_000, "seed": 2, "params": P(de=25_000, unit=1024),
                          "flags": {"seed": 5, "cancel_mod": 7, "empty_mod": 0}}, nslots=64,
                          note="cancelled handles still advance CoDel (pool.js:735-740)"))
    cases.append(run_case("gen_seed3_empty_markers", 5, {"kind": "gen", "n": 40_000, "seed": 3, "params": P(de=25_000, unit=1024),
                          "flags": {"seed": 9, "cancel_mod": 11, "empty_mod": 501}}, batches=[13_333, 26_667], nslots=96))
    cases.append(run_case("gen_target1ms", 1, {"kind": "gen", "n": 30_000, "seed": 4,
                          "params": P(de=10_000, target_ns=1_000_000, unit=4096)}, nslots=33))
    cases.append(run_case("gen_target300ms", 300, {"kind": "gen", "n": 20_000, "seed": 5,
                          "params": P(de=2_000_000, target_ns=300_000_000, unit=256)}, nslots=2))
    cases.append(run_case("gen_sustained_overload", 5, {"kind": "gen", "n": 50_000, "seed": 6,
                          "params": P(de=10_000, unit=1024, all_overloaded_from=9_000)}, batches=[20_000, 30_000], nslots=1024,
                          note="C5 shape: drop state, every claim past drop_next dropped"))
    cases.append(run_case("gen_start_overloaded", 5, {"kind": "gen", "n": 30_000, "seed": 7,
                          "params": P(de=15_000, unit=2048, start_overloaded=True)}, nslots=48))
    cases.append(run_case("gen_backstep", 5, {"kind": "gen", "n": 40_000, "seed": 8, "params": P(de=20_000, unit=2048),
                          "backstep": {"every": 9_973, "ns": 150_000_000}}, nslots=64,
                          note="clock steps back 150 ms: non-monotone, sequential path"))
    cases.append(run_case("gen_many_episodes", 1, {"kind": "gen", "n": 60_000, "seed": 10, "params": P(de=400_000, unit=320, target_ns=1_000_000)}, nslots=128,
                          note="24 s of stream: dozens of dropping episodes, re-entries within 100 ms of drop_next"))
    cases.append(run_case("gen_many_episodes_flags", 1, {"kind": "gen", "n": 60_000, "seed": 11, "params": P(de=700_000, unit=192, target_ns=1_000_000),
                          "flags": {"seed": 3, "cancel_mod": 5, "empty_mod": 4001}}, batches=[1, 29_999, 30_000], nslots=100))
    cases.append(run_case("gen_micro_dense", 5, {"kind": "gen", "n": 50_000, "seed": 9, "params": P(de=1000, unit=4096)}, nslots=256,
                          note="1 us arrivals: long tie stretches in now"))
    return cases


def max_idle_cases():
    out = []
    for target, st, now_ms in [
        (100, None, 0.0),
        (100, {"count": 0, "drop_next": 0, "first_above_time": 0, "dropping": False, "last_empty": 5000.0}, 5500.0),
        (100, {"count": 0, "drop_next": 0, "first_above_time": 0, "dropping": False, "last_empty": 5000.0}, 7000.0),
        (5, {"count": 3, "drop_next": 10.0, "first_above_time": 0, "dropping": False, "last_empty": 1000.0}, 1050.0),
        (5, {"count": 3, "drop_next": 10.0, "first_above_time": 0, "dropping": False, "last_empty": 1000.0}, 1050.0000001),
    ]:
        v = ref.max_idle(target, st, ms(now_ms))
        out.append({"target": target, "state": None if st is None else state_hex(st), "now_ns": ms(now_ms), "expect": hx(v)})
    return out


def plan_cases():
    rec = ref.record_utils_tests()
    assert rec["fail"] == 0, rec
    print("reference test/utils.test.js under SM: %d assertions pass, %d planRebalance calls" % (rec["pass"], len(rec["calls"])))
    rng = np.random.RandomState(12345)
    rand = []
    for t in range(400):
        B = int(rng.randint(1, 9))
        have = [int(x) for x in rng.randint(0, 4, size=B)]
        dead = [bool(x) for x in (rng.rand(B) < (0.0 if t % 4 == 0 else 0.4))]
        target = int(rng.randint(0, 12))
        mx = target + int(rng.randint(0, 6))
        single = bool(t % 3 == 0)
        conns = [["k%d" % i, ["k%dc%d" % (i, j) for j in range(have[i])]] for i in range(B)]
        p = ref.plan(conns, ["k%d" % i for i in range(B) if dead[i]], target, mx, single)
        rand.append({"conns": conns, "dead": ["k%d" % i for i in range(B) if dead[i]], "target": target,
                     "max": mx, "singleton": single, "add": p["add"], "remove": p["remove"]})
    return {"kat": rec["calls"], "kat_assertions": rec["pass"], "random": rand}


def slot_cases():
    out = {"build": [], "reshuffle": [], "shuffle": []}
    for B, S, seed in [(4, 2, 1), (4, 256, 2), (64, 256, tracegen.SEED), (7, 23, 3), (100, 64, 4), (1, 5, 5), (1024, 4096, 6)]:
        us = tracegen.gen_insert_draws(B, seed)
        r = ref.build_slots(B, S, us)
        out["build"].append({"B": B, "S": S, "seed": seed, "keys": r["keys"], "slots": r["slots"]})
    for n, seed in [(2, 1), (5, 2), (64, 3)]:
        us = tracegen.gen_insert_draws(8, seed + 100)
        keys = list(range(n))
        trail = []
        for u in us:
            keys = ref.call("cb_reshuffle(%s,%r)" % (json.dumps(keys), float(u)))
            trail.append(list(keys))
        out["reshuffle"].append({"n": n, "seed": seed + 100, "trail": trail})
    for n, seed in [(1, 1), (8, 2), (33, 3)]:
        us = tracegen.gen_insert_draws(n, seed + 200)
        r = ref.call("cb_shuffle(%s,%s)" % (json.dumps(list(range(n))), json.dumps([float(u) for u in us])))
        out["shuffle"].append({"n": n, "seed": seed + 200, "result": r})
    return out


def main():
    if not ref.available():
        sys.exit("oracle/_ref not built: run `make -C oracle ref` (needs /root/reference)")
    with open(os.path.join(HERE, "codel_golden.json"), "w") as f:
        json.dump({"generator": "tests/golden/gen_golden.py", "engine": "reference lib/*.js under SpiderMonkey 1.7 (oracle/_ref)",
                   "cases": codel_cases(), "max_idle": max_idle_cases()}, f, indent=0, separators=(",", ":"))
    with open(os.path.join(HERE, "plan_rebalance_kat.json"), "w") as f:
        json.dump(plan_cases(), f, separators=(",", ":"))
    with open(os.path.join(HERE, "slots_golden.json"), "w") as f:
        json.dump(slot_cases(), f, separators=(",", ":"))


if __name__ == "__main__":
    main()
