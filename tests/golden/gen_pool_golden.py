#!/usr/bin/env python3
"""tests/golden/gen_pool_golden.py -- golden vectors for the event-driven pool model (SURVEY 8(f)
row 1: claim() enqueue side, idle-first fast path, slot-idle drain, hard timeouts), produced by
running the REFERENCE'S OWN lib/codel.js + lib/queue.js under the SpiderMonkey engine vendored in
the reference tree (oracle/_ref) through cb_pool_sim (oracle/js/driver.js).

Needs /root/reference (this container only); writes tests/golden/pool_sim_golden.json, which holds
the event scripts themselves, so the tests never regenerate them.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref  # noqa: E402
from util import gen_pool_events, state_hex  # noqa: E402  (tests/util.py)

CASES = [
    # name, seed, events, target, backlog (x target), jitter, nslots, idle0
    ("steady_in_band_t10", 2000, 2500, 10, 2.0, 0.1, 13, 0),
    ("steady_in_band_t5", 2002, 2500, 5, 2.7, 0.0, 13, 0),
    ("tight_t2", 2006, 2500, 2, 2.0, 0.5, 7, 0),
    ("deep_backlog_t20", 2008, 2500, 20, 5.0, 0.1, 64, 0),
    ("idle_slots_first_t5", 2011, 1500, 5, 1.3, 0.5, 5, 3),
    ("pauses_and_stalls_t10", 2017, 3000, 10, 2.7, 0.5, 13, 1),
    ("light_load_t1", 2023, 1200, 1, 1.3, 0.1, 3, 2),
    ("overload_t5", 2031, 3000, 5, 5.0, 0.0, 13, 0),
    ("hard_timeouts_t5", 2041, 3000, 5, 12.0, 0.1, 13, 0),
    ("hard_timeouts_t2", 2043, 2500, 2, 9.5, 0.5, 9, 0),
    ("deep_backlog_t10", 2059, 3500, 10, 5.0, 0.1, 13, 0),
    ("deep_backlog_t5_jitter", 2061, 3500, 5, 8.0, 0.5, 31, 0),
]


def main():
    out = []
    for name, seed, nev, tgt, backlog, jitter, nslots, idle0 in CASES:
        ty, a, b = gen_pool_events(seed, nev, tgt, backlog, jitter)
        r = ref.pool_sim(tgt, nslots, ty, a, b, idle0=idle0)
        codes = [0, 0, 0, 0]
        for k in range(1, len(r["res"]), 3):
            codes[r["res"][k]] += 1
        print("%-24s events=%d claims=%d  admitted/codel-drop/skipped/hard-timeout = %s  final count=%s" %
              (name, len(ty), r["counters"]["claim"], codes, r["state"]["count"]))
        out.append({"name": name, "target": tgt, "nslots": nslots, "idle0": idle0,
                    "ev_type": ty, "ev_a": a, "ev_b": b,
                    "expect": {"res": r["res"], "timeouts": r["timeouts"], "state": state_hex(r["state"]),
                               "counters": r["counters"], "idle": r["idle"], "cursor": r["cursor"],
                               "queued": r["queued"]}})
    with open(os.path.join(HERE, "pool_sim_golden.json"), "w") as f:
        json.dump({"generator": "tests/golden/gen_pool_golden.py",
                   "engine": "reference lib/codel.js + lib/queue.js under SpiderMonkey 1.7 (oracle/_ref), cb_pool_sim",
                   "cases": out}, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
