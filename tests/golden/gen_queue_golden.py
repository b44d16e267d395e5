#!/usr/bin/env python3
"""tests/golden/gen_queue_golden.py -- operation scripts run on the REFERENCE'S OWN Queue (lib/queue.js, verbatim,
under the SpiderMonkey engine vendored in the reference tree; cb_queue_ops in oracle/js/driver.js).  Pins the host
mirror node-cueball_b200/queue.py.  Needs /root/reference; writes tests/golden/queue_golden.json."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref  # noqa: E402
from util import gen_queue_ops  # noqa: E402


def main():
    cases = []
    for seed, n in ((5001, 60), (5002, 400), (5003, 1500), (5004, 1500)):
        op, arg = gen_queue_ops(seed, n)
        cases.append({"seed": seed, "op": op, "arg": arg, "out": ref.queue_ops(op, arg)})
    with open(os.path.join(HERE, "queue_golden.json"), "w") as f:
        json.dump({"generator": "tests/golden/gen_queue_golden.py",
                   "engine": "reference lib/queue.js under SpiderMonkey 1.7 (oracle/_ref)", "cases": cases}, f,
                  separators=(",", ":"))


if __name__ == "__main__":
    main()
