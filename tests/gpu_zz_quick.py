#!/usr/bin/env python3
"""tests/gpu_zz_quick.py -- the GPU tests written after the round's last full GPU session, called directly (no
pytest, no torch import: a fresh box pages those in for about a minute) so that they fit in the seconds of GPU
budget that were left.  Appends one JSON line per scenario to gpurun_out/zz_quick.jsonl as it goes."""
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
OUT = open(os.path.join(ROOT, "gpurun_out", "zz_quick.jsonl"), "a")
T0 = time.time()


def note(name, **kw):
    kw.update(name=name, t=round(time.time() - T0, 2))
    OUT.write(json.dumps(kw) + "\n")
    OUT.flush()
    os.fsync(OUT.fileno())
    print(kw, flush=True)


def main():
    note("start")
    import cueball_b200 as cb
    note("imported")
    import test_gpu_pool_mirror as GP
    import test_gpu_zz_codel_load_runs as ZL
    import test_gpu_zz_sizing as ZS
    todo = [("kang_snapshot", lambda: GP.test_kang_snapshot_shape(cb)),
            ("sizing_loop", lambda: ZS.sizing_scenario(cb) and None),
            ("implicit_high_timeout", lambda: ZL.implicit_high_timeout(cb)),
            ("config_c1", lambda: {k: v for k, v in ZL.config_c1(cb).items() if k not in ("delays", "pool")})]
    if os.environ.get("CB_ZZ_ONLY"):
        todo = [t for t in todo if t[0] in os.environ["CB_ZZ_ONLY"].split(",")]
    for d in (() if os.environ.get("CB_ZZ_ONLY") else (300, 5000, 1000, 500, 1500, 2000, 2500)):
        todo.append(("load_run_%d" % d, (lambda d=d: {k: v for k, v in ZL.load_run(cb, d).items() if k not in ("delays", "pool")})))
    bad = 0
    for name, fn in todo:
        t = time.time()
        try:
            r = fn()
            note(name, ok=True, seconds=round(time.time() - t, 2), result=r)
        except BaseException as e:  # noqa: B902
            bad += 1
            note(name, ok=False, seconds=round(time.time() - t, 2), error=repr(e)[:400], tb=traceback.format_exc()[-1500:])
    note("done", failed=bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
