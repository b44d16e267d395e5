#!/usr/bin/env python3
"""tools/e2e_breakdown.py -- where the host-buffer entry point's time goes at a given batch size:
raw pinned H2D / D2H copy times next to cb_admit_batch for several staging chunk sizes."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
chunks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]

from cueball_b200 import tracegen  # noqa: E402  (input generator only)
import bench  # noqa: E402

if os.environ.get("NO_BIND") is None:
    print("bound to %s CPUs next to GPU 0" % bench.bind_to_gpu_numa(0))


def t(f, reps=20):
    f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


dev = torch.device("cuda", 0)
h = torch.empty(2 * n, dtype=torch.int64).pin_memory()
d = torch.empty(2 * n, dtype=torch.int64, device=dev)
ho = torch.empty(5 * n, dtype=torch.uint8).pin_memory()
do = torch.empty(5 * n, dtype=torch.uint8, device=dev)
print("n=%d  H2D %.1f MB: %.3f ms   D2H %.1f MB: %.3f ms" % (
    n, 16 * n / 1e6, t(lambda: d.copy_(h, non_blocking=True)), 5 * n / 1e6, t(lambda: ho.copy_(do, non_blocking=True))))

s, nw = tracegen.gen_trace(n, seed=5, t0=10 ** 12, de=1000, target_ns=5_000_000, unit=1 << 19)
hs = torch.from_numpy(s).pin_memory().numpy()
hn = torch.from_numpy(nw).pin_memory().numpy()
hdec = torch.empty(n, dtype=torch.uint8).pin_memory().numpy()
hslot = torch.empty(n, dtype=torch.int32).pin_memory().numpy()
for ch in chunks:
    if ch:
        os.environ["CB_CHUNK"] = str(ch)
    else:
        os.environ.pop("CB_CHUNK", None)
    import cueball_b200 as cb
    ctx = cb.Context(0)
    ctx.codel_init(5.0)
    ctx.backends_set(64, tracegen.gen_insert_draws(64), 256)

    def call():
        ctx.set_state({"count": 0, "drop_next": 0.0, "first_above_time": 0.0, "dropping": False, "last_empty": None})
        ctx.admit_batch(hs, hn, out_decision=hdec, out_slot=hslot)

    def call_nostate():
        ctx.admit_batch(hs, hn, out_decision=hdec, out_slot=hslot)

    print("chunk=%-8s admit_batch %.3f ms (with set_state %.3f ms) -> %.3g claims/s" % (
        ch or "default", t(call_nostate), t(call), n / t(call_nostate) * 1e3))
    if n > (1 << 22):
        del ctx
        continue
    # the bench's form: rotating slices of one long pinned trace, a handful of calls, each timed
    kb = 10
    sl, nl = tracegen.gen_trace(n * kb, seed=6, t0=10 ** 12, de=1000, target_ns=5_000_000, unit=1 << 19)
    hsl = torch.from_numpy(sl).pin_memory().numpy()
    hnl = torch.from_numpy(nl).pin_memory().numpy()
    per = []
    torch.cuda.synchronize()
    for b in range(kb):
        t0 = time.perf_counter()
        _, _, _, est = ctx.admit_batch(hsl[b * n:(b + 1) * n], hnl[b * n:(b + 1) * n], out_decision=hdec, out_slot=hslot)
        per.append(((time.perf_counter() - t0) * 1e3, est.path, est.drops))
    print("   rotating slices, per call ms(path,drops):", " ".join("%.3f(%d,%d)" % x for x in per))
    del ctx
