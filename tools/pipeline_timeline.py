#!/usr/bin/env python3
"""tools/pipeline_timeline.py [workload] [depth] [batches] -- device timeline of the last `depth`
batches of a pipelined stream (globaltimer stamps the kernels leave in each lane's control block):
when each batch's scan / prepare / commit / emit started and ended, relative to the first."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import cueball_b200 as cb
from cueball_b200 import tracegen

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 3
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 24
w = bench.WORKLOADS[wl]
n = w["n"]
start, now = bench.make_trace(wl, n * nb, 0)
ctx = cb.Context(0)
ctx.codel_init(5)
ctx.backends_set(w["B"], tracegen.gen_insert_draws(w["B"]), w["S"])
ctx.set_pipeline_depth(depth)
dev = torch.device("cuda:0")
ds, dn = torch.from_numpy(start).to(dev), torch.from_numpy(now).to(dev)
dd = torch.empty(n * nb, dtype=torch.uint8, device=dev)
sl = torch.empty(n * nb, dtype=torch.int32, device=dev)
torch.cuda.synchronize()
ps, pn, pd, pl = ds.data_ptr(), dn.data_ptr(), dd.data_ptr(), sl.data_ptr()
for b in range(nb):
    lo = b * n
    ctx.admit_batch_ptr(ps + 8 * lo, pn + 8 * lo, None, n, pd + lo, pl + 4 * lo, None)
ctx.synchronize()
rows = []
for lane in range(depth):
    t = ctx.debug_lane_timestamps(lane)
    rows.append((t[8], lane, t))
rows.sort()
t0 = rows[0][0]
print("%s depth %d, %d batches of %d claims; times in us relative to the first listed batch's scan start" % (wl, depth, nb, n))
print("lane  scan[start end]   prepare[start end]   commit[start handed-over end]   emit[start end]")
prev_commit_end = None
for _, lane, t in rows:
    f = lambda k: (t[k] - t0) / 1e3
    print("%4d  %7.1f %7.1f   %7.1f %7.1f   %7.1f %7.1f %7.1f   %7.1f %7.1f" % (lane, f(8), f(9), f(6), f(4), f(7), f(12), f(5), f(10), f(11)))
