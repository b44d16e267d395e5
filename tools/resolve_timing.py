"""tools/resolve_timing.py -- phase breakdown of the resolve kernel (globaltimer stamps)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import cueball_b200 as cb
from cueball_b200 import tracegen  # noqa: E402  (input generator only)
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
w = bench.WORKLOADS[wl]
n = w["n"]
start, now = bench.make_trace(wl, n * 4, 0)
ctx = cb.Context(0)
ctx.codel_init(5)
ctx.backends_set(w["B"], tracegen.gen_insert_draws(w["B"]), w["S"])
dev = torch.device("cuda:0")
ds, dn = torch.from_numpy(start).to(dev), torch.from_numpy(now).to(dev)
dd = torch.empty(n * 4, dtype=torch.uint8, device=dev)
sl = torch.empty(n * 4, dtype=torch.int32, device=dev)
names = ["prep:entry->pair", "pair->search", "search->merge", "merge->spec", "spec->end", "(gap)", "commit"]
for b in range(4):
    lo, hi = b * n, (b + 1) * n
    ctx.admit_batch_dev(ds[lo:hi], dn[lo:hi], None, dd[lo:hi], sl[lo:hi])
    t = ctx.debug_timestamps()
    seq = [t[6], t[0], t[1], t[2], t[3], t[4], t[7], t[5]]
    st = ctx.stats()
    print(wl, "batch", b, "events", st.n_events, "intervals", st.n_intervals,
          " ".join("%s=%.1fus" % (nm, (seq[i + 1] - seq[i]) / 1e3) for i, nm in enumerate(names)),
          "total=%.1fus" % ((seq[-1] - seq[0]) / 1e3))
