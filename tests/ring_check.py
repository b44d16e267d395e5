#!/usr/bin/env python3
"""tests/ring_check.py -- one logical queue over the GPUs of a box, checked and timed.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
      --master-port 29533 tests/ring_check.py [claims_per_segment] [steps]

Every rank owns segment (k, rank) of step k; the ControlledDelay state and admit cursor go round the
ring of mailboxes over NVLink peer memory (CUDA IPC).  Rank 0 then drains the WHOLE queue with the
oracle (tests/ infrastructure) and every rank compares its decisions and slots bit for bit."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/ -> repo root
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

import cueball_b200 as cb
from cueball_b200 import tracegen

m = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 0  # untimed leading steps (timing runs)
check = os.environ.get("RING_CHECK", "1") != "0"
R = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
tgt, B, S = 5, 4096, 16384
n = m * R * steps
if check:
    start, now = tracegen.gen_trace(n, seed=4242, t0=10 ** 12, de=1 << 9, target_ns=tgt * 1_000_000, unit=1 << 19)
    # this rank's segments, contiguous in its own device buffers
    idx = np.concatenate([np.arange((k * R + rank) * m, (k * R + rank + 1) * m) for k in range(steps)])
    ds, dn = torch.from_numpy(start[idx]).to(dev), torch.from_numpy(now[idx]).to(dev)
else:
    # timing only: every rank makes its own segments (the state still goes round the ring; the
    # segments just do not come from one consistent trace, which costs nothing and checks nothing)
    s_r, n_r = tracegen.gen_trace(m * steps, seed=4242 + 7919 * rank, t0=10 ** 12, de=1 << 9,
                                  target_ns=tgt * 1_000_000, unit=1 << 19)
    ds, dn = torch.from_numpy(s_r).to(dev), torch.from_numpy(n_r).to(dev)
dd = torch.empty(m * steps, dtype=torch.uint8, device=dev)
sl = torch.empty(m * steps, dtype=torch.int32, device=dev)
ctx = cb.Context(local)
ctx.codel_init(tgt)
ctx.backends_set(B, tracegen.gen_insert_draws(B), S)
handle, _ = ctx.ring_mailbox()
handles = [None] * R
dist.all_gather_object(handles, handle)
ctx.ring_connect(R, rank, next_handle=handles[(rank + 1) % R] if R > 1 else None)
stream = torch.cuda.Stream(device=dev)
ctx.set_stream(stream.cuda_stream)
dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for k in range(steps):
    if k == warm:
        ctx.join()
        e0.record(stream)
    lo = k * m
    ctx.admit_batch_ptr(ds.data_ptr() + 8 * lo, dn.data_ptr() + 8 * lo, None, m, dd.data_ptr() + lo, sl.data_ptr() + 4 * lo, None)
ctx.join()
e1.record(stream)
dist.barrier()
torch.cuda.synchronize()
st = ctx.stats()
t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
ok = True
if check:
    from oracle import oracle as O
    cd = O.new_codel(tgt)
    od, os_, ost = O.drain(cd, start, now, None, nslots=S)
    ok = bool(np.array_equal(dd.cpu().numpy(), od[idx]) and np.array_equal(sl.cpu().numpy(), os_[idx]))
    if rank == R - 1:
        ref = cd.as_dict()
        ok = ok and st.cursor == ost.cursor and st.total_drops == ost.drops and st.state.count == ref["count"]
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    ms = float(t.item())
    print({"one_queue_ring": True, "n_gpus": R, "claims_per_segment": m, "steps": steps,
           "ms_per_step": ms / (steps - warm), "claims_per_s": m * R * (steps - warm) / (ms / 1e3),
           "bit_exact_vs_single_queue_oracle": bool(flag.item()) if check else None})
dist.destroy_process_group()
