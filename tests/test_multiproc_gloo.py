"""world_size-2 CPU test (gloo) of the multi-GPU host logic: every rank owns a contiguous
queue segment with its own ControlledDelay and the job's global drop/admit counts come from
ONE sum-allreduce of three integers per step -- exactly what cb_allreduce_counts does with NCCL
on the GPUs (bench.py --gpus N).  On CPU the per-rank admission is stood in by the oracle; the
test checks the sharding (bench.make_trace is per-rank seeded and deterministic) and that the
reduced counts equal the counts of the segments processed one after the other."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, steps, out):
    sys.path.insert(0, ROOT)
    import bench
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, now = bench.make_trace("c4", n * steps, rank)
    cd = O.new_codel(bench.TARGET_MS)
    cursor, tot = 0, np.zeros(3, np.int64)
    for s in range(steps):
        lo, hi = s * n, (s + 1) * n
        _, _, st = O.drain(cd, start[lo:hi], now[lo:hi], nslots=64, cursor=cursor, want_out=False)
        cursor = st.cursor
        counts = torch.tensor([st.drops, st.admits, n], dtype=torch.int64)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)  # the single collective of a step
        tot += counts.numpy()
    if rank == 0:
        out.put(tot.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_segment_sharding_and_count_allreduce():
    sys.path.insert(0, ROOT)
    import bench
    from oracle import oracle as O
    world, n, steps = 2, 40_000, 3
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, steps, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.zeros(3, np.int64)
    segs = []
    for r in range(world):
        start, now = bench.make_trace("c4", n * steps, r)
        segs.append((start, now))
        cd = O.new_codel(bench.TARGET_MS)
        _, _, st = O.drain(cd, start, now, nslots=64, want_out=False)
        want += np.array([st.drops, st.admits, n * steps])
    assert got == want.tolist()
    # segments of different ranks are different traces (per-rank seeds), same rank reproducible
    assert not np.array_equal(segs[0][1], segs[1][1])
    again = bench.make_trace("c4", n * steps, 1)
    assert np.array_equal(again[1], segs[1][1])
