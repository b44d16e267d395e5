"""world_size-2 CPU test (gloo) of the multi-GPU host logic: every rank owns a contiguous
queue segment with its own ControlledDelay and the job's global drop/admit counts come from
ONE sum of three integers per step -- what cb_allreduce_counts does with NCCL and what the peer-memory
boards of cb_peers_* do without a collective on the GPUs (bench.py --gpus N).  On CPU the per-rank admission is stood in by the oracle; the
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
    start, now = (t.numpy() for t in bench.make_trace("c4", n * steps, rank, torch.device("cpu")))
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
        start, now = (t.numpy() for t in bench.make_trace("c4", n * steps, r, torch.device("cpu")))
        segs.append((start, now))
        cd = O.new_codel(bench.TARGET_MS)
        _, _, st = O.drain(cd, start, now, nslots=64, want_out=False)
        want += np.array([st.drops, st.admits, n * steps])
    assert got == want.tolist()
    # segments of different ranks are different traces (per-rank seeds), same rank reproducible
    assert not np.array_equal(segs[0][1], segs[1][1])
    again = bench.make_trace("c4", n * steps, 1, torch.device("cpu"))
    assert np.array_equal(again[1].numpy(), segs[1][1])


def _ring_worker(rank, world, port, m, steps, out):
    """One logical queue over `world` ranks (cb_ring_connect): segment (k, r) follows
    (k, r-1); the ControlledDelay state and the admit cursor are handed from rank to rank.  On the
    GPUs the hand-over is a peer-memory mailbox written by the commit kernel; here it is a gloo
    send/recv and the per-segment admission is stood in by the oracle."""
    sys.path.insert(0, ROOT)
    from cueball_b200 import tracegen
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = m * world * steps
    start, now = tracegen.gen_trace(n, seed=99, t0=10 ** 12, de=1 << 9, target_ns=5_000_000, unit=1 << 13)
    cd = O.new_codel(5)
    box = torch.zeros(8, dtype=torch.float64)  # count, drop_next, first_above, dropping, hasE, E, cursor, seq
    dec_all, slot_all = [], []
    for k in range(steps):
        seq = k * world + rank
        if seq > 0:
            dist.recv(box, src=(rank - 1) % world)
            assert int(box[7].item()) == seq  # the mailbox sequence number the commit kernel waits for
            cd.load({"count": box[0].item(), "drop_next": box[1].item(), "first_above_time": box[2].item(),
                     "dropping": bool(box[3].item()), "last_empty": box[5].item() if box[4].item() else None})
        cursor = int(box[6].item())
        lo = seq * m
        d, s, st = O.drain(cd, start[lo:lo + m], now[lo:lo + m], nslots=97, cursor=cursor)
        dec_all.append(d)
        slot_all.append(s)
        stt = cd.as_dict()
        le = stt["last_empty"]
        box = torch.tensor([stt["count"], stt["drop_next"], stt["first_above_time"], float(stt["dropping"]),
                            0.0 if le is None else 1.0, 0.0 if le is None else le, float(st.cursor), float(seq + 1)],
                           dtype=torch.float64)
        if not (k == steps - 1 and rank == world - 1):
            dist.send(box, dst=(rank + 1) % world)
    out.put((rank, np.concatenate(dec_all).tobytes(), np.concatenate(slot_all).tobytes(), cd.as_dict()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_one_logical_queue_protocol():
    sys.path.insert(0, ROOT)
    from cueball_b200 import tracegen
    from oracle import oracle as O
    world, m, steps = 2, 30_000, 4
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_ring_worker, args=(r, world, port, m, steps, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict()
    for _ in range(world):
        r, d, s, st = out.get(timeout=120)
        got[r] = (np.frombuffer(d, np.uint8), np.frombuffer(s, np.int32), st)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = m * world * steps
    start, now = tracegen.gen_trace(n, seed=99, t0=10 ** 12, de=1 << 9, target_ns=5_000_000, unit=1 << 13)
    cd = O.new_codel(5)
    od, os_, ost = O.drain(cd, start, now, nslots=97)  # ONE sequential drain of the whole queue
    for r in range(world):
        idx = np.concatenate([np.arange((k * world + r) * m, (k * world + r + 1) * m) for k in range(steps)])
        assert np.array_equal(got[r][0], od[idx]) and np.array_equal(got[r][1], os_[idx])
    assert got[world - 1][2] == cd.as_dict()  # the last rank's state is the queue's
    assert od.max() >= 1  # the trace does exercise drops
