"""GPU parity tests of the batch pipeline under streams of small batches: the resolve kernel's F
hand-over (first_above_time handed from batch to batch ahead of the rest of the ControlledDelay
state; through the mailboxes in a ring), every pipeline depth, host-side state changes between
batches, and both launch modes (stream launches / one CUDA graph per lane whose node parameters are
rewritten per batch).  First run on a B200 at the start of round 2 (profiles/r02_validate_pending.txt)."""
import numpy as np
import pytest

import util
from cueball_b200 import tracegen
from oracle import oracle as O

pytestmark = [pytest.mark.gpu]


@pytest.fixture()
def cb():
    import cueball_b200
    return cueball_b200


@pytest.mark.parametrize("depth", [1, 3, 8, 12])
@pytest.mark.parametrize("with_flags", [False, True], ids=["noflags", "flags"])
def test_fchain_streams_of_small_batches(cb, depth, with_flags):
    import torch
    n, m, tgt, nslots = 2_000_000, 31_250, 5, 61
    start, now = tracegen.gen_trace(n, seed=177, de=1_000, unit=1 << 13)
    flags = util.make_flags(n, 5, cancel_mod=6, empty_mod=30_011) if with_flags else None
    ctx = cb.Context(0)
    ctx.codel_init(tgt)
    ctx.backends_set(8, tracegen.gen_insert_draws(8), nslots)
    ctx.set_pipeline_depth(depth)
    dev = torch.device("cuda:0")
    ts, tn = torch.from_numpy(start).to(dev), torch.from_numpy(now).to(dev)
    tf = torch.from_numpy(flags).to(dev) if with_flags else None
    dec = torch.empty(n, dtype=torch.uint8, device=dev)
    slot = torch.empty(n, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    for lo in range(0, n, m):
        k = min(m, n - lo)
        ctx.admit_batch_ptr(ts.data_ptr() + 8 * lo, tn.data_ptr() + 8 * lo, tf.data_ptr() + lo if with_flags else None, k,
                            dec.data_ptr() + lo, slot.data_ptr() + 4 * lo, None)
    st = ctx.stats()
    cd = O.new_codel(tgt)
    od, os_, ost = O.drain(cd, start, now, flags, nslots=nslots)
    assert np.array_equal(dec.cpu().numpy(), od)
    assert np.array_equal(slot.cpu().numpy(), os_)
    util.assert_state_equal({"count": st.state.count, "drop_next": st.state.drop_next,
                             "first_above_time": st.state.first_above_time, "dropping": bool(st.state.dropping),
                             "last_empty": st.state.last_empty if st.state.has_last_empty else None}, cd.as_dict())
    assert st.cursor == ost.cursor and st.total_drops == ost.drops


def test_fchain_after_host_side_state_changes(cb):
    """set_state / empty() between batches reseed the hand-over slot."""
    tgt, nslots, n = 5, 13, 400_000
    start, now = tracegen.gen_trace(n, seed=178, de=2_000, unit=1 << 12, start_overloaded=True)
    ctx = cb.Context(0)
    ctx.codel_init(tgt)
    ctx.backends_set(4, tracegen.gen_insert_draws(4), nslots)
    cd = O.new_codel(tgt)
    state = {"count": 5, "drop_next": float(now[0]) / 1e6 - 30.0, "first_above_time": float(now[0]) / 1e6 - 250.0,
             "dropping": True, "last_empty": None}
    for lo, hi in [(0, 150_000), (150_000, 260_000), (260_000, n)]:
        ctx.set_state(state)
        cd = O.new_codel(tgt, state)
        dec, slot, _, st = ctx.admit_batch(start[lo:hi], now[lo:hi])
        od, os_, _ = O.drain(cd, start[lo:hi], now[lo:hi], None, nslots=nslots, cursor=st.cursor - int((dec == 0).sum()))
        assert np.array_equal(dec, od)
        state = cd.as_dict()
        state["count"] = int(state["count"])


@pytest.mark.parametrize("R", [2, 4])
def test_fchain_one_queue_over_a_ring_of_contexts(cb, R):
    """The F-chain slots travel through the ring mailboxes like the rest of the state."""
    import torch
    tgt, nslots, steps = 5, 97, 6
    rng = np.random.RandomState(500 + R)
    seg = [[int(x) for x in rng.randint(1, 120_000, size=R)] for _ in range(steps)]
    n = sum(sum(x) for x in seg)
    start, now = tracegen.gen_trace(n, seed=291 + R, de=1_500, unit=1 << 13)
    dev = torch.device("cuda:0")
    ts, tn = torch.from_numpy(start).to(dev), torch.from_numpy(now).to(dev)
    dec = torch.empty(n, dtype=torch.uint8, device=dev)
    slot = torch.empty(n, dtype=torch.int32, device=dev)
    ctxs = []
    for r in range(R):
        c = cb.Context(0)
        c.codel_init(tgt)
        c.backends_set(16, tracegen.gen_insert_draws(16), nslots)
        ctxs.append(c)
    boxes = [c.ring_mailbox() for c in ctxs]
    for r, c in enumerate(ctxs):
        c.ring_connect(R, r, next_ptr=boxes[(r + 1) % R][1])
    torch.cuda.synchronize()
    lo = 0
    for k in range(steps):
        for r in range(R):
            m = seg[k][r]
            ctxs[r].admit_batch_ptr(ts.data_ptr() + 8 * lo, tn.data_ptr() + 8 * lo, None, m,
                                    dec.data_ptr() + lo, slot.data_ptr() + 4 * lo, None)
            lo += m
    st = ctxs[R - 1].stats()
    for c in ctxs:
        c.synchronize()
    cd = O.new_codel(tgt)
    od, os_, ost = O.drain(cd, start, now, None, nslots=nslots)
    assert np.array_equal(dec.cpu().numpy(), od)
    assert np.array_equal(slot.cpu().numpy(), os_)
    assert st.cursor == ost.cursor and st.total_drops == ost.drops


def test_three_launches_per_batch():
    """scan, resolve (prepare + commit in one launch), emit: 3 launches per batch."""
    import torch
    import cueball_b200 as cbm
    n, m, tgt, nslots = 1_500_000, 46_875, 5, 61
    start, now = tracegen.gen_trace(n, seed=377, de=1_000, unit=1 << 13)
    flags = util.make_flags(n, 7, cancel_mod=6, empty_mod=30_011)
    for fl in (None, flags):
        ctx = cbm.Context(0)
        ctx.codel_init(tgt)
        ctx.backends_set(8, tracegen.gen_insert_draws(8), nslots)
        dev = torch.device("cuda:0")
        ts, tn = torch.from_numpy(start).to(dev), torch.from_numpy(now).to(dev)
        tf = torch.from_numpy(fl).to(dev) if fl is not None else None
        dec = torch.empty(n, dtype=torch.uint8, device=dev)
        slot = torch.empty(n, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        l0 = ctx.launch_count()
        for lo in range(0, n, m):
            ctx.admit_batch_ptr(ts.data_ptr() + 8 * lo, tn.data_ptr() + 8 * lo, tf.data_ptr() + lo if fl is not None else None,
                                min(m, n - lo), dec.data_ptr() + lo, slot.data_ptr() + 4 * lo, None)
        st = ctx.stats()
        assert ctx.launch_count() - l0 <= 3 * ((n + m - 1) // m) + 1
        cd = O.new_codel(tgt)
        od, os_, ost = O.drain(cd, start, now, fl, nslots=nslots)
        assert np.array_equal(dec.cpu().numpy(), od)
        assert np.array_equal(slot.cpu().numpy(), os_)
        assert st.cursor == ost.cursor and st.total_drops == ost.drops


@pytest.mark.parametrize("graph", ["0", "1"], ids=["streams", "graph"])
def test_launch_modes(monkeypatch, graph):
    """CB_GRAPH=1 forces the graph launch mode (a batch's kernels are one CUDA graph per lane and variant whose
    node parameters are rewritten per batch: one launch call per batch), CB_GRAPH=0 plain stream launches; by
    default the library picks by batch size.  Batches of different sizes, with and without flags."""
    import torch
    monkeypatch.setenv("CB_GRAPH", graph)
    import cueball_b200 as cbm
    n, tgt, nslots = 1_600_000, 5, 61
    start, now = tracegen.gen_trace(n, seed=477, de=1_000, unit=1 << 13)
    flags = util.make_flags(n, 9, cancel_mod=6, empty_mod=30_011)
    rng = np.random.RandomState(1)
    cuts = sorted(set([0, n] + [int(x) for x in rng.randint(1, n, size=40)]))
    for fl in (None, flags):
        ctx = cbm.Context(0)
        ctx.codel_init(tgt)
        ctx.backends_set(8, tracegen.gen_insert_draws(8), nslots)
        dev = torch.device("cuda:0")
        ts, tn = torch.from_numpy(start).to(dev), torch.from_numpy(now).to(dev)
        tf = torch.from_numpy(fl).to(dev) if fl is not None else None
        dec = torch.empty(n, dtype=torch.uint8, device=dev)
        slot = torch.empty(n, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            ctx.admit_batch_ptr(ts.data_ptr() + 8 * lo, tn.data_ptr() + 8 * lo, tf.data_ptr() + lo if fl is not None else None,
                                hi - lo, dec.data_ptr() + lo, slot.data_ptr() + 4 * lo, None)
        st = ctx.stats()
        cd = O.new_codel(tgt)
        od, os_, ost = O.drain(cd, start, now, fl, nslots=nslots)
        assert np.array_equal(dec.cpu().numpy(), od)
        assert np.array_equal(slot.cpu().numpy(), os_)
        assert st.cursor == ost.cursor and st.total_drops == ost.drops


@pytest.mark.parametrize("R,steps", [(2, 5), (3, 80)])
def test_global_counts_over_peer_boards(cb, R, steps):
    """cb_peers_*: every batch's {drops, admits, n} is stored into every rank's board over peer memory and
    summed there; the running totals equal the sum of the ranks' own totals (80 steps wrap the 64-slot boards,
    so the flow-control words are exercised too).  Contexts of one process stand in for the ranks."""
    import torch
    tgt, nslots, m = 5, 61, 20_000
    dev = torch.device("cuda:0")
    ctxs, traces, outs = [], [], []
    for r in range(R):
        c = cb.Context(0)
        c.codel_init(tgt)
        c.backends_set(8, tracegen.gen_insert_draws(8), nslots)
        ctxs.append(c)
        s, n = tracegen.gen_trace(steps * m, seed=900 + r, de=1_000 + 300 * r, unit=1 << 12)
        traces.append((s, n, torch.from_numpy(s).to(dev), torch.from_numpy(n).to(dev)))
        outs.append((torch.empty(steps * m, dtype=torch.uint8, device=dev), torch.empty(steps * m, dtype=torch.int32, device=dev)))
    boards = [c.peers_board() for c in ctxs]
    for r, c in enumerate(ctxs):
        c.peers_connect(R, r, ptrs=[b[1] for b in boards])
    torch.cuda.synchronize()
    for k in range(steps):
        for r, c in enumerate(ctxs):
            _, _, ts, tn = traces[r]
            dec, slot = outs[r]
            lo = k * m
            c.admit_batch_ptr(ts.data_ptr() + 8 * lo, tn.data_ptr() + 8 * lo, None, m, dec.data_ptr() + lo,
                              slot.data_ptr() + 4 * lo, None)
    want_d = want_a = 0
    for r in range(R):
        cd = O.new_codel(tgt)
        od, _, ost = O.drain(cd, traces[r][0], traces[r][1], None, nslots=nslots)
        assert np.array_equal(outs[r][0].cpu().numpy(), od)
        want_d += ost.drops
        want_a += ost.admits
    for c in ctxs:
        g = c.peers_counts()
        assert (g.steps, g.drops, g.admits, g.n) == (steps, want_d, want_a, R * steps * m)
    for c in ctxs:
        c.peers_connect(0, 0)
        c.close()


def test_a_stalled_rank_times_out_instead_of_hanging(cb, monkeypatch):
    """One queue over a ring of two contexts in which rank 0 never submits its segment: rank 1's commit waits on
    its mailbox, gives up after the hand-over timeout (CB_PEER_TIMEOUT_MS; 30 s by default) and the next stats
    read reports the broken chain (CB_ESTATE) instead of hanging the device."""
    import time
    import torch
    monkeypatch.setenv("CB_PEER_TIMEOUT_MS", "200")
    tgt, m = 5, 50_000
    start, now = tracegen.gen_trace(m, seed=77, de=1_000, unit=1 << 12)
    dev = torch.device("cuda:0")
    ts, tn = torch.from_numpy(start).to(dev), torch.from_numpy(now).to(dev)
    dec = torch.empty(m, dtype=torch.uint8, device=dev)
    ctxs = []
    for r in range(2):
        c = cb.Context(0)
        c.codel_init(tgt)
        ctxs.append(c)
    boxes = [c.ring_mailbox() for c in ctxs]
    for r, c in enumerate(ctxs):
        c.ring_connect(2, r, next_ptr=boxes[(r + 1) % 2][1])
    t0 = time.time()
    ctxs[1].admit_batch_ptr(ts.data_ptr(), tn.data_ptr(), None, m, dec.data_ptr(), None, None)  # rank 0 never runs
    with pytest.raises(cb.CueballError) as e:
        ctxs[1].stats()
    assert "chain broken" in str(e.value) and time.time() - t0 < 10.0
    for c in ctxs:
        c.ring_connect(0, 0)
        c.close()
