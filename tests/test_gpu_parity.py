"""GPU parity tests: the CUDA path, called through the C ABI, against
  (a) the committed golden vectors produced by the reference's own JavaScript,
  (b) the C oracle (itself pinned to the reference) on fresh seeded inputs,
  (c) size-independent properties at BASELINE.json's full sizes.
Bit-exact everywhere: decisions, slot/backend indices, CoDel state (doubles by bit pattern)."""
import os

import numpy as np
import pytest

import util
from oracle import oracle as O
from oracle import tracegen

pytestmark = pytest.mark.gpu

G = util.load_golden("codel_golden.json")
PLAN = util.load_golden("plan_rebalance_kat.json")
SLOTS = util.load_golden("slots_golden.json")


@pytest.fixture(scope="module")
def cb():
    import cueball_b200
    return cueball_b200


def gpu_state(st):
    return st.state.as_dict()


def run_case_host(cb, case, ms=False):
    start, now, flags = util.make_trace(case["trace"])
    if ms:
        to_ms = np.vectorize(O.ns_to_ms, otypes=[np.float64])
        start = to_ms(start) if len(start) else start.astype(np.float64)
        now = to_ms(now) if len(now) else now.astype(np.float64)
    ctx = cb.Context(0)
    ctx.codel_init(case["target"])
    if case["init_state"]:
        ctx.set_state(util.state_unhex(case["init_state"]))
    if case["nslots"]:
        B = max(1, min(case["nslots"], 7))
        ctx.backends_set(B, tracegen.gen_insert_draws(B, 5), case["nslots"])
    decs, slots, states, paths = [], [], [], []
    lo = 0
    for b in case["batches"]:
        d, s, _, st = ctx.admit_batch(start[lo:lo + b], now[lo:lo + b], None if flags is None else flags[lo:lo + b])
        decs.append(d)
        slots.append(s)
        states.append(ctx.get_state().as_dict())
        paths.append(st.path)
        lo += b
    return (np.concatenate(decs) if decs else np.zeros(0, np.uint8),
            np.concatenate(slots) if slots else np.zeros(0, np.int32), states, ctx.stats(), paths)


@pytest.mark.parametrize("ms", [False, True], ids=["ns", "ms"])
@pytest.mark.parametrize("case", G["cases"], ids=[c["name"] for c in G["cases"]])
def test_golden_vectors(cb, case, ms):
    dec, slots, states, st, paths = run_case_host(cb, case, ms)
    exp = case["expect"]
    assert util.codes_to_rle(dec) == exp["dec_rle"]
    assert util.crc_i32(slots) == exp["slots_crc32"]
    assert slots[:64].tolist() == exp["slots_head"]
    for got, want in zip(states, exp["states"]):
        util.assert_state_equal(got, util.state_unhex(want), case["name"])
    if sum(case["batches"]):
        assert st.cursor == exp["cursor"]
        assert st.total_drops == exp["drops"] and st.total_admits == exp["admits"]
    nonmono = "backstep" in case["name"] or "backwards" in case["name"]
    if nonmono:
        assert cb.PATH_SEQUENTIAL in paths


@pytest.mark.parametrize("case", G["max_idle"], ids=lambda c: "t%s_%s" % (c["target"], c["now_ns"]))
def test_max_idle(cb, case):
    ctx = cb.Context(0)
    ctx.codel_init(case["target"])
    if case["state"]:
        ctx.set_state(util.state_unhex(case["state"]))
    assert util.same_bits(ctx.max_idle(case["now_ns"]), util.unhx(case["expect"]))


def test_scalar_overloaded_and_empty(cb):
    """ControlledDelay#overloaded one claim at a time == the batch == the oracle."""
    case = next(c for c in G["cases"] if c["name"] == "appB2_reset_reenter")
    start, now, _ = util.make_trace(case["trace"])
    cd = cb.ControlledDelay(5)
    got = "".join("D" if cd.overloaded(int(s), int(n)) else "A" for s, n in zip(start, now))
    assert util.rle(got) == case["expect"]["dec_rle"]
    util.assert_state_equal(cd.context.get_state().as_dict(), util.state_unhex(case["expect"]["states"][-1]))
    cd.empty(int(now[-1]) + 1000)
    assert cd.cd_first_above_time == 0.0 and cd.cd_last_empty == O.ns_to_ms(int(now[-1]) + 1000)


def oracle_run(start, now, flags, target, nslots, state=None, cursor=0):
    cd = O.new_codel(target, state)
    d, s, st = O.drain(cd, start, now, flags, nslots=nslots, cursor=cursor)
    return d, s, cd.as_dict(), st


@pytest.mark.parametrize("seed,n,de,unit,tgt", [
    (21, 300_000, 3_000, 8192, 5), (22, 1_000_001, 1_000, 1 << 15, 5), (23, 2_500_000, 777, 1 << 16, 2),
    (24, 777_777, 40_000, 700, 1), (25, 4_194_304, 1_000, 1 << 16, 5)])
@pytest.mark.parametrize("with_flags", [False, True], ids=["noflags", "flags"])
def test_random_traces_vs_oracle(cb, seed, n, de, unit, tgt, with_flags):
    start, now = tracegen.gen_trace(n, seed=seed, de=de, unit=unit, target_ns=tgt * 1_000_000)
    flags = util.make_flags(n, seed, cancel_mod=9, empty_mod=50_021) if with_flags else None
    nslots = 1024
    ctx = cb.Context(0)
    ctx.codel_init(tgt)
    ctx.backends_set(64, tracegen.gen_insert_draws(64, seed), nslots)
    dec, slot, be, st = ctx.admit_batch(start, now, flags, want_backend=True)
    od, os_, ostate, ost = oracle_run(start, now, flags, tgt, nslots)
    assert st.path == cb.PATH_PARALLEL
    assert np.array_equal(dec, od)
    assert np.array_equal(slot, os_)
    _, table = O.build_slots(64, nslots, tracegen.gen_insert_draws(64, seed))
    assert np.array_equal(be, np.where(os_ >= 0, table[np.maximum(os_, 0)], -1))
    util.assert_state_equal(gpu_state(st), ostate)
    assert (st.drops, st.admits, st.skipped, st.cursor) == (ost.drops, ost.admits, ost.skipped, ost.cursor)


def test_device_resident_api_and_streaming_batches(cb):
    """cb_admit_batch_dev on torch tensors; a stream of batches carries state on the device."""
    import torch
    n, nb, tgt = 3_000_000, 5, 5
    start, now = tracegen.gen_trace(n, seed=31, de=1_000, unit=1 << 15)
    ctx = cb.Context(0)
    ctx.codel_init(tgt)
    ctx.backends_set(16, tracegen.gen_insert_draws(16), 333)
    dev = torch.device("cuda:0")
    ts, tn = torch.from_numpy(start).to(dev), torch.from_numpy(now).to(dev)
    dec = torch.empty(n, dtype=torch.uint8, device=dev)
    slot = torch.empty(n, dtype=torch.int32, device=dev)
    be = torch.empty(n, dtype=torch.int32, device=dev)
    cuts = [0, 1, 600_001, 600_002, 2_000_000, n]  # ragged, unaligned batch boundaries
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        ctx.admit_batch_dev(ts[lo:hi], tn[lo:hi], None, dec[lo:hi], slot[lo:hi], be[lo:hi])
    st = ctx.stats()
    od, os_, ostate, ost = oracle_run(start, now, None, tgt, 333)
    assert np.array_equal(dec.cpu().numpy(), od)
    assert np.array_equal(slot.cpu().numpy(), os_)
    util.assert_state_equal(gpu_state(st), ostate)
    assert st.cursor == ost.cursor and st.total_drops == ost.drops


@pytest.mark.parametrize("depth", [1, 2, 5, 8])
def test_raw_pointer_calls_any_pipeline_depth(cb, depth):
    """admit_batch_ptr (raw device addresses, the native caller's form) over many small batches with
    1..8 batches in flight: decisions, slots and the carried state equal one sequential drain."""
    import torch
    n, m, tgt = 1_200_000, 37_500, 5
    start, now = tracegen.gen_trace(n, seed=77, de=1_000, unit=1 << 13)
    ctx = cb.Context(0)
    ctx.codel_init(tgt)
    ctx.backends_set(8, tracegen.gen_insert_draws(8), 61)
    ctx.set_pipeline_depth(depth)
    dev = torch.device("cuda:0")
    ts, tn = torch.from_numpy(start).to(dev), torch.from_numpy(now).to(dev)
    dec = torch.empty(n, dtype=torch.uint8, device=dev)
    slot = torch.empty(n, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    for lo in range(0, n, m):
        k = min(m, n - lo)
        ctx.admit_batch_ptr(ts.data_ptr() + 8 * lo, tn.data_ptr() + 8 * lo, None, k,
                            dec.data_ptr() + lo, slot.data_ptr() + 4 * lo, None)
    st = ctx.stats()
    od, os_, ostate, ost = oracle_run(start, now, None, tgt, 61)
    assert np.array_equal(dec.cpu().numpy(), od)
    assert np.array_equal(slot.cpu().numpy(), os_)
    util.assert_state_equal(gpu_state(st), ostate)
    assert st.cursor == ost.cursor and st.total_drops == ost.drops


@pytest.mark.parametrize("R,with_flags", [(1, False), (2, False), (3, True), (4, False)])
def test_one_queue_over_a_ring_of_contexts(cb, R, with_flags):
    """one logical queue cut into segments that R contexts drain in turn (segment (k, r)
    follows (k, r-1)); the ControlledDelay state and the admit cursor travel through the ring of device
    mailboxes.  Here the ring members are contexts of one process on one GPU (mailboxes given by
    pointer; across processes the same words are reached through CUDA IPC, tests/ring_check.py).
    Decisions, slots and the final state equal ONE sequential drain of the whole queue."""
    import torch
    tgt, nslots = 5, 97
    rng = np.random.RandomState(1234 + R)
    steps = 6
    seg = [[int(x) for x in rng.randint(1, 150_000, size=R)] for _ in range(steps)]
    seg[1][0] = 1          # a one-claim segment
    seg[2][R - 1] = 333_333
    n = sum(sum(x) for x in seg)
    start, now = tracegen.gen_trace(n, seed=91 + R, de=1_500, unit=1 << 13)
    flags = util.make_flags(n, 17, cancel_mod=7, empty_mod=40_009) if with_flags else None
    dev = torch.device("cuda:0")
    ts, tn = torch.from_numpy(start).to(dev), torch.from_numpy(now).to(dev)
    tf = torch.from_numpy(flags).to(dev) if with_flags else None
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
            ctxs[r].admit_batch_ptr(ts.data_ptr() + 8 * lo, tn.data_ptr() + 8 * lo,
                                    tf.data_ptr() + lo if with_flags else None, m,
                                    dec.data_ptr() + lo, slot.data_ptr() + 4 * lo, None)
            lo += m
    st = ctxs[R - 1].stats()  # the last rank's state is the queue's
    for c in ctxs:
        c.synchronize()
    od, os_, ostate, ost = oracle_run(start, now, flags, tgt, nslots)
    assert np.array_equal(dec.cpu().numpy(), od)
    assert np.array_equal(slot.cpu().numpy(), os_)
    util.assert_state_equal(gpu_state(st), ostate)
    assert st.cursor == ost.cursor and st.total_drops == ost.drops and st.total_admits == ost.admits
    for c in ctxs:
        c.ring_connect(0, 0)


def test_host_pipeline_chunking(cb, monkeypatch):
    """The host-buffer entry point splits into chunks (H2D/compute/D2H overlap): same answer."""
    monkeypatch.setenv("CB_CHUNK", "65536")
    n = 1_000_000
    start, now = tracegen.gen_trace(n, seed=41, de=2_000, unit=1 << 14)
    flags = util.make_flags(n, 41, cancel_mod=5, empty_mod=30_011)
    ctx = cb.Context(0)
    ctx.codel_init(5)
    ctx.backends_set(8, tracegen.gen_insert_draws(8), 100)
    dec, slot, be, st = ctx.admit_batch(start, now, flags, want_backend=True)
    od, os_, ostate, ost = oracle_run(start, now, flags, 5, 100)
    assert np.array_equal(dec, od) and np.array_equal(slot, os_)
    util.assert_state_equal(gpu_state(st), ostate)
    assert (st.drops, st.admits, st.skipped) == (ost.drops, ost.admits, ost.skipped)


def test_sequential_fallback_matches_oracle(cb):
    """Backwards clock inside a batch -> exact one-warp path; also a dense adversarial trace
    (an episode every few claims) that overflows the per-tile event slots."""
    n = 200_000
    start, now = tracegen.gen_trace(n, seed=51, de=20_000, unit=2048)
    now = now - ((np.arange(n) // 9_973) % 2) * 150_000_000
    flags = util.make_flags(n, 51, cancel_mod=6, empty_mod=7_001)
    ctx = cb.Context(0)
    ctx.codel_init(5)
    ctx.backends_set(4, tracegen.gen_insert_draws(4), 17)
    dec, slot, be, st = ctx.admit_batch(start, now, flags, want_backend=True)
    od, os_, ostate, ost = oracle_run(start, now, flags, 5, 17)
    assert st.path == cb.PATH_SEQUENTIAL
    assert np.array_equal(dec, od) and np.array_equal(slot, os_)
    util.assert_state_equal(gpu_state(st), ostate)
    # adversarial but monotone: A A B every 210 ms -> an ok-run every third claim
    m = 30_000
    t = 10_000_000_000 + 210_000_000 * np.arange(m, dtype=np.int64)
    s = t - np.where(np.arange(m) % 3 != 0, 50_000_000, 1_000_000)
    ctx2 = cb.Context(0)
    ctx2.codel_init(5)
    dec2, slot2, _, st2 = ctx2.admit_batch(s, t)
    od2, os2, ostate2, _ = oracle_run(s, t, None, 5, 0)
    assert np.array_equal(dec2, od2)
    util.assert_state_equal(gpu_state(st2), ostate2)
    assert st2.path == cb.PATH_SEQUENTIAL  # event slots overflowed -> exact path


def test_full_size_c3_bit_exact(cb):
    """BASELINE config C3 (16M claims, 1024 backends) vs the C oracle, element for element."""
    n, B, S = 1 << 24, 1024, 4096
    start, now = tracegen.gen_trace(n)
    ctx = cb.Context(0)
    ctx.codel_init(5)
    draws = tracegen.gen_insert_draws(B)
    ctx.backends_set(B, draws, S)
    dec, slot, be, st = ctx.admit_batch(start, now, want_backend=True)
    od, os_, ostate, ost = oracle_run(start, now, None, 5, S)
    assert st.path == cb.PATH_PARALLEL
    assert np.array_equal(dec, od) and np.array_equal(slot, os_)
    _, table = O.build_slots(B, S, draws)
    assert np.array_equal(be, np.where(os_ >= 0, table[np.maximum(os_, 0)], -1))
    util.assert_state_equal(gpu_state(st), ostate)
    # size-independent properties
    assert st.drops + st.admits == n and st.cursor == st.admits
    adm = slot[dec == cb.ADMIT]
    assert np.array_equal(adm, (np.arange(adm.size) % S).astype(np.int32))  # cyclic FIFO of slots
    assert np.all(slot[dec != cb.ADMIT] == -1)
    # splitting the stream anywhere and carrying the state gives the same decisions
    ctx2 = cb.Context(0)
    ctx2.codel_init(5)
    ctx2.backends_set(B, draws, S)
    cut = 9_999_999
    d1, s1, _, _ = ctx2.admit_batch(start[:cut], now[:cut])
    d2, s2, _, st2 = ctx2.admit_batch(start[cut:], now[cut:])
    assert np.array_equal(np.concatenate([d1, d2]), dec) and np.array_equal(np.concatenate([s1, s2]), slot)
    util.assert_state_equal(gpu_state(st2), ostate)


def test_sustained_overload_stream_c5_shape(cb):
    """C5 shape: everything past claim 2^20 overloaded; streamed in chunks, device-resident state."""
    n = 6 << 20
    start, now = tracegen.gen_trace(n, all_overloaded_from=1 << 20)
    ctx = cb.Context(0)
    ctx.codel_init(5)
    ctx.backends_set(4096, tracegen.gen_insert_draws(4096), 16384)
    decs, slots = [], []
    for lo in range(0, n, 1 << 21):
        d, s, _, st = ctx.admit_batch(start[lo:lo + (1 << 21)], now[lo:lo + (1 << 21)])
        decs.append(d)
        slots.append(s)
    od, os_, ostate, ost = oracle_run(start, now, None, 5, 16384)
    assert np.array_equal(np.concatenate(decs), od) and np.array_equal(np.concatenate(slots), os_)
    util.assert_state_equal(gpu_state(st), ostate)
    assert st.state.dropping == 1 and st.total_drops == ost.drops


# ------------------------------------------------------------------ topology / planRebalance
def _names(c):
    return [k for k, _ in c["conns"]]


@pytest.mark.parametrize("c", PLAN["kat"], ids=[c["test"] for c in PLAN["kat"]])
def test_plan_rebalance_reference_kat(cb, c):
    """The reference's own known-answer tests (test/utils.test.js:13-285) through the mirror API."""
    conns = dict((k, list(v)) for k, v in c["conns"])
    dead = dict((k, True) for k in c["dead"])
    plan = cb.planRebalance(conns, dead, c["target"], c["max"], c["singleton"])
    assert plan["add"] == c["add"]
    assert plan["remove"] == c["remove"]


def test_plan_rebalance_random(cb):
    for c in PLAN["random"]:
        conns = dict((k, list(v)) for k, v in c["conns"])
        plan = cb.planRebalance(conns, dict((k, True) for k in c["dead"]), c["target"], c["max"], c["singleton"])
        assert plan["add"] == c["add"] and plan["remove"] == c["remove"], c
    with pytest.raises(AssertionError):
        cb.planRebalance({"b1": []}, {}, 5, 4)
    # larger tables against the C oracle
    rng = np.random.RandomState(3)
    ctx = cb.Context(0)
    for B, tgt, mx in [(1024, 4096, 4096), (4096, 10_000, 16_384), (333, 1000, 1200)]:
        have = rng.randint(0, 5, size=B)
        dead = (rng.rand(B) < 0.3).astype(np.uint8)
        for single in (False, True):
            want, add, rem = ctx.plan_rebalance(have, dead, tgt, mx, single)
            ow, oa, orm = O.plan_rebalance(have, dead, tgt, mx, single)
            assert np.array_equal(want, ow) and np.array_equal(add, oa) and rem == orm


@pytest.mark.parametrize("c", SLOTS["build"], ids=lambda c: "B%d_S%d" % (c["B"], c["S"]))
def test_slot_table_golden(cb, c):
    ctx = cb.Context(0)
    ctx.backends_set(c["B"], tracegen.gen_insert_draws(c["B"], c["seed"]), c["S"])
    keys, slots = ctx.backends_get()
    assert keys.tolist() == c["keys"]
    assert slots.tolist() == c["slots"]


def test_resolver_events_and_reshuffle(cb):
    """'added' one at a time == bulk set; 'removed' and reshuffle follow lib/pool.js:293-313, 501-519."""
    B, S = 40, 100
    draws = tracegen.gen_insert_draws(B, 77)
    a, b = cb.Context(0), cb.Context(0)
    a.backends_set(B, draws, S)
    b.backends_set(0, [], S)
    for k in range(B):
        b.backend_added(k, draws[k])
    ka, sa = a.backends_get()
    kb, sb = b.backends_get()
    assert np.array_equal(ka, kb) and np.array_equal(sa, sb)
    keys = ka.copy()
    for u in tracegen.gen_insert_draws(6, 78):
        a.reshuffle(u)
        keys = O.reshuffle(keys, u)
        assert np.array_equal(a.backends_get()[0], keys)
    a.backend_removed(int(keys[3]))
    keys = np.delete(keys, 3)
    k2, s2 = a.backends_get()
    assert np.array_equal(k2, keys)
    have = np.zeros(B - 1, np.int32)
    _, add, _ = O.plan_rebalance(have, np.zeros(B - 1, np.uint8), S, S)
    assert np.array_equal(s2, keys[add])
    with pytest.raises(cb.CueballError):
        a.backend_removed(10_000)


def test_weighted_extension_is_sane(cb):
    """CB_BALANCE_SRV_WEIGHTED (extension, parity unpinned): only the best priority tier gets
    slots, proportional to weight, and equal weights reduce to an even split."""
    B, S = 12, 120
    table = [(0, 10)] * 4 + [(0, 5)] * 4 + [(1, 100)] * 4
    ctx = cb.Context(0)
    ctx.backends_set(B, tracegen.gen_insert_draws(B, 9), S, table=table, mode=cb.BALANCE_SRV_WEIGHTED)
    keys, slots = ctx.backends_get()
    cnt = np.bincount(slots, minlength=B)
    assert cnt[8:].sum() == 0 and cnt.sum() == S
    assert all(cnt[k] == 20 for k in range(4)) and all(cnt[k] == 10 for k in range(4, 8))


def test_pool_mirror_claim_release(cb):
    """ConnectionPool.claim()/release() surface over a virtual clock (mirrors the shape of the
    reference's test/codel.test.js load run: more claims than slots, CoDel sheds the excess)."""
    clock = [1_000_000_000]
    rng = np.random.RandomState(1)
    pool = cb.ConnectionPool({"targetClaimDelay": 300, "maximum": 2, "clock": lambda: clock[0],
                              "random": rng.random_sample})
    for k in ("b1", "b2"):
        pool.on_resolver_added(k, {"name": k})
    results = {"ok": 0, "timeout": 0, "delays": []}
    held = []

    def mk(t0):
        def cb_(err, hdl, conn):
            if err is not None:
                assert isinstance(err, cb.ClaimTimeoutError)
                results["timeout"] += 1
            else:
                results["ok"] += 1
                results["delays"].append((clock[0] - t0) / 1e6)
                held.append((clock[0] + 50_000_000, hdl))
        return cb_
    with pytest.raises(Exception):
        pool.claim({"timeout": 100}, lambda *a: None)
    for tick in range(500):  # 5 claims / 10 ms for 5 s, 50 ms hold, 2 slots
        clock[0] += 10_000_000
        for _ in range(5):
            pool.claim(mk(clock[0]))
        for item in [h for h in held if h[0] <= clock[0]]:
            held.remove(item)
            item[1].release()
    assert results["ok"] > 0 and results["timeout"] > 0
    stats = pool.getStats()
    assert stats["counters"]["claim"] == 2500
    assert stats["counters"]["claim-timeout"] == results["timeout"]
    # lib/codel.test.js:263-264 bound: mean claim latency within 175 ms of the target
    assert abs(np.mean(results["delays"]) - 300) < 175
