"""tests/util.py -- helpers shared by the parity tests and tests/golden/gen_golden.py."""
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import tracegen  # noqa: E402

GOLDEN = os.path.join(HERE, "golden")
DEC_CHARS = "ADS"  # decision codes 0,1,2


def hx(v):
    return None if v is None else float(v).hex()


def unhx(s):
    return None if s is None else float.fromhex(s)


def state_hex(st):
    return {"count": hx(st["count"]), "drop_next": hx(st["drop_next"]),
            "first_above_time": hx(st["first_above_time"]), "dropping": bool(st["dropping"]),
            "last_empty": hx(st["last_empty"])}


def state_unhex(st):
    if st is None:
        return None
    return {"count": unhx(st["count"]), "drop_next": unhx(st["drop_next"]),
            "first_above_time": unhx(st["first_above_time"]), "dropping": bool(st["dropping"]),
            "last_empty": unhx(st["last_empty"])}


def rle(s):
    out, i = [], 0
    while i < len(s):
        j = i
        while j < len(s) and s[j] == s[i]:
            j += 1
        out.append("%s%d" % (s[i], j - i))
        i = j
    return "".join(out)


def unrle_codes(r):
    """'A3D2' -> uint8 array [0,0,0,1,1]"""
    out, i = [], 0
    while i < len(r):
        ch = r[i]
        j = i + 1
        while j < len(r) and r[j].isdigit():
            j += 1
        out.append(np.full(int(r[i + 1:j]), DEC_CHARS.index(ch), np.uint8))
        i = j
    return np.concatenate(out) if out else np.zeros(0, np.uint8)


def codes_to_rle(dec):
    return rle("".join(DEC_CHARS[int(d)] for d in dec))


def crc_i32(a):
    return zlib.crc32(np.ascontiguousarray(a, dtype=np.int32).tobytes()) & 0xFFFFFFFF


def make_trace(spec):
    """golden trace spec -> (start_ns, now_ns, flags|None) numpy arrays."""
    if spec["kind"] == "explicit":
        start = np.array(spec["start_ns"], dtype=np.int64)
        now = np.array(spec["now_ns"], dtype=np.int64)
    else:
        start, now = tracegen.gen_trace(spec["n"], seed=spec["seed"], **dict(spec["params"]))
    bs = spec.get("backstep")
    if bs:  # alternate blocks of `every` claims are shifted back by `ns`
        blk = (np.arange(len(now)) // bs["every"]) % 2
        now = now - blk.astype(np.int64) * np.int64(bs["ns"])
    flags = None
    fs = spec.get("flags")
    if fs:  # deterministic flags from hash32
        flags = make_flags(len(now), fs["seed"], fs.get("cancel_mod", 0), fs.get("empty_mod", 0))
    return start, now, flags


def make_flags(n, seed, cancel_mod=0, empty_mod=0):
    idx = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        k = ((idx + np.uint64(1)) & np.uint64(0xFFFFFFFF)).astype(np.uint32) * np.uint32(0x9E3779B9)
        h = tracegen._hash32_np(np.uint32(seed & 0xFFFFFFFF) ^ k).astype(np.int64)
    flags = np.ones(n, np.uint8)
    if cancel_mod:
        flags[(h % cancel_mod) == 0] = 0
    if empty_mod:
        flags[((h >> 8) % empty_mod) == 0] |= 2
    return flags


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def same_bits(a, b):
    """bit-for-bit equality of two doubles (None == None)."""
    if a is None or b is None:
        return a is None and b is None
    return np.float64(a).tobytes() == np.float64(b).tobytes()


def assert_state_equal(got, want, ctx=""):
    for k in ("count", "drop_next", "first_above_time", "last_empty"):
        assert same_bits(got[k], want[k]), "%s state.%s: got %r want %r" % (ctx, k, got[k], want[k])
    assert bool(got["dropping"]) == bool(want["dropping"]), "%s state.dropping" % ctx


def gen_pool_events(seed, n_ev, tgt, backlog_mult, jitter):
    """A pool event script (ev_type, ev_a, ev_b as lists; see cbo_pool_sim): an initial burst that
    puts `backlog_mult x target` of sojourn on the queue, then arrivals and slot-idle events at about
    the same rate, so that the sojourn stays between target and the hard timeout long enough for
    CoDel to act; occasional cancels, multi-slot idle events, arrival pauses (queue drains, empty())
    and service stalls."""
    rng = np.random.RandomState(seed)
    ms = 1_000_000
    t = 10 ** 9
    ty, a, b = [], [], []
    nclaims = 0
    svc = float(rng.choice([0.25, 0.5, 1.0]))
    burst = int(backlog_mult * tgt / svc)
    for i in range(burst):
        ty.append(0)
        a.append(t + i * 1000)
        b.append(0)
        nclaims += 1
    tc = t + burst * 1000
    ti = t + int(svc * ms)
    while len(ty) < n_ev:
        if tc <= ti:
            ty.append(0)
            a.append(tc)
            b.append(0)
            nclaims += 1
            tc += max(1, int((svc + rng.uniform(-jitter, jitter) * svc) * ms))
            if rng.rand() < 0.02:
                ty.append(2)
                a.append(int(rng.randint(max(0, nclaims - 40), nclaims)))
                b.append(0)
        else:
            k = 1
            if rng.rand() < 0.05:
                k = int(rng.choice([2, 3]))
            ty.append(1)
            a.append(ti)
            b.append(k)
            ti += max(1, int((svc + rng.uniform(-jitter, jitter) * svc) * ms)) * k
        if rng.rand() < 0.001:
            tc += int(rng.uniform(150, 500)) * ms
        if rng.rand() < 0.001:
            ti += int(rng.uniform(5, 60)) * ms
    return ty, a, b


def gen_lpf_events(seed, n_ev, peak):
    """A pool's life as the low-pass filter sees it: sampler ticks (ty 0: a = connections, b = idleq + initq)
    interleaved with rebalances (ty 1: a = total connections, b = idleq, c = initq, d = waiters).  The load
    follows a seeded random walk with ramps and bursts so that the clamp (lib/pool.js:577-583) switches both ways."""
    rng = np.random.RandomState(seed)
    ty, a, b, c, d = [], [], [], [], []
    conns, busy = int(rng.randint(0, peak // 2 + 1)), 0
    trend = 0
    for _ in range(n_ev):
        if rng.rand() < 0.05:
            trend = int(rng.choice([-2, -1, 0, 1, 2]))
        busy = int(np.clip(busy + trend + rng.randint(-2, 3) + (rng.randint(5, 15) if rng.rand() < 0.03 else 0), 0, peak))
        conns = int(np.clip(max(conns + rng.randint(-1, 2), busy), 0, peak))
        initq = int(rng.randint(0, 3)) if conns > busy else 0
        idle = max(0, conns - busy - initq)
        waiters = int(rng.randint(0, 2 * peak)) if (busy >= conns and rng.rand() < 0.5) else 0
        if rng.rand() < 0.7:
            ty.append(0); a.append(conns); b.append(idle + initq); c.append(0); d.append(0)  # noqa: E702
        else:
            ty.append(1); a.append(conns); b.append(idle); c.append(initq); d.append(waiters)  # noqa: E702
    return ty, a, b, c, d


def gen_queue_ops(seed, n_ops):
    """A script for cb_queue_ops / the Queue mirror: unique pushed values, shifts, removals of nodes that may or may
    not still be linked, forEach passes that remove from inside the callback, peeks."""
    rng = np.random.RandomState(seed)
    op, arg, pushed = [], [], 0
    for _ in range(n_ops):
        r = rng.rand()
        if r < 0.45:
            op.append(0); arg.append(pushed + 1); pushed += 1                    # noqa: E702
        elif r < 0.70:
            op.append(1); arg.append(0)                                          # noqa: E702
        elif r < 0.85:
            op.append(2); arg.append(int(rng.randint(0, max(pushed, 1))))        # noqa: E702
        elif r < 0.95:
            op.append(3); arg.append(int(rng.choice([0, 2, 3, 7])))              # noqa: E702
        else:
            op.append(4); arg.append(0)                                          # noqa: E702
    return op, arg
