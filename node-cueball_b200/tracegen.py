"""tracegen.py -- seeded synthetic claim-queue traces (TEST/BENCH INPUT GENERATOR).

Specification (the JS harness that drives the reference's own code implements the same
generator bit-for-bit, so every implementation sees the same input):

    hash32 = lowbias32;  imul = 32-bit wrapping multiply
    claim i is enqueued at   start_i = t0 + i*de                     [ns]
    phases alternate healthy / overloaded; phase ph has length
        r = hash32(seed ^ imul(ph+1, 0x85ebca6b))
        healthy:    (1 + r%4) * unit claims,   soj = h_i % (0.9*target_ns)
        overloaded: (2 + r%8) * unit claims,   soj = target_ns + h_i % (60*target_ns)
        h_i = hash32(seed ^ imul(i+1, 0x9e3779b9))
    if all_overloaded_from >= 0 every claim from that index on is overloaded
    (sustained overload / drop state);
    claim i is dequeued at   now_i = max(start_i + soj_i, now_{i-1})  [ns, monotone]

This follows SURVEY.md 8(d) "Generator G(N, phases)".  It is NOT part of the
product path; bench.py and the tests use it to make inputs.
"""
import numpy as np

SEED = 0xC0EBA11
M32 = np.uint64(0xFFFFFFFF)


def _hash32_np(x):
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def hash32(x):
    return int(_hash32_np(np.array([x & 0xFFFFFFFF], dtype=np.uint32))[0])


def imul(a, b):
    return (a * b) & 0xFFFFFFFF


def phases(seed, n, unit, start_overloaded=False, all_overloaded_from=-1):
    """[(begin, end, overloaded)] exactly as the JS generator walks them."""
    out = []
    i, ph, over = 0, 0, bool(start_overloaded)
    aof = all_overloaded_from
    while i < n:
        r = hash32(seed ^ imul(ph + 1, 0x85EBCA6B))
        ln = (2 + (r % 8)) * unit if over else (1 + (r % 4)) * unit
        end = i + ln
        if aof >= 0 and i >= aof:
            over = True
            end = n
        elif aof >= 0 and end > aof:
            end = aof
        end = min(end, n)
        out.append((i, end, over))
        i = end
        over = not over
        ph += 1
    return out


def gen_trace(n, seed=SEED, t0=10**9, de=1000, target_ns=5_000_000, unit=1 << 16,
              start_overloaded=False, all_overloaded_from=-1):
    """Returns (start_ns, now_ns) int64 arrays of length n."""
    n = int(n)
    if n == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    idx = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        k = ((idx + np.uint64(1)) & M32).astype(np.uint32) * np.uint32(0x9E3779B9)
        h = _hash32_np(np.uint32(seed & 0xFFFFFFFF) ^ k).astype(np.int64)
    start = np.int64(t0) + idx.astype(np.int64) * np.int64(de)
    healthy_span = (target_ns * 9) // 10
    over_span = target_ns * 60
    soj = h % np.int64(healthy_span)
    for (b, e, over) in phases(seed, n, unit, start_overloaded, all_overloaded_from):
        if over:
            soj[b:e] = np.int64(target_ns) + h[b:e] % np.int64(over_span)
    now = np.maximum.accumulate(start + soj)
    return start, now


def trace_checksum(start, now):
    """Order-sensitive 32-bit checksum; matches cb_trace_checksum in driver.js."""
    c = 0x811C9DC5
    m = 0xFFFFFFFF
    for a, b in zip(start.tolist(), now.tolist()):
        for w in (a & m, (a >> 32) & m, b & m, (b >> 32) & m):
            c = hash32(c ^ w)
    return c


def gen_insert_draws(nbackends, seed=SEED):
    """u_k = xorshift32(seed)_k / 2^32 (exact doubles) -- the injected stand-in for
    Math.random() in lib/pool.js:287."""
    x = seed & 0xFFFFFFFF or 1
    out = np.empty(nbackends, dtype=np.float64)
    for k in range(nbackends):
        x ^= (x << 13) & 0xFFFFFFFF
        x ^= x >> 17
        x ^= (x << 5) & 0xFFFFFFFF
        out[k] = x / 4294967296.0
    return out
