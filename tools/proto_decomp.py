"""tools/proto_decomp.py -- executable specification of the parallel decomposition used
by the CUDA path (DESIGN.md "Decomposition"), checked here on CPU against the C oracle /
golden vectors.  Development aid only: not imported by the product or by the tests'
product path.  The three stages mirror the three kernels:

  scan    : per-claim class c_i in {0,1,2} from the first_above_time recurrence
            (an associative scan) ; events = positions where c_i != c_{i-1}
  resolve : serial chain over events only (entry test, count, drop_next, lower_bound)
  emit    : decisions / slots from the drop intervals
"""
import bisect
import math
import sys

import numpy as np


def scan(now, start, target, flags, F_in):
    n = len(now)
    cls = np.zeros(n, np.int8)
    F = F_in
    for i in range(n):
        above = not ((now[i] - start[i]) < target)
        mprev = bool(flags is not None and i > 0 and (flags[i - 1] & 2))
        Fp = 0.0 if mprev else F
        ok = above and Fp != 0.0 and now[i] >= Fp
        cls[i] = (2 if (now[i] - Fp) >= 100.0 else 1) if ok else 0
        F = (Fp if Fp != 0.0 else now[i] + 100.0) if above else 0.0
    F_out = 0.0 if (flags is not None and n > 0 and (flags[n - 1] & 2)) else F
    events = [(i, int(cls[i])) for i in range(n) if cls[i] != (cls[i - 1] if i > 0 else 0)]
    return cls, events, F_out


def lower_bound(now, lo, hi, D):
    """first index in [lo,hi) with now >= D, else hi (now monotone)"""
    return lo + bisect.bisect_left(now[lo:hi].tolist(), D)


def resolve(now, n, events, D, C, B):
    ivl = []
    if n == 0:
        return ivl, D, C, B
    nev = len(events)
    if nev == 0 or events[0][0] > 0:
        B = False
    k = 0
    while k < nev:
        idx, c = events[k]
        if c == 0:
            B = False
            k += 1
            continue
        # run end
        has2 = (c == 1 and k + 1 < nev and events[k + 1][1] == 2)
        kz = k + (2 if has2 else 1)
        q = (events[kz][0] - 1) if kz < nev else n - 1
        p = idx
        if B:
            d = lower_bound(now, p, q + 1, D)
            if d <= q:
                ivl.append((d, q))
                C += q - d + 1
        else:
            e = -1
            cond1 = (now[p] - D) < 100.0
            if c == 2 or cond1:
                e = p
            elif has2:
                e = events[k + 1][0]
                cond1 = (now[e] - D) < 100.0
            if e >= 0:
                C = (C - 2 if C > 2 else 1) if cond1 else 1
                D = now[e] + 100.0 / math.sqrt(C)
                B = True
                d = lower_bound(now, e + 1, q + 1, D)
                ivl.append((e, e))
                if d <= q:
                    ivl.append((d, q))
                    C += q - d + 1
        k = kz
    return ivl, D, C, B


def run(now, start, target, flags, state, nslots, cursor):
    """state = dict(F, D, C, B); returns decisions, slots, new state"""
    n = len(now)
    if n and np.any(np.diff(now) < 0):
        return None  # non-monotone: sequential path
    cls, events, F_out = scan(now, start, target, flags, state["F"])
    ivl, D, C, B = resolve(now, n, events, state["D"], state["C"], state["B"])
    drop = np.zeros(n, bool)
    for lo, hi in ivl:
        drop[lo:hi + 1] = True
    dec = np.zeros(n, np.uint8)
    slot = np.full(n, -1, np.int32)
    for i in range(n):
        w = True if flags is None else bool(flags[i] & 1)
        if not w:
            dec[i] = 2
        elif drop[i]:
            dec[i] = 1
        else:
            slot[i] = cursor % nslots if nslots else -1
            cursor += 1
    return dec, slot, dict(F=F_out, D=D, C=C, B=B), cursor, len(events), len(ivl)


if __name__ == "__main__":
    sys.path.insert(0, ".")
    sys.path.insert(0, "tests")
    import util
    from oracle import oracle as O
    G = util.load_golden("codel_golden.json")
    to_ms = np.vectorize(O.ns_to_ms, otypes=[np.float64])
    for case in G["cases"]:
        start, now, flags = util.make_trace(case["trace"])
        s_ms = to_ms(start) if len(start) else np.zeros(0)
        n_ms = to_ms(now) if len(now) else np.zeros(0)
        st0 = util.state_unhex(case["init_state"]) or dict(count=0, drop_next=0.0, first_above_time=0.0, dropping=False)
        st = dict(F=st0["first_above_time"], D=st0["drop_next"], C=int(st0["count"]), B=bool(st0["dropping"]))
        lo, cursor, decs, slots, ok, nev, niv = 0, 0, [], [], True, 0, 0
        for bi, b in enumerate(case["batches"]):
            r = run(n_ms[lo:lo + b], s_ms[lo:lo + b], float(case["target"]), None if flags is None else flags[lo:lo + b], st, case["nslots"], cursor)
            if r is None:
                ok = None
                break
            dec, slot, st, cursor, e, v = r
            nev += e
            niv += v
            want = util.state_unhex(case["expect"]["states"][bi])
            assert util.same_bits(st["F"], want["first_above_time"]), (case["name"], st, want)
            assert util.same_bits(st["D"], want["drop_next"]), (case["name"], st, want)
            assert st["C"] == want["count"] and st["B"] == want["dropping"], (case["name"], st, want)
            decs.append(dec)
            slots.append(slot)
            lo += b
        if ok is None:
            print("%-32s non-monotone -> sequential path" % case["name"])
            continue
        dec = np.concatenate(decs) if decs else np.zeros(0, np.uint8)
        assert util.codes_to_rle(dec) == case["expect"]["dec_rle"], case["name"]
        assert util.crc_i32(np.concatenate(slots)) == case["expect"]["slots_crc32"], case["name"]
        print("%-32s OK  events=%d intervals=%d" % (case["name"], nev, niv))
