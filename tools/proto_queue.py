"""tools/proto_queue.py -- executable specification of the device-resident pending-claim queue
(DESIGN.md "What comes next"), checked on CPU against the oracle's
event-driven pool model (tests/test_proto_queue.py).  Development aid: the CUDA kernels of the
`cb_queue_*` entry points follow this file step by step; nothing here runs on the product path.

The reference handles one claim per closure call: `claim()` stamps the handle and either takes an
idle slot or pushes it on `p_waiters` (lib/pool.js:874-885,929-966); every slot that turns idle
runs the drain loop (lib/pool.js:733-757), which calls `ControlledDelay#overloaded` per dequeued
handle with a fresh clock read.  Here the queue is structure-of-arrays (`start_ns`, `deadline_ns`,
`waiting`, `ticket`; FIFO order = array order) and the three operations are data-parallel:

  push(starts)   n claims arrive: `getMaxIdle()` per claim (lib/codel.js:109-118) is elementwise
                 because `cd_last_empty` only changes in an idle event; the first `idle` claims
                 take idle slots (lib/pool.js:934-951), the rest are appended.
  expire(T)      hard timeouts (lib/connection-fsm.js:626-642): `waiting &= T < deadline` over the
                 queued handles that are old enough to have timed out (a prefix found by bisection
                 on start: a handle cannot expire before start + 3 x target).
  idle(T, k)     k slots turn idle at clock T.  With ONE clock value for the whole event the drain
                 loop has a closed form: sojourn = T - start is non-increasing along the FIFO, so
                 the claims at or above target are a PREFIX of length P (bisection with the exact
                 double comparison of lib/codel.js:35-37); along that prefix `canDrop` returns one
                 constant, so `overloaded` (lib/codel.js:56-86) drops either nothing, or exactly the
                 first claim (entering the dropping state pushes `drop_next` beyond T), or every
                 claim of the prefix (already dropping and T >= drop_next: `drop_next` is not
                 advanced while dropping); the first below-target claim resets
                 first_above_time and the dropping flag.  How many handles the k slots consume is a
                 selection: the position of the k-th handle that is still waiting and not dropped
                 (prefix sum over the waiting bits).  The state is advanced over exactly those.

Everything is in the reference's arithmetic: clocks are int64 ns converted with
`utils.currentMillis` (lib/utils.js:198-204), all comparisons on float64 milliseconds.
"""
import math

import numpy as np

INTERVAL = 100.0  # lib/codel.js:16
ADMIT, DROP, SKIPPED, EXPIRED = 0, 1, 2, 3


def ns_to_ms(t):
    """utils.currentMillis() of process.hrtime() = [t // 1e9, t % 1e9]; each op rounds once."""
    t = np.asarray(t, dtype=np.int64)
    secs = t // 1_000_000_000
    ns = t - secs * 1_000_000_000
    return secs.astype(np.float64) * 1000.0 + ns.astype(np.float64) / 1000000.0


class DeviceQueueSpec(object):
    def __init__(self, target, nslots, idle0=0, cursor=0, state=None):
        self.target = float(target)
        self.nslots = int(nslots)
        # ControlledDelay, lib/codel.js:24-32 (+ cd_last_empty, undefined at construction)
        self.F, self.D, self.C, self.B = 0.0, 0.0, 0, False
        self.hasE, self.E = False, 0.0
        if state:
            self.F, self.D = float(state["first_above_time"]), float(state["drop_next"])
            self.C, self.B = int(state["count"]), bool(state["dropping"])
            if state.get("last_empty") is not None:
                self.hasE, self.E = True, float(state["last_empty"])
        self.idle, self.cursor = int(idle0), int(cursor)
        cap = 1024
        self.start_ns = np.zeros(cap, np.int64)
        self.deadline_ns = np.zeros(cap, np.int64)
        self.waiting = np.zeros(cap, np.uint8)
        self.ticket = np.zeros(cap, np.int64)
        self.head = self.tail = 0
        self.next_ticket = 0
        self.where = {}  # ticket -> queue position (host-side map of the handles it hands out)
        self.counters = {"claim": 0, "queued-claim": 0, "max-claim-queue": 0, "claim-timeout": 0}
        self.res = []  # (ticket, code, slot) in resolution order

    # ------------------------------------------------------------------ helpers
    def _grow(self, need):
        cap = self.start_ns.shape[0]
        if self.tail + need <= cap:
            return
        live = self.tail - self.head
        ncap = max(2 * cap, 2 * (live + need))
        for name in ("start_ns", "deadline_ns", "waiting", "ticket"):
            old = getattr(self, name)
            new = np.zeros(ncap, old.dtype)
            new[:live] = old[self.head:self.tail]
            setattr(self, name, new)
        for t, pos in list(self.where.items()):
            self.where[t] = pos - self.head
        self.head, self.tail = 0, live

    def _admit(self, ticket):
        slot = (self.cursor % self.nslots) if self.nslots > 0 else -1
        self.res.append((int(ticket), ADMIT, slot))
        self.cursor += 1

    def state(self):
        return {"count": self.C, "drop_next": self.D, "first_above_time": self.F, "dropping": self.B,
                "last_empty": self.E if self.hasE else None}

    # ------------------------------------------------------------------ claim(): enqueue side
    def push(self, starts_ns):
        """n claims arrive (clock = each claim's own start); returns their tickets."""
        starts_ns = np.asarray(starts_ns, dtype=np.int64)
        n = starts_ns.shape[0]
        tickets = np.arange(self.next_ticket, self.next_ticket + n, dtype=np.int64)
        self.next_ticket += n
        self.counters["claim"] += n  # lib/pool.js:887
        # getMaxIdle, lib/codel.js:109-118 (undefined < x is false)
        now = ns_to_ms(starts_ns)
        bound = self.target * 10.0
        low = (self.E < (now - bound)) if self.hasE else np.zeros(n, bool)
        timeout = np.where(low, self.target * 3.0, bound)
        take = min(self.idle, n)  # lib/pool.js:934-951: idle slots first
        for i in range(take):
            self._admit(tickets[i])
        self.idle -= take
        m = n - take
        if m:
            self._grow(m)
            lo, hi = self.tail, self.tail + m
            self.start_ns[lo:hi] = starts_ns[take:]
            self.deadline_ns[lo:hi] = starts_ns[take:] + (timeout[take:] * 1000000.0).astype(np.int64)
            self.waiting[lo:hi] = 1
            self.ticket[lo:hi] = tickets[take:]
            for j in range(m):
                self.where[int(tickets[take + j])] = lo + j
            self.tail = hi
            self.counters["queued-claim"] += m  # lib/pool.js:963
            # lib/pool.js:962: the queue only grows inside a push, so the high-water mark is its end
            self.counters["max-claim-queue"] = max(self.counters["max-claim-queue"], self.tail - self.head)
        return tickets, timeout

    def cancel(self, ticket):
        pos = self.where.get(int(ticket))
        if pos is not None and pos >= self.head and self.waiting[pos]:
            self.waiting[pos] = 0  # lib/connection-fsm.js:580-587

    # ------------------------------------------------------------------ hard timeouts
    def expire(self, T_ns):
        """lib/connection-fsm.js:626-642 for every queued handle whose timer is due at clock T."""
        h, t = self.head, self.tail
        if h == t:
            return
        # only handles with start <= T - 3*target can be due: a prefix (start is non-decreasing)
        horizon = T_ns - int(self.target * 3.0 * 1000000.0)
        hi = h + int(np.searchsorted(self.start_ns[h:t], horizon, side="right"))
        due = np.flatnonzero((self.waiting[h:hi] != 0) & (self.deadline_ns[h:hi] <= T_ns)) + h
        for pos in due:  # ticket order == queue order
            self.waiting[pos] = 0
            self.res.append((int(self.ticket[pos]), EXPIRED, -1))
        self.counters["claim-timeout"] += int(due.shape[0])

    # ------------------------------------------------------------------ slots turn idle
    def idle_event(self, T_ns, k):
        self.expire(T_ns)
        T = float(ns_to_ms(T_ns))
        h, t = self.head, self.tail
        Q = t - h
        if Q > 0 and k > 0:
            st_ms = ns_to_ms(self.start_ns[h:t])
            # P: claims at or above target form a prefix (sojourn non-increasing).  Exact predicate
            # `!(T - start < target)`; bisection is valid because start_ms is non-decreasing.
            above = ~((T - st_ms) < self.target)
            P = int(np.searchsorted(~above, True))  # first below-target position
            assert above[:P].all() and not above[P:].any(), "enqueue clocks must be non-decreasing"
            w = self.waiting[h:t] != 0
            # ---- the drop pattern along the above-target prefix (lib/codel.js:34-86 at constant now)
            F0, D0, C0, B0 = self.F, self.D, self.C, self.B
            drop_first, drop_from = False, Q  # drops = {0} if drop_first, plus [drop_from, P)
            F1 = F0
            ok = False
            if P > 0:
                if F0 == 0.0:
                    F1 = T + INTERVAL  # claim 0 opens the run, lib/codel.js:39-40
                else:
                    ok = T >= F0       # lib/codel.js:41-42, the same for every claim of the prefix
            enter = (not B0) and ok and (((T - D0) < INTERVAL) or ((T - F0) >= INTERVAL))
            C1, D1 = C0, D0
            if P > 0 and B0 and ok and T >= D0:
                drop_from = 0          # lib/codel.js:65-68 for every claim: drop_next stays put
            elif P > 0 and enter:
                drop_first = True      # lib/codel.js:70-83 at claim 0
                C1 = (C0 - 2 if C0 > 2 else 1) if ((T - D0) < INTERVAL) else 1
                D1 = T + INTERVAL / math.sqrt(C1)
                if T >= D1:            # only if 100/sqrt(count) vanishes against T
                    drop_from = 1
            dropped = np.zeros(Q, bool)
            if drop_first:
                dropped[0] = True
            if drop_from < P:
                dropped[drop_from:P] = True
            # ---- how many handles do k slots consume: position of the k-th admitted one
            a = w & ~dropped
            ca = np.cumsum(a)
            if ca[-1] >= k:
                L = int(np.searchsorted(ca, k)) + 1
                left = 0
            else:
                L = Q
                left = k - int(ca[-1])
            pa = min(L, P)             # above-target claims actually dequeued
            # ---- state after exactly L claims
            if pa > 0:
                self.F = F1
                if B0 and not ok:
                    self.B = False     # lib/codel.js:63-64
                elif B0 and ok and T >= D0:
                    self.C = C0 + pa   # every dequeued claim of the prefix counted
                elif enter:
                    self.B = True
                    self.C = C1 + (max(0, pa - 1) if drop_from == 1 else 0)
                    self.D = D1
            if L > pa:                 # a below-target claim was dequeued: lib/codel.js:37-38, 63-64
                self.F = 0.0
                self.B = False
            # ---- resolutions, in dequeue order
            ndrop = 0
            for i in range(L):
                tk = int(self.ticket[h + i])
                if not w[i]:
                    self.res.append((tk, SKIPPED, -1))     # lib/pool.js:738-740
                elif dropped[i]:
                    self.res.append((tk, DROP, -1))        # lib/pool.js:742-745
                    ndrop += 1
                else:
                    self._admit(tk)                        # lib/pool.js:747
                self.where.pop(tk, None)
            self.counters["claim-timeout"] += ndrop        # lib/connection-fsm.js:633-638
            self.head = h + L
        else:
            left = k
        if left > 0:                   # queue ran empty with slots still idle: lib/pool.js:751-757
            self.E, self.hasE = T, True
            self.F = 0.0
            self.idle += left

    # ------------------------------------------------------------------ replay of an event script
    def run(self, ev_type, ev_a, ev_b):
        for ty, a, b in zip(ev_type, ev_a, ev_b):
            if ty == 0:
                self.expire(int(a))
                self.push(np.array([a], dtype=np.int64))
            elif ty == 1:
                self.idle_event(int(a), int(b))
            else:
                self.cancel(int(a))
        return np.array(self.res, dtype=np.int64).reshape(-1, 3)
