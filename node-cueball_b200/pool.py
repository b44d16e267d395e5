"""ConnectionPool -- the claim()/release() surface of the reference's lib/pool.js for the
claim-queue path, with the pending-claim FIFO and the admission work resident on the GPU.

What is mirrored (reference file:line):
  claim([options,] cb) -> handle{cancel()}          lib/pool.js:859-969
      options.timeout forbidden with targetClaimDelay   :874-878
      timeout = codel.getMaxIdle()                      :880
      idle-first fast path / enqueue on p_waiters       :929-966
  claim handle hard timeout -> ClaimTimeoutError    lib/connection-fsm.js:626-642
  handle.cancel(): waiting -> cancelled (stays in the FIFO, skipped at dequeue),
                   claimed -> released                  lib/connection-fsm.js:580-587
  slot goes idle -> drain loop                      lib/pool.js:715-760
      shift, codel.overloaded(hdl.ch_started), skip non-waiting, timeout() or try()
      queue emptied with the slot still idle -> codel.empty(), slot onto p_idleq
  on_resolver_added / on_resolver_removed / reshuffle   lib/pool.js:285-313, 501-519
  getStats(): counters + the five fields of           lib/pool.js:269-283, 834-857
  toKangSnapshot()                                    lib/pool-monitor.js:91-133
  the low-pass filter of the pool's size (FIRFilter, genTaps, the 5 Hz sampler) and
  _rebalance's target: busy + extras + spares, clamped from below by the filter
                                                    lib/pool.js:44-100, 251-262, 541-592
  handle states waiting/claimed/released/closed/cancelled/failed
                                                    lib/connection-fsm.js:489-808

Every event is ONE device call: claim() is cb_queue_push (enqueue clock, getMaxIdle, idle-first
fast path and the timers that came due, in one kernel), a slot turning idle is cb_queue_idle (the
whole drain of lib/pool.js:733-757 for k slots in closed form: never per claim, never a trial
batch), cancel() is cb_queue_cancel, the timer sweep cb_queue_expire.  The resolution records the
device appends -- (ticket, code, slot) in the order the reference would resolve the handles -- are
turned into the handles' callbacks here.  tests/test_gpu_pool_mirror.py replays the pinned event
scripts of the reference's pool model through this class and compares records and counters exactly.

What is NOT here (out of scope, stays in Node in a real deployment): sockets, the slot and
socket FSMs, retry/backoff and the churn limiter that paces how a rebalance plan is applied
(lib/pool.js:599-658).  The low-pass filter is host arithmetic (128 multiply-adds five times a
second, summed in the reference's order): its INPUT -- busy slots -- is the device queue's counters.  A "slot" here is
an integer (admitted claim j of the pool's life takes slot j mod nslots); its
backend comes from the device-side slot table.  A pool WITHOUT targetClaimDelay has no CoDel
(lib/pool.js:195-200): nothing on its path is accelerated, and it runs as a plain host FIFO.
"""
import collections
import heapq
import math
import time

import numpy as np

from .context import Context
from .errors import ClaimHandleMisusedError, ClaimTimeoutError, NoBackendsError, PoolStoppingError
from .utils import assertClaimDelay

R_ADMIT, R_CODEL_DROP, R_SKIPPED, R_HARD_TIMEOUT = 0, 1, 2, 3  # codes of cb_queue_results

LP_RATE = 5                           # sample rate of the filter, Hz (lib/pool.js:44)
LP_INT = int(round(1000 / LP_RATE))   # one sampling interval, ms (lib/pool.js:46)


def genTaps(count, tc):
    """EMA taps exp(tc * i) normalised to sum 1 (lib/pool.js:64-76), summed in the reference's order."""
    taps = [math.exp(tc * i) for i in range(count)]
    total = 0.0
    for t in taps:
        total += t
    return [t / total for t in taps]


LP_TAPS = genTaps(128, -0.2)          # lib/pool.js:48


class FIRFilter(object):
    """A FIR filter over a circular buffer (lib/pool.js:78-100); get() accumulates newest sample first,
    one rounding per multiply and per add, exactly as the reference's loop does."""

    def __init__(self, taps):
        self.f_taps = taps
        self.f_buf = [0.0] * len(taps)
        self.f_ptr = 0

    def put(self, v):
        self.f_buf[self.f_ptr] = float(v)
        self.f_ptr += 1
        if self.f_ptr == len(self.f_taps):
            self.f_ptr = 0

    def get(self):
        n = len(self.f_taps)
        i = self.f_ptr - 1
        if i < 0:
            i += n
        acc = 0.0
        for j in range(n):
            acc += self.f_buf[i] * self.f_taps[j]
            i -= 1
            if i < 0:
                i += n
        return acc


def rebalanceTarget(lpf, total, idleq, initq, waiters, spares_wanted, maximum):
    """The size _rebalance asks planRebalance for (lib/pool.js:563-586): busy + extras + spares, never
    below what the filter remembers of the recent load.  Returns (target, clamped)."""
    spares = idleq + initq - waiters
    if spares < 0:
        spares = 0
    busy = total - spares
    if busy < 0:
        busy = 0
    extras = waiters - initq
    if extras < 0:
        extras = 0
    target = busy + extras + spares_wanted
    mn = math.ceil(lpf.get())
    clamped = False
    if target < mn * 1.05:
        target = mn
        clamped = True
    if target > maximum:
        target = maximum
    return int(target), clamped


class CueBallClaimHandle(object):
    """lib/connection-fsm.js:489-808 reduced to what the claim queue touches."""

    def __init__(self, pool, callback, started_ns, claim_timeout_ms, ticket):
        self.ch_pool = pool
        self.ch_callback = callback
        self.ch_started_ns = started_ns  # ch_started, :523
        self.ch_claimTimeout = claim_timeout_ms
        self.ch_ticket = ticket
        self.ch_state = "waiting"
        self.ch_slot = None
        self.ch_backend = None

    def isInState(self, s):
        return self.ch_state == s

    def cancel(self):  # :580-587
        if self.ch_state == "waiting":
            self.ch_state = "cancelled"
            self.ch_pool._cancelled(self)
        elif self.ch_state == "claimed":
            self.release()

    def timeout(self):  # :589, :633-638 (the counter is the device's: cb_queue_info.claim_timeout)
        if self.ch_state != "waiting":
            return
        self.ch_state = "failed"
        if self.ch_callback:
            self.ch_callback(ClaimTimeoutError(self.ch_pool), None, None)

    def try_(self, slot, backend):  # try(): :559 -> claiming -> claimed
        assert self.ch_state == "waiting"
        self.ch_state = "claimed"
        self.ch_slot = slot
        self.ch_backend = backend
        if self.ch_callback:
            self.ch_callback(None, self, {"slot": slot, "backend": backend})

    def release(self):  # :615
        if self.ch_state != "claimed":
            raise ClaimHandleMisusedError("release() on a handle in state " + self.ch_state)
        self.ch_state = "released"
        self.ch_pool._slot_released(self.ch_slot)

    def close(self):  # :619
        if self.ch_state != "claimed":
            raise ClaimHandleMisusedError("close() on a handle in state " + self.ch_state)
        self.ch_state = "closed"
        self.ch_pool._slot_released(self.ch_slot)


class ConnectionPool(object):
    _DEADLINE_SLACK = 1024  # stale entries tolerated in the deadline heap before it is rebuilt

    def __init__(self, options):
        assertClaimDelay(options.get("targetClaimDelay"))
        self.p_max = int(options.get("maximum", 0))
        self.p_uuid = options.get("uuid")        # lib/pool.js:137: named in the errors' messages
        self.p_domain = options.get("domain")
        self._clock = options.get("clock") or time.monotonic_ns
        self._random = options.get("random") or np.random.random
        self._ctx = options.get("context") or Context(options.get("device", 0))
        self._capacity = int(options.get("queueCapacity", 1 << 16))
        # how many of the table's slots are connected and idle when the first claim arrives (the reference's
        # connections come up over time; default: all of them)
        self._initial_idle = options.get("initialIdleSlots")
        self.p_codel = None
        if options.get("targetClaimDelay") is not None:
            self._ctx.codel_init(options["targetClaimDelay"])  # lib/pool.js:197-200
            self.p_codel = self._ctx
        self.p_keys = []       # preference list as the caller's keys (device holds indices)
        self._key_ids = {}
        self._id_to_key = {}
        self._next_id = 0
        self._stopping = False
        self._slot_backend = np.zeros(0, np.int32)
        self._handles = {}     # ticket -> handle, until the device resolves it
        self._deadlines = []   # heap of (deadline ns, ticket) of the handles above (entries go stale, see nextDeadline)
        self._claimed = 0      # slots held by claimed handles
        self._busy_slots = {}  # slot -> how many claimed handles hold it (1; the kang snapshot counts them)
        self.p_backends = {}   # key -> backend object as the resolver gave it (lib/pool.js:285-291)
        self.p_spares = int(options.get("spares", 0))
        self.p_lpf = FIRFilter(LP_TAPS)      # lib/pool.js:189
        self.p_lastRebalClamped = False      # lib/pool.js:247
        self.p_lastRebalance = None          # what the last rebalance() asked for (observability)
        self._queue_ready = False
        # the last (ticket, code, slot) records the device produced, in order (observability/tests; bounded)
        self.records = collections.deque(maxlen=int(options.get("recordHistory", 1 << 16)))
        # the pool without CoDel: a plain host FIFO (nothing to accelerate)
        self._host_waiters = []
        self._host_started = False
        self._host_idle = 0
        self._host_counters = {}
        self._host_cursor = 0
        self._ctx.backends_set(0, [], self.p_max)  # empty preference list, `maximum` slots wanted

    # ------------------------------------------------------------ device queue
    def _ensure_queue(self, idle_slots):
        if not self._queue_ready:
            if self._initial_idle is not None:
                idle_slots = self._initial_idle
            self._ctx.queue_create(self._capacity, 4 * self._capacity, int(idle_slots))
            self._queue_ready = True

    def _idle_before_the_queue(self):
        n = len(self._slot_backend)
        return n if self._initial_idle is None else min(int(self._initial_idle), n)

    def _dispatch(self):
        """Turn the device's resolution records into handle transitions, in the device's order."""
        for ticket, code, slot in self._ctx.queue_results().tolist():
            self.records.append((ticket, code, slot))
            h = self._handles.pop(ticket, None)
            if h is None:
                continue
            if code == R_ADMIT:  # lib/pool.js:747 / :934-951
                self._claimed += 1
                self._busy_slots[slot] = self._busy_slots.get(slot, 0) + 1
                h.try_(slot, int(self._slot_backend[slot]) if slot < len(self._slot_backend) else -1)
            elif code in (R_CODEL_DROP, R_HARD_TIMEOUT):  # lib/pool.js:742-745 / connection-fsm.js:626-642
                h.timeout()
            # R_SKIPPED: the handle left 'waiting' before it was dequeued (lib/pool.js:738-740)

    # ------------------------------------------------------------ counters / stats
    def getStats(self):  # lib/pool.js:834-857: exactly these five keys
        if self.p_codel is None:
            return {"counters": dict(self._host_counters), "totalConnections": len(self._slot_backend),
                    "idleConnections": self._host_idle, "pendingConnections": 0,
                    "waiterCount": len(self._host_waiters)}
        if not self._queue_ready:  # no claim yet: the queue is created with this many idle slots (_ensure_queue)
            return {"counters": {}, "totalConnections": len(self._slot_backend),
                    "idleConnections": self._idle_before_the_queue(), "pendingConnections": 0, "waiterCount": 0}
        info = self._ctx.queue_info()
        # a counter exists once it has been bumped (lib/pool.js:269-283)
        counters = {k: v for k, v in info.counters().items() if v}
        return {"counters": counters, "totalConnections": len(self._slot_backend),
                "idleConnections": int(info.idle), "pendingConnections": 0,
                "waiterCount": int(info.tail - info.head)}

    def toKangSnapshot(self):
        """What the pool monitor publishes for a pool (lib/pool-monitor.js:91-133 getPool): backends,
        per-backend connection counts by slot state, dead backends, pool state, counters and the
        spares/maximum options.  A slot of this model is 'busy' while a claimed handle holds it and 'idle'
        otherwise (the reference's other slot states belong to socket management, outside the path);
        the resolver fields are the host's."""
        conns = {k: {} for k in self.p_keys}
        for slot, kid in enumerate(self._slot_backend.tolist()):
            k = self._id_to_key.get(int(kid))
            if k is None:
                continue
            st = "busy" if self._busy_slots.get(slot) else "idle"
            d = conns.setdefault(k, {})
            d[st] = d.get(st, 0) + 1
        return {"backends": dict(self.p_backends), "connections": conns, "dead_backends": [],
                "state": "stopping" if self._stopping else "running", "counters": self.getStats()["counters"],
                "options": {"spares": self.p_spares, "maximum": self.p_max}}

    # ------------------------------------------------------------ sizing (low-pass filter, _rebalance's target)
    def _occupancy(self):
        """(connections, idle slots, waiters) from the device queue's counters; initq is empty in this model
        (a slot of the table is connected)."""
        total = len(self._slot_backend)
        if self.p_codel is None:
            return total, self._host_idle, len(self._host_waiters)
        if not self._queue_ready:
            return total, self._idle_before_the_queue(), 0
        info = self._ctx.queue_info()
        return total, int(info.idle), int(info.tail - info.head)

    def lpTick(self):
        """The body of the pool's 5 Hz sampler (lib/pool.js:251-262; the host calls it every LP_INT ms):
        feeds busy + spares to the filter and, while the last rebalance was clamped, rebalances again."""
        conns, spares, _ = self._occupancy()
        self.p_lpf.put(conns - spares + self.p_spares)
        if self.p_lastRebalClamped:
            return self.rebalance()
        return None

    def rebalance(self):
        """_rebalance up to the plan (lib/pool.js:541-592): the target size, then planRebalance over the
        connections per backend in preference order -- on the device (cb_plan_rebalance).  Returns
        {"target", "clamped", "add": [key...], "remove": [(key, index of the connection)...]} or None when
        the reference returns early (:521-526).  Applying the plan (sockets, the churn limiter of :599-658)
        is the host's."""
        if len(self.p_keys) < 1 or self._stopping:
            return None
        total, idle, waiters = self._occupancy()
        target, clamped = rebalanceTarget(self.p_lpf, total, idle, 0, waiters, self.p_spares, self.p_max)
        self.p_lastRebalClamped = clamped
        ids = [self._key_ids[k] for k in self.p_keys]
        have = [int((self._slot_backend == i).sum()) for i in ids]
        _, add, remove = self._ctx.plan_rebalance(have, [0] * len(ids), target, self.p_max)
        plan = {"target": target, "clamped": clamped, "add": [self.p_keys[int(i)] for i in add],
                "remove": [(self.p_keys[int(k)], int(o)) for k, o in remove]}
        self.p_lastRebalance = plan
        return plan

    # ------------------------------------------------------------ resolver
    def on_resolver_added(self, k, backend=None):  # lib/pool.js:285-291
        kid = self._key_ids.get(k)
        if kid is None:
            kid = self._next_id
            self._next_id += 1
            self._key_ids[k] = kid
            self._id_to_key[kid] = k
        u = float(self._random())
        idx = int(np.floor(u * (len(self.p_keys) + 1)))
        self.p_keys.insert(idx, k)
        self.p_backends[k] = backend if backend is not None else {"key": k}
        self._ctx.backend_added(kid, u, *self._srv(backend))
        self._refresh_slots()

    def on_resolver_removed(self, k):  # lib/pool.js:293-313
        if k not in self._key_ids or k not in self.p_keys:
            raise AssertionError("resolver key %s not found" % k)
        self.p_keys.remove(k)
        self.p_backends.pop(k, None)
        self._ctx.backend_removed(self._key_ids[k])
        self._refresh_slots()

    def reshuffle(self):  # lib/pool.js:501-519
        if len(self.p_keys) <= 1:
            return
        u = float(self._random())
        taken = self.p_keys.pop()
        self.p_keys.insert(int(np.floor(u * (len(self.p_keys) + 1))), taken)
        self._ctx.reshuffle(u)
        self._refresh_slots()

    @staticmethod
    def _srv(backend):
        if not backend:
            return (0, 0)
        return (int(backend.get("priority", 0)), int(backend.get("weight", 0)))

    def _refresh_slots(self):
        keys, slots = self._ctx.backends_get()
        assert [self._id_to_key[int(i)] for i in keys] == self.p_keys, "preference list diverged"
        old_n = len(self._slot_backend)
        self._slot_backend = slots
        new_n = len(slots)
        if new_n == old_n:
            return
        # A topology change in the reference opens / closes sockets; in this model the slots of the
        # new table that no claimed handle holds are idle.  New slots turn idle as an event (they
        # drain waiters, lib/pool.js:715); slots that went away leave the idle queue.
        if self.p_codel is None:
            if self._initial_idle is not None and not self._host_started:
                self._host_idle = min(int(self._initial_idle), new_n)  # the rest connect later: slots_idle()
            else:
                self._host_idle = max(0, new_n - self._claimed)
            self._host_started = True
            self._host_drain(self._clock())
            return
        if not self._queue_ready:
            return  # created with the table's size at the first claim / idle event
        idle_now = int(self._ctx.queue_info().idle)
        want_idle = max(0, new_n - self._claimed)
        if want_idle > idle_now:
            self._ctx.queue_idle(self._clock(), want_idle - idle_now)
            self._dispatch()
        elif want_idle < idle_now:
            self._ctx.queue_trim_idle(idle_now - want_idle)

    # --------------------------------------------------------------- claim
    def claim(self, options=None, cb=None):  # lib/pool.js:859-969
        if callable(options) and cb is None:
            cb, options = options, {}
        options = options or {}
        if self.p_codel is not None and isinstance(options.get("timeout"), (int, float)):
            raise Exception("options.timeout not allowed when targetDelay has been set")  # :875-878
        if self._stopping:  # :889-899
            if cb:
                cb(PoolStoppingError(self), None, None)
            return {"cancel": lambda: None}
        now = self._clock()
        if self.p_codel is None:
            return self._host_claim(options, cb, now)
        if options.get("errorOnEmpty") and len(self.p_keys) < 1:  # :953-957
            handle = CueBallClaimHandle(self, cb, now, float("inf"), -1)
            handle.ch_state = "failed"
            if cb:
                cb(NoBackendsError(self), None, None)
            return handle
        self._ensure_queue(len(self._slot_backend))
        # one kernel: timers that came due, getMaxIdle (:880), idle slot first (:934-951), else enqueue (:960-965)
        ticket, tmo = self._ctx.queue_push([now], want_timeouts=True)
        handle = CueBallClaimHandle(self, cb, now, float(tmo[0]), ticket)
        self._handles[ticket] = handle
        heapq.heappush(self._deadlines, (now + int(float(tmo[0]) * 1e6), ticket))
        if len(self._deadlines) > 2 * len(self._handles) + self._DEADLINE_SLACK:  # resolved handles' entries: drop them in bulk
            self._deadlines = [e for e in self._deadlines if e[1] in self._handles]
            heapq.heapify(self._deadlines)
        self._dispatch()
        return handle

    def nextDeadline(self):
        """The earliest clock value (ns) at which a waiting handle's claim timer fires (its S.timeout(claimTimeout),
        lib/connection-fsm.js:626-642), or None: when a host with its own timer should call expire()."""
        while self._deadlines and self._deadlines[0][1] not in self._handles:
            heapq.heappop(self._deadlines)
        return self._deadlines[0][0] if self._deadlines else None

    def stop(self):
        self._stopping = True

    # ---------------------------------------------------------- events
    def _cancelled(self, handle):
        if self.p_codel is None:
            return  # stays in the host FIFO, skipped at dequeue
        self._ctx.queue_cancel([handle.ch_ticket])
        self._handles.pop(handle.ch_ticket, None)

    def _slot_released(self, slot):
        """A slot FSM returned to 'idle' (lib/connection-fsm.js:1141-1147 -> pool.js:715)."""
        self._claimed -= 1
        if self._busy_slots.get(slot, 0) > 1:
            self._busy_slots[slot] -= 1
        else:
            self._busy_slots.pop(slot, None)
        self.slots_idle(1)

    def slots_idle(self, k, now_ns=None):
        """k slots turned idle at one clock value: the handler of lib/pool.js:715-760 once per slot."""
        k = int(k)
        if k <= 0:
            return
        now = self._clock() if now_ns is None else now_ns
        if self.p_codel is None:
            self._host_idle += k
            self._host_drain(now)
            return
        self._ensure_queue(len(self._slot_backend))
        self._ctx.queue_idle(now, k)
        self._dispatch()

    def expire(self, now_ns=None):
        """Fire the claim timers that have come due (each handle's S.timeout(claimTimeout),
        lib/connection-fsm.js:626-642).  claim() and slots_idle() do this first anyway; a host with
        long checkouts calls it from its own timer."""
        now = self._clock() if now_ns is None else now_ns
        if self.p_codel is None:
            self._host_expire(now)
            return
        if self._queue_ready:
            self._ctx.queue_expire(now)
            self._dispatch()

    # ---------------------------------------------------------- the pool without CoDel (host only)
    def _hc(self, c, n=1):
        self._host_counters[c] = self._host_counters.get(c, 0) + n

    def _host_claim(self, options, cb, now):
        timeout = options["timeout"] if isinstance(options.get("timeout"), (int, float)) else float("inf")
        self._hc("claim")
        self._host_expire(now)
        handle = CueBallClaimHandle(self, cb, now, timeout, -1)
        if self._host_idle > 0:
            self._host_idle -= 1
            self._host_admit(handle)
            return handle
        if options.get("errorOnEmpty") and len(self.p_keys) < 1:
            handle.ch_state = "failed"
            if cb:
                cb(NoBackendsError(self), None, None)
            return handle
        self._host_waiters.append(handle)
        if self._host_counters.get("max-claim-queue", 0) < len(self._host_waiters):
            self._host_counters["max-claim-queue"] = len(self._host_waiters)
        self._hc("queued-claim")
        return handle

    def _host_admit(self, handle):
        n = len(self._slot_backend)
        slot = self._host_cursor % n if n else -1
        self._host_cursor += 1
        self._claimed += 1
        if slot >= 0:
            self._busy_slots[slot] = self._busy_slots.get(slot, 0) + 1
        handle.try_(slot, int(self._slot_backend[slot]) if n else -1)

    def _host_expire(self, now):
        for h in self._host_waiters:
            if h.ch_state == "waiting" and now >= h.ch_started_ns + h.ch_claimTimeout * 1e6:
                self._hc("claim-timeout")
                h.timeout()

    def _host_drain(self, now):
        self._host_expire(now)
        while self._host_idle > 0 and self._host_waiters:
            h = self._host_waiters.pop(0)
            if h.ch_state != "waiting":
                continue
            self._host_idle -= 1
            self._host_admit(h)
