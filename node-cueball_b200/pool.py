"""ConnectionPool -- the claim()/release() surface of the reference's lib/pool.js for the
claim-queue path, with the admission work done by the CUDA library.

What is mirrored (reference file:line):
  claim([options,] cb) -> handle{cancel()}          lib/pool.js:859-969
      options.timeout forbidden with targetClaimDelay   :874-878
      timeout = codel.getMaxIdle()                      :880
      idle-first fast path / enqueue on p_waiters       :929-966
  slot goes idle -> drain loop                      lib/pool.js:715-760
      shift, codel.overloaded(hdl.ch_started), skip non-waiting, timeout() or try()
      queue emptied with the slot still idle -> codel.empty(), slot onto p_idleq
  on_resolver_added / on_resolver_removed / reshuffle   lib/pool.js:285-313, 501-519
  getStats() counters                               lib/pool.js:269-283, 834-857
  handle states waiting/claimed/released/closed/cancelled/failed
                                                    lib/connection-fsm.js:489-808

What is NOT here (out of scope, stays in Node in a real deployment): sockets, the slot and
socket FSMs, retry/backoff, the FIR low-pass filter and the churn limiter.  A "slot" here is
an integer; its backend comes from the device-side slot table.

The drain is batched: when k slots are idle at clock `now`, the waiters are examined in
FIFO order exactly as k consecutive runs of the reference loop would (each run ends at its
first admit).  That prefix is found with one trial device batch over the queue and committed
with a second batch over the prefix, so CoDel's state only ever advances over claims that
were really dequeued.
"""
import time

import numpy as np

from . import _native as N
from .context import Context
from .errors import ClaimHandleMisusedError, ClaimTimeoutError, NoBackendsError, PoolStoppingError
from .queue import Queue, snapshot
from .utils import assertClaimDelay


class CueBallClaimHandle(object):
    """lib/connection-fsm.js:489-808 reduced to what the claim queue touches."""

    def __init__(self, pool, callback, started_ns, claim_timeout_ms):
        self.ch_pool = pool
        self.ch_callback = callback
        self.ch_started_ns = started_ns  # ch_started, :523
        self.ch_claimTimeout = claim_timeout_ms
        self.ch_state = "waiting"
        self.ch_slot = None
        self.ch_backend = None

    def isInState(self, s):
        return self.ch_state == s

    def cancel(self):  # :580
        if self.ch_state == "waiting":
            self.ch_state = "cancelled"

    def timeout(self):  # :589, :633-638
        if self.ch_state != "waiting":
            return
        self.ch_state = "failed"
        self.ch_pool._incrCounter("claim-timeout")
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
    def __init__(self, options):
        assertClaimDelay(options.get("targetClaimDelay"))
        self.p_max = int(options.get("maximum", 0))
        self._clock = options.get("clock") or time.monotonic_ns
        self._random = options.get("random") or np.random.random
        self._ctx = options.get("context") or Context(options.get("device", 0))
        self.p_codel = None
        if options.get("targetClaimDelay") is not None:
            self._ctx.codel_init(options["targetClaimDelay"])  # lib/pool.js:197-200
            self.p_codel = self._ctx
        self.p_waiters = Queue()
        self.p_idleq = Queue()
        self.p_keys = []       # preference list as the caller's keys (device holds indices)
        self._key_ids = {}
        self._next_id = 0
        self.p_counters = {}
        self._pending_idle = 0
        self._stopping = False
        self._slot_backend = np.zeros(0, np.int32)
        self._id_to_key = {}
        self._ctx.backends_set(0, [], self.p_max)  # empty preference list, `maximum` slots wanted

    # ------------------------------------------------------------ counters
    def _incrCounter(self, c, n=1):  # lib/pool.js:269-275
        self.p_counters[c] = self.p_counters.get(c, 0) + n

    def _hwmCounter(self, c, v):  # lib/pool.js:277-283
        if self.p_counters.get(c, 0) < v:
            self.p_counters[c] = v

    def getStats(self):  # lib/pool.js:834-857 (the fields that exist on this path)
        return {"counters": dict(self.p_counters), "totalConnections": self.p_max,
                "idleConnections": self.p_idleq.length, "waiterCount": self.p_waiters.length}

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
        self._ctx.backend_added(kid, u, *self._srv(backend))
        self._refresh_slots()

    def on_resolver_removed(self, k):  # lib/pool.js:293-313
        if k not in self._key_ids or k not in self.p_keys:
            raise AssertionError("resolver key %s not found" % k)
        self.p_keys.remove(k)
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
        old = self._slot_backend
        self._slot_backend = slots
        # (re)build the idle queue for slots that are not claimed: a topology change in the
        # reference closes/opens sockets; in this model every slot of the new table starts idle
        if len(old) != len(slots):
            self.p_idleq = Queue()
            for s in range(len(slots)):
                self.p_idleq.push(s)

    # --------------------------------------------------------------- claim
    def claim(self, options=None, cb=None):  # lib/pool.js:859-969
        if callable(options) and cb is None:
            cb, options = options, {}
        options = options or {}
        if self.p_codel is not None:
            if isinstance(options.get("timeout"), (int, float)):
                raise Exception("options.timeout not allowed when targetDelay has been set")  # :875-878
            timeout = self._ctx.max_idle(self._clock())  # :880
        elif isinstance(options.get("timeout"), (int, float)):
            timeout = options["timeout"]
        else:
            timeout = float("inf")
        self._incrCounter("claim")  # :887
        if self._stopping:
            if cb:
                cb(PoolStoppingError(), None, None)
            return {"cancel": lambda: None}
        handle = CueBallClaimHandle(self, cb, self._clock(), timeout)
        # tryNext, :929-966
        while self.p_idleq.length > 0:
            slot = self.p_idleq.shift()
            handle.try_(slot, int(self._slot_backend[slot]))
            return handle
        if options.get("errorOnEmpty") and len(self.p_keys) < 1:
            handle.ch_state = "failed"
            if cb:
                cb(NoBackendsError(), None, None)
            return handle
        self.p_waiters.push(handle)
        self._hwmCounter("max-claim-queue", self.p_waiters.length)
        self._incrCounter("queued-claim")
        return handle

    def stop(self):
        self._stopping = True

    # ---------------------------------------------------------- drain loop
    def _slot_released(self, slot):
        """A slot FSM returned to 'idle' (lib/connection-fsm.js:1141-1147 -> pool.js:715)."""
        self._drain([slot], self._clock())

    def slots_idle(self, slots, now_ns=None):
        """Several slots turned idle at (about) the same clock value."""
        self._drain(list(slots), self._clock() if now_ns is None else now_ns)

    def _drain(self, slots, now_ns):
        k = len(slots)
        if k == 0:
            return
        handles = self.p_waiters.take_all()
        n = len(handles)
        if n == 0:  # lib/pool.js:751-757
            if self.p_codel is not None:
                self._ctx.codel_empty(now_ns)
            for s in slots:
                self.p_idleq.push(s)
            return
        start, flags = snapshot(handles)
        now = np.full(n, now_ns, dtype=np.int64)
        if self.p_codel is not None:
            saved = self._ctx.get_state().as_dict()
            dec, _, _, _ = self._ctx.admit_batch(start, now, flags, want_slot=False)
            adm = np.flatnonzero(dec == N.ADMIT)
            if adm.size >= k:
                L = int(adm[k - 1]) + 1
            else:
                L = n
            if L < n:  # commit only the prefix that k runs of the loop would have dequeued
                saved["count"] = int(saved["count"])
                self._ctx.set_state(saved)
                dec, _, _, _ = self._ctx.admit_batch(start[:L], now[:L], flags[:L], want_slot=False)
        else:
            dec = np.where(flags & 1, N.ADMIT, N.SKIPPED).astype(np.uint8)
            adm = np.flatnonzero(dec == N.ADMIT)
            L = int(adm[k - 1]) + 1 if adm.size >= k else n
            dec = dec[:L]
        free = list(slots)
        for i in range(L):
            h = handles[i]
            if dec[i] == N.SKIPPED:
                continue
            if dec[i] == N.DROP:
                h.timeout()
                continue
            s = free.pop(0)
            h.try_(s, int(self._slot_backend[s]) if s < len(self._slot_backend) else -1)
        for h in handles[L:]:  # untouched tail goes back in FIFO order
            self.p_waiters.push(h)
        if free:  # queue drained with slots still idle: lib/pool.js:751-757
            if self.p_codel is not None:
                self._ctx.codel_empty(now_ns)
            for s in free:
                self.p_idleq.push(s)
