"""ControlledDelay -- same surface as the reference's lib/codel.js, computed on the B200.

    cd = ControlledDelay(targetClaimDelay)        lib/codel.js:24-32
    cd.overloaded(start_ns[, now_ns]) -> bool     lib/codel.js:56-86
    cd.empty([now_ns])                            lib/codel.js:91-94
    cd.getMaxIdle([now_ns]) -> ms                 lib/codel.js:109-118

Timestamps are process.hrtime()-style nanoseconds; the clock defaults to
time.monotonic_ns (utils.currentMillis reads process.hrtime, lib/utils.js:198-204).
overloaded_batch() is the batch form the drop-in pool uses (one device call per batch,
never per claim).
"""
import time

from .context import Context


class ControlledDelay(object):
    def __init__(self, targetClaimDelay, context=None, clock=None):
        self._ctx = context if context is not None else Context()
        self._ctx.codel_init(targetClaimDelay)  # asserts like utils.assertClaimDelay
        self.cd_targdelay = targetClaimDelay
        self._clock = clock or time.monotonic_ns

    @property
    def context(self):
        return self._ctx

    def overloaded(self, start_ns, now_ns=None):
        return self._ctx.overloaded(start_ns, self._clock() if now_ns is None else now_ns)

    def overloaded_batch(self, start, now, flags=None, **kw):
        return self._ctx.admit_batch(start, now, flags, **kw)

    def empty(self, now_ns=None):
        self._ctx.codel_empty(self._clock() if now_ns is None else now_ns)

    def getMaxIdle(self, now_ns=None):
        return self._ctx.max_idle(self._clock() if now_ns is None else now_ns)

    # the reference's fields, read back from the device
    def _st(self):
        return self._ctx.get_state()

    cd_first_above_time = property(lambda self: self._st().first_above_time)
    cd_drop_next = property(lambda self: self._st().drop_next)
    cd_count = property(lambda self: self._st().count)
    cd_dropping = property(lambda self: bool(self._st().dropping))

    @property
    def cd_last_empty(self):
        st = self._st()
        return st.last_empty if st.has_last_empty else None
