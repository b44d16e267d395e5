"""Error names of the reference (lib/errors.js) for the conditions this path raises."""


class ClaimTimeoutError(Exception):
    """A claim dropped by CoDel or timed out in the queue (lib/errors.js:35-43;
    raised from CueBallClaimHandle state 'failed', lib/connection-fsm.js:633-638)."""

    def __init__(self, pool=None):
        Exception.__init__(self, "Timed out while waiting for connection in pool")
        self.pool = pool


class NoBackendsError(Exception):
    """lib/errors.js: no backends available (lib/pool.js:953-957)."""


class PoolStoppingError(Exception):
    """lib/errors.js: claim() on a stopping pool (lib/pool.js:889-899)."""


class ClaimHandleMisusedError(Exception):
    """lib/errors.js: release()/close() on a handle in the wrong state."""
