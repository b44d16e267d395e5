"""Error names and messages of the reference (lib/errors.js) for the conditions this path raises.  The reference
formats the pool's uuid and domain into the message (VError); the mirror's pools carry `p_uuid` / `p_domain` when
the host gives them."""


def _ident(pool):
    if pool is None:
        return ""
    uuid, dom = getattr(pool, "p_uuid", None), getattr(pool, "p_domain", None)
    return " %s (%s)" % (uuid, dom) if uuid or dom else ""


class ClaimTimeoutError(Exception):
    """A claim dropped by CoDel or timed out in the queue (lib/errors.js:35-43;
    raised from CueBallClaimHandle state 'failed', lib/connection-fsm.js:633-638)."""

    def __init__(self, pool=None):
        Exception.__init__(self, "Timed out while waiting for connection in pool" + _ident(pool))
        self.pool = pool


class NoBackendsError(Exception):
    """lib/errors.js:45-54: no backends available (lib/pool.js:953-957)."""

    def __init__(self, pool=None):
        Exception.__init__(self, "No backends available in pool" + _ident(pool))
        self.pool = pool


class PoolStoppingError(Exception):
    """lib/errors.js:71-79: claim() on a stopping pool (lib/pool.js:889-899)."""

    def __init__(self, pool=None):
        Exception.__init__(self, "Pool%s is stopping and cannot take new requests" % _ident(pool))
        self.pool = pool


class ClaimHandleMisusedError(Exception):
    """lib/errors.js:25-33: the handle used as if it were the socket; here also release()/close() on a handle that
    does not hold a connection."""

    def __init__(self, detail=None):
        Exception.__init__(self, "CueBall claim handle used as if it was a socket (Check the order and number of "
                                 "arguments in your claim callbacks)" + (": " + detail if detail else ""))
