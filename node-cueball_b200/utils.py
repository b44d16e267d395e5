"""utils -- the hot-path subset of the reference's lib/utils.js.

planRebalance keeps the reference's signature and result shape (lib/utils.js:239-393):
    plan = planRebalance(connections, dead, target, max[, singleton])
    connections: dict key -> list of connection objects (insertion-ordered, like a JS object)
    dead:        dict key -> True
    plan:        {'add': [key...], 'remove': [connection...]}
The plan itself is computed by the CUDA library (cb_plan_rebalance).
"""
import time

from .context import Context

_default_ctx = None


def _ctx():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


def currentMillis():
    """lib/utils.js:198-204 over a monotone ns clock (same three roundings)."""
    t = time.monotonic_ns()
    return float((t // 10**9) * 1000) + (t % 10**9) / 1000000.0


def planRebalance(connections, dead, target, max, singleton=False, context=None):
    ctx = context or _ctx()
    if not isinstance(connections, dict):
        raise AssertionError("connections (object) is required")
    keys = list(connections.keys())
    have = [len(connections[k] or []) for k in keys]
    deadv = [1 if dead.get(k) is True else 0 for k in keys]
    try:
        want, add, rem = ctx.plan_rebalance(have, deadv, target, max, singleton)
    except Exception as e:  # the reference asserts (AssertionError) on bad target/max
        if getattr(e, "code", None) == -1:
            raise AssertionError(str(e))
        raise
    return {"add": [keys[i] for i in add], "remove": [connections[keys[k]][o] for k, o in rem]}


def shuffle(array, draws=None, context=None):
    """lib/utils.js:207-217: Fisher-Yates in place, returns the array.  `draws` (one float in [0,1) per
    element, in the order the reference would call Math.random()) default to random.random(); the
    permutation itself is computed by the CUDA library (cb_shuffle) on the element indices."""
    import random
    n = len(array)
    if draws is None:
        draws = [random.random() for _ in range(n)]
    perm = (context or _ctx()).shuffle(list(range(n)), draws)
    items = [array[int(i)] for i in perm]
    array[:] = items
    return array


def assertClaimDelay(delay):
    """lib/utils.js:188-195"""
    if delay is None:
        return
    if isinstance(delay, bool) or not isinstance(delay, (int, float)) or delay != delay or delay in (float("inf"), float("-inf")):
        raise AssertionError("options.targetClaimDelay (finite) is required")
    if not delay > 0:
        raise AssertionError("options.targetClaimDelay > 0")
    if delay != int(delay):
        raise AssertionError("options.targetClaimDelay")
