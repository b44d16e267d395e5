"""node-cueball_b200 -- the B200-native claim-queue admission path of node-cueball.

Host-side mirror of the reference's interface for this path (names as in the reference):

    ControlledDelay          lib/codel.js
    Queue                    lib/queue.js
    planRebalance            lib/utils.js:239-393
    ConnectionPool           lib/pool.js (claim / release / resolver events / getStats / kang snapshot /
                             the sizing loop: FIRFilter, lpTick, rebalance)
    StaticIpResolver, srvKey lib/resolver.js:1157-1171,1387-1456 (backend identity, event stream)
    ClaimTimeoutError ...    lib/errors.js

All arithmetic of the claim path runs in libcueball_b200.so (hand-written sm_100a CUDA behind the C ABI in
include/cueball_b200.h); the one host computation is the pool's 5 Hz low-pass filter (pool.FIRFilter).  There is no CPU fallback: importing works anywhere, but creating a
Context needs the built library and a B200.

The directory name carries a hyphen (it is the repo's mandated package directory); import
it as `cueball_b200` through the loader module at the repository root.
"""
from . import _native
from ._native import (ADMIT, BALANCE_SRV_WEIGHTED, BALANCE_UNIFORM, DROP, F_EMPTY_AFTER, F_WAITING,
                      PATH_PARALLEL, PATH_SEQUENTIAL, SKIPPED, CueballError)
from .codel import ControlledDelay
from .context import Context
from .errors import ClaimHandleMisusedError, ClaimTimeoutError, NoBackendsError, PoolStoppingError
from .pool import ConnectionPool, CueBallClaimHandle, FIRFilter
from .queue import Queue
from .resolver import DeviceResolverState, StaticIpResolver, process_backends, srvKey
from .utils import assertClaimDelay, currentMillis, planRebalance

__all__ = ["ADMIT", "DROP", "SKIPPED", "F_WAITING", "F_EMPTY_AFTER", "PATH_PARALLEL", "PATH_SEQUENTIAL",
           "BALANCE_UNIFORM", "BALANCE_SRV_WEIGHTED", "CueballError", "ControlledDelay", "Context",
           "ClaimHandleMisusedError", "ClaimTimeoutError", "NoBackendsError", "PoolStoppingError",
           "ConnectionPool", "CueBallClaimHandle", "FIRFilter", "Queue", "assertClaimDelay", "currentMillis",
           "planRebalance", "StaticIpResolver", "DeviceResolverState", "process_backends", "srvKey", "_native"]
