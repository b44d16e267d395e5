"""ctypes binding of libcueball_b200.so (the C ABI in include/cueball_b200.h).

There is no CPU fallback: if the shared library is missing, or no sm_100 device is
present, loading / cb_create fails loudly.  The product never touches the CPU checkers.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcueball_b200.so")

CB_OK, CB_EINVAL, CB_ECUDA, CB_ENOMEM, CB_ESTATE, CB_ENCCL = 0, -1, -2, -3, -4, -5
ADMIT, DROP, SKIPPED = 0, 1, 2
F_WAITING, F_EMPTY_AFTER = 1, 2
PATH_PARALLEL, PATH_SEQUENTIAL = 0, 1
BALANCE_UNIFORM, BALANCE_SRV_WEIGHTED = 0, 1
K_SCAN, K_RESOLVE, K_EMIT, K_COUNT = 0, 1, 2, 3
SRVKEY_LEN = 28

# every symbol include/cueball_b200.h declares (tests check the library exports them all)
SYMBOLS = [
    "cb_create", "cb_destroy", "cb_last_error", "cb_abi_version", "cb_set_stream", "cb_set_pipeline_depth", "cb_join", "cb_synchronize",
    "cb_host_alloc", "cb_host_free", "cb_codel_init", "cb_codel_get_state", "cb_codel_set_state",
    "cb_codel_empty", "cb_codel_max_idle", "cb_admit_batch", "cb_admit_batch_ms", "cb_admit_batch_dev",
    "cb_admit_batch_ms_dev", "cb_batch_stats_get", "cb_codel_overloaded", "cb_codel_overloaded_ms", "cb_backends_set",
    "cb_backend_added", "cb_backend_removed", "cb_reshuffle", "cb_backends_get", "cb_plan_rebalance", "cb_shuffle",
    "cb_comm_unique_id", "cb_comm_init", "cb_allreduce_counts", "cb_profile_enable", "cb_profile_get",
    "cb_launch_count", "cb_debug_timestamps", "cb_debug_lane_timestamps",
    "cb_srv_keys", "cb_resolver_process",
    "cb_ring_mailbox", "cb_ring_connect", "cb_peers_board", "cb_peers_connect", "cb_peers_counts",
    "cb_queue_create", "cb_queue_push", "cb_queue_cancel", "cb_queue_expire", "cb_queue_idle", "cb_queue_trim_idle", "cb_queue_results",
    "cb_queue_info_get",
]


class CodelState(C.Structure):
    _fields_ = [("target_ms", C.c_double), ("first_above_time", C.c_double), ("drop_next", C.c_double),
                ("count", C.c_uint64), ("dropping", C.c_int32), ("has_last_empty", C.c_int32),
                ("last_empty", C.c_double)]

    def as_dict(self):
        return {"count": float(self.count), "drop_next": self.drop_next,
                "first_above_time": self.first_above_time, "dropping": bool(self.dropping),
                "last_empty": self.last_empty if self.has_last_empty else None}


class BatchStats(C.Structure):
    _fields_ = [("n", C.c_uint64), ("drops", C.c_uint64), ("admits", C.c_uint64), ("skipped", C.c_uint64),
                ("cursor", C.c_uint64), ("total_drops", C.c_uint64), ("total_admits", C.c_uint64),
                ("global_drops", C.c_uint64), ("global_admits", C.c_uint64), ("global_n", C.c_uint64),
                ("state", CodelState), ("path", C.c_int32), ("n_events", C.c_uint32),
                ("n_intervals", C.c_uint32), ("reserved", C.c_uint32)]


class GlobalCounts(C.Structure):
    _fields_ = [("steps", C.c_uint64), ("drops", C.c_uint64), ("admits", C.c_uint64), ("n", C.c_uint64)]


class QueueInfo(C.Structure):
    _fields_ = [("head", C.c_uint64), ("tail", C.c_uint64), ("next_ticket", C.c_uint64), ("idle", C.c_uint64),
                ("claim", C.c_uint64), ("queued_claim", C.c_uint64), ("max_claim_queue", C.c_uint64),
                ("claim_timeout", C.c_uint64), ("n_results", C.c_uint64), ("error", C.c_uint32),
                ("reserved", C.c_uint32)]

    def counters(self):
        """The reference's counter names (lib/pool.js:269-283)."""
        return {"claim": self.claim, "queued-claim": self.queued_claim, "max-claim-queue": self.max_claim_queue,
                "claim-timeout": self.claim_timeout}


class Backend(C.Structure):
    _fields_ = [("priority", C.c_int32), ("weight", C.c_int32)]


class CueballError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "cueball_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """Load the shared library (built in-tree by __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback for this path)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, dbl, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double, C.c_size_t
    L.cb_create.argtypes = [C.POINTER(vp), C.c_int]
    L.cb_destroy.argtypes = [vp]
    L.cb_destroy.restype = None
    L.cb_last_error.argtypes = [vp]
    L.cb_last_error.restype = C.c_char_p
    L.cb_abi_version.restype = C.c_int
    L.cb_set_stream.argtypes = [vp, vp]
    L.cb_synchronize.argtypes = [vp]
    L.cb_set_pipeline_depth.argtypes = [vp, C.c_int]
    L.cb_join.argtypes = [vp]
    L.cb_host_alloc.argtypes = [vp, sz]
    L.cb_host_alloc.restype = vp
    L.cb_host_free.argtypes = [vp, vp]
    L.cb_host_free.restype = None
    L.cb_codel_init.argtypes = [vp, dbl]
    L.cb_codel_get_state.argtypes = [vp, C.POINTER(CodelState)]
    L.cb_codel_set_state.argtypes = [vp, C.POINTER(CodelState)]
    L.cb_codel_empty.argtypes = [vp, i64]
    L.cb_codel_max_idle.argtypes = [vp, i64, C.POINTER(dbl)]
    for name in ("cb_admit_batch", "cb_admit_batch_ms"):
        getattr(L, name).argtypes = [vp, vp, vp, vp, sz, vp, vp, vp, C.POINTER(BatchStats)]
    for name in ("cb_admit_batch_dev", "cb_admit_batch_ms_dev"):
        getattr(L, name).argtypes = [vp, vp, vp, vp, sz, vp, vp, vp]
    L.cb_batch_stats_get.argtypes = [vp, C.POINTER(BatchStats)]
    L.cb_codel_overloaded.argtypes = [vp, i64, i64, C.POINTER(C.c_int)]
    L.cb_codel_overloaded_ms.argtypes = [vp, dbl, dbl, C.POINTER(C.c_int)]
    L.cb_backends_set.argtypes = [vp, vp, i32, vp, i32, C.c_int]
    L.cb_backend_added.argtypes = [vp, i32, vp, dbl]
    L.cb_backend_removed.argtypes = [vp, i32]
    L.cb_reshuffle.argtypes = [vp, dbl]
    L.cb_shuffle.argtypes = [vp, vp, sz, vp]
    L.cb_backends_get.argtypes = [vp, C.POINTER(i32), vp, C.POINTER(i32), vp]
    L.cb_plan_rebalance.argtypes = [vp, vp, vp, i32, i32, i32, C.c_int, vp, vp, C.POINTER(i32), vp, vp,
                                    C.POINTER(i32)]
    L.cb_comm_unique_id.argtypes = [vp, C.POINTER(sz)]
    L.cb_comm_init.argtypes = [vp, C.c_int, C.c_int, vp, sz]
    L.cb_allreduce_counts.argtypes = [vp]
    L.cb_profile_enable.argtypes = [vp, C.c_int]
    L.cb_profile_get.argtypes = [vp, C.POINTER(dbl), C.POINTER(u64)]
    L.cb_launch_count.argtypes = [vp]
    L.cb_launch_count.restype = u64
    L.cb_debug_timestamps.argtypes = [vp, C.POINTER(u64)]
    L.cb_debug_lane_timestamps.argtypes = [vp, C.c_int, C.POINTER(u64)]
    L.cb_ring_mailbox.argtypes = [vp, vp, C.POINTER(sz), C.POINTER(vp)]
    L.cb_ring_connect.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.cb_srv_keys.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(i32), C.POINTER(C.c_char_p), sz, vp]
    L.cb_resolver_process.argtypes = [vp, vp, sz, vp, vp, C.POINTER(i32), vp, vp, C.POINTER(i32), C.POINTER(i32)]
    L.cb_peers_board.argtypes = [vp, vp, C.POINTER(sz), C.POINTER(vp)]
    L.cb_peers_connect.argtypes = [vp, C.c_int, C.c_int, vp, C.POINTER(vp)]
    L.cb_peers_counts.argtypes = [vp, C.POINTER(GlobalCounts)]
    L.cb_queue_create.argtypes = [vp, sz, sz, u64]
    L.cb_queue_push.argtypes = [vp, vp, sz, C.POINTER(i64), vp]
    L.cb_queue_cancel.argtypes = [vp, vp, sz]
    L.cb_queue_expire.argtypes = [vp, i64]
    L.cb_queue_idle.argtypes = [vp, i64, u64]
    L.cb_queue_trim_idle.argtypes = [vp, u64, C.POINTER(u64)]
    L.cb_queue_results.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.cb_queue_info_get.argtypes = [vp, C.POINTER(QueueInfo)]
    _lib = L
    return L


def check(ctx, rc):
    if rc != CB_OK:
        msg = lib().cb_last_error(ctx)
        raise CueballError(rc, msg.decode() if msg else "?")
    return rc
