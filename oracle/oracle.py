"""oracle/oracle.py -- ctypes front-end of oracle/liboracle.so (the plain-C
restatement of the reference's claim-queue hot path).

TEST INFRASTRUCTURE ONLY (see oracle/cueball_oracle.h).  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")

ADMIT, DROP, SKIPPED = 0, 1, 2
F_WAITING, F_EMPTY_AFTER = 1, 2


class Codel(C.Structure):
    _fields_ = [("cd_targdelay", C.c_double), ("cd_first_above_time", C.c_double),
                ("cd_drop_next", C.c_double), ("cd_count", C.c_double),
                ("cd_dropping", C.c_int), ("cd_has_last_empty", C.c_int),
                ("cd_last_empty", C.c_double)]

    def as_dict(self):
        return {"count": self.cd_count, "drop_next": self.cd_drop_next,
                "first_above_time": self.cd_first_above_time,
                "dropping": bool(self.cd_dropping),
                "last_empty": self.cd_last_empty if self.cd_has_last_empty else None}

    def load(self, st):
        self.cd_count = float(st["count"])
        self.cd_drop_next = float(st["drop_next"])
        self.cd_first_above_time = float(st["first_above_time"])
        self.cd_dropping = 1 if st["dropping"] else 0
        le = st.get("last_empty")
        self.cd_has_last_empty = 0 if le is None else 1
        self.cd_last_empty = 0.0 if le is None else float(le)


class Stats(C.Structure):
    _fields_ = [("drops", C.c_uint64), ("admits", C.c_uint64), ("skipped", C.c_uint64),
                ("cursor", C.c_uint64)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.cbo_ns_to_ms.restype = C.c_double
        L.cbo_ns_to_ms.argtypes = [C.c_int64]
        L.cbo_current_millis.restype = C.c_double
        L.cbo_current_millis.argtypes = [C.c_int64, C.c_int64]
        L.cbo_codel_init.argtypes = [C.POINTER(Codel), C.c_double]
        L.cbo_codel_overloaded.argtypes = [C.POINTER(Codel), C.c_double, C.c_double]
        L.cbo_codel_empty.argtypes = [C.POINTER(Codel), C.c_double]
        L.cbo_codel_get_max_idle.restype = C.c_double
        L.cbo_codel_get_max_idle.argtypes = [C.POINTER(Codel), C.c_double]
        for name in ("cbo_drain_batch", "cbo_drain_batch_flat", "cbo_drain_batch_ms"):
            getattr(L, name).argtypes = [C.POINTER(Codel), C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_size_t, C.c_int64, C.c_uint64, C.c_void_p,
                                         C.c_void_p, C.POINTER(Stats)]
        L.cbo_plan_rebalance.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32),
                                         C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
        L.cbo_build_slots.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.cbo_pref_reshuffle.argtypes = [C.c_void_p, C.c_int32, C.c_double]
        L.cbo_shuffle.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.cbo_pool_sim.argtypes = [C.POINTER(Codel), C.c_int64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(PoolResult)]
        _lib = L
    return _lib


class PoolResult(C.Structure):
    _fields_ = [("n_res", C.c_uint64), ("idle", C.c_uint64), ("cursor", C.c_uint64), ("queued", C.c_uint64),
                ("claim", C.c_uint64), ("queued_claim", C.c_uint64), ("max_claim_queue", C.c_uint64),
                ("claim_timeout", C.c_uint64)]

    def counters(self):
        return {"claim": self.claim, "queued-claim": self.queued_claim, "max-claim-queue": self.max_claim_queue,
                "claim-timeout": self.claim_timeout}


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def new_codel(target, state=None):
    cd = Codel()
    lib().cbo_codel_init(C.byref(cd), float(target))
    if state:
        cd.load(state)
    return cd


def ns_to_ms(t):
    return lib().cbo_ns_to_ms(int(t))


def drain(cd, start, now, flags=None, nslots=0, cursor=0, variant="flat", want_out=True):
    """Run the drain loop; returns (decision u8[n], slot i32[n], Stats)."""
    L = lib()
    ms = start.dtype == np.float64
    start = np.ascontiguousarray(start)
    now = np.ascontiguousarray(now)
    n = start.shape[0]
    if flags is not None:
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
    dec = np.empty(n, np.uint8) if want_out else None
    slot = np.empty(n, np.int32) if want_out else None
    st = Stats()
    fn = L.cbo_drain_batch_ms if ms else {"flat": L.cbo_drain_batch_flat, "list": L.cbo_drain_batch}[variant]
    rc = fn(C.byref(cd), _p(start), _p(now), _p(flags), n, int(nslots), int(cursor), _p(dec), _p(slot), C.byref(st))
    if rc != 0:
        raise RuntimeError("oracle drain failed")
    return dec, slot, st


def max_idle(cd, now_ms):
    return lib().cbo_codel_get_max_idle(C.byref(cd), float(now_ms))


def plan_rebalance(have, dead, target, mx, singleton=False):
    """Returns (want, add, remove[(key, ordinal)]) or raises on assertion failure."""
    have = np.ascontiguousarray(have, dtype=np.int32)
    dead = np.ascontiguousarray(dead, dtype=np.uint8)
    B = have.shape[0]
    want = np.zeros(max(B, 1), np.int32)
    add = np.zeros(max(int(mx), int(have.sum()), 1) + 1, np.int32)
    rk = np.zeros(max(int(have.sum()), 1), np.int32)
    ro = np.zeros(max(int(have.sum()), 1), np.int32)
    na, nr = C.c_int32(0), C.c_int32(0)
    rc = lib().cbo_plan_rebalance(_p(have), _p(dead), B, int(target), int(mx), int(bool(singleton)),
                                  _p(want), _p(add), C.byref(na), _p(rk), _p(ro), C.byref(nr))
    if rc != 0:
        raise AssertionError("planRebalance: target must be >= 0 and max >= target")
    return want[:B].copy(), add[:na.value].copy(), list(zip(rk[:nr.value].tolist(), ro[:nr.value].tolist()))


def build_slots(B, S, draws):
    draws = np.ascontiguousarray(draws, dtype=np.float64)
    keys = np.zeros(max(B, 1), np.int32)
    slots = np.zeros(max(S, 1), np.int32)
    n = lib().cbo_build_slots(int(B), int(S), _p(draws), _p(keys), _p(slots))
    if n < 0:
        raise RuntimeError("build_slots failed")
    return keys[:B].copy(), slots[:n].copy()


def reshuffle(keys, u):
    keys = np.ascontiguousarray(keys, dtype=np.int32).copy()
    lib().cbo_pref_reshuffle(_p(keys), keys.shape[0], float(u))
    return keys


def shuffle(arr, draws):
    arr = np.ascontiguousarray(arr, dtype=np.int32).copy()
    draws = np.ascontiguousarray(draws, dtype=np.float64)
    lib().cbo_shuffle(_p(arr), arr.shape[0], _p(draws))
    return arr


def pool_sim(cd, nslots, ev_type, ev_a, ev_b, idle0=0, cursor=0):
    """The event-driven pool model (claim / slots idle / cancel); returns (res[k,3] int64 rows of
    (ticket, code, slot), timeouts_ms per claim, PoolResult).  cd is advanced in place."""
    ev_type = np.ascontiguousarray(ev_type, dtype=np.uint8)
    ev_a = np.ascontiguousarray(ev_a, dtype=np.int64)
    ev_b = np.ascontiguousarray(ev_b, dtype=np.int64)
    nclaims = int((ev_type == 0).sum())
    res = np.zeros(6 * max(nclaims, 1), np.int64)
    tmo = np.zeros(max(nclaims, 1), np.float64)
    out = PoolResult()
    rc = lib().cbo_pool_sim(C.byref(cd), int(nslots), int(idle0), int(cursor), _p(ev_type), _p(ev_a), _p(ev_b),
                            ev_type.shape[0], _p(res), _p(tmo), C.byref(out))
    if rc != 0:
        raise RuntimeError("oracle pool_sim failed")
    return res[:3 * out.n_res].reshape(-1, 3).copy(), tmo[:nclaims].copy(), out


def lpf_taps():
    """genTaps(128, -0.2) of lib/pool.js:64-76."""
    t = np.zeros(128, np.float64)
    lib().cbo_lpf_taps(_p(t))
    return t


def lpf_run(spares, mx, ty, a, b, c, d):
    """The pool's sampler / _rebalance target events (see cueball_oracle.h); returns rows of
    (target, clamped, lpf.get()) for the ty-1 events."""
    ty = np.ascontiguousarray(ty, dtype=np.uint8)
    arrs = [np.ascontiguousarray(x, dtype=np.int64) for x in (a, b, c, d)]
    out = np.zeros(3 * max(int((ty == 1).sum()), 1), np.float64)
    L = lib()
    L.cbo_lpf_run.restype = C.c_size_t
    k = L.cbo_lpf_run(C.c_int64(int(spares)), C.c_int64(int(mx)), _p(ty), _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), _p(arrs[3]),
                      C.c_size_t(ty.shape[0]), _p(out))
    return out[:3 * k].reshape(-1, 3).copy()
