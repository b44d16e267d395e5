"""EmuContext -- TEST INFRASTRUCTURE: the slice of cueball_b200.Context that the host mirror
`ConnectionPool` (node-cueball_b200/pool.py) calls, served by the library's own kernels compiled for the
HOST (tests/emu: csrc/cb_queue.cuh and csrc/cb_topology.cuh under the execution-model emulation) instead of
by the GPU.  It lets the `-m "not gpu"` suite replay the pinned pool-model scripts through the REAL pool.py
(handles, callbacks, getStats, the kang snapshot) in this container; on the GPU box the same scripts go
through the real C ABI (tests/test_gpu_pool_mirror.py).  Nothing in the product imports this module, and the
product's Context has no such mode: without the CUDA library and a device cb_create fails.

The host-side glue mirrored here is what csrc/cueball_b200.cu does around the kernels:
rebuild_slots() = plan_rebalance_kernel(have = 0, dead = 0, target = max = the wanted slots) +
slot_table_kernel; cb_queue_create rounds the capacity up to a power of two >= 1024; cb_queue_push reads
next_ticket before the kernel; cb_queue_trim_idle is a bookkeeping word."""
import ctypes as C
import types

import numpy as np

import test_pipeline_kernels_emulated as TP
import test_queue_kernels_emulated as TQ
import test_topology_kernels_emulated as TT


class EmuContext(object):
    KEY_CAP = 256

    def __init__(self):
        self.Q = TQ.load_emu()
        self.Q.emu_queue_set_nslots.argtypes = [C.c_void_p, C.c_longlong]
        self.Q.emu_queue_trim_idle.argtypes = [C.c_void_p, C.c_uint64]
        self.Q.emu_queue_trim_idle.restype = C.c_uint64
        self.T = TT.load_emu()
        self.target = None
        self.q = None
        self.pipe = None
        self.keys = np.zeros(self.KEY_CAP, np.int32)
        self.nkeys = np.zeros(1, np.int32)
        self.want_slots = 0
        self.slot_backend = np.zeros(0, np.int32)
        self.launches = 0

    # ------------------------------------------------------------------ CoDel
    def codel_init(self, target_ms):
        from cueball_b200.utils import assertClaimDelay
        assertClaimDelay(target_ms)  # cb_codel_init returns CB_EINVAL for what utils.assertClaimDelay rejects
        self.target = float(target_ms)
        if getattr(self, "pipe", None):
            self.P.emu_pipe_destroy(self.pipe)
        self.pipe = None

    def _pipe(self):
        """The scalar and batch entry points run on the emulated batch pipeline (tests/emu/pipeline_emu.cpp).  In the
        library they share ONE DevState with the queue; the two emulation libraries each keep their own, so one
        emulated context serves either the queue events or the scalar/batch calls, not a mix."""
        assert self.q is None, "this emulated context already serves the queue (separate state)"
        if self.pipe is None:
            self.P = TP.load_emu()
            self.P.emu_pipe_overloaded.argtypes = [C.c_void_p, C.c_longlong, C.c_longlong]
            self.P.emu_pipe_overloaded_ms.argtypes = [C.c_void_p, C.c_double, C.c_double]
            self.P.emu_pipe_empty.argtypes = [C.c_void_p, C.c_longlong]
            self.P.emu_pipe_max_idle.argtypes = [C.c_void_p, C.c_longlong]
            self.P.emu_pipe_max_idle.restype = C.c_double
            self.pipe = self.P.emu_pipe_create(self.target, 0)
        return self.pipe

    def overloaded(self, start_ns, now_ns):
        p = self._pipe()
        self.launches += 1
        return bool(self.P.emu_pipe_overloaded(p, int(start_ns), int(now_ns)))

    def overloaded_ms(self, start_ms, now_ms):
        p = self._pipe()
        self.launches += 1
        return bool(self.P.emu_pipe_overloaded_ms(p, float(start_ms), float(now_ms)))

    def codel_empty(self, now_ns):
        p = self._pipe()
        self.launches += 1
        self.P.emu_pipe_empty(p, int(now_ns))

    def max_idle(self, now_ns):
        p = self._pipe()
        self.launches += 1
        return float(self.P.emu_pipe_max_idle(p, int(now_ns)))

    def get_state(self):
        from cueball_b200 import _native as N
        p = self._pipe()
        d, cursor, _, _ = TP.get_state(self.P, p)
        st = N.CodelState()
        st.target_ms, st.count, st.drop_next, st.first_above_time = self.target, int(d["count"]), d["drop_next"], d["first_above_time"]
        st.dropping = 1 if d["dropping"] else 0
        st.has_last_empty = 0 if d["last_empty"] is None else 1
        st.last_empty = 0.0 if d["last_empty"] is None else d["last_empty"]
        self._cursor = cursor
        return st

    def set_state(self, d):
        le = d.get("last_empty")
        p = self._pipe()
        self.P.emu_pipe_set_state(p, float(d["first_above_time"]), float(d["drop_next"]), int(d["count"]),
                                  1 if d["dropping"] else 0, 0 if le is None else 1, 0.0 if le is None else float(le), 0)

    def admit_batch(self, start, now, flags=None, **kw):
        """cb_admit_batch / cb_admit_batch_ms on host arrays: (decision, slot, None, stats)."""
        start, now = TP.aligned(np.ascontiguousarray(start)), TP.aligned(np.ascontiguousarray(now))
        n = start.shape[0]
        dec, slot = TP.aligned(np.zeros(n, np.uint8)), TP.aligned(np.zeros(n, np.int32))
        info = (C.c_uint64 * 4)()
        f = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
        p = self._pipe()
        fn = self.P.emu_pipe_admit_ms if start.dtype == np.float64 else self.P.emu_pipe_admit
        fn(p, start.ctypes.data, now.ctypes.data, None if f is None else f.ctypes.data, n,
           len(self.slot_backend), dec.ctypes.data, slot.ctypes.data, info)
        self.launches += 3
        st = types.SimpleNamespace(n=n, drops=int((dec == 1).sum()), admits=int((dec == 0).sum()),
                                   skipped=int((dec == 2).sum()), path=int(info[0]), n_events=int(info[1]),
                                   n_intervals=int(info[2]), state=self.get_state())
        return dec.copy(), slot.copy(), None, st

    # --------------------------------------------------------------- topology
    def _rebuild(self):
        B, S = int(self.nkeys[0]), int(self.want_slots)
        if B <= 0 or S <= 0:
            self.slot_backend = np.zeros(0, np.int32)
        else:
            have, dead = np.zeros(B, np.int32), np.zeros(B, np.uint8)
            want, add = np.zeros(B, np.int32), np.zeros(S + 1, np.int32)
            na, nr = np.zeros(1, np.int32), np.zeros(1, np.int32)
            self.T.emu_plan_rebalance(have.ctypes.data, dead.ctypes.data, B, S, S, 0, want.ctypes.data, add.ctypes.data,
                                      na.ctypes.data, None, None, nr.ctypes.data)
            table = np.zeros(S + 8, np.int32)
            ns = np.zeros(1, np.int32)
            self.T.emu_slot_table(self.keys.ctypes.data, add.ctypes.data, na.ctypes.data, table.ctypes.data, S + 8, ns.ctypes.data)
            self.slot_backend = table[:int(ns[0])].copy()
            self.launches += 2
        if self.q:
            self.Q.emu_queue_set_nslots(self.q, len(self.slot_backend))

    def backends_set(self, nbackends, insert_u, nslots, table=None, mode=0):
        assert table is None and mode == 0, "the emulated context serves the uniform mode only"
        self.want_slots = int(nslots)
        self.nkeys[0] = int(nbackends)
        if nbackends:
            us = np.ascontiguousarray(insert_u, dtype=np.float64)
            self.T.emu_pref_build(self.keys.ctypes.data, int(nbackends), us.ctypes.data)
        self._rebuild()

    def backend_added(self, key, u, priority=0, weight=0):
        assert 0 <= key and self.nkeys[0] + 1 < self.KEY_CAP and 0.0 <= u < 1.0
        self.T.emu_pref_insert(self.keys.ctypes.data, self.nkeys.ctypes.data, int(key), float(u))
        self._rebuild()

    def backend_removed(self, key):
        if self.T.emu_pref_remove(self.keys.ctypes.data, self.nkeys.ctypes.data, int(key)) != 0:
            raise AssertionError("resolver key %d not found" % key)
        self._rebuild()

    def reshuffle(self, u):
        if self.nkeys[0] <= 1:
            return
        self.T.emu_pref_reshuffle(self.keys.ctypes.data, self.nkeys.ctypes.data, float(u))
        self._rebuild()

    def plan_rebalance(self, have, dead, target, mx, singleton=False):
        """cb_plan_rebalance: plan_rebalance_kernel over caller-supplied connection counts."""
        have = np.ascontiguousarray(have, dtype=np.int32)
        dead = np.ascontiguousarray(dead, dtype=np.uint8)
        B = have.shape[0]
        tot = max(int(have.sum()) if B else 0, 0) + 1
        want, add = np.zeros(max(B, 1), np.int32), np.zeros(max(int(mx), tot) + 1, np.int32)
        rk, ro = np.zeros(tot, np.int32), np.zeros(tot, np.int32)
        na, nr = np.zeros(1, np.int32), np.zeros(1, np.int32)
        self.T.emu_plan_rebalance(have.ctypes.data, dead.ctypes.data, B, int(target), int(mx), int(bool(singleton)),
                                  want.ctypes.data, add.ctypes.data, na.ctypes.data, rk.ctypes.data, ro.ctypes.data,
                                  nr.ctypes.data)
        self.launches += 1
        return want[:B].copy(), add[:na[0]].copy(), list(zip(rk[:nr[0]].tolist(), ro[:nr[0]].tolist()))

    def shuffle(self, arr, us):
        """cb_shuffle: utils.shuffle of an int32 array with injected draws (shuffle_kernel)."""
        a = np.array(arr, dtype=np.int32)
        us = np.ascontiguousarray(us, dtype=np.float64)
        if a.shape[0] < 2:
            return a
        out = np.zeros_like(a)
        self.T.emu_shuffle(a.ctypes.data, out.ctypes.data, us.ctypes.data, a.shape[0], 3)
        self.launches += 1
        return out

    def backends_get(self):
        return self.keys[:int(self.nkeys[0])].copy(), self.slot_backend.copy()

    # ------------------------------------------------------ pending-claim queue
    def queue_create(self, capacity, result_capacity=None, idle_slots=0):
        assert self.target is not None, "cb_codel_init has not been called"
        assert self.pipe is None, "this emulated context already serves the scalar/batch calls (separate state)"
        cap = 1024
        while cap < capacity:
            cap <<= 1
        if self.q:
            self.Q.emu_queue_destroy(self.q)
        self.q = self.Q.emu_queue_create(self.target, len(self.slot_backend), cap, int(result_capacity or 2 * capacity),
                                         int(idle_slots))
        self._res_cap = int(result_capacity or 2 * capacity)

    def queue_info(self):
        from cueball_b200 import _native as N
        raw = TQ.QueueDev()
        self.Q.emu_queue_info(self.q, C.byref(raw))
        assert raw.err == 0, "queue error flags %#x" % raw.err
        out = N.QueueInfo()
        C.memmove(C.byref(out), C.byref(raw), C.sizeof(out))
        return out

    def queue_push(self, start_ns, want_timeouts=False):
        start = np.ascontiguousarray(start_ns, dtype=np.int64)
        first = int(self.queue_info().next_ticket)
        tmo = np.empty(start.shape[0], np.float64) if want_timeouts else None
        self.Q.emu_queue_push(self.q, start.ctypes.data, start.shape[0], None if tmo is None else tmo.ctypes.data)
        self.launches += 1
        return first, tmo

    def queue_cancel(self, tickets):
        t = np.ascontiguousarray(tickets, dtype=np.int64)
        self.Q.emu_queue_cancel(self.q, t.ctypes.data, t.shape[0])
        self.launches += 1

    def queue_expire(self, now_ns):
        self.Q.emu_queue_expire(self.q, int(now_ns))
        self.launches += 1

    def queue_idle(self, now_ns, k_slots):
        self.Q.emu_queue_idle(self.q, int(now_ns), int(k_slots))
        self.launches += 1

    def queue_trim_idle(self, k):
        return int(self.Q.emu_queue_trim_idle(self.q, int(k)))

    def queue_results(self):
        buf = np.empty((max(self._res_cap, 1), 3), np.int64)
        n = self.Q.emu_queue_results(self.q, buf.ctypes.data, buf.shape[0])
        return buf[:n].copy()

    # ------------------------------------------------------------------ state
    def stats(self):
        F, D, E = C.c_double(), C.c_double(), C.c_double()
        Cn, cur = C.c_uint64(), C.c_uint64()
        B, hasE = C.c_int(), C.c_int()
        self.Q.emu_queue_get_state(self.q, C.byref(F), C.byref(D), C.byref(Cn), C.byref(B), C.byref(hasE), C.byref(E),
                                   C.byref(cur))
        state = types.SimpleNamespace(count=Cn.value, drop_next=D.value, first_above_time=F.value, dropping=B.value,
                                      has_last_empty=hasE.value, last_empty=E.value)
        return types.SimpleNamespace(cursor=cur.value, state=state)

    def launch_count(self):
        return self.launches

    def close(self):
        if self.q:
            self.Q.emu_queue_destroy(self.q)
            self.q = None
