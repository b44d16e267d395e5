"""Context: one cb_ctx (one pool's claim queue on one B200) with numpy / torch helpers.

Host-side mirror of the state a reference ConnectionPool keeps for this path
(lib/pool.js:191-200: p_waiters, p_codel, p_keys): the arithmetic all happens in the CUDA
library behind include/cueball_b200.h.
"""
import ctypes as C

import numpy as np

from . import _native as N


def _np_ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context(object):
    def __init__(self, device=0):
        self._L = N.lib()
        self._h = C.c_void_p()
        rc = self._L.cb_create(C.byref(self._h), int(device))
        if rc != N.CB_OK:
            msg = self._L.cb_last_error(None)
            raise N.CueballError(rc, msg.decode() if msg else "cb_create failed")
        self.device = int(device)

    _n_backends_known = 0  # backends in the context's resolver map (sizes cb_resolver_process's 'removed' buffers)

    def close(self):
        if getattr(self, "_h", None):
            self._L.cb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        return N.check(self._h, rc)

    # ------------------------------------------------------------------ CoDel
    def codel_init(self, target_ms):
        self._ck(self._L.cb_codel_init(self._h, float(target_ms)))

    def get_state(self):
        st = N.CodelState()
        self._ck(self._L.cb_codel_get_state(self._h, C.byref(st)))
        return st

    def set_state(self, d):
        """d: dict(count, drop_next, first_above_time, dropping, last_empty|None)"""
        st = N.CodelState()
        st.count = int(d["count"])
        st.drop_next = float(d["drop_next"])
        st.first_above_time = float(d["first_above_time"])
        st.dropping = 1 if d["dropping"] else 0
        le = d.get("last_empty")
        st.has_last_empty = 0 if le is None else 1
        st.last_empty = 0.0 if le is None else float(le)
        self._ck(self._L.cb_codel_set_state(self._h, C.byref(st)))

    def codel_empty(self, now_ns):
        self._ck(self._L.cb_codel_empty(self._h, int(now_ns)))

    def max_idle(self, now_ns):
        out = C.c_double()
        self._ck(self._L.cb_codel_max_idle(self._h, int(now_ns), C.byref(out)))
        return out.value

    def overloaded(self, start_ns, now_ns):
        out = C.c_int()
        self._ck(self._L.cb_codel_overloaded(self._h, int(start_ns), int(now_ns), C.byref(out)))
        return bool(out.value)

    def overloaded_ms(self, start_ms, now_ms):
        """ControlledDelay#overloaded with the reference's own double-millisecond clocks."""
        out = C.c_int()
        self._ck(self._L.cb_codel_overloaded_ms(self._h, float(start_ms), float(now_ms), C.byref(out)))
        return bool(out.value)

    # ---------------------------------------------------------------- batches
    def admit_batch(self, start, now, flags=None, want_backend=False, want_slot=True, want_decision=True,
                    out_decision=None, out_slot=None, out_backend=None):
        """Host (numpy) buffers through the host-buffer C entry point.
        start/now: int64 ns or float64 ms.  Returns (decision, slot, backend, BatchStats).
        out_* let the caller supply (e.g. pinned) output arrays."""
        start = np.ascontiguousarray(start)
        now = np.ascontiguousarray(now)
        if start.dtype != now.dtype or start.shape != now.shape or start.ndim != 1:
            raise ValueError("start/now must be 1-D arrays of the same dtype and length")
        n = start.shape[0]
        if start.dtype == np.int64:
            fn = self._L.cb_admit_batch
        elif start.dtype == np.float64:
            fn = self._L.cb_admit_batch_ms
        else:
            raise TypeError("timestamps must be int64 (ns) or float64 (ms)")
        if flags is not None:
            flags = np.ascontiguousarray(flags, dtype=np.uint8)
            if flags.shape != start.shape:
                raise ValueError("flags length mismatch")
        dec = out_decision if out_decision is not None else (np.empty(n, np.uint8) if want_decision else None)
        slot = out_slot if out_slot is not None else (np.empty(n, np.int32) if want_slot else None)
        be = out_backend if out_backend is not None else (np.empty(n, np.int32) if want_backend else None)
        for a, dt in ((dec, np.uint8), (slot, np.int32), (be, np.int32)):
            if a is not None and (a.dtype != dt or a.shape[0] < n or not a.flags["C_CONTIGUOUS"]):
                raise ValueError("bad output array")
        st = N.BatchStats()
        self._ck(fn(self._h, _np_ptr(start), _np_ptr(now), _np_ptr(flags), n, _np_ptr(dec), _np_ptr(slot),
                    _np_ptr(be), C.byref(st)))
        return dec, slot, be, st

    def admit_batch_dev(self, start, now, flags=None, decision=None, slot=None, backend=None):
        """torch CUDA tensors (queue resident in HBM); asynchronous on the context's stream."""
        n = start.numel()
        import torch
        if start.dtype == torch.int64:
            fn = self._L.cb_admit_batch_dev
        elif start.dtype == torch.float64:
            fn = self._L.cb_admit_batch_ms_dev
        else:
            raise TypeError("timestamps must be int64 (ns) or float64 (ms)")

        def p(t):
            return None if t is None else C.c_void_p(t.data_ptr())
        self._ck(fn(self._h, p(start), p(now), p(flags), n, p(decision), p(slot), p(backend)))

    def admit_batch_ptr(self, start, now, flags, n, decision, slot, backend, ns=True):
        """Same call with raw device addresses (ints, 0/None for absent arrays): what a native
        caller of the C ABI does, without the per-call tensor bookkeeping of admit_batch_dev."""
        fn = self._L.cb_admit_batch_dev if ns else self._L.cb_admit_batch_ms_dev
        rc = fn(self._h, start, now, flags or None, n, decision or None, slot or None, backend or None)
        if rc:
            self._ck(rc)

    def stats(self):
        st = N.BatchStats()
        self._ck(self._L.cb_batch_stats_get(self._h, C.byref(st)))
        return st

    def set_stream(self, cuda_stream_handle):
        self._ck(self._L.cb_set_stream(self._h, C.c_void_p(cuda_stream_handle) if cuda_stream_handle else None))

    def synchronize(self):
        self._ck(self._L.cb_synchronize(self._h))

    def join(self):
        """The caller's stream waits for every batch enqueued so far (stream-ordered)."""
        self._ck(self._L.cb_join(self._h))

    def set_pipeline_depth(self, depth):
        self._ck(self._L.cb_set_pipeline_depth(self._h, int(depth)))

    # --------------------------------------------------------------- topology
    def backends_set(self, nbackends, insert_u, nslots, table=None, mode=N.BALANCE_UNIFORM):
        u = np.ascontiguousarray(insert_u, dtype=np.float64)
        tbl = None
        if table is not None:
            arr = (N.Backend * nbackends)()
            for k, (pr, w) in enumerate(table):
                arr[k].priority, arr[k].weight = int(pr), int(w)
            tbl = C.cast(arr, C.c_void_p)
        self._ck(self._L.cb_backends_set(self._h, tbl, int(nbackends), _np_ptr(u), int(nslots), int(mode)))

    def backend_added(self, key, u, priority=0, weight=0):
        be = N.Backend(int(priority), int(weight))
        self._ck(self._L.cb_backend_added(self._h, int(key), C.cast(C.pointer(be), C.c_void_p), float(u)))

    def backend_removed(self, key):
        self._ck(self._L.cb_backend_removed(self._h, int(key)))

    def reshuffle(self, u):
        self._ck(self._L.cb_reshuffle(self._h, float(u)))

    def shuffle(self, arr, us):
        """utils.shuffle (lib/utils.js:207-217) of an int32 array with injected draws; returns the permuted copy."""
        a = np.array(arr, dtype=np.int32)
        us = np.ascontiguousarray(us, dtype=np.float64)
        if us.shape[0] < a.shape[0]:
            raise ValueError("one draw per element is required")
        self._ck(self._L.cb_shuffle(self._h, _np_ptr(a), a.shape[0], _np_ptr(us)))
        return a

    def backends_get(self):
        nk, ns = C.c_int32(), C.c_int32()
        self._ck(self._L.cb_backends_get(self._h, C.byref(nk), None, C.byref(ns), None))
        keys = np.zeros(max(nk.value, 1), np.int32)
        slots = np.zeros(max(ns.value, 1), np.int32)
        self._ck(self._L.cb_backends_get(self._h, C.byref(nk), _np_ptr(keys), C.byref(ns), _np_ptr(slots)))
        return keys[:nk.value].copy(), slots[:ns.value].copy()

    def plan_rebalance(self, have, dead, target, mx, singleton=False):
        have = np.ascontiguousarray(have, dtype=np.int32)
        dead = np.ascontiguousarray(dead, dtype=np.uint8)
        B = have.shape[0]
        want = np.zeros(max(B, 1), np.int32)
        add = np.zeros(max(int(mx), 0) + 1, np.int32)
        tot = max(int(have.sum()) if B else 0, 0) + 1
        rk, ro = np.zeros(tot, np.int32), np.zeros(tot, np.int32)
        na, nr = C.c_int32(), C.c_int32()
        self._ck(self._L.cb_plan_rebalance(self._h, _np_ptr(have), _np_ptr(dead), B, int(target), int(mx),
                                           int(bool(singleton)), _np_ptr(want), _np_ptr(add), C.byref(na),
                                           _np_ptr(rk), _np_ptr(ro), C.byref(nr)))
        return want[:B].copy(), add[:na.value].copy(), list(zip(rk[:nr.value].tolist(), ro[:nr.value].tolist()))

    # -------------------------------------------------------------- multi-GPU
    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(128)
        sz = C.c_size_t(128)
        rc = N.lib().cb_comm_unique_id(buf, C.byref(sz))
        if rc != N.CB_OK:
            raise N.CueballError(rc, (N.lib().cb_last_error(None) or b"?").decode())
        return buf.raw[:sz.value]

    def comm_init(self, nranks, rank, uid):
        self._ck(self._L.cb_comm_init(self._h, int(nranks), int(rank), C.c_char_p(uid), len(uid)))

    def allreduce_counts(self):
        self._ck(self._L.cb_allreduce_counts(self._h))

    # ------------------------------------------------------------ measurement
    def profile_enable(self, on=True):
        # True / 1: CUDA events around each kernel; 2: launch ONLY the scan kernel (the caller times a burst of
        # back-to-back launches with its own events on the context's stream); False / 0: off
        self._ck(self._L.cb_profile_enable(self._h, int(on)))

    def profile_get(self):
        ms = (C.c_double * N.K_COUNT)()
        cnt = (C.c_uint64 * N.K_COUNT)()
        self._ck(self._L.cb_profile_get(self._h, ms, cnt))
        return list(ms), list(cnt)

    def debug_timestamps(self):
        out = (C.c_uint64 * 16)()
        self._ck(self._L.cb_debug_timestamps(self._h, out))
        return list(out)

    # ------------------------------------------------ pending-claim queue in HBM
    def queue_create(self, capacity, result_capacity=None, idle_slots=0):
        self._ck(self._L.cb_queue_create(self._h, int(capacity), int(result_capacity or 2 * capacity), int(idle_slots)))

    def queue_push(self, start_ns, want_timeouts=False):
        """n x claim(): returns (first_ticket, timeouts_ms or None)."""
        start = np.ascontiguousarray(start_ns, dtype=np.int64)
        first = C.c_int64()
        tmo = np.empty(start.shape[0], np.float64) if want_timeouts else None
        self._ck(self._L.cb_queue_push(self._h, _np_ptr(start), start.shape[0], C.byref(first), _np_ptr(tmo)))
        return first.value, tmo

    def queue_cancel(self, tickets):
        t = np.ascontiguousarray(tickets, dtype=np.int64)
        self._ck(self._L.cb_queue_cancel(self._h, _np_ptr(t), t.shape[0]))

    def queue_expire(self, now_ns):
        self._ck(self._L.cb_queue_expire(self._h, int(now_ns)))

    def queue_idle(self, now_ns, k_slots):
        self._ck(self._L.cb_queue_idle(self._h, int(now_ns), int(k_slots)))

    def queue_trim_idle(self, k):
        out = C.c_uint64()
        self._ck(self._L.cb_queue_trim_idle(self._h, int(k), C.byref(out)))
        return int(out.value)

    def queue_results(self):
        """Resolution records since the last call: int64 array of (ticket, code, slot) rows."""
        info = self.queue_info()
        buf = np.empty((max(int(info.n_results), 1), 3), np.int64)
        n = C.c_size_t()
        self._ck(self._L.cb_queue_results(self._h, _np_ptr(buf), buf.shape[0], C.byref(n)))
        return buf[:n.value].copy()

    def queue_info(self):
        out = N.QueueInfo()
        self._ck(self._L.cb_queue_info_get(self._h, C.byref(out)))
        return out

    # ------------------------------------------------ one queue over several GPUs
    def ring_mailbox(self):
        """(ipc_handle: bytes, local_ptr: int) of this context's mailbox."""
        buf = C.create_string_buffer(64)
        ln = C.c_size_t(64)
        ptr = C.c_void_p()
        self._ck(self._L.cb_ring_mailbox(self._h, buf, C.byref(ln), C.byref(ptr)))
        return buf.raw[:ln.value], int(ptr.value)

    def ring_connect(self, nranks, rank, next_handle=None, next_ptr=None):
        hb = C.create_string_buffer(next_handle, len(next_handle)) if next_handle else None
        self._ck(self._L.cb_ring_connect(self._h, int(nranks), int(rank), hb,
                                         C.c_void_p(next_ptr) if next_ptr else None))

    # ------------------------------------------------ resolver ingestion
    def srv_keys(self, backends):
        """srvKey() of a list of {name, port, address} (lib/resolver.js:1157-1171), SHA-1 on the device."""
        n = len(backends)
        if n == 0:
            return []
        names = (C.c_char_p * n)(*[str(b["name"]).encode() for b in backends])
        addrs = (C.c_char_p * n)(*[str(b["address"]).encode() for b in backends])
        ports = (C.c_int32 * n)(*[int(b["port"]) for b in backends])
        out = C.create_string_buffer(n * N.SRVKEY_LEN)
        self._ck(self._L.cb_srv_keys(self._h, names, ports, addrs, n, out))
        raw = out.raw
        return [raw[k * N.SRVKEY_LEN:(k + 1) * N.SRVKEY_LEN].decode() for k in range(n)]

    def resolver_process(self, keys):
        """state_process (lib/resolver.js:1024-1108) on the context's backend map.  Returns
        (removed [(key, id)], added [(key, id, input row)], empty_set)."""
        n = len(keys)
        blob = "".join(keys).encode()
        assert len(blob) == n * N.SRVKEY_LEN
        cap = n + self._n_backends_known + 1
        rid = (C.c_int32 * cap)()
        rkeys = C.create_string_buffer(cap * N.SRVKEY_LEN)
        aid, arow = (C.c_int32 * (n + 1))(), (C.c_int32 * (n + 1))()
        nr, na, empty = C.c_int32(), C.c_int32(), C.c_int32()
        self._ck(self._L.cb_resolver_process(self._h, blob, n, rid, rkeys, C.byref(nr), aid, arow, C.byref(na), C.byref(empty)))
        removed = [(rkeys.raw[j * N.SRVKEY_LEN:(j + 1) * N.SRVKEY_LEN].decode(), rid[j]) for j in range(nr.value)]
        added = [(keys[arow[j]], aid[j], arow[j]) for j in range(na.value)]
        self._n_backends_known += na.value - nr.value
        return removed, added, bool(empty.value)

    # ------------------------------------------------ per-batch global counts over peer memory
    def peers_board(self):
        """(ipc_handle: bytes, local_ptr: int) of this context's count board."""
        buf = C.create_string_buffer(64)
        ln = C.c_size_t(64)
        ptr = C.c_void_p()
        self._ck(self._L.cb_peers_board(self._h, buf, C.byref(ln), C.byref(ptr)))
        return buf.raw[:ln.value], int(ptr.value)

    def peers_connect(self, nranks, rank, handles=None, ptrs=None):
        """handles: list of nranks 64-byte IPC handles (other processes); ptrs: list of nranks device
        pointers (contexts of this process).  nranks == 0 disconnects."""
        hb = None
        if handles:
            raw = b"".join((h if h is not None else b"\0" * 64).ljust(64, b"\0") for h in handles)
            hb = C.create_string_buffer(raw, len(raw))
        pa = None
        if ptrs:
            pa = (C.c_void_p * len(ptrs))(*[C.c_void_p(p) if p else None for p in ptrs])
        self._ck(self._L.cb_peers_connect(self._h, int(nranks), int(rank), hb, pa))

    def peers_counts(self):
        out = N.GlobalCounts()
        self._ck(self._L.cb_peers_counts(self._h, C.byref(out)))
        return out

    def debug_lane_timestamps(self, lane):
        out = (C.c_uint64 * 16)()
        self._ck(self._L.cb_debug_lane_timestamps(self._h, int(lane), out))
        return list(out)

    def launch_count(self):
        return int(self._L.cb_launch_count(self._h))
