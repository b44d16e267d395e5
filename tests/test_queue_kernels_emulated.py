"""The CUDA kernels of the device-resident pending-claim queue (node-cueball_b200/csrc/cb_queue.cuh),
compiled for the HOST under the execution-model emulation of tests/emu (OS threads for CUDA threads,
barriers for __syncthreads, slot exchange for warp shuffles) and run against the pinned pool-model
oracle.  This checks the kernels' control flow and arithmetic -- the part that can be checked
without a GPU; the hardware run itself is tests/test_gpu_queue_unvalidated.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import util
from oracle import oracle as O

EMU = os.path.join(util.ROOT, "tests", "emu")
SO = os.path.join(EMU, "libqueue_emu.so")
G = util.load_golden("pool_sim_golden.json")


class QueueDev(C.Structure):
    _fields_ = [("head", C.c_uint64), ("tail", C.c_uint64), ("next_ticket", C.c_uint64), ("idle", C.c_uint64),
                ("claim", C.c_uint64), ("queued_claim", C.c_uint64), ("max_claim_queue", C.c_uint64),
                ("claim_timeout", C.c_uint64), ("n_res", C.c_uint64), ("err", C.c_uint32), ("pad", C.c_uint32)]


def load_emu():
    """Build (if stale) and load the emulated queue library."""
    srcs = [os.path.join(EMU, "queue_emu.cpp"), os.path.join(EMU, "emu_runtime.cpp"), os.path.join(EMU, "cuda_emu.h"),
            os.path.join(util.ROOT, "node-cueball_b200", "csrc", "cb_queue.cuh"),
            os.path.join(util.ROOT, "node-cueball_b200", "csrc", "cb_device.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-pthread", "-ffp-contract=off",
                               "-I", os.path.join(EMU, "stub"), "-o", SO, srcs[0], srcs[1]])
    L = C.CDLL(SO)
    L.emu_queue_create.restype = C.c_void_p
    L.emu_queue_create.argtypes = [C.c_double, C.c_longlong, C.c_uint64, C.c_uint64, C.c_uint64]
    L.emu_queue_destroy.argtypes = [C.c_void_p]
    L.emu_queue_set_state.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_uint64, C.c_int, C.c_int, C.c_double]
    L.emu_queue_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 7
    L.emu_queue_push.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    L.emu_queue_cancel.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.emu_queue_expire.argtypes = [C.c_void_p, C.c_longlong]
    L.emu_queue_idle.argtypes = [C.c_void_p, C.c_longlong, C.c_uint64]
    L.emu_queue_results.restype = C.c_uint64
    L.emu_queue_results.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.emu_queue_info.argtypes = [C.c_void_p, C.POINTER(QueueDev)]
    return L


@pytest.fixture(scope="module")
def emu():
    return load_emu()


def run_emulated(L, target, nslots, ty, a, b, idle0, state=None, batch_pushes=False):
    q = L.emu_queue_create(float(target), nslots, 1 << 13, 1 << 15, idle0)
    if state:
        le = state.get("last_empty")
        L.emu_queue_set_state(q, state["first_above_time"], state["drop_next"], int(state["count"]),
                              int(bool(state["dropping"])), 0 if le is None else 1, 0.0 if le is None else le)
    out = []
    buf = np.zeros((1 << 15, 3), np.int64)

    def flush():
        n = L.emu_queue_results(q, buf.ctypes.data, buf.shape[0])
        out.append(buf[:n].copy())

    i = 0
    n_ev = len(ty)
    while i < n_ev:
        if ty[i] == 0:
            j = i + 1
            if batch_pushes:
                while j < n_ev and ty[j] == 0:
                    j += 1
            starts = np.array(a[i:j], dtype=np.int64)
            L.emu_queue_push(q, starts.ctypes.data, starts.shape[0], None)
            i = j
        elif ty[i] == 1:
            L.emu_queue_idle(q, int(a[i]), int(b[i]))
            flush()
            i += 1
        else:
            tk = np.array([a[i]], dtype=np.int64)
            L.emu_queue_cancel(q, tk.ctypes.data, 1)
            i += 1
    flush()
    info = QueueDev()
    L.emu_queue_info(q, C.byref(info))
    F, D, E = C.c_double(), C.c_double(), C.c_double()
    Cn, cur = C.c_uint64(), C.c_uint64()
    B, hasE = C.c_int(), C.c_int()
    L.emu_queue_get_state(q, C.byref(F), C.byref(D), C.byref(Cn), C.byref(B), C.byref(hasE), C.byref(E), C.byref(cur))
    L.emu_queue_destroy(q)
    st = {"count": Cn.value, "drop_next": D.value, "first_above_time": F.value, "dropping": bool(B.value),
          "last_empty": E.value if hasE.value else None}
    return np.concatenate(out) if out else np.zeros((0, 3), np.int64), info, st, cur.value


def check(L, target, nslots, ty, a, b, idle0, state=None, batch_pushes=False):
    cd = O.new_codel(target, state)
    want, _, o = O.pool_sim(cd, nslots, ty, a, b, idle0=idle0)
    got, info, st, cursor = run_emulated(L, target, nslots, ty, a, b, idle0, state, batch_pushes)
    assert info.err == 0
    assert got.tolist() == want.tolist()
    assert {"claim": info.claim, "queued-claim": info.queued_claim, "max-claim-queue": info.max_claim_queue,
            "claim-timeout": info.claim_timeout} == o.counters()
    assert (info.idle, info.tail - info.head, cursor) == (o.idle, o.queued, o.cursor)
    util.assert_state_equal(st, cd.as_dict())


FULL = os.environ.get("CB_EMU_FULL") == "1"
# An emulated launch costs ~20 ms (64 OS threads, condition-variable barriers), so the default run
# replays the first events of a few scripts; CB_EMU_FULL=1 replays everything (12 golden scripts, 6
# random ones: ~16 minutes; all pass).
QUICK_CASES = ["hard_timeouts_t5", "deep_backlog_t5_jitter", "idle_slots_first_t5"]
CASES = G["cases"] if FULL else [c for c in G["cases"] if c["name"] in QUICK_CASES]
HEAD = None if FULL else 450


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_emulated_kernels_on_golden_scripts(emu, case):
    ty, a, b = case["ev_type"][:HEAD], case["ev_a"][:HEAD], case["ev_b"][:HEAD]
    check(emu, case["target"], case["nslots"], ty, a, b, case["idle0"])


@pytest.mark.parametrize("seed", range(6) if FULL else [3])
def test_emulated_kernels_on_random_scripts(emu, seed):
    rng = np.random.RandomState(8100 + seed)
    tgt = int(rng.choice([1, 2, 5, 10, 20]))
    ty, a, b = util.gen_pool_events(8200 + seed, 1500 if FULL else 400, tgt, float(rng.choice([2.0, 5.0, 9.0, 12.0])),
                                    float(rng.choice([0.0, 0.1, 0.5])))
    check(emu, tgt, int(rng.choice([1, 7, 64])), ty, a, b, int(rng.randint(0, 4)))


def test_emulated_kernels_mass_drop_events(emu):
    """Hundreds to thousands of handles resolved by one event, starting inside a dropping episode;
    the arrivals go in as two batched pushes (no timer is due while they arrive)."""
    ms = 1_000_000
    t0 = 50 * 10 ** 9
    ty, a, b = [], [], []
    for i in range(2250):
        ty.append(0); a.append(t0 + i * 20_000); b.append(0)  # noqa: E702
    ty.append(1); a.append(t0 + 45 * ms); b.append(50)        # noqa: E702
    ty.append(2); a.append(2100); b.append(0)                 # noqa: E702
    for i in range(2250, 3000):
        ty.append(0); a.append(t0 + i * 20_000); b.append(0)  # noqa: E702
    ty.append(1); a.append(t0 + 61 * ms); b.append(700)       # noqa: E702
    ty.append(1); a.append(t0 + 400 * ms); b.append(5000)     # noqa: E702
    state = {"count": 7, "drop_next": (t0 + 10 * ms) / 1e6, "first_above_time": (t0 - 200 * ms) / 1e6,
             "dropping": True, "last_empty": None}
    check(emu, 5, 13, ty, a, b, 0, state, batch_pushes=True)
