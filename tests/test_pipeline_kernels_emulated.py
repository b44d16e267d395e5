"""The batch pipeline's CUDA kernels (scan -> prepare -> commit -> emit, csrc/cb_scan.cuh,
cb_resolve.cuh, cb_emit.cuh) compiled for the HOST under the execution-model emulation of tests/emu
and compared with the oracle: decisions, slots and the carried ControlledDelay state, batch after
batch.  The same kernels run on the GPU in tests/test_gpu_parity.py; this file keeps their control
flow and arithmetic under test on machines without one (and lets a kernel change be checked before
GPU time is spent on it).  Batches without waiting-flags only."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import util
from cueball_b200 import tracegen
from oracle import oracle as O

EMU = os.path.join(util.ROOT, "tests", "emu")
SO = os.path.join(EMU, "libpipeline_emu.so")
FULL = os.environ.get("CB_EMU_FULL") == "1"


def load_emu():
    """Build (if stale) and load the emulated pipeline library."""
    csrc = os.path.join(util.ROOT, "node-cueball_b200", "csrc")
    srcs = [os.path.join(EMU, "pipeline_emu.cpp"), os.path.join(EMU, "emu_runtime.cpp"), os.path.join(EMU, "cuda_emu.h")] + \
           [os.path.join(csrc, f) for f in ("cb_device.cuh", "cb_scan.cuh", "cb_resolve.cuh", "cb_emit.cuh", "cb_peers.cuh")]
    # CB_EMU_DEFS="-DCB_SCAN_WARPS=1 -DCB_SCAN_NST=2" (with CB_EMU_SCAN_GRID=1) makes one warp walk many tiles
    # through a short stage ring: the fuzz runs use it to exercise the scan kernel's streaming loop
    defs = os.environ.get("CB_EMU_DEFS", "").split()
    so = SO if not defs else SO.replace(".so", "_" + "".join(c for c in "".join(defs) if c.isalnum()) + ".so")
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-rdynamic", "-fPIC", "-shared", "-pthread",
                               "-ffp-contract=off"] + defs + ["-I", os.path.join(EMU, "stub"), "-o", so, srcs[0], srcs[1]])
    L = C.CDLL(so)
    L.emu_pipe_create.restype = C.c_void_p
    L.emu_pipe_create.argtypes = [C.c_double, C.c_int]
    L.emu_pipe_destroy.argtypes = [C.c_void_p]
    L.emu_pipe_ring.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.emu_pipe_set_state.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_uint64, C.c_int, C.c_int, C.c_double, C.c_uint64]
    L.emu_pipe_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 9
    L.emu_pipe_admit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p,
                                 C.c_void_p, C.c_void_p]
    L.emu_pipe_admit_ms.argtypes = L.emu_pipe_admit.argtypes
    return L


@pytest.fixture(scope="module")
def emu():
    return load_emu()


def aligned(arr):
    """a 16-byte aligned copy (the kernels take their vector path only for aligned inputs)"""
    raw = np.empty(arr.nbytes + 16, np.uint8)
    off = (-raw.ctypes.data) % 16
    out = raw[off:off + arr.nbytes].view(arr.dtype)
    out[:] = arr
    return out


def get_state(L, p):
    F, D, E = C.c_double(), C.c_double(), C.c_double()
    Cn, cur, td, ta = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    B, hasE = C.c_int(), C.c_int()
    L.emu_pipe_get_state(p, C.byref(F), C.byref(D), C.byref(Cn), C.byref(B), C.byref(hasE), C.byref(E), C.byref(cur),
                         C.byref(td), C.byref(ta))
    return ({"count": Cn.value, "drop_next": D.value, "first_above_time": F.value, "dropping": bool(B.value),
             "last_empty": E.value if hasE.value else None}, cur.value, td.value, ta.value)


def run_batches(L, start, now, cuts, target, nslots, state=None, unaligned=False, fchain=0, flags=None):
    p = L.emu_pipe_create(float(target), int(fchain))
    if state:
        le = state.get("last_empty")
        L.emu_pipe_set_state(p, state["first_above_time"], state["drop_next"], int(state["count"]),
                             int(bool(state["dropping"])), 0 if le is None else 1, 0.0 if le is None else le, 0)
    n = start.shape[0]
    dec = aligned(np.full(n + 16, 255, np.uint8))
    slot = aligned(np.full(n + 16, -7, np.int32))
    paths = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        s = aligned(start[lo:hi]) if not unaligned else start[lo:hi].copy()
        w = aligned(now[lo:hi]) if not unaligned else now[lo:hi].copy()
        d = aligned(np.zeros(hi - lo, np.uint8))
        sl = aligned(np.zeros(hi - lo, np.int32))
        info = (C.c_uint64 * 4)()
        f = np.ascontiguousarray(flags[lo:hi]) if flags is not None else None
        admit = L.emu_pipe_admit_ms if start.dtype == np.float64 else L.emu_pipe_admit
        admit(p, s.ctypes.data, w.ctypes.data, f.ctypes.data if f is not None else None, hi - lo, nslots,
              d.ctypes.data, sl.ctypes.data, info)
        dec[lo:hi] = d
        slot[lo:hi] = sl
        paths.append(int(info[0]))
    st = get_state(L, p)
    L.emu_pipe_destroy(p)
    return dec[:n].copy(), slot[:n].copy(), st, paths


def check(L, start, now, cuts, target, nslots, state=None, expect_parallel=True, fchain=0, flags=None):
    cd = O.new_codel(target, state)
    od, os_, ost = O.drain(cd, start, now, flags, nslots=nslots)
    dec, slot, (st, cursor, td, ta), paths = run_batches(L, start, now, cuts, target, nslots, state, fchain=fchain,
                                                         flags=flags)
    assert np.array_equal(dec, od)
    assert np.array_equal(slot, os_)
    util.assert_state_equal(st, cd.as_dict())
    assert (cursor, td, ta) == (ost.cursor, ost.drops, ost.admits)
    if expect_parallel:
        assert all(p == 0 for p in paths), paths
    return od


# one resolve kernel since round 2 (prepare + commit in one launch, F hand-over); the `fchain` argument
# of the helpers is ignored by the emulation and kept only for the fuzz scripts' call sites
FCHAIN = pytest.mark.parametrize("fchain", [0], ids=["resolve"])


@FCHAIN
def test_small_trace_in_ragged_batches(emu, fchain):
    n = 30_000 if FULL else 9_000
    start, now = tracegen.gen_trace(n, seed=5, de=8_000, unit=1 << 10)
    od = check(emu, start, now, [0, 1, 2_049, 4_096, n], 5, 97, fchain=fchain)
    assert 0 < (od == 1).sum() < n  # drops and admits both occur


@FCHAIN
def test_batch_starting_inside_a_dropping_episode(emu, fchain):
    n = 7_000
    start, now = tracegen.gen_trace(n, seed=9, de=20_000, unit=1 << 11, start_overloaded=True)
    state = {"count": 9, "drop_next": float(now[0]) / 1e6 - 20.0, "first_above_time": float(now[0]) / 1e6 - 300.0,
             "dropping": True, "last_empty": None}
    check(emu, start, now, [0, 3_000, n], 5, 13, state, fchain=fchain)


@FCHAIN
def test_batches_cut_inside_long_overload_runs(emu, fchain):
    """Batch boundaries inside above-target runs: inherited stretches with both class crossings,
    batches that are one above-target run end to end (first_above_time passes through), episodes
    carried across several batches."""
    n = 12_288
    start, now = tracegen.gen_trace(n, seed=21, de=30_000, unit=1 << 10, all_overloaded_from=1_500)
    cuts = [0, 1_000, 2_048, 2_300, 4_096, 6_200, 8_192, 10_000, n]
    od = check(emu, start, now, cuts, 5, 29, fchain=fchain)
    assert (od == 1).sum() > 1_000


@FCHAIN
def test_waiting_flags_and_empty_markers(emu, fchain):
    """Cancelled handles (skipped after CoDel has advanced) and codel.empty() markers, including
    markers on the last claim of a batch (the next batch starts with first_above_time reset)."""
    n = 10_240
    start, now = tracegen.gen_trace(n, seed=33, de=60_000, unit=1 << 10, all_overloaded_from=2_500)
    flags = util.make_flags(n, 33, cancel_mod=5, empty_mod=3_001)
    cuts = [0, 700, 2_048, 4_500, 8_192, n]
    flags[cuts[2] - 1] |= 2  # markers exactly at batch boundaries
    flags[cuts[4] - 1] |= 2
    od = check(emu, start, now, cuts, 5, 23, fchain=fchain, flags=flags)
    assert (od == 2).sum() > 1_000 and (od == 1).sum() > 100


@FCHAIN
def test_millisecond_clock_entry_points(emu, fchain):
    """The _ms variants of the kernels (clocks already in double milliseconds, as the reference holds
    them): same decisions as the ns variants on the converted trace."""
    n = 8_200
    start, now = tracegen.gen_trace(n, seed=71, de=50_000, unit=1 << 10, all_overloaded_from=3_000)
    to_ms = np.vectorize(O.ns_to_ms, otypes=[np.float64])
    sm, nm = to_ms(start), to_ms(now)
    flags = util.make_flags(n, 71, cancel_mod=7, empty_mod=2_003)
    for fl in (None, flags):
        cd = O.new_codel(5)
        od, os_, ost = O.drain(cd, sm, nm, fl, nslots=41)
        dec, slot, (st, cursor, td, ta), paths = run_batches(emu, sm, nm, [0, 2_048, 5_000, n], 5, 41, fchain=fchain, flags=fl)
        assert paths == [0, 0, 0]
        assert np.array_equal(dec, od) and np.array_equal(slot, os_)
        util.assert_state_equal(st, cd.as_dict())
        assert (cursor, td) == (ost.cursor, ost.drops)


@FCHAIN
def test_non_monotone_clock_takes_the_sequential_path(emu, fchain):
    n = 5_000
    start, now = tracegen.gen_trace(n, seed=11, de=10_000, unit=1 << 10)
    now = now.copy()
    now[2_500:] -= 150_000_000  # the clock steps back 150 ms inside the batch
    start = np.minimum(start, now)
    cd = O.new_codel(5)
    od, os_, ost = O.drain(cd, start, now, None, nslots=31)
    cuts = [0, 1_000, n, ]
    dec, slot, (st, cursor, td, ta), paths = run_batches(emu, start, now, cuts, 5, 31, fchain=fchain)
    assert paths == [0, 1]
    assert np.array_equal(dec, od) and np.array_equal(slot, os_)
    util.assert_state_equal(st, cd.as_dict())


@FCHAIN
@pytest.mark.parametrize("R", [2, 3])
def test_one_queue_over_a_ring_of_contexts(emu, fchain, R):
    """cb_ring_connect's hand-over (state and cursor through the next context's mailbox; with the
    F-chain variant also first_above_time, one slot per step) across R emulated contexts: segment
    (k, r) follows (k, r-1); the result equals ONE sequential drain."""
    tgt, nslots, steps = 5, 19, 4
    rng = np.random.RandomState(40 + R)
    seg = [[int(x) for x in rng.randint(300, 2_600, size=R)] for _ in range(steps)]
    n = sum(sum(x) for x in seg)
    start, now = tracegen.gen_trace(n, seed=61 + R, de=40_000, unit=1 << 10, all_overloaded_from=1_200)
    pipes = [emu.emu_pipe_create(float(tgt), int(fchain)) for _ in range(R)]
    for r in range(R):
        emu.emu_pipe_ring(pipes[r], pipes[(r + 1) % R], R, r)
    dec = np.zeros(n, np.uint8)
    slot = np.zeros(n, np.int32)
    lo = 0
    for k in range(steps):
        for r in range(R):
            m = seg[k][r]
            s, w = aligned(start[lo:lo + m]), aligned(now[lo:lo + m])
            d, sl = aligned(np.zeros(m, np.uint8)), aligned(np.zeros(m, np.int32))
            info = (C.c_uint64 * 4)()
            emu.emu_pipe_admit(pipes[r], s.ctypes.data, w.ctypes.data, None, m, nslots, d.ctypes.data, sl.ctypes.data, info)
            assert info[0] == 0
            dec[lo:lo + m] = d
            slot[lo:lo + m] = sl
            lo += m
    st, cursor, td, ta = get_state(emu, pipes[R - 1])  # the last rank's state is the queue's
    for p in pipes:
        emu.emu_pipe_destroy(p)
    cd = O.new_codel(tgt)
    od, os_, ost = O.drain(cd, start, now, None, nslots=nslots)
    assert np.array_equal(dec, od) and np.array_equal(slot, os_)
    util.assert_state_equal(st, cd.as_dict())
    assert (cursor, td, ta) == (ost.cursor, ost.drops, ost.admits)
    assert (od == 1).sum() > 500


def test_a_few_fuzz_cases(emu):
    """Twenty cases of tests/fuzz_emulated_pipeline.py (odd sizes, ragged cuts, carried-in states, flags,
    both commit variants); the full fuzz is run by hand (`python tests/fuzz_emulated_pipeline.py 600`)."""
    out = subprocess.run([sys.executable, os.path.join(util.ROOT, "tests", "fuzz_emulated_pipeline.py"), "0", "100000", "20"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "iterations 20 bad 0" in out.stdout, out.stdout[-2000:]
