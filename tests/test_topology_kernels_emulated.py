"""The topology kernels (csrc/cb_topology.cuh: preference list insert / remove / reshuffle,
planRebalance, slot table) compiled for the HOST under tests/emu and checked against the reference's
own planRebalance known-answer tests (tests/golden/plan_rebalance_kat.json, recorded from
test/utils.test.js under the reference engine), the randomized reference plans, and the oracle's
preference-list operations."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import util
from cueball_b200 import tracegen
from oracle import oracle as O

EMU = os.path.join(util.ROOT, "tests", "emu")
SO = os.path.join(EMU, "libtopology_emu.so")
PLAN = util.load_golden("plan_rebalance_kat.json")
VP = C.c_void_p


def load_emu():
    srcs = [os.path.join(EMU, "topology_emu.cpp"), os.path.join(EMU, "emu_runtime.cpp"), os.path.join(EMU, "cuda_emu.h"),
            os.path.join(util.ROOT, "node-cueball_b200", "csrc", "cb_topology.cuh"),
            os.path.join(util.ROOT, "node-cueball_b200", "csrc", "cb_device.cuh"),
            os.path.join(util.ROOT, "node-cueball_b200", "csrc", "cb_resolver.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-rdynamic", "-fPIC", "-shared", "-pthread",
                               "-ffp-contract=off", "-I", os.path.join(EMU, "stub"), "-o", SO, srcs[0], srcs[1]])
    L = C.CDLL(SO)
    L.emu_pref_build.argtypes = [VP, C.c_int, VP]
    L.emu_pref_insert.argtypes = [VP, VP, C.c_int32, C.c_double]
    L.emu_pref_remove.argtypes = [VP, VP, C.c_int32]
    L.emu_pref_remove.restype = C.c_int
    L.emu_pref_reshuffle.argtypes = [VP, VP, C.c_double]
    L.emu_plan_rebalance.argtypes = [VP, VP, C.c_int, C.c_int, C.c_int, C.c_int, VP, VP, VP, VP, VP, VP]
    L.emu_slot_table.argtypes = [VP, VP, VP, VP, C.c_int, VP]
    L.emu_shuffle.argtypes = [VP, VP, VP, C.c_int, C.c_int]
    L.emu_srv_keys.argtypes = [VP, VP, C.c_int, VP]
    L.emu_weighted_want.argtypes = [VP, C.c_int, VP, VP, C.c_int, VP, VP, VP]
    return L


@pytest.fixture(scope="module")
def emu():
    return load_emu()


def plan_emulated(L, c):
    names = [k for k, _ in c["conns"]]
    have = np.array([len(v) for _, v in c["conns"]], np.int32)
    dead = np.array([1 if k in c["dead"] else 0 for k in names], np.uint8)
    B = len(names)
    want = np.zeros(max(B, 1), np.int32)
    add = np.zeros(max(int(c["max"]), int(have.sum()), 1) + 1, np.int32)
    rk = np.zeros(max(int(have.sum()), 1), np.int32)
    ro = np.zeros(max(int(have.sum()), 1), np.int32)
    na, nr = np.zeros(1, np.int32), np.zeros(1, np.int32)
    L.emu_plan_rebalance(have.ctypes.data, dead.ctypes.data, B, int(c["target"]), int(c["max"]), int(bool(c["singleton"])),
                         want.ctypes.data, add.ctypes.data, na.ctypes.data, rk.ctypes.data, ro.ctypes.data, nr.ctypes.data)
    conns = dict((k, v) for k, v in c["conns"])
    return ([names[i] for i in add[:na[0]]], [conns[names[k]][o] for k, o in zip(rk[:nr[0]], ro[:nr[0]])])


@pytest.mark.parametrize("c", PLAN["kat"], ids=[c["test"] for c in PLAN["kat"]])
def test_plan_rebalance_kernel_on_the_reference_kats(emu, c):  # test/utils.test.js:13-285
    add, rem = plan_emulated(emu, c)
    assert add == c["add"] and rem == c["remove"]


def test_plan_rebalance_kernel_on_random_reference_plans(emu):
    for c in PLAN["random"][::5]:
        add, rem = plan_emulated(emu, c)
        assert add == c["add"] and rem == c["remove"], c


def test_preference_list_and_slot_table_kernels(emu):
    B, S = 150, 420  # several 64-thread chunks per shift
    us = np.ascontiguousarray(tracegen.gen_insert_draws(B, 7), np.float64)
    keys = np.zeros(B + 8, np.int32)
    emu.emu_pref_build(keys.ctypes.data, B, us.ctypes.data)
    okeys, oslots = O.build_slots(B, S, us)
    assert keys[:B].tolist() == okeys.tolist()
    # slot table from the oracle's plan.add (planRebalance({}, {}, S, S) over the preference list)
    have = np.zeros(B, np.int32)
    dead = np.zeros(B, np.uint8)
    want = np.zeros(B, np.int32)
    add = np.zeros(S + 1, np.int32)
    na, nr = np.zeros(1, np.int32), np.zeros(1, np.int32)
    rk, ro = np.zeros(1, np.int32), np.zeros(1, np.int32)
    emu.emu_plan_rebalance(have.ctypes.data, dead.ctypes.data, B, S, S, 0, want.ctypes.data, add.ctypes.data,
                           na.ctypes.data, rk.ctypes.data, ro.ctypes.data, nr.ctypes.data)
    table = np.full(S + 12, -5, np.int32)
    ns = np.zeros(1, np.int32)
    emu.emu_slot_table(keys.ctypes.data, add.ctypes.data, na.ctypes.data, table.ctypes.data, S + 12, ns.ctypes.data)
    assert ns[0] == S and table[:S].tolist() == oslots.tolist() and (table[S:] == -1).all()
    # single events: removed (present and absent), added, reshuffle -- against a Python list
    ref = keys[:B].tolist()
    n = np.array([B], np.int32)
    rng = np.random.RandomState(3)
    for step in range(30):
        r = rng.rand()
        if r < 0.4 and len(ref) > 1:
            k = int(rng.choice(ref))
            assert emu.emu_pref_remove(keys.ctypes.data, n.ctypes.data, k) == 0
            ref.remove(k)
        elif r < 0.5:
            assert emu.emu_pref_remove(keys.ctypes.data, n.ctypes.data, 10_000 + step) == -1
        elif r < 0.8 and len(ref) < B:
            k, u = 20_000 + step, float(rng.rand())
            emu.emu_pref_insert(keys.ctypes.data, n.ctypes.data, k, u)
            ref.insert(int(np.floor(u * (len(ref) + 1))), k)            # lib/pool.js:287-288
        else:
            u = float(rng.rand())
            emu.emu_pref_reshuffle(keys.ctypes.data, n.ctypes.data, u)  # lib/pool.js:501-519
            taken = ref.pop()
            ref.insert(int(np.floor(u * (len(ref) + 1))), taken)
        assert n[0] == len(ref) and keys[:n[0]].tolist() == ref


def shuffle_emulated(L, arr, us, ctas):
    arr = np.ascontiguousarray(arr, np.int32)
    us = np.ascontiguousarray(us, np.float64)
    out = np.full(arr.shape[0], -7, np.int32)
    L.emu_shuffle(arr.ctypes.data, out.ctypes.data, us.ctypes.data, arr.shape[0], ctas)
    return out


def test_shuffle_kernel_golden_and_random(emu):
    """utils.shuffle (lib/utils.js:207-217): the per-position walk of shuffle_kernel against the results recorded
    from the reference engine, then against the oracle's loop at sizes around the CTA width (64 here) and the
    2048-entry staging tile, with one CTA and with several."""
    slots = util.load_golden("slots_golden.json")
    for c in slots["shuffle"]:
        us = tracegen.gen_insert_draws(c["n"], c["seed"])
        assert shuffle_emulated(emu, np.arange(c["n"]), us, 1).tolist() == c["result"]
    rng = np.random.RandomState(77)
    for n in (1, 2, 3, 63, 64, 65, 127, 300, 2047, 2048, 2049, 2113, 4500):
        us = rng.rand(n)
        if n > 4:
            us[rng.randint(n)] = 0.0                      # j = 0
            us[rng.randint(n)] = np.nextafter(1.0, 0.0)   # j = i (a transposition with itself)
        arr = rng.randint(-1000, 1000, size=n).astype(np.int32)
        want = O.shuffle(arr, us)
        for ctas in (1, 3):
            assert np.array_equal(shuffle_emulated(emu, arr, us, ctas), want), (n, ctas)


def test_srv_key_kernel_against_hashlib(emu):
    """csrc/cb_resolver.cuh: base64(sha1(message)) per thread -- message lengths across every padding boundary
    (55/56, 63/64, 119/120 bytes), the srvKey strings of lib/resolver.js:1157-1171 among them."""
    import base64
    import hashlib
    rng = np.random.RandomState(17)
    msgs = [b"", b"a", b"svc.example.com||8080||10.0.0.1", b"_http._tcp.foo||80||2001:db8::1"]
    msgs += [bytes(rng.randint(0, 256, n).astype(np.uint8)) for n in list(range(50, 70)) + list(range(115, 125)) + [200, 300]]
    msgs += [("backend%d.dc.example.com||%d||10.%d.%d.%d" % (i, 1000 + i, i % 256, (7 * i) % 256, (13 * i) % 256)).encode()
             for i in range(120)]   # more than one CTA
    blob = np.frombuffer(b"".join(msgs) + b"\0" * 16, dtype=np.uint8).copy()
    offs = np.zeros(len(msgs) + 1, np.uint32)
    offs[1:] = np.cumsum([len(m) for m in msgs])
    keys = np.zeros(len(msgs) * 28, np.uint8)
    emu.emu_srv_keys(blob.ctypes.data, offs.ctypes.data, len(msgs), keys.ctypes.data)
    got = [keys[28 * k:28 * k + 28].tobytes().decode() for k in range(len(msgs))]
    assert got == [base64.b64encode(hashlib.sha1(m).digest()).decode() for m in msgs]


def weighted_spec(keys, prio, weight, S):
    """The documented rule of the CB_BALANCE_SRV_WEIGHTED extension (no reference behaviour exists),
    stated independently of the kernel: only the backends of the numerically lowest priority take slots; S slots are
    apportioned by weight with the largest-remainder method, ties to the earlier position in the preference list; a tier
    whose weights are all zero is uniform."""
    best = min(prio[k] for k in keys)
    tier = [i for i, k in enumerate(keys) if prio[k] == best]
    wsum = sum(weight[keys[i]] for i in tier)
    w = {i: (weight[keys[i]] if wsum > 0 else 1) for i in tier}
    tot = wsum if wsum > 0 else len(tier)
    want = [0] * len(keys)
    for i in tier:
        want[i] = S * w[i] // tot
    left = S - sum(want)
    order = sorted(tier, key=lambda i: (-(S * w[i] % tot), i))
    for i in order[:left]:
        want[i] += 1
    return want


def test_weighted_want_kernel_against_its_specification(emu):
    rng = np.random.RandomState(23)
    for trial in range(60):
        B = int(rng.randint(1, 40))
        S = int(rng.randint(0, 300))
        keys = rng.permutation(B).astype(np.int32)
        prio = rng.randint(0, 3, B).astype(np.int32)
        weight = (rng.randint(0, 20, B) * (rng.rand(B) < (0.0 if trial % 7 == 0 else 0.8))).astype(np.int32)
        want = np.zeros(B, np.int32)
        add = np.zeros(S + 1, np.int32)
        na = np.zeros(1, np.int32)
        emu.emu_weighted_want(keys.ctypes.data, B, prio.ctypes.data, weight.ctypes.data, S, want.ctypes.data,
                              add.ctypes.data, na.ctypes.data)
        spec = weighted_spec(keys.tolist(), prio.tolist(), weight.tolist(), S)
        assert want.tolist() == spec, (trial, B, S)
        assert na[0] == S == sum(spec)
        assert add[:S].tolist() == [i for i, c in enumerate(spec) for _ in range(c)]
