"""Dry run of bench.py's CUDA arm on the CPU: real torch (CPU tensors), torch.cuda / torch.distributed
replaced by stand-ins, and the native context replaced by a stand-in that drains every batch with the
oracle -- so the control flow, the in-run parity comparison, every extra leg (N = 1 and N > 1) and the
JSON line are exercised here, and a slip in the script shows up before the GPU box runs it at round end.
Nothing is measured.  Also: the device-side trace generator of bench.py against the numpy generator."""
import contextlib
import ctypes as C
import json
import os
import sys
import types

import numpy as np
import pytest

import util  # noqa: F401  (puts the repo root on sys.path)
from cueball_b200 import tracegen
from oracle import oracle as O

bench = None  # imported by the `fakes` fixture: importing bench.py sets CUDA / NCCL environment defaults, which
              # must not leak into a pytest process that goes on to run the GPU tests


def _arr(ptr, n, ctype, dtype):
    return np.frombuffer((ctype * n).from_address(ptr), dtype=dtype)


class FakeContext(object):
    """cueball_b200.Context with the oracle inside (tests/ may use it; the product never does)."""
    launches = 0
    world = 1

    def __init__(self, dev):
        self.cd, self.cursor, self.td, self.ta, self.nslots, self.tgt = None, 0, 0, 0, 0, 5
        self.peers = 0

    @staticmethod
    def comm_unique_id():
        return b"x" * 128

    def codel_init(self, t):
        self.tgt = t
        self.cd = O.new_codel(t)

    def backends_set(self, B, u, S):
        self.nslots = S

    def set_stream(self, s): pass                                   # noqa: E704
    def comm_init(self, w, r, i): pass                              # noqa: E704
    def allreduce_counts(self): pass                                # noqa: E704
    def join(self): pass                                            # noqa: E704
    def synchronize(self): pass                                     # noqa: E704
    def close(self): pass                                           # noqa: E704
    def profile_enable(self, on): pass                              # noqa: E704
    def ring_mailbox(self): return b"h" * 64, 1                     # noqa: E704
    def ring_connect(self, n, r, next_handle=None, next_ptr=None): pass  # noqa: E704
    def peers_board(self): return b"b" * 64, 1                      # noqa: E704

    def peers_connect(self, n, r, handles=None, ptrs=None):
        self.peers = n

    def peers_counts(self):
        w = FakeContext.world
        return types.SimpleNamespace(steps=1, drops=self.td * w, admits=self.ta * w, n=(self.td + self.ta) * w)

    def set_state(self, st):
        self.cd = O.new_codel(self.tgt, st)

    def _drain(self, s, nw, dec, slot):
        od, os_, ost = O.drain(self.cd, s, nw, None, nslots=self.nslots, cursor=self.cursor)
        self.cursor = ost.cursor
        self.td += ost.drops
        self.ta += ost.admits
        if dec is not None:
            dec[:] = od
        if slot is not None:
            slot[:] = os_
        return ost

    def admit_batch_ptr(self, s, nw, f, n, d, sl, b):
        FakeContext.launches += 3
        self._drain(_arr(s, n, C.c_int64, np.int64), _arr(nw, n, C.c_int64, np.int64),
                    _arr(d, n, C.c_uint8, np.uint8), _arr(sl, n, C.c_int32, np.int32))

    def admit_batch_dev(self, s, nw, f, d, sl):
        FakeContext.launches += 3
        self._drain(s.numpy(), nw.numpy(), d.numpy(), sl.numpy())

    def admit_batch(self, hs, hn, out_decision=None, out_slot=None):
        ost = self._drain(hs, hn, out_decision, out_slot)
        return out_decision, out_slot, None, types.SimpleNamespace(drops=ost.drops)

    def launch_count(self):
        return FakeContext.launches

    def stats(self):
        d = self.cd.as_dict()
        st = types.SimpleNamespace(count=int(d["count"]), drop_next=d["drop_next"], first_above_time=d["first_above_time"],
                                   dropping=int(d["dropping"]))
        return types.SimpleNamespace(total_drops=self.td, total_admits=self.ta, cursor=self.cursor, state=st, path=0,
                                     global_drops=self.td, global_admits=self.ta, global_n=self.td + self.ta)

    def profile_get(self):
        return [0.1, 0.05, 0.02], [10, 10, 10]


class FakeEvent(object):
    def __init__(self, enable_timing=False): pass                   # noqa: E704
    def record(self, stream=None): pass                             # noqa: E704
    def elapsed_time(self, other): return 2.0                       # noqa: E704


@pytest.fixture
def fakes(monkeypatch):
    global bench
    torch = pytest.importorskip("torch")
    for var in ("CUDA_DEVICE_MAX_CONNECTIONS", "NCCL_NTHREADS", "NCCL_MAX_CTAS"):
        monkeypatch.setenv(var, os.environ.get(var, ""))  # restored (or removed again) after the test
        if not os.environ[var]:
            monkeypatch.delenv(var)
    import bench as _bench
    bench = _bench
    import torch.distributed as dist
    cpu = torch.device("cpu")
    monkeypatch.setattr(torch, "device", lambda *a, **k: cpu)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(torch.cuda, "set_device", lambda i: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "Stream", lambda device=None: types.SimpleNamespace(cuda_stream=0x1234))
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(dist, "init_process_group", lambda backend, device_id=None: None)
    monkeypatch.setattr(dist, "barrier", lambda: None)
    monkeypatch.setattr(dist, "destroy_process_group", lambda: None)
    monkeypatch.setattr(dist, "broadcast_object_list", lambda obj, src=0: None)

    def all_gather_object(out, obj):
        for i in range(len(out)):
            out[i] = obj
    monkeypatch.setattr(dist, "all_gather_object", all_gather_object)

    def all_reduce(t, op=None):  # every rank holds what this one holds
        if op is None or op == dist.ReduceOp.SUM:
            t.mul_(FakeContext.world)
    monkeypatch.setattr(dist, "all_reduce", all_reduce)
    fcb = types.ModuleType("cueball_b200")
    fcb.Context = FakeContext
    fcb._native = types.SimpleNamespace(K_SCAN=0)
    fcb.tracegen = tracegen
    monkeypatch.setitem(sys.modules, "cueball_b200", fcb)
    monkeypatch.setattr(bench, "WORKLOADS", {k: dict(v, n=8192, gen=dict(v["gen"], unit=1 << 11,
                                                                          **({"all_overloaded_from": 4096} if k == "c5" else {})))
                                             for k, v in bench.WORKLOADS.items()})
    monkeypatch.setattr(bench, "bind_to_gpu_numa", lambda i: 4)
    monkeypatch.setattr(bench, "ClockSampler", lambda i, n=1: types.SimpleNamespace(start=lambda: None, stop=lambda: {"sm_mhz": 1}))
    monkeypatch.setattr(bench, "pcie_rates", lambda dev, a, b: {"h2d_gbs": 50.0, "d2h_gbs": 50.0,
                                                                   "duplex_gbs": {"h2d": 45.0, "d2h": 20.0}})
    monkeypatch.setattr(bench, "stress_leg", lambda local, dev, stream: {"value": 1.0, "bit_exact_vs_oracle": True})
    FakeContext.launches = 0
    FakeContext.world = 1
    return types.SimpleNamespace(torch=torch)


REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "parity")


def test_single_gpu_line(fakes, monkeypatch, capsys):
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "4", "--warmup", "3"])
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    bench.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for k in REQUIRED + ("cpu_baseline", "c2", "nonmonotone_stress"):
        assert k in line, k
    assert line["config"]["workload"] == "c3" and line["n_gpus"] == 1
    assert line["parity"]["ok"] and line["c2"]["parity"]["ok"]
    assert line["roofline"]["bound"] == "hbm" and 0 < line["roofline"]["frac"]
    assert line["e2e"]["h2d_bytes_per_step"] == 16 * 8192 and line["e2e"]["frac_of_pcie"] > 0
    assert line["gpu_launches"] == 3 * 4
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 1
    assert set(line["config"]) == set(bench.workload_config("c3", 1))


def test_multi_gpu_line(fakes, monkeypatch, capsys):
    FakeContext.world = 2
    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("LOCAL_RANK", "0")
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", "2", "--steps", "3", "--warmup", "3"])
    bench.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for k in REQUIRED + ("global_counts", "nccl_allreduce", "one_queue", "c4", "c5"):
        assert k in line, k
    assert line["n_gpus"] == 2 and line["config"]["workload"] == "c3"
    assert line["config"]["claims_per_step_per_gpu"] == 8192  # the same per-GPU batch at every N
    assert line["global_counts"]["equals_sum_of_ranks"]
    assert line["parity"]["ok"] and line["c4"]["parity"]["ok"] and line["c5"]["parity"]["ok"]
    assert "per_step" in line["nccl_allreduce"] and "per_window" in line["nccl_allreduce"]


def test_other_ranks_print_nothing(fakes, monkeypatch, capsys):
    FakeContext.world = 2
    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setenv("LOCAL_RANK", "1")
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", "2", "--steps", "3", "--warmup", "3", "--no-extra"])
    bench.main()
    assert capsys.readouterr().out.strip() == ""


def test_reference_arm_shares_the_config(monkeypatch, capsys):
    import bench as b
    from oracle import ref
    monkeypatch.setattr(b, "REF_SAMPLE", {k: 4096 for k in b.REF_SAMPLE})
    monkeypatch.setattr(ref, "available", lambda: False)  # the C port stands in: quick
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--steps", "2", "--warmup", "3"])
    monkeypatch.delenv("RANK", raising=False)
    b.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["config"] == b.workload_config("c3", 1)
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["value"] == line["value"]


def test_reference_arm_runs_one_pool_per_gpu(monkeypatch, capsys):
    """--impl reference --gpus 2: the job is two independent pools, so two reference pools run at once (one core each);
    the line's config is the CUDA arm's, its value the aggregate over the slowest pool's time."""
    import bench as b
    from oracle import ref
    monkeypatch.setattr(b, "REF_SAMPLE", {k: 8192 for k in b.REF_SAMPLE})
    monkeypatch.setattr(ref, "available", lambda: False)  # the C port stands in: quick
    monkeypatch.setitem(b.WORKLOADS, "c3", dict(b.WORKLOADS["c3"], n=1 << 16))  # the full-step C port leg stays small
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "3"])
    monkeypatch.delenv("RANK", raising=False)
    b.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    cfg = b.workload_config("c3", 2)
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["config"] == cfg
    assert cfg["parallelism"] == "queue segment per GPU"
    assert line["cpu_baseline"]["cores"] == 2 and line["cpu_baseline"]["kind"] == "port"
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"] and line["gpu_launches"] == 0
    # rank 1 of the same launch prints nothing
    monkeypatch.setenv("RANK", "1")
    b.main()
    assert capsys.readouterr().out.strip() == ""


@pytest.mark.parametrize("kw", [{}, {"all_overloaded_from": 5000}])
def test_device_trace_generator_matches_numpy(kw):
    """bench.gen_trace_torch (run on the CPU here) against tracegen.gen_trace, bit for bit: whole traces and
    windows generated without their prefix (the look-back bound of the running maximum)."""
    torch = pytest.importorskip("torch")
    import bench as b
    n, unit = 700_000, 1 << 14
    s_ref, n_ref = tracegen.gen_trace(n, seed=12345, t0=b.T0_NS, de=b.DE_NS, target_ns=b.TARGET_MS * 1_000_000, unit=unit, **kw)
    cpu = torch.device("cpu")
    s, nw = b.gen_trace_torch(0, n, 12345, cpu, unit, piece=1 << 17, **kw)
    assert np.array_equal(s.numpy(), s_ref) and np.array_equal(nw.numpy(), n_ref)
    for lo, hi in [(400_000, 520_000), (1, 2), (699_000, 700_000), (310_000, 640_123)]:
        s, nw = b.gen_trace_torch(lo, hi, 12345, cpu, unit, piece=1 << 16, **kw)
        assert np.array_equal(s.numpy(), s_ref[lo:hi]) and np.array_equal(nw.numpy(), n_ref[lo:hi])
