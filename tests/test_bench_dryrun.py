"""Dry run of bench.py's CUDA arm with torch / torch.distributed / the native context replaced by
stand-ins: the control flow, the timed passes (N = 1 and N > 1, every collective mode) and the JSON
line are exercised on CPU, so that a slip in the script shows up here and not on the GPU box at
round end.  Nothing is measured."""
import argparse
import contextlib
import json
import os
import sys
import types

import numpy as np
import pytest

import util  # noqa: F401  (puts the repo root on sys.path)

bench = None  # imported by the `fakes` fixture: importing bench.py sets CUDA / NCCL environment defaults, which
              # must not leak into a pytest process that goes on to run the GPU tests


class FakeTensor(object):
    def __init__(self, arr):
        self.a = np.asarray(arr)

    def to(self, dev):
        return self

    def data_ptr(self):
        return 0x7f0000000000

    def __getitem__(self, k):
        return FakeTensor(self.a[k])

    def pin_memory(self):
        return self

    def numpy(self):
        return self.a

    def item(self):
        return self.a.reshape(-1)[0].item()

    def tolist(self):
        return self.a.tolist()


class FakeEvent(object):
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass

    def elapsed_time(self, other):
        return 2.0


def make_fake_torch(log):
    t = types.ModuleType("torch")
    t.uint8, t.int32, t.int64, t.float64 = np.uint8, np.int32, np.int64, np.float64
    t.device = lambda kind, idx=0: "%s:%d" % (kind, idx)
    t.from_numpy = lambda a: FakeTensor(a)
    t.empty = lambda n, dtype=None, device=None: FakeTensor(np.zeros(int(n), dtype))
    t.tensor = lambda v, dtype=None, device=None: FakeTensor(np.array(v, dtype=dtype))
    cuda = types.ModuleType("torch.cuda")
    cuda.set_device = lambda i: None
    cuda.synchronize = lambda: None
    cuda.Event = FakeEvent
    cuda.Stream = lambda device=None: types.SimpleNamespace(cuda_stream=0x1234)
    cuda.stream = lambda s: contextlib.nullcontext()
    t.cuda = cuda
    dist = types.ModuleType("torch.distributed")
    dist.ReduceOp = types.SimpleNamespace(MAX="max", SUM="sum")
    dist.init_process_group = lambda backend, device_id=None: log.append("init")
    dist.barrier = lambda: None
    dist.all_reduce = lambda tensor, op=None: log.append("all_reduce")
    dist.broadcast_object_list = lambda obj, src=0: None
    dist.destroy_process_group = lambda: log.append("destroy")
    t.distributed = dist
    return t, dist


class FakeContext(object):
    launches = 0

    def __init__(self, dev):
        self.log = []

    @staticmethod
    def comm_unique_id():
        return b"x" * 128

    def codel_init(self, t): pass                                   # noqa: E704
    def backends_set(self, B, u, S): pass                           # noqa: E704
    def set_stream(self, s): pass                                   # noqa: E704
    def comm_init(self, w, r, i): pass                              # noqa: E704
    def allreduce_counts(self): self.log.append("cb_allreduce")     # noqa: E704
    def join(self): pass                                            # noqa: E704
    def synchronize(self): pass                                     # noqa: E704
    def close(self): pass                                           # noqa: E704
    def set_state(self, st): pass                                   # noqa: E704
    def profile_enable(self, on): pass                              # noqa: E704

    def admit_batch_ptr(self, s, nw, f, n, d, sl, b):
        FakeContext.launches += 4

    def admit_batch_dev(self, *a):
        FakeContext.launches += 4

    def admit_batch(self, hs, hn, out_decision=None, out_slot=None):
        return out_decision, out_slot, None, types.SimpleNamespace(drops=1)

    def launch_count(self):
        return FakeContext.launches

    def stats(self):
        return types.SimpleNamespace(total_drops=10, total_admits=30, global_drops=10)

    def profile_get(self):
        return [0.1, 0.05, 0.02], [10, 10, 10]


@pytest.fixture
def fakes(monkeypatch):
    global bench
    for var in ("CUDA_DEVICE_MAX_CONNECTIONS", "NCCL_NTHREADS", "NCCL_MAX_CTAS"):
        monkeypatch.setenv(var, os.environ.get(var, ""))  # restored (or removed again) after the test
        if not os.environ[var]:
            monkeypatch.delenv(var)
    import bench as _bench
    bench = _bench
    log = []
    ft, fdist = make_fake_torch(log)
    fcb = types.ModuleType("cueball_b200")
    fcb.Context = FakeContext
    fcb._native = types.SimpleNamespace(K_SCAN=0)
    monkeypatch.setitem(sys.modules, "torch", ft)
    monkeypatch.setitem(sys.modules, "torch.distributed", fdist)
    monkeypatch.setitem(sys.modules, "cueball_b200", fcb)
    monkeypatch.setattr(bench, "WORKLOADS", {k: dict(v, n=4096) for k, v in bench.WORKLOADS.items()})
    monkeypatch.setattr(bench, "cpu_baseline", lambda wl: {"value": 1.0, "unit": "claims/s", "cores": 1, "kind": "port", "sample": "x"})
    monkeypatch.setattr(bench, "bind_to_gpu_numa", lambda i: 4)
    monkeypatch.setattr(bench, "ClockSampler", lambda i: types.SimpleNamespace(start=lambda: None, stop=lambda: {"sm_mhz": 1}))
    FakeContext.launches = 0
    return log


REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "cpu_baseline")


def test_single_gpu_line(fakes, monkeypatch):
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    line = bench.run_cuda(argparse.Namespace(gpus=1, steps=6, warmup=3), "c2")
    json.dumps(line)
    for k in REQUIRED:
        assert k in line, k
    assert line["gpu_launches"] == 4 * 6 and line["n_gpus"] == 1
    assert line["config"]["collective"] is None and "per_step_collective" not in line
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])


@pytest.mark.parametrize("mode", ["window", "step", "none"])
def test_multi_gpu_line(fakes, monkeypatch, mode):
    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("LOCAL_RANK", "0")
    monkeypatch.setenv("CB_BENCH_COLLECTIVE", mode)
    line = bench.run_cuda(argparse.Namespace(gpus=2, steps=5, warmup=3), "c4")
    json.dumps(line)
    assert line["n_gpus"] == 2 and line["value"] > 0
    if mode == "window":
        assert line["global_counts"] == {"drops": 10, "admits": 30, "claims": 40}
        assert line["per_step_collective"]["value"] > 0
    else:
        assert "per_step_collective" not in line
    assert (line["config"]["collective"] is None) == (mode == "none")
    assert "destroy" in fakes


def test_non_zero_rank_prints_nothing(fakes, monkeypatch):
    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setenv("LOCAL_RANK", "1")
    assert bench.run_cuda(argparse.Namespace(gpus=2, steps=4, warmup=3), "c4") is None


def test_main_default_invocation_prints_one_json_line(fakes, monkeypatch, capsys):
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    monkeypatch.delenv("RANK", raising=False)
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    bench.main()
    out = [ln for ln in capsys.readouterr().out.splitlines() if ln.strip()]
    assert len(out) == 1
    line = json.loads(out[0])
    assert line["config"]["workload"] == "c2" and line["steps"] == 200 and line["warmup"] >= 3
    assert line["c3"]["value"] > 0 and "roofline" in line["c3"]


def test_main_driver_style_flags(fakes, monkeypatch, capsys):
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", "1", "--steps", "7", "--warmup", "1"])
    bench.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["steps"] == 7 and line["warmup"] == 3  # W >= 3 is enforced
