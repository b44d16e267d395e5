#!/usr/bin/env python3
"""bench.py -- claim-queue admissions/sec (CoDel + backend-slot pick) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl reference] [--no-extra]

A "step" is one pass of the hot path (scan -> resolve -> emit kernels behind cb_admit_batch_dev)
over one batch of synthetic pending claims.  Workloads are the BASELINE.json configs
(SURVEY.md 8(d)); sizes are PER GPU and the same at every N (weak scaling):

    c3  2^24 claims / step, 1024 backends, 4096 slots    (configs[2]: the largest single-GPU config;
                                                          the default at every N)
    c2  2^20 claims / step,   64 backends,  256 slots    (configs[1]; extra key `c2` at N=1)
    c4  2^23 claims / step / GPU, 4096 backends          (configs[3]: 64M over 8 GPUs; extra key at N>1)
    c5  2^24 claims / step / GPU, sustained overload     (configs[4]; extra key at N>1)

The (W+K) batches of a run are consecutive segments of ONE seeded trace, generated on the device
before the timed region (same generator as cueball_b200/tracegen.py, checked bit for bit in
tests/), so every step reads cold data (the run's inputs are many times the 126 MB L2) and the
ControlledDelay state chains from batch to batch on the device.

`value` = claims decided per second, whole job (all ranks), inputs resident in HBM, CUDA events on
the launching stream, max over ranks.  Nothing in the timed region touches the host: no stats read,
no torch collective.  N > 1: one process per GPU, each rank drains its own queue segment (no
data-path collective); the GLOBAL drop/admit counts exist after every batch, device-side, through
the peer-memory boards of cb_peers_* (each batch stores its counts into its rank's board; an extra CTA
of its emit kernel reads the other ranks' rows over NVLink).  The north star's literal "single NCCL allreduce" is timed
separately (`nccl_allreduce`).  `one_queue` is the other multi-GPU mode: ONE logical queue whose
ControlledDelay state goes round a ring of device mailboxes, checked bit for bit against one
sequential drain on a window.
`parity` = the timed outputs against the C oracle in the same run: every decision and slot of the
W+K steps, the drop/admit totals and the final ControlledDelay state (bit patterns).
`e2e` = the same metric through the host-buffer C entry point (pinned host arrays in, decisions and
slots back on the host every step), with the measured PCIe copy rates next to it.
`roofline` is for the dominant kernel (the scan kernel: 16 algorithmic bytes per claim)
from CUDA events recorded around each kernel on the launching stream in a separate profiling pass.
`cpu_baseline` is the C restatement of the reference path (oracle/, pinned bit-for-bit to the
reference's own JS) on one host core, on the full step size.

--impl reference times the reference's OWN JavaScript (lib/queue.js + lib/codel.js verbatim) under
the SpiderMonkey engine vendored in the reference tree (oracle/_ref): one pool's drain is a sequential
state machine on one core (Node itself is single threaded); with --gpus N the job is N independent
pools, so N reference processes run at once on N cores.  Each pool drains a bounded sample of the same
workload per step (`config.reference_sample_claims_per_step`, the same key in both arms).
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import tempfile
import time

# more hardware work queues than the default 8, so that the pipeline's streams do not share one
# (must be set before CUDA initialises)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# the NCCL collective of the `nccl_allreduce` leg is a 24-byte sum: one small CTA (see cb_comm_init);
# NCCL reads these once per process, so they are set before torch initialises it
if os.environ.get("CB_BENCH_NCCL_DEFAULTS") is None:
    os.environ.setdefault("NCCL_NTHREADS", "128")
    os.environ.setdefault("NCCL_MAX_CTAS", "1")

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from cueball_b200 import tracegen  # noqa: E402  (input generator only)

WORKLOADS = {
    #        claims/step/GPU  backends  slots   generator extras
    # healthy/overloaded phases of (1..4)x / (2..9)x 0.52 s at 1 us arrivals: several dropping
    # episodes per run, healthy stretches longer than the 300 ms the backlog needs to drain
    "c2": dict(n=1 << 20, B=64, S=256, gen={"unit": 1 << 19}),
    "c3": dict(n=1 << 24, B=1024, S=4096, gen={"unit": 1 << 19}),
    "c4": dict(n=1 << 23, B=4096, S=16384, gen={"unit": 1 << 19}),
    "c5": dict(n=1 << 24, B=4096, S=16384, gen={"unit": 1 << 19, "all_overloaded_from": 1 << 20}),
}
TARGET_MS = 5
T0_NS, DE_NS = 10 ** 9, 1000
ALG_BYTES_SCAN = 16      # start_ns 8 + now_ns 8 (read once by the scan kernel)
ALG_BYTES_STEP = 21      # + decision 1 + slot 4
REF_SAMPLE = {"c2": 1 << 20, "c3": 1 << 19, "c4": 1 << 19, "c5": 1 << 19}  # claims per step of the JS reference arm
REF_QUEUE = 1 << 18      # the vendored engine's 64 MB heap holds one queue of 2^18 pending claims: a step's sample is
                         # drained as consecutive queues of this depth
METRIC = "claim-queue admissions/sec (CoDel + backend-slot pick)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_config(wl, world):
    """The part of `config` that names the workload: identical in the CUDA arm and the reference arm."""
    w = WORKLOADS[wl]
    return {"workload": wl, "claims_per_step_per_gpu": w["n"], "backends": w["B"], "slots": w["S"],
            "target_ms": TARGET_MS, "interval_ms": 100,
            "parallelism": "queue segment per GPU" if world > 1 else "1 GPU",
            "reference_sample_claims_per_step": REF_SAMPLE[wl]}


def bind_to_gpu_numa(index):
    """Pin this process to the CPUs next to GPU `index` (what numactl / a production host does):
    pinned host buffers are then allocated on the GPU's own NUMA node, and the end-to-end leg
    measures the PCIe link instead of the inter-socket fabric.  Returns the CPU count or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        cpus = sorted(set(cpus) & os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, n_gpus=1):
        self.ids = ",".join(str(gpu_index + k) for k in range(max(1, n_gpus)))  # every GPU of the job (rank 0 samples)
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", self.ids, "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
            time.sleep(0.25)  # the sampler is up before the timed region starts
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return None
        time.sleep(0.12)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = {}, [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.setdefault(c[0], []).append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return None
        med = {g: float(np.median(v)) for g, v in sm.items()}
        out = {"sm_mhz": min(med.values()), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
               "samples": sum(len(v) for v in sm.values())}
        if len(med) > 1:
            out["sm_mhz_by_gpu"] = med  # sm_mhz is the slowest GPU's median
        return out


# ------------------------------------------------------------------------------ trace on the device
def gen_trace_torch(lo, hi, seed, dev, unit, all_overloaded_from=-1, start_overloaded=False, piece=1 << 24):
    """Claims [lo, hi) of the trace tracegen.gen_trace(hi, seed=seed, t0=T0_NS, de=DE_NS, ...) would
    produce, as int64 CUDA tensors, without generating the claims before `lo`: now_i is a running
    maximum of start_j + soj_j, and since start advances DE_NS per claim while soj <= 61 * target,
    only the last 61*target/DE_NS claims can hold the maximum.  (tests/test_bench_dryrun.py compares
    this with the numpy generator bit for bit, on the CPU.)"""
    import torch
    target_ns = TARGET_MS * 1_000_000
    healthy_span, over_span = (target_ns * 9) // 10, target_ns * 60
    back = (target_ns + over_span) // DE_NS + 2
    a = max(0, lo - back)
    ph = tracegen.phases(seed, hi, unit, start_overloaded, all_overloaded_from)
    M32 = 0xFFFFFFFF
    start_out = torch.empty(hi - lo, dtype=torch.int64, device=dev)
    now_out = torch.empty(hi - lo, dtype=torch.int64, device=dev)
    carry = None
    for p0 in range(a, hi, piece):
        p1 = min(hi, p0 + piece)
        idx = torch.arange(p0, p1, dtype=torch.int64, device=dev)
        k = (((idx + 1) & M32) * 0x9E3779B9) & M32
        x = (k ^ (seed & M32)) & M32
        x = x ^ (x >> 16)
        x = (x * 0x7FEB352D) & M32
        x = x ^ (x >> 15)
        x = (x * 0x846CA68B) & M32
        h = x ^ (x >> 16)
        start = T0_NS + idx * DE_NS
        soj = h % healthy_span
        for (b, e, over) in ph:
            if over and e > p0 and b < p1:
                bb, ee = max(b, p0) - p0, min(e, p1) - p0
                soj[bb:ee] = target_ns + h[bb:ee] % over_span
        now = torch.cummax(start + soj, 0).values
        if carry is not None:
            now = torch.clamp(now, min=carry)
        carry = int(now[-1].item())
        s0 = max(lo, p0)
        if p1 > s0:
            start_out[s0 - lo:p1 - lo] = start[s0 - p0:]
            now_out[s0 - lo:p1 - lo] = now[s0 - p0:]
        del idx, k, x, h, start, soj, now
    return start_out, now_out


def rank_seed(rank, salt=0):
    return tracegen.SEED + 7919 * rank + salt


def make_trace(wl, total_claims, rank, dev, salt=0):
    """Rank `rank`'s queue segment of workload `wl`: per-rank seeded, deterministic (int64 tensors on `dev`)."""
    w = WORKLOADS[wl]
    return gen_trace_torch(0, total_claims, rank_seed(rank, salt), dev, w["gen"]["unit"], w["gen"].get("all_overloaded_from", -1))


# ------------------------------------------------------------------------------ reference arm
def run_reference(args, wl):
    """The reference's own CPU implementation of the path on this box's host cores.  One pool's drain is a
    sequential state machine on one thread (Node itself is single threaded); the --gpus N job is N independent pools
    (a queue segment per GPU), whose CPU counterpart is N reference processes, one pool each, running at once on N cores."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle as O
    from oracle import ref
    w = WORKLOADS[wl]
    n = REF_SAMPLE[wl]
    pools = max(1, args.gpus)
    params = dict(t0=T0_NS, de=DE_NS, target_ns=TARGET_MS * 1_000_000, unit=w["gen"].get("unit", 1 << 16),
                  start_overloaded=False, all_overloaded_from=w["gen"].get("all_overloaded_from", -1))
    times, decided = [], 0
    kind = "reference" if ref.available() else "port"
    parts = max(1, n // REF_QUEUE)
    m = n // parts

    def one_pool(step, pool):
        """One pool's sample of this step: `parts` consecutive queues of m pending claims; seconds in the drain loop."""
        dt = 0.0
        for part in range(parts):
            seed = tracegen.SEED + (step * pools + pool) * parts + part
            if kind == "reference":
                r = ref.bench(seed, m, params, TARGET_MS, w["S"])  # its own OS process (the engine's shell)
                dt += r["drain_ms"] / 1e3
            else:
                s, nw = tracegen.gen_trace(m, seed=seed, **{k: v for k, v in params.items()})
                cd = O.new_codel(TARGET_MS)
                t0 = time.perf_counter()
                O.drain(cd, s, nw, nslots=w["S"], variant="list", want_out=True)  # ctypes releases the GIL
                dt += time.perf_counter() - t0
        return dt

    with ThreadPoolExecutor(max_workers=pools) as ex:
        for step in range(args.warmup + args.steps):
            dts = list(ex.map(lambda p: one_pool(step, p), range(pools)))
            if step >= args.warmup:
                times.append(max(max(dts), 1e-9))  # the pools run at once: the step takes as long as the slowest
                decided += m * parts * pools
    total = sum(times)
    value = decided / total
    # context for the ratio: SpiderMonkey 1.7 is an interpreter from 2007; a JIT-compiling engine (V8 under
    # Node.js, which is not installed here) would land between this number and the compiled C restatement
    # of the same loop over the same linked-list queue, timed here on ONE FULL step of the CUDA arm's size
    port = None
    try:
        nfull = min(w["n"], 1 << 24)
        s, nw = tracegen.gen_trace(nfull, seed=tracegen.SEED, **{k: v for k, v in params.items()})
        best = None
        for _ in range(2):
            cd = O.new_codel(TARGET_MS)
            t0 = time.perf_counter()
            O.drain(cd, s, nw, nslots=w["S"], variant="list", want_out=True)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        port = nfull / best
    except Exception:
        pass
    sample = ("%d pool(s) at once, one OS process and core each; per pool and step %d claims of the %s trace spec as %d "
              "queue(s) of %d pending claims (same generator, JS side; the engine's heap holds one such queue; the CUDA "
              "arm's step is %d claims per GPU); timed: the drain loop lib/pool.js:733-753 = Queue.shift + "
              "ControlledDelay.overloaded per claim, the slowest pool of each step" % (pools, m * parts, wl, parts, m, w["n"]))
    line = {
        "impl": "reference", "metric": METRIC,
        "value": value, "unit": "claims/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / max(len(times), 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64+int64", "data": "synthetic",
        "config": workload_config(wl, args.gpus),
        "engine": ("reference lib/queue.js+lib/codel.js verbatim under SpiderMonkey 1.7 (vendored in the reference "
                   "tree; Node.js is not installed)" if kind == "reference" else "C restatement (oracle/), linked-list queue"),
        "cpu_baseline": {"value": value, "unit": "claims/s", "cores": pools, "host_cpus": os.cpu_count(), "kind": kind,
                         "sample": sample,
                         "c_restatement_linked_list_claims_per_s": port,
                         "c_restatement_sample": "one full step of the CUDA arm (%d claims), linked-list queue, 1 core" % min(w["n"], 1 << 24)},
        "e2e": {"value": value, "unit": "claims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------ CUDA arm
def cpu_baseline(wl, start, now):
    """The oracle's C restatement (flat-array variant = the fastest honest CPU form), 1 core, on one
    full step of the same workload (host arrays handed in by the caller)."""
    from oracle import oracle as O
    w = WORKLOADS[wl]
    n = start.shape[0]
    best = None
    for _ in range(3):
        cd = O.new_codel(TARGET_MS)
        t0 = time.perf_counter()
        O.drain(cd, start, now, nslots=w["S"], variant="flat", want_out=True)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    cd = O.new_codel(TARGET_MS)
    t0 = time.perf_counter()
    m = min(n, 1 << 22)
    O.drain(cd, start[:m], now[:m], nslots=w["S"], variant="list", want_out=True)
    dl = time.perf_counter() - t0
    return {"value": n / best, "unit": "claims/s", "cores": 1, "kind": "port",
            "sample": "one full step (%d claims) of the %s trace, best of 3; C restatement of lib/codel.js + lib/pool.js:733-753 "
                      "over flat arrays (oracle/cueball_oracle.c, -O2 -ffp-contract=off)" % (n, wl),
            "linked_list_variant_claims_per_s": m / dl}


class Run(object):
    """One rank's context, trace and buffers for one workload."""

    def __init__(self, wl, K, W, world, rank, local, dev, stream, seed_salt=0):
        import torch

        import cueball_b200 as cb
        self.cb, self.torch = cb, torch
        self.wl, self.K, self.W, self.world, self.rank, self.dev, self.stream = wl, K, W, world, rank, dev, stream
        w = WORKLOADS[wl]
        self.w, self.n, self.nb = w, w["n"], W + K
        self.d_start, self.d_now = make_trace(wl, self.n * self.nb, rank, dev, seed_salt)
        self.d_dec = torch.empty(self.n * self.nb, dtype=torch.uint8, device=dev)
        self.d_slot = torch.empty(self.n * self.nb, dtype=torch.int32, device=dev)
        self.ctx = cb.Context(local)
        self.ctx.codel_init(TARGET_MS)
        self.ctx.backends_set(w["B"], tracegen.gen_insert_draws(w["B"]), w["S"])
        self.ctx.set_stream(stream.cuda_stream)
        torch.cuda.synchronize()
        self.p = (self.d_start.data_ptr(), self.d_now.data_ptr(), self.d_dec.data_ptr(), self.d_slot.data_ptr())

    def step(self, b):
        # one C-ABI call per step on raw device addresses, as a native host (the N-API shim) makes it
        lo, n = b * self.n, self.n
        ps, pn, pd, psl = self.p
        self.ctx.admit_batch_ptr(ps + 8 * lo, pn + 8 * lo, None, n, pd + lo, psl + 4 * lo, None)

    def reset(self):
        self.ctx.set_state({"count": 0, "drop_next": 0.0, "first_above_time": 0.0, "dropping": False, "last_empty": None})

    def close(self):
        self.ctx.synchronize()
        self.ctx.close()
        del self.d_start, self.d_now, self.d_dec, self.d_slot
        self.torch.cuda.empty_cache()


def barrier(world):
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world, dev):
    import torch
    import torch.distributed as dist
    if world == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(xs, world, dev):
    import torch
    import torch.distributed as dist
    if world == 1:
        return [int(x) for x in xs]
    t = torch.tensor(list(xs), dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(x) for x in t.tolist()]


def _realtime(on, saved=None):
    """SCHED_FIFO for the enqueueing thread while the timed steps are submitted (best effort: needs CAP_SYS_NICE)."""
    try:
        if on:
            saved = (os.sched_getscheduler(0), os.sched_getparam(0))
            os.sched_setscheduler(0, os.SCHED_FIFO, os.sched_param(10))
            return saved
        if saved is not None:
            os.sched_setscheduler(0, saved[0], saved[1])
    except (OSError, AttributeError, PermissionError):
        pass
    return None


def timed_pass(run, per_step=None):
    """W untimed + K timed steps; (device ms of the K steps on this rank, host ms/step to enqueue,
    kernels launched by the K steps).  Nothing between the two events touches the host except the
    enqueueing itself.  per_step: optional callable after every step (e.g. an NCCL allreduce)."""
    torch = run.torch
    W, K, ctx, stream = run.W, run.K, run.ctx, run.stream
    idle = os.environ.get("CB_BENCH_IDLE_RANKS") == "1" and run.rank != 0  # diagnosis: only rank 0 works
    # The last two of the W warm-up steps run AFTER the barrier, device-ordered in front of the first event (cb_join:
    # a stream wait, no host wait): a cross-rank barrier leaves the GPU idle for up to milliseconds, and the timed
    # steps should meet the clocks and memory state of a running pipeline, not of an idle GPU.
    rewarm = 0 if os.environ.get("CB_BENCH_REWARM") == "0" else min(2, W)
    for b in range(W - rewarm):
        if idle:
            break
        run.step(b)
        if per_step:
            per_step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the K steps are ~1 ms of GPU time fed by this thread: keep the collector and the scheduler out of the way
    gc_was = gc.isenabled()
    gc.disable()
    ctx.synchronize()
    barrier(run.world)
    rt = _realtime(True)
    # Nothing slow between here and the first timed step: with the count boards connected the ranks' batches are
    # coupled on the device (a batch's counts need every rank's post), so the re-warm steps end at the same moment
    # on every rank -- the first event is ordered behind them and must already be in the stream by then, or the
    # ranks whose host was early would time the others' lateness as their own steps.
    for b in range(W - rewarm, W):
        if idle:
            break
        run.step(b)
        if per_step:
            per_step()
    ctx.join()
    e0.record(stream)
    lc0 = ctx.launch_count()
    th0 = time.perf_counter()
    for b in range(W, W + K):
        if idle:
            break
        run.step(b)
        if per_step:
            per_step()
    hms = (time.perf_counter() - th0) * 1e3 / K  # host time to enqueue one step (no waiting)
    nlaunch = ctx.launch_count() - lc0            # this library's kernels launched by the K timed steps
    ctx.join()  # the timing stream waits for the last batch of the pipeline (stream-ordered, no host wait)
    e1.record(stream)
    _realtime(False, rt)
    if gc_was:
        gc.enable()
    barrier(run.world)
    ctx.synchronize()
    return e0.elapsed_time(e1), hms, nlaunch


def parity_check(run, full_limit=1 << 29):
    """The outputs of the W+K steps just run against the C oracle (one sequential drain from the
    initial state): every decision and slot, the totals, the final ControlledDelay state."""
    from oracle import oracle as O
    st = run.ctx.stats()
    total = run.n * run.nb
    t0 = time.perf_counter()
    hs, hn = run.d_start.cpu().numpy(), run.d_now.cpu().numpy()
    cd = O.new_codel(TARGET_MS)
    od, os_, ost = O.drain(cd, hs, hn, None, nslots=run.w["S"])
    lo = 0 if total <= full_limit else total - run.n
    dec = run.d_dec[lo:].cpu().numpy()
    slot = run.d_slot[lo:].cpu().numpy()
    ok_dec = bool(np.array_equal(dec, od[lo:]))
    ok_slot = bool(np.array_equal(slot, os_[lo:]))
    ref = cd.as_dict()
    got = {"count": float(st.state.count), "drop_next": st.state.drop_next, "first_above_time": st.state.first_above_time,
           "dropping": bool(st.state.dropping)}
    ok_state = all(np.float64(got[k]).tobytes() == np.float64(ref[k]).tobytes() for k in ("count", "drop_next", "first_above_time")) \
        and got["dropping"] == ref["dropping"]
    ok_tot = (st.total_drops == ost.drops and st.total_admits == ost.admits and st.cursor == ost.cursor)
    kd = int((od[run.W * run.n:] == 1).sum())
    return {"ok": bool(ok_dec and ok_slot and ok_state and ok_tot), "decisions": ok_dec, "slots": ok_slot, "state_bits": ok_state,
            "totals": ok_tot, "claims_compared": int(total - lo), "claims_drained_by_oracle": int(total),
            "timed_steps_drops": kd, "timed_steps_admits": int(run.K * run.n - kd),
            "oracle": "oracle/cueball_oracle.c cbo_drain_batch_flat, one sequential drain of the W+K steps",
            "seconds": round(time.perf_counter() - t0, 2)}, (hs, hn)


def roofline(run, wl, ms, traffic):
    """The dominant kernel's launch duration, with CUDA events on the launching stream, same inputs as the timed
    steps.  `scan`: two events around a burst of back-to-back scan launches over min(K, 20) DIFFERENT batches (cold
    data each) divided by the launch count -- one kernel alone between two events also measures its launch
    latency (a few microseconds here).  `scan_isolated`, `resolve`, `emit`: events around each kernel of batches
    run strictly one after the other."""
    cb, torch = run.cb, run.torch
    nbk = min(run.K, 20)
    run.reset()
    run.ctx.profile_enable(1)
    for b in range(run.W, run.W + nbk):
        lo, hi = b * run.n, (b + 1) * run.n
        run.ctx.admit_batch_dev(run.d_start[lo:hi], run.d_now[lo:hi], None, run.d_dec[lo:hi], run.d_slot[lo:hi])
    kms, kcnt = run.ctx.profile_get()
    scan_iso = kms[cb._native.K_SCAN] / max(kcnt[cb._native.K_SCAN], 1)
    run.ctx.profile_enable(2)  # scan kernel only
    for b in range(run.W, run.W + 2):
        run.step(b)
    run.ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(run.stream)
    for b in range(run.W, run.W + nbk):
        run.step(b)
    e1.record(run.stream)
    torch.cuda.synchronize()
    scan_ms = e0.elapsed_time(e1) / nbk
    run.ctx.profile_enable(0)
    run.reset()
    peak, peak_src = peaks()
    achieved = ALG_BYTES_SCAN * run.n / (scan_ms / 1e3) / 1e9
    step_gbs = ALG_BYTES_STEP * run.n * run.K / (ms / 1e3) / 1e9
    return {"bound": "hbm", "kernel": "cb::scan_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic.get(wl), "peak_source": peak_src,
            "alg_bytes_per_claim": ALG_BYTES_SCAN, "alg_bytes_per_launch": ALG_BYTES_SCAN * run.n,
            "timing": "CUDA events on the launching stream around %d back-to-back scan launches over %d different batches" % (nbk, nbk),
            "kernel_ms": {"scan": scan_ms, "scan_isolated": scan_iso, "resolve": kms[1] / max(kcnt[1], 1),
                          "emit": kms[2] / max(kcnt[2], 1)},
            "step": {"alg_bytes_per_claim": ALG_BYTES_STEP, "achieved": step_gbs, "frac": step_gbs / peak}}


def pcie_rates(dev, stream_a, stream_b):
    """Measured pinned-memory copy rates of this GPU's PCIe link (GB/s): H2D alone, D2H alone, both at once."""
    import torch
    nb_in, nb_out = 256 << 20, 80 << 20
    hi, ho = torch.empty(nb_in, dtype=torch.uint8).pin_memory(), torch.empty(nb_out, dtype=torch.uint8).pin_memory()
    di, do = torch.empty(nb_in, dtype=torch.uint8, device=dev), torch.empty(nb_out, dtype=torch.uint8, device=dev)

    def t(fn, reps=4):
        best = None
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return best

    def h2d():
        with torch.cuda.stream(stream_a):
            di.copy_(hi, non_blocking=True)

    def d2h():
        with torch.cuda.stream(stream_b):
            ho.copy_(do, non_blocking=True)

    a, b = t(h2d), t(d2h)
    both = t(lambda: (h2d(), d2h()))
    return {"h2d_gbs": nb_in / a / 1e9, "d2h_gbs": nb_out / b / 1e9,
            "duplex_gbs": {"h2d": nb_in / both / 1e9, "d2h": nb_out / both / 1e9,
                           "note": "256 MiB in and 80 MiB out at once (the path's 16:5 byte ratio)"}}


def e2e_leg(run, host_trace, pcie):
    """Host buffers through the public entry point: pinned inputs in, decisions and slots back, every step."""
    torch = run.torch
    n, W = run.n, run.W
    ke = min(run.K, 10 if n > (1 << 21) else 100)  # about 50 ms of calls either way
    hs_all, hn_all = host_trace
    hs = torch.from_numpy(hs_all[W * n:(W + ke) * n]).pin_memory().numpy()
    hn = torch.from_numpy(hn_all[W * n:(W + ke) * n]).pin_memory().numpy()
    hdec = torch.empty(n, dtype=torch.uint8).pin_memory().numpy()
    hslot = torch.empty(n, dtype=torch.int32).pin_memory().numpy()
    run.ctx.set_stream(None)
    run.reset()
    for _ in range(3):  # warm the staging buffers, the copy engines and the clocks
        run.ctx.admit_batch(hs[:n], hn[:n], out_decision=hdec, out_slot=hslot)
    run.reset()
    barrier(run.world)
    t0 = time.perf_counter()
    drops = 0
    for b in range(ke):
        _, _, _, est = run.ctx.admit_batch(hs[b * n:(b + 1) * n], hn[b * n:(b + 1) * n], out_decision=hdec, out_slot=hslot)
        drops += est.drops  # the step's result is read on the host
    barrier(run.world)
    te = max_over_ranks(time.perf_counter() - t0, run.world, run.dev)
    run.ctx.set_stream(run.stream.cuda_stream)
    h2d, d2h = 16 * n, 5 * n + 176
    step_s = te / ke
    out = {"value": run.world * n * ke / te, "unit": "claims/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "steps": ke, "ms_per_step": step_s * 1e3, "host_memory": "pinned inputs and outputs", "drops_read_on_host": int(drops)}
    if pcie:
        bound_s = max(h2d / (pcie["duplex_gbs"]["h2d"] * 1e9), d2h / (pcie["duplex_gbs"]["d2h"] * 1e9))
        out["pcie"] = pcie
        out["frac_of_pcie"] = bound_s / step_s
        out["pcie_bound_ms_per_step"] = bound_s * 1e3
    return out


def load_traffic():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


def device_leg(wl, K, W, world, rank, local, dev, stream, want_parity=True, want_e2e=True, want_roof=True, pcie=None,
               peers=False, sampler=None, want_cpu=False):
    """One workload on this rank: timed pass (+ optional in-run parity, roofline, e2e).  Returns a dict (rank 0
    fills the line from it)."""
    import torch.distributed as dist
    run = Run(wl, K, W, world, rank, local, dev, stream)
    peers = peers and os.environ.get("CB_BENCH_PEERS", "1") != "0"  # measurement knob: the same pass without the boards
    if peers and world > 1:
        h, _ = run.ctx.peers_board()
        hs = [None] * world
        dist.all_gather_object(hs, h)
        run.ctx.peers_connect(world, rank, handles=hs)
    elif peers and os.environ.get("CB_BENCH_PEERS") == "self":  # one rank posting to its own board (cost of the post kernel)
        run.ctx.peers_connect(1, 0)
    if sampler:
        sampler.start()
    ms, host_ms, launches = timed_pass(run)
    clocks = sampler.stop() if sampler else None
    ms_rank = ms
    ms = max_over_ranks(ms, world, dev)
    out = {"ms": ms, "host_ms": host_ms, "launches": sum_over_ranks([launches], world, dev)[0], "clocks": clocks,
           "value": world * run.n * K / (ms / 1e3)}
    if world > 1:  # every rank's own device time of the K steps (the line's ms_per_step is the maximum)
        import torch
        t = torch.zeros(world, dtype=torch.float64, device=dev)
        t[rank] = ms_rank / K
        dist.all_reduce(t)
        out["ms_per_step_by_rank"] = [round(x, 5) for x in t.tolist()]
    st = run.ctx.stats()
    out["admit_fraction"] = st.total_admits / max(st.total_admits + st.total_drops, 1)
    if peers and world > 1 and os.environ.get("CB_PEERS_DEBUG"):
        run.ctx.peers_connect(0, 0)  # measurement variants of the post: nothing to read back
    elif peers and world > 1:
        g = run.ctx.peers_counts()  # running global totals after this rank's last batch, from its own board
        loc = sum_over_ranks([st.total_drops, st.total_admits, run.n * run.nb], world, dev)
        out["global_counts"] = {"drops": int(g.drops), "admits": int(g.admits), "claims": int(g.n), "steps": int(g.steps),
                                "equals_sum_of_ranks": [int(g.drops), int(g.admits), int(g.n)] == loc,
                                "how": "cb_peers_*: per-batch counts in each rank's board, the other ranks' rows read over NVLink peer "
                                       "memory (device-side, after every batch, inside the timed region); no collective"}
        run.ctx.peers_connect(0, 0)
    host_trace = None
    if want_parity:
        par, host_trace = parity_check(run)
        oks = sum_over_ranks([1 if par["ok"] else 0], world, dev)[0]
        par["ranks_ok"] = oks
        par["ok"] = bool(oks == world)
        out["parity"] = par
    if want_roof:
        out["roofline"] = roofline(run, wl, ms, load_traffic())
    if want_cpu and rank == 0 and host_trace is not None:
        n = run.n
        out["cpu_baseline"] = cpu_baseline(wl, host_trace[0][W * n:(W + 1) * n], host_trace[1][W * n:(W + 1) * n])
    if want_e2e:
        if host_trace is None:
            host_trace = (run.d_start.cpu().numpy(), run.d_now.cpu().numpy())
        out["e2e"] = e2e_leg(run, host_trace, pcie)
    run.close()
    return out


def nccl_leg(wl, K, W, world, rank, local, dev, stream):
    """The north star's literal form: the ranks' counts summed by NCCL (cb_allreduce_counts: one 24-byte
    ncclAllReduce on a side stream, device-ordered after the batch's resolve kernel, no host wait)
    -- once per K-step window and, separately, after every batch."""
    import torch.distributed as dist

    import cueball_b200 as cb
    run = Run(wl, K, W, world, rank, local, dev, stream, seed_salt=31)
    obj = [cb.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    run.ctx.comm_init(world, rank, obj[0])
    out = {}
    # (a) once per window: the last timed batch's counts (each rank's batch counts; a cumulative sum is the
    # running totals of cb_batch_stats) -- ONE collective per K steps
    calls = {"i": 0}

    def last_only():
        calls["i"] += 1
        if calls["i"] <= W or calls["i"] == W + K:  # the warm-up steps' calls set up the communicator's channels
            run.ctx.allreduce_counts()
    ms, _, _ = timed_pass(run, per_step=last_only)
    ms = max_over_ranks(ms, world, dev)
    g = run.ctx.stats()
    out["per_window"] = {"value": world * run.n * K / (ms / 1e3), "unit": "claims/s", "ms_per_step": ms / K,
                         "collective": "1 ncclAllReduce(3 x u64) per %d-step window, asynchronous on a side stream" % K,
                         "last_batch_global": {"drops": int(g.global_drops), "admits": int(g.global_admits), "claims": int(g.global_n)}}
    # (b) after every batch
    run.reset()
    ms2, _, _ = timed_pass(run, per_step=run.ctx.allreduce_counts)
    ms2 = max_over_ranks(ms2, world, dev)
    out["per_step"] = {"value": world * run.n * K / (ms2 / 1e3), "unit": "claims/s", "ms_per_step": ms2 / K,
                       "collective": "cb_allreduce_counts = 1 ncclAllReduce(3 x u64) after every batch, on a side stream "
                                     "(an 8-rank 24-byte allreduce is ~80 us and they serialise on their communicator)"}
    run.close()
    return out


def ring_leg(wl, K, W, world, rank, local, dev, stream):
    """ONE logical queue over the GPUs: segment (k, r) of the queue is batch k of rank r; the ControlledDelay
    state and the admit cursor travel round a ring of device mailboxes over NVLink peer memory.  Checked bit
    for bit against one sequential drain (C oracle) of the queue's first steps."""
    import torch
    import torch.distributed as dist

    import cueball_b200 as cb
    from oracle import oracle as O
    w = WORKLOADS[wl]
    m, nb = w["n"], W + K
    seed = tracegen.SEED + 101
    unit = w["gen"]["unit"]
    # this rank's segments of the one global trace, generated in place (see gen_trace_torch)
    ds = torch.empty(m * nb, dtype=torch.int64, device=dev)
    dn = torch.empty(m * nb, dtype=torch.int64, device=dev)
    for k in range(nb):
        lo = (k * world + rank) * m
        s, nw = gen_trace_torch(lo, lo + m, seed, dev, unit)
        ds[k * m:(k + 1) * m], dn[k * m:(k + 1) * m] = s, nw
        del s, nw
    dd = torch.empty(m * nb, dtype=torch.uint8, device=dev)
    sl = torch.empty(m * nb, dtype=torch.int32, device=dev)
    ctx = cb.Context(local)
    ctx.codel_init(TARGET_MS)
    ctx.backends_set(w["B"], tracegen.gen_insert_draws(w["B"]), w["S"])
    handle, _ = ctx.ring_mailbox()
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    ctx.ring_connect(world, rank, next_handle=handles[(rank + 1) % world] if world > 1 else None)
    ctx.set_stream(stream.cuda_stream)
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gc_was = gc.isenabled()
    gc.disable()
    rt = _realtime(True)
    for k in range(nb):
        if k == W:
            ctx.join()
            e0.record(stream)
        lo = k * m
        ctx.admit_batch_ptr(ds.data_ptr() + 8 * lo, dn.data_ptr() + 8 * lo, None, m, dd.data_ptr() + lo, sl.data_ptr() + 4 * lo, None)
    ctx.join()
    e1.record(stream)
    _realtime(False, rt)
    if gc_was:
        gc.enable()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    # window check: the queue's first kc steps, drained sequentially by the oracle on every rank's host
    kc = max(1, min(nb, (1 << 27) // (m * world)))
    gs, gn = gen_trace_torch(0, kc * world * m, seed, dev, unit)
    hs, hn = gs.cpu().numpy(), gn.cpu().numpy()
    del gs, gn
    cd = O.new_codel(TARGET_MS)
    od, os_, _ = O.drain(cd, hs, hn, None, nslots=w["S"])
    ok = True
    for k in range(kc):
        lo = (k * world + rank) * m
        ok = ok and bool(np.array_equal(dd[k * m:(k + 1) * m].cpu().numpy(), od[lo:lo + m]))
        ok = ok and bool(np.array_equal(sl[k * m:(k + 1) * m].cpu().numpy(), os_[lo:lo + m]))
    oks = sum_over_ranks([1 if ok else 0], world, dev)[0]
    ctx.synchronize()
    ctx.ring_connect(0, 0)
    ctx.close()
    del ds, dn, dd, sl
    torch.cuda.empty_cache()
    return {"value": world * m * K / (ms / 1e3), "unit": "claims/s", "ms_per_step": ms / K, "claims_per_segment": m,
            "segments_per_step": world, "workload": wl,
            "bit_exact_vs_one_sequential_drain": bool(oks == world),
            "checked": "decisions and slots of the queue's first %d steps (%d claims) on every rank against ONE sequential "
                       "drain by the C oracle" % (kc, kc * world * m),
            "mode": "one logical queue: ControlledDelay state + admit cursor round a ring of device mailboxes over NVLink "
                    "peer memory (cb_ring_*), F hand-over ahead of the state; no collective"}


def stress_leg(local, dev, stream):
    """SURVEY 8(d)'s non-monotone stress trace: the clock steps back 150 ms every 2^20 claims, so every batch takes
    the exact sequential path (one warp runs lib/codel.js:56-86 claim by claim -- the count-2 / sqrt(count)
    branch of :72-79 is only reachable this way).  Full parity against the oracle; the rate is the fallback's."""
    import torch

    import cueball_b200 as cb
    from oracle import oracle as O
    n, nb = 1 << 21, 3
    start, now = tracegen.gen_trace(n * nb, seed=tracegen.SEED + 5, target_ns=TARGET_MS * 1_000_000, unit=1 << 18,
                                    all_overloaded_from=1 << 16)
    now = now.copy()
    for j in range(1 << 20, n * nb, 1 << 20):
        now[j:] -= 150_000_000  # the clock steps back 150 ms
        now[j:] = np.maximum(now[j:], start[j:] - 10 ** 12)  # keep it an int64 clock; sojourn may go negative
    ds, dn = torch.from_numpy(start).to(dev), torch.from_numpy(now).to(dev)
    dd = torch.empty(n * nb, dtype=torch.uint8, device=dev)
    sl = torch.empty(n * nb, dtype=torch.int32, device=dev)
    ctx = cb.Context(local)
    ctx.codel_init(TARGET_MS)
    ctx.backends_set(64, tracegen.gen_insert_draws(64), 256)
    ctx.set_stream(stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for b in range(nb):
        lo = b * n
        ctx.admit_batch_ptr(ds.data_ptr() + 8 * lo, dn.data_ptr() + 8 * lo, None, n, dd.data_ptr() + lo, sl.data_ptr() + 4 * lo, None)
    ctx.join()
    e1.record(stream)
    torch.cuda.synchronize()
    st = ctx.stats()
    cd = O.new_codel(TARGET_MS)
    od, os_, ost = O.drain(cd, start, now, None, nslots=256)
    ref = cd.as_dict()
    ok = bool(np.array_equal(dd.cpu().numpy(), od) and np.array_equal(sl.cpu().numpy(), os_) and st.total_drops == ost.drops
              and np.float64(st.state.drop_next).tobytes() == np.float64(ref["drop_next"]).tobytes()
              and float(st.state.count) == ref["count"])
    ms = e0.elapsed_time(e1)
    ctx.close()
    return {"value": n * nb / (ms / 1e3), "unit": "claims/s", "claims": n * nb, "path": int(st.path),
            "max_count_reached": float(ref["count"]), "bit_exact_vs_oracle": ok,
            "note": "backwards clock: every batch runs the exact one-warp state machine (CB_PATH_SEQUENTIAL); this is "
                    "the fallback's rate, not the streaming path's"}


def run_cuda(args, wl, extra=True):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = bind_to_gpu_numa(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    # a real (non-legacy-default) stream: the library's kernels and the timing events share it
    stream = torch.cuda.Stream(device=dev)
    assert stream.cuda_stream != 0
    K, W = args.steps, args.warmup

    fast = os.environ.get("CB_BENCH_FAST") == "1"  # diagnosis only: the timed pass and nothing else (the line is incomplete)
    pcie = pcie_rates(dev, torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)) if (rank == 0 or world > 1) and not fast else None
    sampler = ClockSampler(local, world) if rank == 0 and os.environ.get("CB_BENCH_SAMPLER", "1") != "0" else None
    main = device_leg(wl, K, W, world, rank, local, dev, stream, want_parity=not fast, want_e2e=not fast, want_roof=not fast, pcie=pcie,
                      peers=True, sampler=sampler, want_cpu=(extra and world == 1 and not fast))
    w = WORKLOADS[wl]
    cfg = workload_config(wl, world)
    line = {
        "metric": METRIC, "value": main["value"], "unit": "claims/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": main["ms"] / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64+int64", "data": "synthetic",
        "config": cfg,
        "details": {
            "l2": "inputs larger than L2: %d consecutive batches of one trace resident in HBM (%.0f MB), each read once"
                  % (W + K, 16.0 * w["n"] * (W + K) / 1e6),
            "state": "ControlledDelay state chained on device across steps",
            "pipeline": "up to 8 batches in flight, one lane (stream) each; the scan is launched at low priority, "
                        "resolve/emit at high; commits ordered by a device-side chain word",
            "global_counts": "per batch, device-side, over peer-memory boards (cb_peers_*)" if world > 1 else None,
            "timed_region": "barrier + synchronize, the last 2 warm-up steps (device-ordered, so the timed steps start on a busy GPU), "
                            "CUDA event, the K steps, cb_join (stream-ordered), CUDA event, barrier + synchronize; no host wait and "
                            "no torch collective inside",
            "admit_fraction": main["admit_fraction"],
            "cpu_affinity": ("%d CPUs of the GPU's NUMA node" % numa_cpus) if numa_cpus else "unbound"},
        "roofline": main.get("roofline"), "e2e": main.get("e2e"), "gpu_launches": main["launches"], "clocks": main["clocks"],
        "host_enqueue_ms_per_step": main["host_ms"], "parity": main.get("parity"),
    }
    if "global_counts" in main:
        line["global_counts"] = main["global_counts"]
    if "ms_per_step_by_rank" in main:
        line["ms_per_step_by_rank"] = main["ms_per_step_by_rank"]
    if "cpu_baseline" in main:
        line["cpu_baseline"] = main["cpu_baseline"]
    if extra:
        if world == 1:
            # configs[1]: 1M claims per step (launch/latency-bound: 21 MB is 3.2 us at the roofline)
            k2, w2 = 200, 10
            c2 = device_leg("c2", k2, w2, world, rank, local, dev, stream, want_parity=True, want_e2e=True, want_roof=True,
                            pcie=pcie)
            line["c2"] = {"value": c2["value"], "unit": "claims/s", "ms_per_step": c2["ms"] / k2, "steps": k2, "warmup": w2,
                          "host_enqueue_ms_per_step": c2["host_ms"], "gpu_launches": c2["launches"], "roofline": c2["roofline"],
                          "e2e": c2["e2e"], "parity": c2["parity"], "config": workload_config("c2", 1)}
            line["nonmonotone_stress"] = stress_leg(local, dev, stream)
        else:
            line["nccl_allreduce"] = nccl_leg(wl, K, W, world, rank, local, dev, stream)
            line["one_queue"] = ring_leg("c4", K, W, world, rank, local, dev, stream)
            for xw in ("c4", "c5"):
                x = device_leg(xw, K, W, world, rank, local, dev, stream, want_parity=True, want_e2e=False, want_roof=(xw == "c4"),
                               peers=True)
                line[xw] = {"value": x["value"], "unit": "claims/s", "ms_per_step": x["ms"] / K, "gpu_launches": x["launches"],
                            "parity": x["parity"], "global_counts": x.get("global_counts"), "roofline": x.get("roofline"),
                            "admit_fraction": x["admit_fraction"], "config": workload_config(xw, world)}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the cpu_baseline leg and the extra keys")
    args = ap.parse_args()
    wl = args.workload or "c3"
    dk, dw = {"c2": (200, 10), "c3": (20, 5), "c4": (30, 5), "c5": (20, 5)}[wl]
    if args.impl == "reference":
        dk, dw = 8, 3  # seconds of single-threaded JS per step
    args.steps = args.steps if args.steps is not None else dk
    args.warmup = max(3, args.warmup if args.warmup is not None else dw)
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            run_reference(args, wl)
        return
    line = run_cuda(args, wl, extra=not args.no_extra)
    if line is not None:
        print(json.dumps(line))


if __name__ == "__main__":
    main()
