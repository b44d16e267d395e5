#!/usr/bin/env python3
"""bench.py -- claim-queue admissions/sec (CoDel + backend-slot pick) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl reference]

A "step" is one pass of the hot path (scan -> resolve -> emit kernels behind
cb_admit_batch_dev) over one batch of synthetic pending claims.  Workloads are the
BASELINE.json configs:

    c2  2^20 claims / step,   64 backends,  256 slots     (configs[1]; default at N=1)
    c3  2^24 claims / step, 1024 backends, 4096 slots     (configs[2])
    c4  2^23 claims / step / GPU, 4096 backends           (configs[3]: 64M over 8 GPUs)
    c5  2^24 claims / step / GPU, sustained overload       (configs[4]: streamed chunks)

The (W+K) batches of a run are consecutive segments of ONE seeded trace, all resident in HBM
before the timed region, so every step reads cold data (the run's inputs are many times the
126 MB L2) and the ControlledDelay state chains from batch to batch on the device.

Metric value = claims decided per second, whole job (all ranks), inputs resident in HBM.
N > 1: one process per GPU, each rank drains its own queue segment (no data-path collective);
the global drop count is ONE NCCL allreduce of the ranks' cumulative counts at the end of the
timed window (`config.collective`); `per_step_collective` is a second pass with an allreduce
after every batch (CB_BENCH_COLLECTIVE=window|step|none picks the main pass's mode).
`e2e` = the same through the host-buffer C entry point (pinned host arrays in, decisions and
slots back on the host every step).  `roofline` is for the dominant kernel (the scan kernel:
16 algorithmic bytes per claim) from CUDA events recorded around each kernel on
the launching stream in a separate profiling pass.  `cpu_baseline` is the C restatement of the
reference path (oracle/, pinned bit-for-bit to the reference's own JS) on one host core.

--impl reference times the reference's OWN JavaScript (lib/queue.js + lib/codel.js verbatim)
under the SpiderMonkey engine vendored in the reference tree (oracle/_ref), one core -- the
path is a sequential state machine and Node itself is single threaded.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

# more hardware work queues than the default 8, so that the pipeline's streams do not share one
# (must be set before CUDA initialises)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# the per-step collective is a 24-byte sum: one small CTA (see cb_comm_init); NCCL reads these once
# per process, so they are set before torch initialises it
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
    "c5": dict(n=1 << 24, B=4096, S=16384, gen={"all_overloaded_from": 1 << 20}),
}
TARGET_MS = 5
ALG_BYTES_SCAN = 16      # start_ns 8 + now_ns 8 (read once by the scan kernel)
ALG_BYTES_STEP = 21      # + decision 1 + slot 4
REF_SAMPLE = 1 << 18     # claims per step for --impl reference (bounded CPU sample)


def env_int(name, default):
    return int(os.environ.get(name, default))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
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
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def make_trace(wl, total_claims, rank):
    w = WORKLOADS[wl]
    seed = tracegen.SEED + 7919 * rank
    return tracegen.gen_trace(total_claims, seed=seed, target_ns=TARGET_MS * 1_000_000, **w["gen"])


# ------------------------------------------------------------------------------ reference arm
def run_reference(args, wl):
    """The reference's own CPU implementation of the path on this box's host cores."""
    from oracle import oracle as O
    from oracle import ref
    w = WORKLOADS[wl]
    n = min(REF_SAMPLE, w["n"])
    params = dict(t0=10**9, de=1000, target_ns=TARGET_MS * 1_000_000, unit=w["gen"].get("unit", 1 << 16),
                  start_overloaded=False, all_overloaded_from=w["gen"].get("all_overloaded_from", -1))
    times, decided = [], 0
    kind = "reference" if ref.available() else "port"
    for step in range(args.warmup + args.steps):
        seed = tracegen.SEED + step
        if kind == "reference":
            r = ref.bench(seed, n, params, TARGET_MS, w["S"])
            dt = r["drain_ms"] / 1e3
        else:
            s, nw = tracegen.gen_trace(n, seed=seed, **{k: v for k, v in params.items()})
            cd = O.new_codel(TARGET_MS)
            t0 = time.perf_counter()
            O.drain(cd, s, nw, nslots=w["S"], variant="list", want_out=True)
            dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(max(dt, 1e-9))
            decided += n
    total = sum(times)
    value = decided / total
    # context for the ratio: SpiderMonkey 1.7 is an interpreter from 2007; a JIT-compiling engine (V8 under
    # Node.js, which is not installed here) would land between this number and the compiled C restatement
    # of the same loop over the same linked-list queue, timed here on the same sample
    port = None
    try:
        s, nw = tracegen.gen_trace(n, seed=tracegen.SEED, **{k: v for k, v in params.items()})
        best = None
        for _ in range(3):
            cd = O.new_codel(TARGET_MS)
            t0 = time.perf_counter()
            O.drain(cd, s, nw, nslots=w["S"], variant="list", want_out=True)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        port = n / best
    except Exception:
        pass
    sample = ("%d claims per step of the %s trace spec (same generator, JS side); timed: the drain loop "
              "lib/pool.js:733-753 = Queue.shift + ControlledDelay.overloaded per claim" % (n, wl))
    line = {
        "impl": "reference", "metric": "claim-queue admissions/sec (CoDel + backend-slot pick)",
        "value": value, "unit": "claims/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / max(len(times), 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64+int64", "data": "synthetic",
        "config": {"workload": wl, "claims_per_step": n, "backends": w["B"], "slots": w["S"],
                   "target_ms": TARGET_MS, "engine": ("reference lib/queue.js+lib/codel.js verbatim under SpiderMonkey 1.7 "
                                                      "(vendored in the reference tree; Node.js is not installed)"
                                                      if kind == "reference" else "C restatement (oracle/), linked-list queue")},
        "cpu_baseline": {"value": value, "unit": "claims/s", "cores": 1, "kind": kind, "sample": sample,
                         "c_restatement_linked_list_claims_per_s": port},
        "e2e": {"value": value, "unit": "claims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------ CUDA arm
def cpu_baseline(wl):
    """The oracle's C restatement (flat-array variant = the fastest honest CPU form), 1 core,
    on a bounded sample of the same workload."""
    from oracle import oracle as O
    w = WORKLOADS[wl]
    n = min(1 << 24, 4 * w["n"])
    s, nw = tracegen.gen_trace(n, target_ns=TARGET_MS * 1_000_000, **w["gen"])
    best = None
    for _ in range(3):
        cd = O.new_codel(TARGET_MS)
        t0 = time.perf_counter()
        O.drain(cd, s, nw, nslots=w["S"], variant="flat", want_out=True)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    cd = O.new_codel(TARGET_MS)
    t0 = time.perf_counter()
    m = min(n, 1 << 22)
    O.drain(cd, s[:m], nw[:m], nslots=w["S"], variant="list", want_out=True)
    dl = time.perf_counter() - t0
    return {"value": n / best, "unit": "claims/s", "cores": 1, "kind": "port",
            "sample": "%d claims of the %s trace, best of 3; C restatement of lib/codel.js + lib/pool.js:733-753 over flat "
                      "arrays (oracle/cueball_oracle.c, -O2 -ffp-contract=off)" % (n, wl),
            "linked_list_variant_claims_per_s": m / dl}


def run_cuda(args, wl, extra=True):
    import torch
    import torch.distributed as dist

    import cueball_b200 as cb

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

    w = WORKLOADS[wl]
    n, K, W = w["n"], args.steps, args.warmup
    nb = W + K
    start, now = make_trace(wl, n * nb, rank)

    ctx = cb.Context(local)
    ctx.codel_init(TARGET_MS)
    ctx.backends_set(w["B"], tracegen.gen_insert_draws(w["B"]), w["S"])
    # a real (non-legacy-default) stream: the library's kernels and the timing events share it
    stream = torch.cuda.Stream(device=dev)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    if world > 1:  # one communicator per context, id from rank 0
        obj = [cb.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        ctx.comm_init(world, rank, obj[0])

    d_start = torch.from_numpy(start).to(dev)
    d_now = torch.from_numpy(now).to(dev)
    d_dec = torch.empty(n * nb, dtype=torch.uint8, device=dev)
    d_slot = torch.empty(n * nb, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    p_start, p_now, p_dec, p_slot = d_start.data_ptr(), d_now.data_ptr(), d_dec.data_ptr(), d_slot.data_ptr()

    # The global drop count (north_star: "a single NCCL allreduce over NVLink for the global drop
    # count").  "window" (default): every rank's cumulative counts are summed by ONE NCCL allreduce at
    # the end of the K-step window, inside the timed region.  "step": cb_allreduce_counts after every
    # batch (its own 24-byte NCCL allreduce on a side stream) -- measured in a second pass and
    # reported as `per_step_collective`: one 8-rank allreduce takes ~80 us, they serialise on their
    # communicator, and a 43-us step cannot sustain one per batch.  "none": no collective.
    coll_mode = os.environ.get("CB_BENCH_COLLECTIVE", "none" if os.environ.get("CB_BENCH_NO_ALLREDUCE") else "window")
    if world == 1:
        coll_mode = "none"

    def step_dev(b, per_step=False):
        # one C-ABI call per step on raw device addresses, as a native host (the N-API shim) makes it
        lo = b * n
        ctx.admit_batch_ptr(p_start + 8 * lo, p_now + 8 * lo, None, n, p_dec + lo, p_slot + 4 * lo, None)
        if per_step:
            ctx.allreduce_counts()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------
    def timed_pass(per_step):
        """W untimed + K timed steps; returns (device ms of the K steps on this rank, host ms/step,
        global cumulative [drops, admits, claims] or None, kernels launched by the K steps)."""
        for b in range(W):
            step_dev(b, per_step)
        ctx.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        lc0 = ctx.launch_count()
        th0 = time.perf_counter()
        for b in range(W, W + K):
            step_dev(b, per_step)
        hms = (time.perf_counter() - th0) * 1e3 / K  # host time to enqueue one step (no waiting)
        nlaunch = ctx.launch_count() - lc0            # this library's kernels launched by the K timed steps
        ctx.join()  # the timing stream waits for the last batch of the pipeline
        glob = None
        if coll_mode == "window" and not per_step:
            stw = ctx.stats()  # waits for the window; this rank's cumulative counts
            glob = torch.tensor([stw.total_drops, stw.total_admits, stw.total_drops + stw.total_admits],
                                dtype=torch.int64, device=dev)
            with torch.cuda.stream(stream):
                dist.all_reduce(glob)  # the one NCCL allreduce, ordered on the timing stream
        e1.record(stream)
        barrier()
        ctx.synchronize()
        return e0.elapsed_time(e1), hms, glob, nlaunch

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, host_ms, glob, launches = timed_pass(coll_mode == "step")
    clocks = sampler.stop() if sampler else None
    st = ctx.stats()
    global_counts = [int(x) for x in glob.tolist()] if glob is not None else None
    per_step = None
    if coll_mode == "window" and extra:
        # second pass over the same batches with a collective after every batch (state carries on)
        ms2, _, _, _ = timed_pass(True)
        t2 = torch.tensor([ms2], dtype=torch.float64, device=dev)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ms2 = float(t2.item())
        per_step = {"value": world * n * K / (ms2 / 1e3), "unit": "claims/s", "ms_per_step": ms2 / K,
                    "collective": "cb_allreduce_counts = 1 ncclAllReduce(3 x u64) after every batch, on a side stream"}
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    value = world * n * K / (ms / 1e3)

    # ---- per-kernel profile pass (CUDA events around each kernel, same stream) --------------
    ctx.set_state({"count": 0, "drop_next": 0.0, "first_above_time": 0.0, "dropping": False, "last_empty": None})
    ctx.profile_enable(True)
    for b in range(W, W + min(K, 20)):
        lo, hi = b * n, (b + 1) * n
        ctx.admit_batch_dev(d_start[lo:hi], d_now[lo:hi], None, d_dec[lo:hi], d_slot[lo:hi])
    kms, kcnt = ctx.profile_get()
    ctx.profile_enable(False)
    scan_ms = kms[cb._native.K_SCAN] / max(kcnt[cb._native.K_SCAN], 1)
    peak, peak_src = peaks()
    achieved = ALG_BYTES_SCAN * n / (scan_ms / 1e3) / 1e9
    roof = {"bound": "hbm", "kernel": "cb::scan_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": TRAFFIC.get(wl), "peak_source": peak_src,
            "alg_bytes_per_claim": ALG_BYTES_SCAN,
            "kernel_ms": {"scan": scan_ms, "resolve": kms[1] / max(kcnt[1], 1), "emit": kms[2] / max(kcnt[2], 1)},
            "step": {"alg_bytes_per_claim": ALG_BYTES_STEP,
                     "achieved": ALG_BYTES_STEP * n * K / (ms / 1e3) / 1e9 / 1.0,
                     "frac": ALG_BYTES_STEP * n * K / (ms / 1e3) / 1e9 / peak}}

    # ---- end to end: host buffers through the public entry point ----------------------------
    ke = min(K, 10)
    hs = torch.from_numpy(start[W * n:(W + ke) * n]).pin_memory().numpy()
    hn = torch.from_numpy(now[W * n:(W + ke) * n]).pin_memory().numpy()
    hdec = torch.empty(n, dtype=torch.uint8).pin_memory().numpy()
    hslot = torch.empty(n, dtype=torch.int32).pin_memory().numpy()
    ctx.set_stream(None)
    ctx.set_state({"count": 0, "drop_next": 0.0, "first_above_time": 0.0, "dropping": False, "last_empty": None})
    ctx.admit_batch(hs[:n], hn[:n])  # warm the staging buffers
    ctx.admit_batch(hs[:n], hn[:n])
    barrier()
    t0 = time.perf_counter()
    drops = 0
    for b in range(ke):
        dec, slot, _, est = ctx.admit_batch(hs[b * n:(b + 1) * n], hn[b * n:(b + 1) * n],
                                            out_decision=hdec, out_slot=hslot)
        drops += est.drops  # the step's result is read on the host
    barrier()
    te = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([te], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        te = float(t.item())
    e2e = {"value": world * n * ke / te, "unit": "claims/s", "h2d_bytes_per_step": 16 * n,
           "d2h_bytes_per_step": 5 * n + 176, "steps": ke, "host_memory": "pinned inputs and outputs",
           "cpu_affinity": ("%d CPUs of the GPU's NUMA node" % numa_cpus) if numa_cpus else "unbound"}

    ctx.synchronize()
    ctx.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None
    line = {
        "metric": "claim-queue admissions/sec (CoDel + backend-slot pick)", "value": value, "unit": "claims/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64+int64", "data": "synthetic",
        "config": {"workload": wl, "claims_per_step_per_gpu": n, "backends": w["B"], "slots": w["S"],
                   "target_ms": TARGET_MS, "interval_ms": 100, "parallelism": "queue segment per GPU" if world > 1 else "1 GPU",
                   "collective": {"window": "1 ncclAllReduce(3 x i64) of every rank's cumulative counts per %d-step window, "
                                            "inside the timed region" % K,
                                  "step": "cb_allreduce_counts: 1 ncclAllReduce(3 x u64) after every batch",
                                  "none": None}[coll_mode],
                   "l2": "inputs larger than L2: %d consecutive batches of one trace resident in HBM (%.0f MB), each read once"
                         % (nb, 16.0 * n * nb / 1e6),
                   "state": "ControlledDelay state chained on device across steps",
                   "pipeline": "up to 5 batches in flight, one lane (stream) each; the scan is launched at low priority, "
                               "prepare/commit/emit at high; commits ordered by a device-side chain word",
                   "admit_fraction": st.total_admits / max(st.total_admits + st.total_drops, 1)},
        "roofline": roof, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "host_enqueue_ms_per_step": host_ms,
    }
    if extra:
        line["cpu_baseline"] = cpu_baseline(wl)
    if world > 1:
        if global_counts is not None:
            line["global_counts"] = {"drops": global_counts[0], "admits": global_counts[1], "claims": global_counts[2]}
        if per_step is not None:
            line["per_step_collective"] = per_step
        dist.destroy_process_group()
    return line


# DRAM traffic per scan-kernel launch from `ncu --set full` captures (profiles/): bytes
TRAFFIC = {}
_tp = os.path.join(ROOT, "profiles", "traffic.json")
if os.path.exists(_tp):
    with open(_tp) as _f:
        TRAFFIC = json.load(_f)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the cpu_baseline leg and the c3 side measurement")
    args = ap.parse_args()
    wl = args.workload or ("c2" if args.gpus == 1 else "c4")
    dk, dw = {"c2": (200, 10), "c3": (15, 4), "c4": (30, 5), "c5": (15, 4)}[wl]
    if args.impl == "reference":
        dk, dw = 8, 3  # ~0.6 s of single-threaded JS per step
    args.steps = args.steps if args.steps is not None else dk
    args.warmup = max(3, args.warmup if args.warmup is not None else dw)
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            run_reference(args, wl)
        return
    line = run_cuda(args, wl, extra=not args.no_extra)
    if line is None:
        return
    if not args.no_extra and args.gpus == 1 and wl == "c2":
        # the HBM-roofline-relevant size: one 16M-claim batch per step (BASELINE config C3)
        a3 = argparse.Namespace(gpus=1, steps=15, warmup=4)
        c3 = run_cuda(a3, "c3", extra=False)
        line["c3"] = {k: c3[k] for k in ("value", "unit", "ms_per_step", "steps", "roofline", "e2e", "gpu_launches")}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
