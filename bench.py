#!/usr/bin/env python3
"""bench.py -- min-snap trajectories/sec (16-segment, xyz, fp64) on N B200s vs the host CPU.

Workload (BASELINE.json configs[1]): 4 096 trajectories x 16 segments, derivative order 4
(degree-7 polynomials), fp64, random-walk waypoints (rng 1234), V_avg 5 m/s, sample spacing 1 m, the reference's
shipped penalty weights (path_weight 1e-7, vel_zero_weight 0.01) so that every row of the hot path runs: time
allocation, pass-1 solve, worst-deviation search, penalised solve, the 10-step reweighting loop, coefficient
recovery and the distance-thresholded sampler.  A "step" is one full GenerateTrajectoryMatrix pass over one batch.
The plain-weights variant (path_weight = vel_zero_weight = 0: one solve, HBM-bound) is measured as well and reported
under "variants".  Two more objects describe the same step widened into the neighbouring stages of getPlan (SURVEY.md
section 8f): "wgs84_frame" (the sampled rows leave as WGS84 lon/lat/alt) and "leader_chain" (WGS84 waypoints -> ENU ->
minimum snap -> samples -> cost-map lookup -> altitude optimisation -> WGS84 rows, all on the device).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's own CPU code (oracle/_ref)
    python bench.py --impl rows-cpu                                # CPU baselines of the section-8f stages (oracle ports)

Prints ONE JSON line (rank 0).  See the task contract for the keys; numbers are never taken under a profiler.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "min-snap trajectories/sec (16-seg, xyz, fp64)"
UNIT = "trajectories/s"
B_DEFAULT, NS, ORDER = 4096, 16, 4
ROTATE = 8  # distinct input/output buffer sets cycled through so a step never finds its data in L2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "rows-cpu"])
    ap.add_argument("--weights", default="shipped", choices=["shipped", "plain"])
    ap.add_argument("--batch", type=int, default=B_DEFAULT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--configs", default="all",
                    help="comma list out of cfg1,cfg3,cfg4,cfg5 (the other BASELINE.json configs, reported under "
                         "'variants'), 'all' or 'none'; cfg3 and cfg5 are ONE batch sharded over the ranks")
    ap.add_argument("--no-parity", action="store_true", help="skip the whole-batch parity check of the cpu_baseline leg")
    ap.add_argument("--streams", type=int, default=4,
                    help="solver handles (each with its own CUDA stream) per GPU that consecutive steps alternate between "
                         "(a divisor of 8; measured on one B200: 2 and 3 -> 63.4 M trajectories/s, 4 -> 65.7 M, 8 -> 65.1 M)")
    return ap.parse_args()


def workload_name(weights, B):
    w = "shipped penalty weights pw=1e-7 vw=0.01 (1 + up to 11 solves/trajectory)" if weights == "shipped" else \
        "plain weights pw=vw=0 (1 solve/trajectory)"
    return f"cfg2: {B} trajectories x {NS} segments, order {ORDER} (degree 7), fp64, V_avg 5, sample 1 m, {w}"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock / throttle reasons through NVML (falls back to nvidia-smi) while the GPU is under load."""

    def __init__(self, index):
        self.index, self.samples, self._stop = index, [], threading.Event()
        self.max_mhz = None
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self._stop.is_set():
                try:
                    util = nv.nvmlDeviceGetUtilizationRates(h).gpu
                except Exception:
                    util = -1
                self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), int(get_reasons(h)), util))
                time.sleep(0.02)
        except Exception as e:  # pragma: no cover - depends on the box
            self.error = repr(e)

    def start(self):
        self._t.start()
        return self

    def stop(self):
        self._stop.set()
        self._t.join(timeout=2)
        names = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown"}
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": getattr(self, "error", "no samples")}
        busy = [s for s in self.samples if s[2] > 0] or self.samples
        reasons = sorted({n for s in busy for bit, n in names.items() if s[1] & bit})
        return {"sm_mhz": statistics.median(s[0] for s in busy), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(busy)}


# ------------------------------------------------------------------------------------------------ reference arm
def ref_config(weights):
    from oracle import ref

    pw, vw = (1e-7, 0.01) if weights == "shipped" else (0.0, 0.0)
    return ref.RefConfig(order=ORDER, path_weight=pw, vel_zero_weight=vw, V_avg=5.0, min_time_s=0.1, sample_distance=1.0)


def cpu_rate(weights, wp, n_traj, threads):
    """Trajectories/s of the reference CPU code on `n_traj` trajectories with `threads` OpenMP threads."""
    from oracle import ref

    off = np.arange(n_traj + 1, dtype=np.int64) * (NS + 1)   # CSR over waypoint rows: NS + 1 points per trajectory
    t0 = time.perf_counter()
    counts, used, _ = ref.generate_batch(off, wp[: n_traj * (NS + 1)], ref_config(weights), nthreads=threads, kind="fast")
    dt = time.perf_counter() - t0
    return n_traj / dt, used, int(counts.sum())


def cpu_baseline(weights, wp, budget_s=12.0):
    """Bounded sample of the same workload on all host cores (+ the single-threaded as-shipped call sequence)."""
    cores = os.cpu_count() or 1
    r1, _, _ = cpu_rate(weights, wp, 8, 1)                       # single thread, as shipped: one call per trajectory
    n = int(min(wp.shape[0] // (NS + 1), max(cores * 2, r1 * cores * budget_s * 0.6)))
    rN, used, _ = cpu_rate(weights, wp, n, cores)
    return {"value": rN, "unit": UNIT, "cores": used, "kind": "reference",
            "sample": f"first {n} of the {wp.shape[0] // (NS + 1)} trajectories, one GenerateTrajectoryMatrix call each, "
                      f"OpenMP over trajectories on {used} threads; unmodified reference minimum_snap.cpp built against "
                      f"the oracle's Eigen shim -- NOT real Eigen (not installable here): the shim's inverse()/operator* "
                      f"skip zero multiplicands, which makes this baseline faster than a naive dense backend "
                      f"(conservative for the ratio); -O3 -march=x86-64-v3/v4; single_thread_value = 8 trajectories, "
                      f"1 thread, as-shipped call sequence",
            "sample_trajectories": n, "single_thread_value": r1,
            "structured_cpu": structured_cpu_baseline(weights, wp, cores)}


def structured_cpu_baseline(weights, wp, cores):
    """The "good CPU" line: the library's own structured O(ns) algorithm -- the sequential kernel set
    of csrc/msnap_generic.cuh -- compiled for the host (oracle/structured_cpu.cpp), whole batch, OpenMP over trajectories.
    Separates what the algorithm buys (this line over `value`) from what the GPU buys (the bench value over this line)."""
    try:
        from cs_pathplan_b200 import workloads
        from oracle import structured_ref as sr

        cfg = workloads.synthetic_config(4, weights)
        B = wp.shape[0] // (NS + 1)

        def rate(threads, reps):
            best = None
            for _ in range(reps):
                t0 = time.perf_counter()
                r = sr.generate_batch(wp, cfg, ns=NS, threads=threads)
                dt = time.perf_counter() - t0
                best = dt if best is None or dt < best else best
            return B / best, int(r["sample_offset"][-1])

        rN, rows = rate(cores, 3)
        r1, _ = rate(1, 1)
        return {"value": rN, "unit": UNIT, "cores": cores, "single_thread_value": r1, "rows": rows, "kind": "port",
                "sample": f"all {B} trajectories, best of 3 calls; the library's sequential kernel set (block-tridiagonal "
                          f"elimination, reweighting loop, two-pass sampler) compiled for the host with -O3 -march=x86-64-v3, "
                          f"OpenMP over segments / trajectories on {cores} threads; same outputs as the GPU path (rows "
                          f"within 1e-7 m, counts equal: tests/test_gpu_full_parity.py)"}
    except Exception as e:  # noqa: BLE001 -- a baseline must never take the GPU number down with it
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": repr(e)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cs_pathplan_b200 import workloads
    from oracle import ref

    if not ref.available("fast"):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (make -C oracle needs /root/reference)"}))
        return
    wp, _ = workloads.cfg2(B=args.batch)
    cores = os.cpu_count() or 1
    r1, _, _ = cpu_rate(args.weights, wp, 8, 1)
    # bounded sample per step: about one second of all-core work, less if K is large (whole run within ~2 minutes)
    step_s = min(1.0, 120.0 / max(1, args.steps + args.warmup))
    n = int(min(args.batch, max(cores, r1 * cores * step_s * 0.6)))
    for _ in range(args.warmup):
        cpu_rate(args.weights, wp, n, cores)
    t0 = time.perf_counter()
    used = cores
    for _ in range(args.steps):
        _, used, _ = cpu_rate(args.weights, wp, n, cores)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = (f"each step = a {n}-trajectory subsample (the first {n}) of the {args.batch}-trajectory batch, one "
              f"GenerateTrajectoryMatrix call each, OpenMP over trajectories on {used} threads; value = subsample / time; "
              f"unmodified reference minimum_snap.cpp + the oracle's zero-skipping Eigen shim (not real Eigen)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.weights, args.batch), "sample_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_rows_cpu(args):
    """CPU baselines of the two neighbouring stages, timed on this host's cores on bounded samples:
    the C port of the WGS84 <-> ENU transforms (oracle/geo_port.c, -O3, 1 thread and OpenMP over rows) and a banded-
    Cholesky stand-in for the altitude optimiser's per-trajectory SimplicialLDLT loop (oracle/alt_oracle.py, 1 thread).
    Reported beside scripts/geo_bench.py and scripts/alt_bench.py; not a target."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from cs_pathplan_b200 import workloads
    from oracle import alt_oracle as ao
    from oracle import geo

    cores = os.cpu_count() or 1
    ref = np.array([109.56059880227296, 40.86719901015758, 0.0])
    enu = workloads.enu_rows(2_000_000)
    out = {"impl": "rows-cpu", "cores": cores, "geo": {"sample_rows": enu.shape[0], "kind": "port (oracle/geo_port.c, -O3)"}}
    lla = geo.enu_to_wgs84_batch(enu, ref, threads=cores, fast=True)
    for name, fn, src in (("enu_to_wgs84", geo.enu_to_wgs84_batch, enu), ("wgs84_to_enu", geo.wgs84_to_enu_batch, lla)):
        for threads in (1, cores):
            fn(src[:1000], ref, threads=threads, fast=True)
            t0 = time.perf_counter()
            fn(src, ref, threads=threads, fast=True)
            out["geo"][f"{name}_rows_per_s_{'1thread' if threads == 1 else 'all_cores'}"] = src.shape[0] / (time.perf_counter() - t0)
    rows, off, elev = workloads.sampled_rows(4096)
    m = 256
    t0 = time.perf_counter()
    for b in range(m):
        sl = slice(int(off[b]), int(off[b + 1]))
        ao.optimize_segment_altitude_enu_banded(rows[sl], ao.shipped_params(), elev[sl])
    out["alt"] = {"trajectories_per_s_1thread": m / (time.perf_counter() - t0), "sample_trajectories": m,
                  "kind": "port (oracle/alt_oracle.py: numpy assembly + LAPACK dpbsv per solve, Python driver)"}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------ our arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    from cs_pathplan_b200 import TrajectoryGeneratorTool, roofline, workloads

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # keep this rank's host threads and the pinned buffers they first touch on the GPU's own NUMA node (matters for the
    # end-to-end figure when several ranks share the box; MSNAP_BENCH_NUMA=0 switches it off)
    numa_cpus = None
    all_cpus = os.sched_getaffinity(0)
    if os.environ.get("MSNAP_BENCH_NUMA", "1") != "0":
        from cs_pathplan_b200.hostpin import pin_to_gpu_numa

        numa_cpus = pin_to_gpu_numa(local)
    if world > 1:
        # rank 0 prints ONE JSON line on stdout: NCCL's version banner (printed at NCCL_DEBUG=VERSION and above) and any
        # debug output it is asked for go to stderr
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG", None)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # Consecutive steps (independent batches) alternate between S solver handles, each on its own explicit CUDA stream,
    # so that the solve kernel of one step overlaps the sampler kernel of the previous one -- the way a service that
    # has several batches in flight would drive the library.  The timing events are recorded on stream 0, which waits
    # for the other streams' last step before the end event.  (Passing torch's legacy default stream would hand the
    # library a NULL stream, which it replaces by its own: events recorded there would not see the kernels at all.)
    S = max(1, args.streams)
    if ROTATE % S:
        S = 2  # a buffer set must always be used by the same stream
    tools = [TrajectoryGeneratorTool(local) for _ in range(S)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(S)]
    for t_, s_ in zip(tools, streams):
        t_.set_stream(s_.cuda_stream)
        assert s_.cuda_stream != 0
    tool, stream = tools[0], streams[0]
    B, m = args.batch, 2 * ORDER
    n_seg = B * NS

    class Bufs:
        pass

    def make_set(seed, cfg):
        s = Bufs()
        wp_h, _ = workloads.cfg2(B=B, seed=seed)
        s.wp_h = torch.from_numpy(wp_h).pin_memory()
        s.wp = s.wp_h.to(dev)
        s.cap = tool.sample_bound(cfg, wp_h, ns=NS)
        s.times = torch.empty(n_seg, dtype=torch.float64, device=dev)
        s.coeff = torch.empty(n_seg * 3 * m, dtype=torch.float64, device=dev)
        s.max_dev = torch.empty(B, dtype=torch.float64, device=dev)
        s.vw = torch.empty(B, dtype=torch.float64, device=dev)
        s.iters = torch.empty(B, dtype=torch.int32, device=dev)
        s.off = torch.empty(B + 1, dtype=torch.int64, device=dev)
        s.samples = torch.empty((s.cap, 3), dtype=torch.float64, device=dev)
        s.flags = torch.empty(B, dtype=torch.int32, device=dev)
        return s

    def step(cfg, s, which=0):
        tools[which].generate_batch_dev(cfg, s.wp, s.off, s.samples, ns=NS, times=s.times, coeff=s.coeff,
                                        max_dev=s.max_dev, iters=s.iters, vw_final=s.vw, flags=s.flags)

    def timed_run(cfg, sets, steps, warmup, n_streams):
        # W untimed steps -- and at least one per handle: a handle's first call sizes its workspace (allocation + sync)
        for i in range(max(warmup, n_streams)):
            step(cfg, sets[i % ROTATE], i % n_streams)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        done = [torch.cuda.Event() for _ in range(n_streams)]
        l0 = sum(t_.launch_count for t_ in tools)
        e0.record(stream)
        for k in range(1, n_streams):
            streams[k].wait_event(e0)                      # nothing of the timed region starts before e0
        for i in range(steps):
            step(cfg, sets[i % ROTATE], i % n_streams)     # buffer set i % 8 is only ever used by stream i % n_streams
        for k in range(1, n_streams):
            done[k].record(streams[k])
            stream.wait_event(done[k])                     # e1 fires after the last step of every stream
        e1.record(stream)
        e1.synchronize()
        torch.cuda.synchronize()
        launches = sum(t_.launch_count for t_ in tools) - l0
        ms = max_over_ranks(e0.elapsed_time(e1))
        barrier()
        return ms, launches

    results = {}
    sampler = ClockSampler(local).start() if rank == 0 else None
    fp64_peak = tool.measure_fp64_peak()
    for weights in ([args.weights] + [w for w in ("shipped", "plain") if w != args.weights]):
        cfg = workloads.synthetic_config(ORDER, weights)
        sets = [make_set(1234 + 1000 * rank + r, cfg) for r in range(ROTATE)]
        headline = weights == args.weights
        steps = args.steps if headline else min(args.steps, 500)
        ms, launches = timed_run(cfg, sets, steps, max(args.warmup, 3), S)
        ms_1, _ = timed_run(cfg, sets, min(steps, 300), 3, 1) if S > 1 else (ms / steps * min(steps, 300), 0)
        latency_ms = ms_1 / min(steps, 300)                # one step at a time on one stream
        tot_samples = int(sets[0].off[-1].item())
        iters0 = sets[0].iters.to(torch.int64)
        total_solves = int((iters0 + 1).sum().item())
        flags_bad = int((sets[0].flags != 0).sum().item())
        # roofline pass: per-kernel CUDA-event timing on the launching stream (separate from the headline timing)
        tool.profile_begin()
        prof_steps = 20
        for i in range(prof_steps):
            step(cfg, sets[i % ROTATE])
        prof = tool.profile_end()
        dom = max((k for k in prof if k in ("k_fused_solve", "k_sample_scan")), key=lambda k: prof[k]["total_ms"])
        dom_ms = prof[dom]["total_ms"] / prof_steps        # per step (a kernel may launch more than once per step)
        all_ms = sum(v["total_ms"] for v in prof.values()) / prof_steps
        t_np = sets[0].times.cpu().numpy()
        cand = int(np.sum(np.floor((t_np + 1e-12) / np.minimum(0.1, t_np / 10.0) + 1e-9)))
        abytes = roofline.algorithmic_bytes(B, n_seg, ORDER, tot_samples)
        aflops = roofline.algorithmic_flops(B, n_seg, ORDER, cfg.path_weight > 0, total_solves, cand)
        use_pw = cfg.path_weight > 0
        kbytes = {k: roofline.kernel_bytes(k, B, n_seg, ORDER, tot_samples) for k in ("k_fused_solve", "k_sample_scan")}
        kflops = {k: roofline.kernel_flops(k, B, n_seg, ORDER, use_pw, total_solves, cand)
                  for k in ("k_fused_solve", "k_sample_scan")}
        footprint = ROTATE * (abytes + 24 * (sets[0].cap - tot_samples))
        # end to end through the host-pointer C ABI: pinned host inputs -> H2D, solve, D2H of every result.  As in the
        # device-resident run, consecutive steps alternate between the S handles -- here each handle is driven by its own
        # host thread (the call is synchronous and releases the GIL), so one step's upload and kernels overlap the
        # previous step's download; every step still moves all of its bytes inside the timed region.
        def make_out():
            o = {k: torch.empty(shape, dtype=dt).pin_memory().numpy() for k, shape, dt in (
                ("times", (n_seg,), torch.float64), ("coeff", (n_seg, 3, m), torch.float64),
                ("max_dev", (B,), torch.float64), ("iters", (B,), torch.int32), ("vw_final", (B,), torch.float64),
                ("sample_offset", (B + 1,), torch.int64), ("samples", (sets[0].cap, 3), torch.float64))}
            o["flags"] = torch.zeros(B, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
            o["best_s"] = torch.zeros(n_seg, dtype=torch.int32).pin_memory().numpy()
            return o

        outs = [make_out() for _ in range(S)]
        out = outs[0]
        wp_np = [s.wp_h.numpy() for s in sets]
        e2e_steps = max(3 * S, min(steps, 30 * S))
        e2e_steps -= e2e_steps % S

        def e2e_worker(which, n, outputs):
            res = None
            for i in range(which, n, S):
                res = tools[which].generate_batch(cfg, wp_np[i % ROTATE], ns=NS, capacity=sets[0].cap, out=outs[which],
                                                  stats=False, outputs=outputs)
            return res

        def e2e_run(n, outputs):
            if S == 1:
                return e2e_worker(0, n, outputs)
            th = [threading.Thread(target=e2e_worker, args=(k, n, outputs)) for k in range(1, S)]
            for t_ in th:
                t_.start()
            res = e2e_worker(0, n, outputs)
            for t_ in th:
                t_.join()
            return res

        def e2e_measure(outputs):
            e2e_run(3 * S, outputs)
            barrier()
            t0 = time.perf_counter()
            r = e2e_run(e2e_steps, outputs)
            torch.cuda.synchronize()
            sec = max_over_ranks(time.perf_counter() - t0)
            barrier()
            t0 = time.perf_counter()
            for i in range(6):
                tool.generate_batch(cfg, wp_np[i % ROTATE], ns=NS, capacity=sets[0].cap, out=out, stats=False, outputs=outputs)
            single_ms = (time.perf_counter() - t0) / 6 * 1e3   # one call at a time on one handle
            d2h = int(r.samples.nbytes + out["sample_offset"].nbytes + out["flags"].nbytes)
            if outputs == "all":
                d2h += int(out["times"].nbytes + out["coeff"].nbytes + out["max_dev"].nbytes + out["iters"].nbytes +
                           out["vw_final"].nbytes + out["best_s"].nbytes)
            return dict(value=world * B * e2e_steps / sec, unit=UNIT, h2d_bytes_per_step=wp_np[0].nbytes, d2h_bytes_per_step=d2h,
                        steps=e2e_steps, ms_per_step=sec / e2e_steps * 1e3, host_threads=S, single_call_ms=single_ms,
                        d2h_GBps_per_gpu=d2h / (sec / e2e_steps) / 1e9,
                        host_cpus_pinned=None if numa_cpus is None else len(numa_cpus))

        # headline e2e = the reference call's own outputs: GenerateTrajectoryMatrix returns the sampled rows and nothing else
        # (ms.hpp:60-61; the caller copies exactly those, cpp:4464-4470).  The all-outputs figure (coefficients, times,
        # decisions: 1.7x the bytes) is kept beside it.
        e2e_main = e2e_measure("samples")
        e2e_main["outputs"] = ("the reference call's outputs: sampled rows [S,3] + CSR offsets + per-trajectory flags "
                               "(optional arrays passed as NULL)")
        e2e_all = e2e_measure("all")
        e2e_all["outputs"] = "every array of include/msnap.h: + coefficients, segment times, max_dev, iters, final weight, best_s"
        e2e_main["all_outputs"] = e2e_all
        # the same step with the sampled rows leaving as WGS84 [lon, lat, alt] (getPlan's
        # enuToWGS84_Batch, uavPathPlanning.cpp:3699, applied on the device by k_enu_to_wgs84 after the sampler)
        geo = None
        if headline:
            origin = np.array([109.56059880227296, 40.86719901015758, 0.0])   # readme.md:11, the uav31_0 origin
            for t_ in tools:
                t_.set_sample_frame("wgs84", origin)
            g_steps = min(steps, 500)
            g_ms, g_launches = timed_run(cfg, sets, g_steps, 3, S)
            tool.profile_begin()
            for i in range(prof_steps):
                step(cfg, sets[i % ROTATE])
            gprof = tool.profile_end()
            for t_ in tools:
                t_.set_sample_frame("enu")
            kname = next(k for k in gprof if k.startswith("k_enu_to_wgs84"))
            geo = dict(value=world * B * g_steps / (g_ms * 1e-3), unit=UNIT, ms_per_step=g_ms / g_steps,
                       launches_per_step=g_launches / g_steps, kernel=kname,
                       kernel_ms_per_step=gprof[kname]["total_ms"] / prof_steps, rows_per_step=tot_samples,
                       algorithmic_bytes_per_launch=48 * tot_samples)
        # SURVEY 8f ranks 1 + 2 together: getPlan's leader chain on the device (cpp:2640 -> 3684 -> 3712-3729 -> 3699):
        # WGS84 waypoints -> ENU -> minimum snap -> sampled rows -> cost-map lookup -> altitude optimisation -> WGS84 rows
        chain = None
        if headline:
            from cs_pathplan_b200 import shipped_altitude_params

            alt_p = shipped_altitude_params()
            gx = torch.arange(128, dtype=torch.float64, device=dev)
            grid = (-60.0 + 25.0 * torch.sin(gx[None, :] / 9.0) * torch.cos(gx[:, None] / 7.0)).to(torch.float32).contiguous()
            g_res, g_ox, g_oy = 8.0, -512.0, 512.0            # 128 x 128 cells of 8 m around the origin, terrain -85..-35 m
            N_FOLLOWERS = 6
            for s_ in sets:
                s_.lla_wp = torch.from_numpy(tool.enuToWGS84_Batch(s_.wp_h.numpy(), origin)).to(dev)
                s_.elev = torch.empty(s_.cap, dtype=torch.float64, device=dev)
                s_.lla = torch.empty((s_.cap, 3), dtype=torch.float64, device=dev)
                s_.solves = torch.empty(B, dtype=torch.int32, device=dev)
                s_.followers = torch.empty((N_FOLLOWERS * s_.cap, 3), dtype=torch.float64, device=dev)

            f_dist, f_rows = tool.formation_parameters()        # config.yaml defaults: 50 m, 8 rows per column
            d_starts = torch.from_numpy(np.column_stack([origin[0] + 1e-3 * np.arange(N_FOLLOWERS), np.full(N_FOLLOWERS, origin[1]),
                                                         np.zeros(N_FOLLOWERS)])).to(dev)

            def chain_step(cfg_, s_, which=0):
                t_ = tools[which]
                t_.generate_batch_dev(cfg_, s_.lla_wp, s_.off, s_.samples, ns=NS, times=s_.times, coeff=s_.coeff,
                                      max_dev=s_.max_dev, iters=s_.iters, vw_final=s_.vw, flags=s_.flags)
                t_.cost_map_lookup_dev(grid, g_res, g_ox, g_oy, s_.samples, s_.elev, n_rows=s_.off[B:])
                t_.altitude_optimize_batch_dev(alt_p, s_.off, s_.samples, s_.elev, solves=s_.solves)
                t_.followers_dev(s_.samples, s_.off, s_.followers, 1, f_dist, N_FOLLOWERS, f_rows, "wgs84", origin, d_starts)
                t_.enu_to_wgs84_dev(origin, s_.samples, s_.lla, n_rows=s_.off[B:])

            for t_ in tools:
                t_.set_waypoint_frame("wgs84", origin)
            step_main = step
            step = chain_step
            c_steps = min(steps, 300)
            c_ms, c_launches = timed_run(cfg, sets, c_steps, 3, S)
            tool.profile_begin()
            for i in range(prof_steps):
                chain_step(cfg, sets[i % ROTATE])
            cprof = tool.profile_end()
            step = step_main
            for t_ in tools:
                t_.set_waypoint_frame("enu")
            chain = dict(value=world * B * c_steps / (c_ms * 1e-3), unit=UNIT, ms_per_step=c_ms / c_steps,
                         launches_per_step=c_launches / c_steps, rows_per_step=int(sets[0].off[-1].item()),
                         mean_altitude_solves=float(sets[0].solves.double().mean().item()),
                         kernels_ms_per_step={k: v["total_ms"] / prof_steps for k, v in cprof.items()})
        results[weights] = dict(
            geo=geo, chain=chain, ms=ms, steps=steps, launches=launches, value=world * B * steps / (ms * 1e-3), latency_ms=latency_ms,
            e2e=e2e_main,
            prof=prof, dom=dom, dom_ms=dom_ms, all_kernels_ms=all_ms, abytes=abytes, aflops=aflops, kbytes=kbytes,
            kflops=kflops,
            samples=tot_samples, candidates=cand, mean_iters=float(iters0.double().mean().item()), flags_bad=flags_bad,
            footprint=footprint)
        del sets
        torch.cuda.empty_cache()

    # ------------------------------------------------------------------------------------------ the other configs
    # BASELINE.json configs[0], [2], [3], [4] on the same line ("variants").  cfg3 and cfg5 are ONE batch sharded by
    # trajectory index over the ranks (contiguous for cfg3, balanced on the segment count for cfg5; no collective on the
    # data path): strong scaling, value = whole batch / max-over-ranks time, per-rank times reported (load imbalance).
    from cs_pathplan_b200 import sharding

    L2_BYTES = 126e6
    flush_buf = torch.empty(64 << 20, dtype=torch.float32, device=dev)   # 256 MB > L2

    def gather_floats(x):
        if world == 1:
            return [float(x)]
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    def config_variant(wp_full, cfg, ns=None, seg_offset=None, steps=10, warmup=3, sdo=-1.0, vo=-1.0):
        wp_l, ns_l, so_l, (b0, b1) = sharding.shard_batch(wp_full, rank, world, ns=ns, seg_offset=seg_offset)
        Bl = b1 - b0
        n_seg_l = Bl * ns_l if ns_l else int(so_l[-1])
        m_ = 2 * cfg.order
        wp_l = np.ascontiguousarray(wp_l)
        d_wp = torch.from_numpy(wp_l).to(dev)
        d_so = None if so_l is None else torch.from_numpy(np.ascontiguousarray(so_l)).to(dev)
        cap = tool.sample_bound(cfg, wp_l, ns=ns_l, seg_offset=so_l, v_avg_override=vo)
        o = dict(times=torch.empty(n_seg_l, dtype=torch.float64, device=dev),
                 coeff=torch.empty(n_seg_l * 3 * m_, dtype=torch.float64, device=dev),
                 max_dev=torch.empty(Bl, dtype=torch.float64, device=dev),
                 vw_final=torch.empty(Bl, dtype=torch.float64, device=dev),
                 iters=torch.empty(Bl, dtype=torch.int32, device=dev),
                 flags=torch.empty(Bl, dtype=torch.int32, device=dev))
        off = torch.empty(Bl + 1, dtype=torch.int64, device=dev)
        samples = torch.empty((cap, 3), dtype=torch.float64, device=dev)

        def one():
            tool.generate_batch_dev(cfg, d_wp, off, samples, ns=ns_l, seg_offset=d_so, sample_distance_override=sdo,
                                    v_avg_override=vo, **o)

        for _ in range(warmup):
            one()
        torch.cuda.synchronize()
        rows_l = int(off[-1].item())
        abytes_l = roofline.algorithmic_bytes(Bl, n_seg_l, cfg.order, rows_l)
        small = abytes_l < 2 * L2_BYTES      # the batch would stay in L2 between steps: flush it, time step by step
        barrier()
        l0 = tool.launch_count
        if small:
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            with torch.cuda.stream(stream):
                for a, b in ev:
                    flush_buf.zero_()
                    a.record(stream)
                    one()
                    b.record(stream)
            torch.cuda.synchronize()
            ms_l = sum(a.elapsed_time(b) for a, b in ev) / steps
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                one()
            e1.record(stream)
            e1.synchronize()
            torch.cuda.synchronize()
            ms_l = e0.elapsed_time(e1) / steps
        launches = (tool.launch_count - l0) / steps
        per_rank = gather_floats(ms_l)
        ms = max(per_rank)
        barrier()
        tool.profile_begin()
        for _ in range(3):
            one()
        prof = tool.profile_end()
        torch.cuda.synchronize()
        rows = [int(v) for v in gather_floats(rows_l)]
        segs = [int(v) for v in gather_floats(n_seg_l)]
        trajs = [int(v) for v in gather_floats(Bl)]
        solves_l = float((o["iters"].to(torch.int64) + 1).sum().item())
        solves = sum(gather_floats(solves_l))
        flags_bad = int(sum(gather_floats(float((o["flags"] != 0).sum().item()))))
        base, total_rows = sharding.global_sample_base(rows_l)       # this rank's first row in the global CSR output
        B_tot, seg_tot = sum(trajs), sum(segs)
        t_np = o["times"].cpu().numpy()
        cand_l = float(np.sum(np.floor((t_np + 1e-12) / np.minimum(0.1, t_np / 10.0) + 1e-9)))
        cand = sum(gather_floats(cand_l))
        ab = roofline.algorithmic_bytes(B_tot, seg_tot, cfg.order, total_rows)
        af = roofline.algorithmic_flops(B_tot, seg_tot, cfg.order, cfg.path_weight > 0, solves, cand)
        return dict(value=B_tot / (ms * 1e-3), unit=UNIT, ms_per_step=ms, steps=steps,
                    scaling="strong: ONE batch sharded by trajectory index over the ranks, no collective",
                    per_rank_ms=per_rank, imbalance_worst_over_mean=ms / (sum(per_rank) / len(per_rank)),
                    per_rank_trajectories=trajs, per_rank_segments=segs, per_rank_rows=rows,
                    segments_per_s=seg_tot / (ms * 1e-3), rows_per_step=total_rows, launches_per_step=launches,
                    mean_reweight_iters=solves / B_tot - 1.0, flagged_trajectories=flags_bad,
                    l2_policy=("256 MB flush write between steps, steps timed one by one" if small else
                               f"inputs + outputs {abytes_l / 1e6:.0f} MB per GPU > 126 MB L2, steps back to back"),
                    algorithmic_bytes=ab, algorithmic_flops=af,
                    kernels_ms_per_step={k: v["total_ms"] / 3 for k, v in prof.items()} if rank == 0 else None)

    want = {"cfg1", "cfg3", "cfg4", "cfg5"} if args.configs == "all" else \
        (set() if args.configs == "none" else set(args.configs.split(",")))
    if world > 1:
        want &= {"cfg3", "cfg5"}            # only the configs BASELINE.json shards; the others are single-GPU lines
    extra = {}
    if "cfg3" in want:
        wp3, ns3 = workloads.cfg3()
        for w_ in ("shipped", "plain"):
            v = config_variant(wp3, workloads.synthetic_config(ORDER, w_), ns=ns3, steps=10 if w_ == "shipped" else 20)
            v["workload"] = (f"cfg3: ONE batch of {1 << 20} trajectories x 8 segments, order 4, fp64, rng 1235, {w_} weights, "
                             f"contiguous shards over {world} GPU(s)")
            extra["cfg3_sharded" if w_ == "shipped" else "cfg3_sharded_plain"] = v
        del wp3
    if "cfg5" in want:
        wp5, so5 = workloads.cfg5()
        for w_ in ("shipped", "plain"):
            v = config_variant(wp5, workloads.synthetic_config(ORDER, w_, sample_distance=0.0), seg_offset=so5,
                               steps=10 if w_ == "shipped" else 20)
            v["workload"] = (f"cfg5: ONE batch of 65536 trajectories, 2-256 segments (log-uniform, rng 1237), order 4, dense "
                             f"10 Hz output (sample_distance 0), {w_} weights, shards balanced on the segment count over "
                             f"{world} GPU(s)")
            extra["cfg5_sharded" if w_ == "shipped" else "cfg5_sharded_plain"] = v
        del wp5
    if "cfg4" in want:
        wp4, ns4 = workloads.cfg4()
        for w_ in ("shipped", "plain"):
            v = config_variant(wp4, workloads.synthetic_config(ORDER, w_), ns=ns4, steps=20)
            v["workload"] = f"cfg4: 1024 boustrophedon patrols x 512 segments, order 4, fp64, rng 1236, {w_} weights"
            extra["cfg4" if w_ == "shipped" else "cfg4_plain"] = v
        del wp4
    if "cfg1" in want:
        wp1, cfg1, sdo1, vo1 = workloads.cfg1()
        lat = []
        for _ in range(25):
            t0 = time.perf_counter()
            s1 = tool.GenerateTrajectoryMatrix(wp1, cfg1, sdo1, vo1)
            lat.append((time.perf_counter() - t0) * 1e3)
        rep = 4096
        v = config_variant(np.tile(wp1, (rep, 1)), cfg1, ns=wp1.shape[0] - 1, steps=50, sdo=sdo1, vo=vo1)
        v["workload"] = ("cfg1: uav31_0, the 7 ENU leader waypoints of readme.md:14-20, shipped YAML (order 2), getPlan's "
                         f"overrides (sample distance 300 m, leader_speed 30 m/s); value = the same trajectory replicated {rep}x")
        v["single_call_ms"] = statistics.median(lat[5:])
        v["single_call"] = ("B = 1 through the drop-in call TrajectoryGeneratorTool::GenerateTrajectoryMatrix with host "
                            "buffers (sample bound + H2D + kernels + D2H), wall clock, median of 20 calls")
        v["single_call_samples"] = int(s1.shape[0])
        # The same B = 1 call with device-resident buffers: launched kernel by kernel, and as ONE CUDA-graph replay (the
        # *_dev entry points only enqueue work, tests/test_gpu_graph.py): what a caller that plans every tick would do.
        try:
            ns1 = wp1.shape[0] - 1
            d_wp1 = torch.from_numpy(wp1).to(dev)
            cap1 = tool.sample_bound(cfg1, wp1, ns=ns1, v_avg_override=vo1)
            off1 = torch.zeros(2, dtype=torch.int64, device=dev)
            rows1 = torch.zeros((cap1, 3), dtype=torch.float64, device=dev)

            def one1():
                tool.generate_batch_dev(cfg1, d_wp1, off1, rows1, ns=ns1, sample_distance_override=sdo1, v_avg_override=vo1)

            def timed1(fn, n=200):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(stream):
                    for _ in range(5):
                        fn()
                    e0.record(stream)
                    for _ in range(n):
                        fn()
                    e1.record(stream)
                e1.synchronize()
                return e0.elapsed_time(e1) / n

            eager_ms = timed1(one1)
            l0 = tool.launch_count
            one1()
            n_launch1 = tool.launch_count - l0
            torch.cuda.synchronize()
            g1 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1, stream=stream):
                one1()
            graph_ms = timed1(g1.replay)
            torch.cuda.synchronize()
            v["dev_call_ms"] = {"eager": eager_ms, "cuda_graph_replay": graph_ms, "kernels_per_call": n_launch1,
                                "samples": int(off1[1].item()),
                                "note": "B = 1, device-resident waypoints and rows, back-to-back calls on one stream, CUDA "
                                        "events; replay = the call captured once into a CUDA graph"}
            del g1
        except Exception as exc:  # never lose the line over the extra figure
            v["dev_call_ms"] = {"error": repr(exc)[:200]}
        extra["cfg1"] = v
    del flush_buf
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, peak_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured)") if "hbm_gbs" in peaks else \
            (6650.0, "B200_PROFILING.md fallback")
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")))
        except Exception:
            pass

        def roof(r, weights):
            # the dominant kernel against ITS OWN algorithmic bytes / flops per launch (one launch per step)
            t = r["dom_ms"] * 1e-3
            kb, kf = r["kbytes"][r["dom"]], r["kflops"][r["dom"]]
            gbs = kb / t / 1e9
            tf = kf / t / 1e12
            tr = traffic.get(weights, {}).get(r["dom"]) if isinstance(traffic, dict) else None
            t_all = r["all_kernels_ms"] * 1e-3
            # the binding roofline = whichever of (bytes / HBM peak, flops / FP64 peak) gives the larger time floor
            hbm = {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "peak_source": peak_src}
            fp64 = {"achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak,
                    "peak_source": "msnap_measure_fp64_peak (own DFMA micro-benchmark, measured this run)"}
            bound = "fp64" if kf / (fp64_peak * 1e12) > kb / (hbm_peak * 1e9) else "hbm"
            bind, other = (fp64, hbm) if bound == "fp64" else (hbm, fp64)
            return {"bound": bound, "kernel": r["dom"], "achieved": bind["achieved"], "peak": bind["peak"],
                    "unit": bind["unit"], "frac": bind["frac"], "traffic": tr, "peak_source": bind["peak_source"],
                    "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture "
                                      "(profiles/dominant_kernel_traffic.json)" if tr else None,
                    "non_binding": dict(other, bound="hbm" if bound == "fp64" else "fp64"),
                    "kernel_ms_per_step": r["dom_ms"], "all_kernels_ms_per_step": r["all_kernels_ms"],
                    "algorithmic_bytes_per_launch": kb, "algorithmic_flops_per_launch": kf,
                    "whole_step": {"algorithmic_bytes": r["abytes"], "algorithmic_flops": r["aflops"],
                                   "hbm_frac": r["abytes"] / t_all / 1e9 / hbm_peak,
                                   "fp64_frac": r["aflops"] / t_all / 1e12 / fp64_peak},
                    "kernels": {k: {"ms_per_step": v["total_ms"] / 20,
                                    "hbm_frac": r["kbytes"][k] / (v["total_ms"] / 20 * 1e-3) / 1e9 / hbm_peak,
                                    "fp64_frac": r["kflops"][k] / (v["total_ms"] / 20 * 1e-3) / 1e12 / fp64_peak}
                                for k, v in r["prof"].items() if k in r["kbytes"]},
                    "kernels_ms_per_step": {k: v["total_ms"] / 20 for k, v in r["prof"].items()}}

        h = results[args.weights]
        line = {
            "metric": METRIC, "value": h["value"], "unit": UNIT, "n_gpus": world, "steps": h["steps"],
            "warmup": max(args.warmup, 3), "ms_per_step": h["ms"] / h["steps"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.weights, B), "trajectories_per_gpu_per_step": B,
                       "l2_policy": f"inputs and outputs rotate over {ROTATE} distinct buffer sets "
                                    f"({h['footprint'] / 1e6:.0f} MB > 126 MB L2); workspace is reused",
                       "streams_per_gpu": S, "single_stream_ms_per_step": h["latency_ms"],
                       "pipelining": f"consecutive steps (independent batches) alternate between {S} solver handles / CUDA "
                                     f"streams per GPU; ms_per_step = timed region / steps",
                       "samples_per_step": h["samples"], "mean_reweight_iters": h["mean_iters"],
                       "flagged_trajectories": h["flags_bad"], "parallelism": f"trajectory-sharded x{world}, no collective"},
            "e2e": h["e2e"], "gpu_launches": h["launches"], "clocks": clocks, "roofline": roof(h, args.weights),
            "variants": {w: {"value": r["value"], "ms_per_step": r["ms"] / r["steps"],
                             "single_stream_ms_per_step": r["latency_ms"], "e2e": r["e2e"],
                             "roofline": roof(r, w), "workload": workload_name(w, B)}
                         for w, r in results.items() if w != args.weights},
        }
        for k_, v_ in extra.items():
            t_ = v_["ms_per_step"] * 1e-3
            v_["roofline_whole_step"] = {"hbm_frac": v_["algorithmic_bytes"] / t_ / 1e9 / (hbm_peak * world),
                                         "fp64_frac": v_["algorithmic_flops"] / t_ / 1e12 / (fp64_peak * world),
                                         "hbm_peak_GBps": hbm_peak, "fp64_peak_TFLOPs": fp64_peak}
            line["variants"][k_] = v_
        if h.get("geo"):
            g = dict(h["geo"])
            gbs = g["algorithmic_bytes_per_launch"] / (g["kernel_ms_per_step"] * 1e-3) / 1e9
            g.update(note="same workload and driver as `value`, sampled rows converted to WGS84 lon/lat/alt on the device "
                          "(msnap_set_sample_frame); SURVEY.md section 8f rank 1",
                     kernel_hbm_GBps=gbs, kernel_hbm_frac=gbs / hbm_peak,
                     kernel_rows_per_s=g["rows_per_step"] / (g["kernel_ms_per_step"] * 1e-3))
            line["wgs84_frame"] = g
        if h.get("chain"):
            c = dict(h["chain"])
            c["note"] = ("getPlan's leader chain per batch, device resident, same multi-stream driver as `value`: WGS84 waypoints -> "
                         "ENU (k_wgs84_to_enu) -> minimum snap + sampler -> cost-map lookup -> altitude optimisation (shipped "
                         "config.yaml parameters, synthetic 128 x 128 terrain grid) -> 6 followers in V formation as WGS84 rows "
                         "-> leader WGS84 rows; SURVEY.md section 8f ranks 1-3")
            line["leader_chain"] = c
        if world == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, all_cpus)     # the CPU legs use every core of the box again
            try:
                from oracle import ref

                wp, _ = workloads.cfg2(B=B)
                if ref.available("fast"):
                    line["cpu_baseline"] = cpu_baseline(args.weights, wp)
                else:
                    line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                            "sample": "oracle/_ref not built on this box"}
            except Exception as e:  # the baseline must never take the GPU number down with it
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": repr(e)}
            # Whole-batch parity, outside every timed region: ALL trajectories of the timed batch (workloads.cfg2(), the
            # headline weights) through the unmodified reference (oracle/_ref, parity build) -- sample counts, every row,
            # the reweighting decisions and the coefficients (oracle/parity.py).
            if not args.no_parity:
                try:
                    from oracle import parity

                    cfg_h = workloads.synthetic_config(ORDER, args.weights)
                    res_h = tool.generate_batch(cfg_h, wp, ns=NS)
                    pr = parity.batch_parity(res_h, wp, np.arange(B + 1, dtype=np.int64) * NS, cfg_h)
                    pr["against"] = ("oracle/_ref/libmsnap_ref.so = unmodified minimum_snap.cpp, -O2 without FMA "
                                     "contraction, GenerateTrajectoryMatrix + the reweighting loop per trajectory; bars: "
                                     "counts equal, rows <= 1e-6 m, coefficients <= 1e-8 (position-scaled), decisions "
                                     "equal; max_row_err_m / max_coeff_err are over the trajectories within the bars of "
                                     "the reference, `reference_unsound` = the rest (the reference's dense inverse of M "
                                     "loses up to 12 digits there), checked against 40-digit arithmetic (oracle/parity.py)")
                    pr["ok"] = bool(pr["unexplained"] == 0 and pr["count_mismatch"] == 0 and
                                    pr.get("iters_mismatch", 0) == 0 and pr.get("time_mismatch", 0) == 0)
                    line["parity"] = pr
                except Exception as e:
                    line["parity"] = {"checked": 0, "error": repr(e)}
        print(json.dumps(line))
    for t_ in tools:
        t_.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "rows-cpu":
        run_rows_cpu(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
