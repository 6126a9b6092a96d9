#!/usr/bin/env python
"""bench.py — particle-steps/s of physim's `astro ! verlet` step on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W                  (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W (CPU oracle = the reference arm)

Workload (config.workload = "c3"): BASELINE.json configs[2] — cube n=1,000,000 seed=1 spin=500 + 4
stars (m=1e5) ! astro theta=1.3 (e default 1.0) ! verlet, dt = 1e-6 (pipeline default), synthetic
bodies from physim_b200.generators (the reference's distributions, our seeds).

  value      device-resident steps (state in HBM), CUDA events on the launching stream, max over ranks
  e2e        the same step through the C ABI with HOST Entity buffers (pb200_verlet_step_fused):
             pack -> H2D -> tree/force/verlet kernels -> D2H -> unpack, wall clock, every step
  roofline   the kernel with the largest share of the device step, timed live with CUDA event
             pairs around each of its launches; algorithmic bytes from DESIGN.md's table
  cpu_baseline  the CPU oracle (fp64 restatement of the reference, 1 thread like the reference's
             simulation thread) on a bounded number of steps of the same workload
  direct_sum tiled all-pairs kernel on 2^24 bodies (BASELINE configs[3]) sampled over a target
             slice: interactions/s and fraction of the FP32 peak (19 flop per interaction)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# stdout carries exactly ONE JSON line: everything else written to fd 1 by this process or by native libraries
# (NCCL prints its version banner there) is sent to stderr
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


WORKLOADS = {
    # name: (generator args, element, theta, e, dt)
    "c3": dict(n=1_000_000, stars=4, element="astro", theta=1.3, e=1.0, dt=1e-6,
               desc="cube n=1000000 seed=1 spin=500 + 4 stars ! astro theta=1.3 ! verlet"),
    "c1": dict(n=100_000, stars=2, element="astro2", theta=1.5, e=0.5, dt=1e-5,
               desc="cube n=100000 seed=1 spin=1000 + 2 stars ! astro2 theta=1.5 e=0.5 ! verlet"),
    "c5s": dict(n=4_194_304, element="astro2", theta=0.7, e=0.5, dt=1e-6,
                desc="cube n=4194304 seed=1 ! astro2 theta=0.7 e=0.5 ! verlet (configs[4] at 1/16 size)"),
    "c3o": dict(n=1_000_000, stars=4, element="astro2", theta=1.5, e=0.5, dt=1e-5,
                desc="cube n=1000000 seed=1 spin=500 + 4 stars ! astro2 theta=1.5 e=0.5 ! verlet"),
}

# Algorithmic HBM bytes per body per launch, SURVEY.md §8(d)'s rows (DESIGN.md §4 maps kernels to rows):
#   extent 24 R | encode 24 R + 8 + 4 W | sort pass 12 R + 12 W | gather 4 + 32 R + 16 W |
#   cell build (units, scan, cell table, centres of mass) 76 | accelerations out 16 R + 12 W | verlet 120-132
# The cell-build row is charged IN FULL to cells_kernel (the kernel `roofline` names at c3) although
# unit / scan / kids / climb share it: the fraction reported for cells_kernel is therefore an upper bound
# on its own share, and `roofline_step` below carries the whole 0.55 kB/body against the whole step.
ALG_BYTES_PER_BODY = {
    "extent_kernel": 24, "encode_kernel": 36, "encode_bucket_kernel": 36, "sort_onesweep_pass": 24,
    "sort_local_kernel": 24 + 52, "gather_kernel": 52, "cells_kernel": 76, "walk_kernel": 28,
    "verlet_kernel": 132, "verlet_lean_kernel": 32 + 32 + 16 + 32, "verlet_velocity_kernel": 32 + 32 + 32,
    "to_soa_kernel": 32 + 24, "bbox_kernel": 24,
    # inside the cell-build row (no separate figure in §8(d)); compulsory bytes of each, for the per-kernel list only
    "unit_kernel": 8 + 32 + 2 + 4 + 1.1, "scan_lookback_kernel": 8, "kids_kernel": 1.5 * 9, "climb_kernel": 0.09 * 188,
    # sharded step (N > 1): HBM bytes on the rank that runs the kernel; the peer stores travel over NVLink
    "walk_sharded_kernel": 32 + 20, "shard_scatter_kernel": 20 + 32, "top_export_kernel": 0.2, "top_build_kernel": 0.4,
}
# kernels of the sharded step that run over ALL bodies on every rank (the others see one rank's share)
REPLICATED_KERNELS = {"encode_bucket_kernel", "shard_scatter_kernel", "verlet_lean_kernel", "verlet_kernel", "extent_kernel"}
STEP_BYTES_PER_BODY = 550  # "Sum ~ 0.55 kB/body-step"
FLOP_PER_INTERACTION = 19
NCU_KERNELS = "r02_ncu_c3_kernels.json"  # per-kernel DRAM traffic of the committed ncu --set full capture


def make_state(w):
    from physim_b200 import generators as gen
    if w.get("stars") == 4:
        return gen.headline_pipeline(w["n"], seed=1, spin=500.0)
    if w.get("stars") == 2:
        return gen.readme_pipeline(w["n"], seed=1, spin=1000.0)
    return gen.cube(w["n"], seed=1)


def workload(name, world=1, scaling="weak"):
    """The named workload; under weak scaling (N > 1) the SAME simulation grows to N x the body count, one
    sharded simulation over the N GPUs (BASELINE's own multi-GPU configs are larger problems on more GPUs)."""
    w = dict(WORKLOADS[name])
    if world > 1 and scaling == "weak":
        w["n"] = w["n"] * world
        w["desc"] = w["desc"].replace("n=%d" % WORKLOADS[name]["n"], "n=%d" % w["n"]) + \
            f" [weak scaling: {WORKLOADS[name]['n']} bodies per GPU x {world} GPUs, one simulation]"
    return w


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured", d
        except Exception:
            pass
    return 6650.0, "fallback", {}


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, sampled in-process through NVML
    (same counters as the nvidia-smi line of B200_PROFILING.md, without its start-up latency)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, gpu_index=0, period=0.005):
        self.gpu, self.period = gpu_index, period
        self.samples, self.max_mhz, self.err = [], None, None
        self.window = None  # (t0, t1) of the timed region, perf_counter clock
        self._stop = threading.Event()
        self.thread = None

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self._stop.is_set():
                mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                bits = int(get_reasons(h))
                self.samples.append((time.perf_counter(), mhz, bits))
                time.sleep(self.period)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self._stop.set()
        if self.thread:
            self.thread.join(timeout=2)
        rows = self.samples
        if self.window:  # samples inside the timed region (one NVML round trip is ~10-20 ms)
            inside = [r for r in rows if self.window[0] <= r[0] <= self.window[1]]
            # a region shorter than the sampling period: the samples bracketing it
            rows = inside or sorted(rows, key=lambda r: min(abs(r[0] - self.window[0]), abs(r[0] - self.window[1])))[:2]
        sm = [r[1] for r in rows]
        mask = 0
        for r in rows:
            mask |= r[2]
        reasons = [n for bit, n in self.REASONS.items() if mask & bit]
        out = {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz,
               "samples": len(sm), "samples_total": len(self.samples), "reasons": reasons}
        if self.err:
            out["error"] = self.err
        return out


def cpu_info():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return model, os.cpu_count()


def run_reference(args, w):
    """The reference arm: the CPU oracle (fp64 restatement of the Rust reference; rustc is not
    available so the reference itself cannot be built) on the host cores.  One thread: the
    reference's whole step runs on its single simulation thread (pipeline.rs:134)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding as ob
    ob.build()
    state = make_state(w)
    n = len(state)
    cur = state
    # bounded sample: about 90 s of CPU work (a 1 M-body step takes 0.6-0.9 s on one core), never more than asked
    steps_run = max(2, min(args.steps, int(100e6 // max(n, 1)) + 1))
    warm_run = min(args.warmup, 2)
    for _ in range(warm_run):
        cur, _ = ob.run_pipeline(w["element"], cur, w["theta"], w["e"], w["dt"], 1)
    t0 = time.perf_counter()
    cur, secs = ob.run_pipeline(w["element"], cur, w["theta"], w["e"], w["dt"], steps_run)
    dt = time.perf_counter() - t0
    value = n * steps_run / dt
    model, cores = cpu_info()
    line = {
        "impl": "reference", "metric": "particle-steps/s", "value": value, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": steps_run, "steps_requested": args.steps, "warmup": warm_run,
        "ms_per_step": dt / steps_run * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "description": w["desc"], "n_bodies": n,
                   "element": w["element"], "theta": w["theta"], "e": w["e"], "dt": w["dt"]},
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": 1, "kind": "port",
                         "sample": f"{steps_run} full steps of the {n}-body workload after {warm_run} warm-up "
                                   f"(bounded: ~90 s of CPU work; value is per step)",
                         "phases_s_per_step": {"build": secs[0] / steps_run, "walk_force": secs[1] / steps_run,
                                               "integrate_copies": secs[2] / steps_run},
                         "cpu": model, "host_cores": cores},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_ours(args, w):
    import torch
    from physim_b200 import api
    import physim_b200._build as b

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        b.build()
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    if api.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: physim_b200 has no CPU fallback")
    api.lib().pb200_set_device(local_rank)
    torch.cuda.set_device(local_rank)

    state = make_state(w)
    n = len(state)
    hbm_peak, peak_kind, peaks = measured_peaks()

    # ---------------- device-resident steps (value) ----------------
    # one GPU: pb200_sim_*; several: pb200_msim_* (csrc/multi.cu) - one rank per process here, the NCCL id made
    # by rank 0 and handed round with a torch.distributed broadcast (plumbing); every collective of the step is
    # issued by the library itself on its own stream
    stream = torch.cuda.Stream(device=local_rank)
    if world == 1:
        sim = api.Sim(w["element"], theta=w["theta"], e=w["e"], dt=w["dt"])
        sim.set_stream(stream.cuda_stream)
    else:
        sim = make_msim(api, torch, dist, w, rank, world, local_rank)
    sim.upload(state)

    def steps(k):
        sim.run(k)  # k steps enqueued back to back; one host sync at the end

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    steps(max(args.warmup, 3))
    launches0 = sim.stats()["kernel_launches"]
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
        torch.cuda.synchronize()
    t_region0 = time.perf_counter()
    if world == 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        steps(args.steps)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    else:
        ms = sim.run_timed(args.steps)  # CUDA events on the library's stream of this rank; max over ranks below
        torch.cuda.synchronize()
    sampler.window = (t_region0, time.perf_counter())
    if dist:
        dist.barrier()
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if dist:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    st = sim.stats()
    launches = st["kernel_launches"] - launches0 - (1 if world == 1 else 0)  # the stats call itself counts interactions
    value = n * args.steps / (ms * 1e-3)

    # ---------------- per-kernel device times (roofline) ----------------
    sim.profile(True)
    prof_steps = 5
    steps(prof_steps)
    torch.cuda.synchronize()
    report = sim.profile_report()
    sim.profile(False)
    tot_ms = sum(k["ms"] for k in report) or 1.0
    top = max(report, key=lambda k: k["ms"])
    per_launch_ms = top["ms"] / top["launches"]
    def bodies_of(kernel):  # bodies one launch of `kernel` on one rank handles
        return n if (world == 1 or kernel in REPLICATED_KERNELS) else n // world

    alg_bytes = ALG_BYTES_PER_BODY.get(top["kernel"], 0) * bodies_of(top["kernel"])
    achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9 if alg_bytes else None
    traffic = None  # dram read + write bytes per launch, from the committed ncu --set full capture
    try:
        if args.workload == "c3":
            nc = json.load(open(os.path.join(ROOT, "profiles", NCU_KERNELS)))[top["kernel"]]
            traffic = (nc["dram_read_mb_per_launch"] + nc["dram_write_mb_per_launch"]) * 1e6
    except Exception:
        pass
    roofline = {"kernel": top["kernel"], "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                "unit": "GB/s", "frac": (achieved / hbm_peak) if achieved else None, "traffic": traffic,
                "traffic_source": f"profiles/{NCU_KERNELS} (ncu --set full, cold L2)" if traffic else None,
                "peak_source": peak_kind, "avg_launch_ms": per_launch_ms,
                "share_of_step": top["ms"] / tot_ms,
                "alg_bytes_per_launch": alg_bytes,
                "alg_bytes_source": "SURVEY.md §8(d) row for this kernel x the bodies one launch handles on one rank "
                                    "(cells_kernel: the whole 76 B/body cell-build row)",
                "kernels": [{"kernel": k["kernel"], "launches_per_step": k["launches"] / prof_steps,
                             "ms_per_step": k["ms"] / prof_steps,
                             "gbs": (ALG_BYTES_PER_BODY.get(k["kernel"], 0) * bodies_of(k["kernel"]) / (k["ms"] / k["launches"] * 1e-3) / 1e9)
                             if k["ms"] > 0 else None}
                            for k in sorted(report, key=lambda k: -k["ms"])]}

    line = {
        "metric": "particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "description": w["desc"], "n_bodies": n,
                   "element": w["element"], "theta": w["theta"], "e": w["e"], "dt": w["dt"],
                   "n_cells": st["n_cells"], "interactions_per_step": st["interactions"],
                   "parallelism": ("one simulation over %d GPUs: key encoding sharded by body index, tree build + walk "
                                   "sharded by Morton key range (cuts at level-K cells, rebalanced every step); keys, "
                                   "level-K cell records and accelerations stored into the peers' HBM from inside the "
                                   "kernels over NVLink (epoch flags, no collective call per step), remote cells read "
                                   "through peer pointers in the walk, integrator state replicated" % world)
                   if world > 1 else "single GPU",
                   "bodies_per_gpu": n // world,
                   "l2": "no flush: per-step working set (~0.4 GB) exceeds the 126 MB L2; steps run "
                         "back to back as in the simulation loop",
                   "precision": "keys/tree/acceptance/integrator fp64, force law fp32"},
        "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks,
        # the whole step against HBM: SURVEY §8(d)'s 0.55 kB/body over the measured step time
        "roofline_step": {"bound": "hbm", "achieved": STEP_BYTES_PER_BODY * n / (ms / args.steps * 1e-3) / 1e9,
                          "peak": hbm_peak * world, "unit": "GB/s",
                          "frac": STEP_BYTES_PER_BODY * n / (ms / args.steps * 1e-3) / 1e9 / (hbm_peak * world),
                          "alg_bytes_per_body": STEP_BYTES_PER_BODY, "n_gpus": world,
                          "peak_note": "aggregate over the GPUs of the run"},
    }

    if rank == 0 and world == 1 and not args.skip_extras:
        line["e2e"] = measure_e2e(api, state, w, args)
        line["direct_sum"] = measure_direct(api, args, hbm_peak)
        line["cpu_baseline"] = measure_cpu(state, w)
        line["integrators"] = measure_integrators(api, state, w)
        line["other_workloads"] = measure_other_workloads(api, args)
    if world > 1:
        bodies, cells = sim.rank_counts()
        line["sharding"] = {"sharded_steps": st.get("sharded_steps"), "replicated_steps": st.get("replicated_steps"),
                            "replays": st["replays"],
                            "bodies_per_rank": None if bodies is None else [int(b) for b in bodies],
                            "cells_per_rank": None if cells is None else [int(c) for c in cells]}
    if world > 1 and not args.skip_extras:
        ds = measure_direct_multi(api, rank, world, local_rank, dist, torch)
        if rank == 0:
            line["direct_sum"] = ds
    if world > 1:
        e2e_multi = measure_e2e_multi(sim, state, n, rank, dist, torch)
        if rank == 0:
            line["e2e"] = e2e_multi
    if (world == 1 or world >= 8) and not args.skip_extras:
        if world > 1:
            sim.close()
        sim = state = None  # (frees the 1-GPU handle: its buffers go with the last reference)
        bl = measure_bh_large(api, torch, dist, rank, world, local_rank)
        if rank == 0:
            line["bh_large"] = bl
    if rank == 0:
        emit(line)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def measure_e2e(api, state, w, args):
    """The calls a physim pipeline makes, with HOST Entity buffers (two persistent arrays, as pipeline.rs:100-173):
      e2e           fused verlet step through the C ABI, the state copied in AND out every step
      resident      the same entry point with the shim's `resident` opt-in: input checked on a sample against the
                    previous output and not re-uploaded, whole Entity records DMA'd into the page-locked new_state
      dropin        the composition stock physim runs: pb200_integrator_step whose acc_fn calls the plugin's
                    `{element}_get_api()->transform` (two uploads, two downloads per step)"""
    n = len(state)
    steps = args.steps

    def loop(step_fn, k):
        bufs = [state.copy(), state.copy()]
        cur = 0
        for _ in range(max(args.warmup, 3)):
            step_fn(bufs[cur], bufs[cur ^ 1])
            cur ^= 1
        t0 = time.perf_counter()
        for _ in range(k):
            step_fn(bufs[cur], bufs[cur ^ 1])
            cur ^= 1
        return (time.perf_counter() - t0) / k, bufs[cur]

    el = api.TransformElement(w["element"], theta=w["theta"], e=w["e"])
    v = api.Verlet()
    dt, cur = loop(lambda a, b: v.integrate_fused(a, el, w["dt"], out=b), steps)
    vs = v.stats()
    vr = api.Verlet()
    vr.set_resident(True)
    # one persistent new_state, as physim's pipeline has (pipeline.rs:100-103); the state handed in is what the
    # previous call wrote (physim passes a clone of it - the clone is the pipeline's own cost, outside this API)
    res_buf = state.copy()
    for _ in range(max(args.warmup, 3)):
        vr.integrate_fused(res_buf, el, w["dt"], out=res_buf)
    t0 = time.perf_counter()
    for _ in range(steps):
        vr.integrate_fused(res_buf, el, w["dt"], out=res_buf)
    dt_res = (time.perf_counter() - t0) / steps
    rs = vr.stats()
    hits, misses = vr.resident_counts()
    vd = api.Verlet()
    dt_drop, _ = loop(lambda a, b: vd.integrate_dropin(a, el, w["dt"], out=b), max(5, steps // 5))
    # transform-only boundary (astro*_transform through `*_get_api`)
    acc = el.transform(cur)
    t1 = time.perf_counter()
    for _ in range(5):
        acc = el.transform(cur, acc)
    dt_tr = (time.perf_counter() - t1) / 5
    return {"value": n / dt, "unit": "particle-steps/s", "ms_per_step": dt * 1e3,
            "h2d_bytes_per_step": n * 33, "d2h_bytes_per_step": n * 24,
            "api": "pb200_verlet_step_fused(host Entity[n] -> host Entity[n]), state copied in and out every step "
                   "(H2D {x,y,z,m,fixed}; D2H {x,y,z}: the regular step's velocity (x' - x)/dt is taken on the host from "
                   "the caller's own input, verlet.rs:68-70, same bits)",
            "device_ms": {"h2d": vs["ms_h2d"], "force": vs["ms_force"], "integrate": vs["ms_integrate"],
                          "d2h": vs["ms_d2h"]},
            "host_ms": {"pack": vs["ms_host_pack"], "unpack": vs["ms_host_unpack"], "call": vs["ms_wall"]},
            "resident": {"value": n / dt_res, "ms_per_step": dt_res * 1e3, "h2d_bytes_per_step": 0,
                         "d2h_bytes_per_step": n * 80, "steps_without_upload": hits, "steps_with_upload": misses,
                         "api": "pb200_verlet_step_fused after pb200_verlet_set_resident(v, 1): input verified on a "
                                "2048-entity sample against the previous output, whole Entity records DMA'd into the "
                                "caller's page-locked new_state",
                         "device_ms": {"h2d": rs["ms_h2d"], "force": rs["ms_force"], "integrate": rs["ms_integrate"],
                                       "d2h": rs["ms_d2h"]}},
            "dropin": {"value": n / dt_drop, "ms_per_step": dt_drop * 1e3,
                       "h2d_bytes_per_step": n * (33 + 56 + 24), "d2h_bytes_per_step": n * (16 + 24),
                       "api": f"pb200_integrator_step(acc_fn = {w['element']}_get_api()->transform): what the Rust "
                              "shim's verlet runs when gravity is a separate plugin element"},
            "transform_only": {"api": f"{w['element']}_get_api()->transform", "ms_per_call": dt_tr * 1e3,
                               "h2d_bytes": n * 33, "d2h_bytes": n * 16}}


def measure_direct(api, args, hbm_peak):
    """BASELINE configs[3]: 2^24 bodies, astro2 theta=0 (== all pairs) e=0.5; a 1/64 target slice."""
    from physim_b200 import generators as gen
    n = 1 << 24
    state = gen.cube(n, seed=1)
    n_t = 2 * 296 * 1024  # 2 full waves of the 4-targets-per-thread CTAs (2 CTAs x 148 SMs per wave)
    sim = api.Sim("astro2", theta=0.0, e=0.5, dt=1e-6)
    sim.upload(state)
    sim.set_targets(0, n_t)
    sim.run_timed(1)
    steps = 2
    sampler = ClockSampler(0, period=0.02)
    sampler.start()
    ms = sim.run_timed(steps)
    clocks = sampler.stop()
    inter = n_t * n * steps
    rate = inter / (ms * 1e-3)
    fp32_probe = api.probe_fp32_tflops()
    import torch
    props = torch.cuda.get_device_properties(0)
    sms = props.multi_processor_count
    nominal = sms * 128 * 2 * 1.965e9 / 1e12
    tf = rate * FLOP_PER_INTERACTION / 1e12
    return {"workload": "cube n=16777216 ! astro2 theta=0 e=0.5 (all pairs)",
            "sample": f"targets [0, {n_t}) x all {n} sources, {steps} evaluations",
            "interactions_per_s": rate, "ms_per_evaluation": ms / steps,
            "full_step_s_extrapolated": n * n / rate, "clocks": clocks,
            "roofline": {"bound": "fp32", "achieved": tf, "unit": "TFLOP/s", "flop_per_interaction": 19,
                         "peak": nominal, "peak_source": f"{sms} SMs x 128 lanes x 2 x 1.965 GHz (max boost)",
                         "frac": tf / nominal, "ffma_probe_tflops": fp32_probe,
                         "frac_of_probe": tf / fp32_probe if fp32_probe > 0 else None}}


def make_msim(api, torch, dist, w, rank, world, local_rank):
    """pb200_msim_* handle of this process's rank; the communicator id travels by a torch.distributed broadcast."""
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.from_numpy(api.comm_unique_id().copy()).cuda()
    dist.broadcast(idt, 0)
    return api.MultiSim(w["element"], theta=w["theta"], e=w["e"], dt=w["dt"], world=world, rank=rank,
                        device=local_rank, comm_id=idt.cpu().numpy())


def measure_direct_multi(api, rank, world, local_rank, dist, torch):
    """BASELINE configs[3]: 2^24 bodies, astro2 theta=0 (all pairs), targets sharded by body index over the
    ranks, the accelerations all-gathered inside the step (csrc/multi.cu step_direct).  With 8 ranks ONE FULL
    STEP is run and timed (every target x every source, gather and verlet included, ~15 s); with fewer ranks a
    1/8 problem slice would take minutes, so 2^21 targets per rank are timed instead and the step extrapolated."""
    from physim_b200 import generators as gen
    n = 1 << 24
    full = world >= 8
    state = gen.cube(n, seed=1)
    w = dict(element="astro2", theta=0.0, e=0.5, dt=1e-6)
    if full:
        ms_ = make_msim(api, torch, dist, w, rank, world, local_rank)
        ms_.upload(state)
        t_ms = ms_.run_timed(1)        # first step (first-step verlet formula)
        t_ms = ms_.run_timed(1)
        t = torch.tensor([t_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms = float(t.item())
        rate = n * n / (t_ms * 1e-3)
        sample = f"one full step: {n} targets x {n} sources over {world} ranks, all-gather of accelerations and verlet inside"
        ms_.close()
    else:
        sim = api.Sim("astro2", theta=0.0, e=0.5, dt=1e-6, device=local_rank)
        sim.upload(state)
        per = n // world
        n_t = min(2 * 296 * 1024, per)
        sim.set_targets(rank * per, rank * per + n_t)
        sim.run_timed(1)
        torch.cuda.synchronize()
        dist.barrier()
        t_ms = sim.run_timed(2) / 2
        t = torch.tensor([t_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms = float(t.item())
        rate = world * n_t * n / (t_ms * 1e-3)
        sample = f"per rank: {n_t} of its {per} targets x all {n} sources, 2 evaluations; max over ranks (extrapolated)"
    props = torch.cuda.get_device_properties(local_rank)
    nominal = world * props.multi_processor_count * 128 * 2 * 1.965e9 / 1e12
    tf = rate * FLOP_PER_INTERACTION / 1e12
    return {"workload": "cube n=16777216 ! astro2 theta=0 e=0.5 ! verlet (all pairs), targets sharded over %d GPUs" % world,
            "sample": sample, "full_step_measured": full, "interactions_per_s": rate,
            "step_s": n * n / rate,
            "roofline": {"bound": "fp32", "achieved": tf, "unit": "TFLOP/s", "flop_per_interaction": 19,
                         "peak": nominal, "peak_source": f"{world} x {props.multi_processor_count} SMs x 128 x 2 x 1.965 GHz",
                         "frac": tf / nominal}}


def measure_e2e_multi(sim, state, n, rank, dist, torch, steps=20):
    """N > 1: the host boundary of the multi-GPU loop.  The state lives on the GPUs (the only form pb200_msim_* has:
    there is no per-step upload to skip), and EVERY step the new state is read back into host Entity records from
    rank 0 (pb200_msim_download: D2H of positions and velocities, 64 B/body, unpacked into the 80-byte records) -
    what a renderer or csvsink behind a multi-GPU integrator would receive.  Wall clock between barriers."""
    host = state.copy()
    for _ in range(3):
        sim.run(1)
        if rank == 0:
            sim.download(host)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.run(1)
        if rank == 0:
            sim.download(host)
    torch.cuda.synchronize()
    dist.barrier()
    dt = (time.perf_counter() - t0) / steps
    return {"value": n / dt, "unit": "particle-steps/s", "ms_per_step": dt * 1e3, "h2d_bytes_per_step": 0,
            "d2h_bytes_per_step": n * 64,
            "api": "pb200_msim_run(1) + pb200_msim_download(host Entity[n]) every step: state resident on the GPUs, "
                   "read back from rank 0 into host Entity records each step"}


def measure_bh_large(api, torch, dist, rank, world, local_rank):
    """BASELINE configs[4]: cube n=2^26 ! astro2 theta=0.7 e=0.5 ! verlet, the bodies made on the device from the
    reference's ChaCha8 `cube seed=1` stream.  One GPU: the single-GPU step (global radix sort).  8 GPUs: one
    simulation sharded by Morton key range (2 and 4 GPUs would hold more bodies per rank than the shared-memory
    bucket sort of the sharded build takes, so they are not run)."""
    n = 1 << 26
    w = dict(element="astro2", theta=0.7, e=0.5, dt=1e-6)
    out = {"workload": "cube n=67108864 seed=1 ! astro2 theta=0.7 e=0.5 ! verlet", "n_gpus": world}
    try:
        if world == 1:
            sim = api.Sim(w["element"], theta=w["theta"], e=w["e"], dt=w["dt"])
        else:
            sim = make_msim(api, torch, dist, w, rank, world, local_rank)
        sim.generate_cube(n, seed=1)
        sim.run(3)  # first step (replicated, plans the shards) + two regular steps
        steps = 16  # (half a checkpoint interval of the resident loop: its 3 x 2 GB state copies weigh as in a long run)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        ms = sim.run_timed(steps)
        if dist:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        st = sim.stats()
        out.update({"ms_per_step": ms / steps, "value": n * steps / (ms * 1e-3), "unit": "particle-steps/s",
                    "steps": steps, "n_cells": st["n_cells"], "interactions_per_step": st["interactions"],
                    "interactions_per_s": st["interactions"] * steps / (ms * 1e-3)})
        sim.profile(True)   # per-kernel device times (event pairs around every launch), 3 more steps
        sim.run(3)
        torch.cuda.synchronize()
        rep = sim.profile_report()
        sim.profile(False)
        out["kernels_ms_per_step"] = {k["kernel"]: round(k["ms"] / 3, 4) for k in sorted(rep, key=lambda k: -k["ms"])}
        if world > 1:
            bodies, _ = sim.rank_counts()
            out["sharding"] = {"sharded_steps": st.get("sharded_steps"), "replicated_steps": st.get("replicated_steps"),
                               "replays": st["replays"],
                               "bodies_per_rank": None if bodies is None else [int(b) for b in bodies]}
            sim.close()
        del sim
    except Exception as e:  # noqa: BLE001  (the headline line must still be printed)
        out["error"] = repr(e)
    return out


def measure_other_workloads(api, args):
    """The other BASELINE configurations that fit one GPU, device-resident (state in HBM), so that they appear in
    the driver's record: c1 (configs[0]), c3o (north_star's target config: 1 M bodies, astro2 theta=1.5, with its
    own end-to-end and CPU legs), c5s (configs[4] at 1/16 size)."""
    out = {}
    for name, steps in (("c1", 400), ("c3o", 100), ("c5s", 20)):
        w = workload(name)
        state = make_state(w)
        n = len(state)
        sim = api.Sim(w["element"], theta=w["theta"], e=w["e"], dt=w["dt"])
        sim.upload(state)
        sim.run_timed(5)
        ms = sim.run_timed(steps)
        st = sim.stats()
        row = {"description": w["desc"], "n_bodies": n, "ms_per_step": ms / steps, "value": n * steps / (ms * 1e-3),
               "unit": "particle-steps/s", "steps": steps, "n_cells": st["n_cells"],
               "interactions_per_step": st["interactions"], "sort_mode": st["sort_mode"]}
        del sim
        if name == "c3o":
            el = api.TransformElement(w["element"], theta=w["theta"], e=w["e"])
            v = api.Verlet()
            bufs = [state.copy(), state.copy()]
            cur = 0
            for _ in range(3):
                v.integrate_fused(bufs[cur], el, w["dt"], out=bufs[cur ^ 1])
                cur ^= 1
            k = 30
            t0 = time.perf_counter()
            for _ in range(k):
                v.integrate_fused(bufs[cur], el, w["dt"], out=bufs[cur ^ 1])
                cur ^= 1
            dt = (time.perf_counter() - t0) / k
            row["e2e"] = {"value": n / dt, "unit": "particle-steps/s", "ms_per_step": dt * 1e3,
                          "h2d_bytes_per_step": n * 33, "d2h_bytes_per_step": n * 24,
                          "api": "pb200_verlet_step_fused(host Entity[n] -> host Entity[n])"}
            row["cpu_baseline"] = measure_cpu(state, w, budget=4e6)
        out[name] = row
    return out


def measure_integrators(api, state, w):
    """the same workload under `euler` and `rk4` (four force evaluations per step),
    device-resident, next to the CPU oracle's loop (1 thread, bounded number of steps)."""
    from oracle import binding as ob
    out = {}
    n = len(state)
    for name, gpu_steps, cpu_steps in (("euler", 40, 3), ("rk4", 20, 2)):
        sim = api.Sim(w["element"], theta=w["theta"], e=w["e"], dt=w["dt"])
        sim.set_integrator(name)
        sim.upload(state)
        sim.run_timed(3)
        ms = sim.run_timed(gpu_steps)
        t0 = time.perf_counter()
        ob.run_pipeline_with(name, w["element"], state, w["theta"], w["e"], w["dt"], cpu_steps)
        cpu = time.perf_counter() - t0
        out[name] = {"value": n * gpu_steps / (ms * 1e-3), "unit": "particle-steps/s", "ms_per_step": ms / gpu_steps,
                     "cpu_baseline": {"value": n * cpu_steps / cpu, "cores": 1, "kind": "port",
                                      "sample": f"{cpu_steps} full steps"}}
    return out


def measure_cpu(state, w, budget=12e6):
    from oracle import binding as ob
    ob.build()
    n = len(state)
    steps = max(2, min(10, int(budget // max(n, 1))))  # about 10-20 s of CPU work
    t0 = time.perf_counter()
    _, secs = ob.run_pipeline(w["element"], state, w["theta"], w["e"], w["dt"], steps)
    dt = time.perf_counter() - t0
    model, cores = cpu_info()
    return {"value": n * steps / dt, "unit": "particle-steps/s", "cores": 1, "kind": "port",
            "sample": f"{steps} full steps of the same {n}-body workload",
            "phases_s_per_step": {"build": secs[0] / steps, "walk_force": secs[1] / steps,
                                  "integrate_copies": secs[2] / steps},
            "cpu": model, "host_cores": cores,
            "note": "1 thread: the reference runs the whole step on one simulation thread (pipeline.rs:134)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--skip-extras", action="store_true", help="skip e2e / direct-sum / cpu legs")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = N x the workload's bodies in one sharded simulation (default), strong = the same bodies")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")) if args.impl == "ours" else args.gpus
    w = workload(args.workload, max(world, 1), args.scaling)
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
