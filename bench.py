#!/usr/bin/env python
"""Benchmark of the hot path: FMM evaluation throughput, BASELINE.json metric
"FMM eval Mtargets/s (1M src biharmonic3d)" on config #3 (1M-source biharmonic3d interpolant
sampled at ~10M grid points).

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun)
  python bench.py --impl reference ...                     (CPU arm: the oracle port, host cores)

One step = one pass of the hot path over one batch: set_weights(1M weights) +
set_target_points(10M targets) + evaluate() -> 10M values; i.e. upward pass (P2M, M2M, multipole
DFT), target tree build, M2L, L2L, L2P, P2P and the scatter to caller order, every step.
`value` is measured with inputs/outputs resident in HBM; `e2e` through the same public calls
with pinned HOST buffers (H2D of targets + weights and D2H of the result inside the timed
region).

N > 1 GPUs: STRONG scaling of the ONE 10M-target grid (BASELINE.json config #3, "sharded over 2/4/8 B200").
The level-4 cells of the octree are partitioned over the ranks by Morton key range, balanced by target
count; a rank holds all sources + weights and only ITS targets (it copies only its slab H2D / D2H in the
e2e leg), computes the multipoles it owns or needs, and the ranks exchange the level-4 expansions with one
NCCL all-gather per step (plt_eval_set_partition).  `value` = all 10M targets / max-over-ranks
time.  The former replica number (every rank its own full grid) is kept as `weak_replicas`.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SOURCES = 1_000_000
GRID = (216, 216, 215)
METRIC = "FMM eval Mtargets/s (1M src biharmonic3d)"
UNIT = "Mtargets/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--accuracy", type=float, default=float("inf"),
                    help="evaluation accuracy (reference default: infinity -> order 6)")
    ap.add_argument("--n-sources", type=int, default=N_SOURCES)
    ap.add_argument("--grid", type=int, nargs=3, default=list(GRID))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    ap.add_argument("--no-fit", action="store_true", help="skip the auxiliary config #2 fit (FGMRES + RAS) timing")
    ap.add_argument("--no-sampler", action="store_true", help="skip the per-layer isosurface sampler timing")
    ap.add_argument("--cpu-sample-layers", type=int, default=216,
                    help="x-layers of the target grid (central slab) the CPU arm evaluates per step "
                         "(default: all of them, ~10 s of CPU work on 16 cores)")
    return ap.parse_args()


def workload(args, rank=0, world=1):
    from polatory_b200 import workloads as wl
    return wl.c3_isosurface_field(args.n_sources, tuple(args.grid))


def config_dict(args, n_src, n_trg, cfg, world):
    return {
        "workload": "config #3: isosurface field evaluation, biharmonic3d (s=1, c=0) interpolant with "
                    f"{n_src} sources (unit-sphere surface + normal-offset SDF points) sampled at "
                    f"{n_trg} grid targets ({'x'.join(map(str, args.grid))}) over 1.1 x bbox"
                    + (f"; ONE grid sharded over {world} ranks by Morton key range of the level-4 cells" if world > 1 else ""),
        "rbf": "bh3", "kernel_kind": "K", "dim": 3,
        "accuracy": "inf" if np.isinf(args.accuracy) else args.accuracy,
        "tree_height": cfg.get("tree_height"), "order": cfg.get("order"), "d": cfg.get("d"),
        "step": "set_weights + set_target_points + evaluate (upward pass, target tree, M2L, L2L, L2P, P2P)",
        "cache_policy": "inputs larger than L2 (240 MB targets, >4 GB expansions per step)",
        "parallelism": f"targets sharded by Morton range over {world} GPU(s); sources + weights replicated; "
                       "all-gather of the level-4 multipole expansions" if world > 1 else "single GPU",
    }


# ---------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def bind_to_gpu_numa_node(gpu_index):
    """One process per GPU: run on the cores next to the rank's GPU so that the pinned host buffers of the
    end-to-end leg are first-touched on the local NUMA node (8 ranks copying 250 MB per step each otherwise
    share one memory controller / PCIe root)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ---------------------------------------------------------------------------------------
def cpu_sample(args, src, w, trg, lo, hi):
    """Bounded sample of the same workload for the CPU arm: all sources, the central x-slab of
    the target grid, the tree height of the FULL problem."""
    g = args.grid
    layers = min(args.cpu_sample_layers, g[0])
    x0 = (g[0] - layers) // 2
    per_layer = g[1] * g[2]
    sub = trg[x0 * per_layer:(x0 + layers) * per_layer]
    from oracle import fmm as ofmm
    height = ofmm.tree_height(3, max(len(src), len(trg)))
    desc = (f"oracle port (oracle/fmm_oracle.c, OpenMP): all {len(src)} sources -> central slab of "
            f"{layers} x-layers = {len(sub)} of the {len(trg)} targets, tree height {height} of the full "
            f"problem, order/d as the GPU arm; whole evaluate() incl. tree build and upward pass")
    return sub, height, desc


def run_cpu_once(args, src, w, sub, lo, hi, height, order, d):
    from oracle import fmm as ofmm
    t0 = time.perf_counter()
    out = ofmm.fmm("bh3", [1.0, 0.0], 3, 0, lo, hi, src, sub, w, order, d, height)
    dt = time.perf_counter() - t0
    return dt, out


def reference_arm(args):
    """`--impl reference`: the reference's CPU implementation of the path on the host cores.
    The reference itself cannot be built in this image (ScalFMM3 / Eigen absent), so
    this is the oracle port, labelled as such."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must use every host core it can.
    # The variable is read when libgomp initialises, i.e. before the oracle library is loaded.
    if os.environ.get("OMP_NUM_THREADS", "") in ("", "1") or "TORCHELASTIC_RUN_ID" in os.environ:
        os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity")
                                            else (os.cpu_count() or 1))
    from oracle import fmm as ofmm
    src, w, trg, lo, hi = workload(args)
    order, d = (6, -1) if np.isinf(args.accuracy) else ((12, 8) if args.accuracy == 0 else (10, -1))
    sub, height, desc = cpu_sample(args, src, w, trg, lo, hi)
    for _ in range(args.warmup):
        run_cpu_once(args, src, w, sub, lo, hi, height, order, d)
    times = [run_cpu_once(args, src, w, sub, lo, hi, height, order, d)[0] for _ in range(args.steps)]
    t = float(np.mean(times))
    value = len(sub) / t / 1e6
    cores = ofmm.num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, len(src), len(trg), {"tree_height": height, "order": order, "d": d}, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        bind_to_gpu_numa_node(physical_gpu_index(local_rank))
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    import polatory_b200 as pb
    from polatory_b200 import _lib

    src, w, trg, lo, hi = workload(args, rank, world)
    n_src, n_trg_global = len(src), len(trg)
    ev = pb.make_fmm_evaluator(pb.make_rbf("bh3", [1.0, 0.0]), pb.Bbox(lo, hi))
    ev.set_accuracy(args.accuracy)
    trg_full = trg
    shard = None
    if world > 1:
        from polatory_b200.parallel import DEFAULT_CUT_LEVEL, partition_keys
        cut = DEFAULT_CUT_LEVEL[3]
        keys = ev.point_keys(trg, cut)
        key_begin = partition_keys(keys, world, 3, cut)
        own = np.nonzero((keys >= key_begin[rank]) & (keys < key_begin[rank + 1]))[0]
        trg = np.ascontiguousarray(trg[own])
        height = pb.fmm.tree_height(3, max(n_src, n_trg_global))   # src/fmm/utility.hpp:12-16 on the GLOBAL counts
        ev.force_config(0, -1, height)
        ev.set_partition(rank, world, cut, key_begin, group=dist.group.WORLD)
        shard = {"cut_level": cut, "key_begin": [int(k) for k in key_begin], "targets_this_rank": int(len(own))}
    n_trg = len(trg)
    d_src = torch.from_numpy(src).to(dev)
    d_w = torch.from_numpy(w).to(dev)
    d_trg = torch.from_numpy(trg).to(dev)
    d_out = torch.empty(n_trg, dtype=torch.float64, device=dev)
    ev.set_source_points(d_src)

    def step_device():
        ev.set_weights(d_w)
        ev.set_target_points(d_trg)
        ev.evaluate(d_out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()  # the library issues on the legacy default stream = torch's current stream
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize()

    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    launches0 = ev.launch_count()
    ms_total = timed(step_device, args.steps)
    launches = ev.launch_count() - launches0
    clocks = sampler.stop()
    phases = ev.phase_times()
    cfg = ev.config()
    work = ev.work_stats()   # of the plan of the timed steps (the host-buffer leg below rebuilds plans per slab)
    ms_step = ms_total / args.steps
    value = n_trg_global / (ms_step * 1e-3) / 1e6
    allgathers = ev.allgather_count()

    # ---- e2e: same calls, pinned host buffers, H2D + D2H inside the timed region ----
    h_trg = torch.from_numpy(trg).pin_memory()
    h_w = torch.from_numpy(w).pin_memory()
    h_out = torch.empty(n_trg, dtype=torch.float64).pin_memory()

    def step_host():
        # interpolation::Evaluator::evaluate(points) = set_target_points + evaluate (evaluator.hpp:83-87), as one ABI
        # call: on one GPU the library streams the host targets in slabs (copy-in / evaluate / copy-out pipeline)
        ev.set_weights(h_w.numpy())
        ev.evaluate_points(h_trg.numpy(), h_out.numpy())

    if args.no_e2e:
        e2e_ms, e2e_value, host_ok = None, None, None
    else:
        for _ in range(2):
            step_host()
        e2e_ms = timed(step_host, args.steps) / args.steps
        e2e_value = n_trg_global / (e2e_ms * 1e-3) / 1e6
        host_ok = bool(np.allclose(h_out.numpy()[:1000], d_out[:1000].cpu().numpy(), rtol=0, atol=0))

    # ---- roofline of the dominant kernel (live CUDA-event time of the last timed step) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    import ctypes
    tf = ctypes.c_double(0.0)
    _lib.load().plt_measure_fp64_peak(ctypes.byref(tf))
    fp64_peak = float(tf.value)
    dominant = max(phases, key=phases.get) if phases else None
    roofline = roofline_for(dominant, phases, cfg, n_src, n_trg, hbm_peak, hbm_src, fp64_peak, work)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, n_src, n_trg_global, cfg, world),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms,
                # summed over the ranks: every rank copies its own target slab + the weights in, its own values out
                "h2d_bytes_per_step": int(trg_full.nbytes + world * w.nbytes), "d2h_bytes_per_step": int(8 * n_trg_global),
                "matches_device_path": host_ok},
        "gpu_launches": int(launches),
        "phases_ms": {k: round(v, 4) for k, v in phases.items()},
        "roofline": roofline,
        "fp64_peak_tflops_measured": fp64_peak,
    }

    line["phase_rooflines"] = phase_rooflines(phases, cfg, n_trg, hbm_peak, fp64_peak, work)
    if world > 1:
        line["multi_gpu"] = multi_gpu_report(args, ev, dist, dev, world, rank, shard, d_w, d_trg, d_out, d_src,
                                             trg_full, allgathers, ms_step, timed, barrier)
    if world == 1 and not args.no_fit:
        line["fit"] = fit_timing(src, dev)
    if world == 1 and not args.no_sampler:
        line["isosurface_sampler"] = sampler_timing(args, src, w, d_trg, lo, hi, d_out)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sub, height, desc = cpu_sample(args, src, w, trg, lo, hi)
        order, d = cfg["order"], cfg["d"]
        dt, out_cpu = run_cpu_once(args, src, w, sub, lo, hi, height, order, d)
        from oracle import fmm as ofmm
        # parity of the bench's own output against the oracle on the sample
        g = args.grid
        layers = min(args.cpu_sample_layers, g[0])
        x0 = (g[0] - layers) // 2
        got = d_out[x0 * g[1] * g[2]:(x0 + layers) * g[1] * g[2]].cpu().numpy()
        rel = float(np.max(np.abs(got - out_cpu)) / np.max(np.abs(out_cpu)))
        line["cpu_baseline"] = {"value": len(sub) / dt / 1e6, "unit": UNIT, "cores": ofmm.num_threads(),
                                "kind": "port", "sample": desc, "seconds": dt,
                                "gpu_vs_cpu_port_max_rel_diff": rel}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sampler_timing(args, src, w, d_trg, lo, hi, d_ref):
    """Config #3 the way the isosurface generator drives it (include/polatory/isosurface/rmt/lattice.hpp:421-445):
    the same 10M lattice nodes in per-layer batches against fixed centres + weights.  The source tree and the
    multipole spectra stay resident (the reference rebuilds them per batch, src/fmm/fmm_evaluator.hpp:107-109), so a
    batch is target-side work only.  Tree height per batch = max(n_src, n_batch) rule -> 7 (the one-shot grid: 8)."""
    import torch
    import polatory_b200 as pb
    from polatory_b200.evaluator import RbfFieldFunction
    from polatory_b200.operator import Model
    g = args.grid
    per_layer = g[1] * g[2]
    field = RbfFieldFunction(Model(pb.make_rbf("bh3", [1.0, 0.0]), poly_degree=-1), src, w)
    field.set_evaluation_bbox(pb.Bbox(lo, hi))
    out = {}
    for layers in (1, 8):
        n_b = (g[0] + layers - 1) // layers
        res = None
        for rep in range(2):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for b in range(n_b):
                res = field(d_trg[b * layers * per_layer:(b + 1) * layers * per_layer])
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        out[f"{layers}_layer_batches"] = {"batches": n_b, "targets_per_batch": layers * per_layer, "total_ms": ms,
                                          "Mtargets_per_s": g[0] * per_layer / (ms * 1e-3) / 1e6,
                                          "config": field.evaluator.a[0].config()}
    return out


def multi_gpu_report(args, ev, dist, dev, world, rank, shard, d_w, d_trg, d_out, d_src, trg_full, allgathers,
                     ms_step, timed, barrier):
    """Evidence for the sharded run: per-rank device times and shard sizes, the NCCL exchange count, parity of
    the N-GPU result against the same rank evaluating its targets alone (no partition), and the former replica
    (weak-scaling) figure."""
    import torch
    # per-rank time of one more step (device events, this rank only)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    ev.set_weights(d_w)
    ev.set_target_points(d_trg)
    ev.evaluate(d_out)
    e1.record()
    torch.cuda.synchronize()
    mine = torch.zeros(world, dtype=torch.float64, device=dev)
    mine[rank] = e0.elapsed_time(e1)
    dist.all_reduce(mine)
    counts = torch.zeros(world, dtype=torch.float64, device=dev)
    counts[rank] = shard["targets_this_rank"]
    dist.all_reduce(counts)
    up = ev.phase_times()
    upward_ms = torch.tensor([sum(up.get(k, 0.0) for k in ("p2m", "m2m", "m2hat")), up.get("allgather", 0.0)],
                             dtype=torch.float64, device=dev)
    dist.all_reduce(upward_ms, op=dist.ReduceOp.MAX)
    sharded = d_out.clone()
    # 1-GPU result of the same targets: partition removed, same (global) tree height
    ev.set_partition(0, 1)
    ev.set_weights(d_w)
    ev.set_target_points(d_trg)
    ev.evaluate(d_out)
    diff = torch.stack([(sharded - d_out).abs().max(), d_out.abs().max()])
    dist.all_reduce(diff, op=dist.ReduceOp.MAX)
    # replicas: every rank the full grid (the round-1 weak-scaling number)
    d_full = torch.from_numpy(trg_full).to(dev)
    d_out_full = torch.empty(len(trg_full), dtype=torch.float64, device=dev)

    def step_full():
        ev.set_weights(d_w)
        ev.set_target_points(d_full)
        ev.evaluate(d_out_full)

    for _ in range(2):
        step_full()
    ms_full = timed(step_full, max(2, args.steps // 2)) / max(2, args.steps // 2)
    return {"cut_level": shard["cut_level"], "key_begin": shard["key_begin"],
            "targets_per_rank": [int(c) for c in counts.cpu()],
            "step_ms_per_rank": [round(float(v), 3) for v in mine.cpu()],
            "upward_ms_max": float(upward_ms[0]), "allgather_ms_max": float(upward_ms[1]),
            "nccl_allgathers_per_step": 1, "nccl_allgathers_total_rank0": int(allgathers),
            "n_gpu_vs_1_gpu_max_rel_diff": float(diff[0] / diff[1]),
            "weak_replicas": {"value": world * len(trg_full) / (ms_full * 1e-3) / 1e6, "unit": UNIT,
                              "ms_per_step": ms_full, "what": "every rank its own full 10M-target grid (round-1 number)"}}


def fit_timing(points, dev, tol=1e-4):
    """The fit twice in this process: `first_run` carries the one-time costs of a process (CUDA module loading of ~40
    kernels, pool growth), the headline keys are the second run (every handle, tree, factor and Krylov basis is
    still created from scratch inside it)."""
    first = fit_once(points, dev, tol)
    second = fit_once(points, dev, tol)
    second["first_run"] = {k: first[k] for k in ("wall_s", "operator_setup_s", "ras_setup_s", "solve_s", "iterations")}
    return second


def fit_once(points, dev, tol=1e-4):
    """Second half of BASELINE.json's metric: wall-time of the 1M-point fit (config #2: bh3 SDF centres,
    degree 0, absolute tolerance 1e-4, evaluator accuracy tol / 100), end to end from host arrays, in the
    reference's configuration (include/polatory/interpolation/solver.hpp:40-41,60): the matvec operator at
    accuracy 0 (-> order 12, d 8), a separate residual evaluator at the user's accuracy (its accuracy search
    included), RAS set-up (coarse points, domains, batched factorisations), FGMRES iterations with the FMM
    matvec and the RAS preconditioner, convergence by the reference's ResidualEvaluator (exact sums on <= 1024
    sampled data points, then every point through the fast evaluator).  The final residual on the exact sample
    is asserted here."""
    import torch
    import polatory_b200 as pb
    from polatory_b200.operator import Model, Operator, ResidualEvaluator, solve
    from polatory_b200.ras import RasPreconditioner
    n = len(points)
    third = (n + 2) // 3
    values = np.concatenate([np.zeros(third), np.full(third, 1e-2), np.full(n - 2 * third, -1e-2)])
    model = Model(pb.make_rbf("bh3", [1.0, 0.0]), poly_degree=0, nugget=0.0)
    bbox = pb.Bbox(points.min(axis=0), points.max(axis=0))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    op = Operator(model, bbox, 0.0, 0.0)                      # solver.hpp:40
    res_op = Operator(model, bbox, tol / 100.0, tol / 100.0)  # solver.hpp:41
    op.set_points(points)
    res_op.set_points(points)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    pc = RasPreconditioner(model, points)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    w, iters = solve(op, values, tol, 100, preconditioner=pc.apply, residual_op=res_op)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    # acceptance (test/interpolation/test_fitter.cpp:57-64): the exact residual on the reference's sample
    chk = ResidualEvaluator(res_op)
    chk.set_values(torch.from_numpy(values).to(dev))
    ok, res, _, _ = chk.converged(w, tol)
    assert ok and res <= tol, f"fit did not meet the tolerance: residual {res}"
    # steady-state cost of the two operators of an iteration
    x = torch.cat([torch.from_numpy(values).to(dev), torch.zeros(1, dtype=torch.float64, device=dev)])
    y = torch.empty_like(x)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    op.apply(w, y)
    pc.apply(x, y)
    e[0].record()
    op.apply(w, y)
    e[1].record()
    pc.apply(x, y)
    e[2].record()
    torch.cuda.synchronize()
    ph = op.a[0].phase_times()
    return {"workload": f"config #2 fit: {n} bh3 centres (sphere surface + normal offsets, values 0 / +-1e-2), degree 0, "
                        f"tolerance {tol} absolute, accuracy tol/100, FGMRES + RAS, reference configuration "
                        f"(matvec at accuracy 0, separate residual evaluator)",
            "wall_s": t3 - t0, "operator_setup_s": t1 - t0, "ras_setup_s": t2 - t1, "solve_s": t3 - t2,
            "ras_setup_breakdown_s": {k: round(v, 3) for k, v in pc.setup_seconds.items()},
            "iterations": iters, "levels": pc.n_levels, "domains": [f.n_dom if f else 1 for f in pc.fine],
            "residual_max_abs": res, "matvec_config": op.a[0].config(),
            "residual_evaluator_config": res_op.a[0].config(), "matvec_ms": e[0].elapsed_time(e[1]),
            "ras_apply_ms": e[1].elapsed_time(e[2]),
            "matvec_phases_ms": {k: round(v, 4) for k, v in ph.items()}}


def phase_rooflines(phases, cfg, n_trg, hbm_peak, fp64_peak, work):
    """Roofline fractions of the secondary phases whose algorithmic work is known from the device counters
    (DESIGN.md section 5): the pruned inverse DFT and the fused leaf pass against HBM, the near field against FP64
    (SURVEY 8d counting rule: 12 flop per bh3 pair)."""
    out = {}
    try:
        p = cfg.get("order") or 0
        F, P = (2 * p - 1) ** 2 * p, p ** 3
        if phases.get("m2l_idft") and stats.get("m2l_target_cells"):
            gbs = (16.0 * F + 8.0 * P) * stats["m2l_target_cells"] / (phases["m2l_idft"] * 1e-3) / 1e9
            out["m2l_idft"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak}
        if phases.get("l2l_l2p_leaf"):
            # 8 P bytes of parent local per 2^dim leaves + compact leaf M2L result + 8 (D + 1) bytes per target
            nb = 8.0 * P * stats.get("m2l_target_cells", 0) + 8.0 * (3 + 1) * n_trg
            gbs = nb / (phases["l2l_l2p_leaf"] * 1e-3) / 1e9
            out["l2l_l2p_leaf"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                   "frac": gbs / hbm_peak, "note": "shared-memory bound (LDS.128 of the expansions), DESIGN.md"}
        if phases.get("p2p") and stats.get("p2p_pairs"):
            tf = 12.0 * stats["p2p_pairs"] / (phases["p2p"] * 1e-3) / 1e12
            out["p2p"] = {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s",
                          "frac": tf / fp64_peak if fp64_peak else None,
                          "note": "12 flop per pair by the counting rule; 12 FP64-pipe instructions per pair in SASS"}
    except Exception as e:  # diagnostics only
        out["error"] = str(e)
    return out


def measured_traffic(kernel):
    """DRAM bytes per step of a kernel from this round's `ncu --set full` capture
    (profiles/traffic.json, written by tools/ncu_summary.py traffic); None when not captured."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t.get(kernel, {}).get("dram_bytes_per_step")
    except Exception:
        return None


def roofline_for(name, phases, cfg, n_src, n_trg, hbm_peak, hbm_src, fp64_peak, stats):
    """Algorithmic work of the dominant kernel / its CUDA-event time
    (sum over the kernel's launches of one step, events on the launching stream)."""
    if not name:
        return None
    ms = phases[name]
    p = cfg.get("order") or 0
    out = {"kernel": name, "ms": ms, "work": stats}
    F = (2 * p - 1) ** 2 * p
    P = p ** 3
    if name == "m2l_hadamard" and stats.get("m2l_pairs"):
        flops = 8.0 * F * stats["m2l_pairs"]  # one complex multiply-add per frequency and pair
        ach = flops / (ms * 1e-3) / 1e12
        out.update({"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": ach / fp64_peak if fp64_peak else None, "traffic": measured_traffic("m2l_hadamard"),
                    "peak_source": "DFMA-chain microbenchmark in this run (plt_measure_fp64_peak)",
                    "algorithmic": f"8 flop x F={F} frequencies x {stats['m2l_pairs']} M2L pairs"})
    elif name == "m2l_idft" and stats.get("m2l_target_cells"):
        nb = (16.0 * F + 8.0 * P) * stats["m2l_target_cells"]
        ach = nb / (ms * 1e-3) / 1e9
        out.update({"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                    "traffic": None, "peak_source": hbm_src,
                    "algorithmic": f"(16 F + 8 P) bytes x {stats['m2l_target_cells']} target cells"})
    else:
        nb = 8.0 * (3 + 1) * n_trg
        ach = nb / (ms * 1e-3) / 1e9
        out.update({"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                    "traffic": None, "peak_source": hbm_src, "algorithmic": "32 bytes per target"})
    return out


if __name__ == "__main__":
    main()
