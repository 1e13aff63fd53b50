#!/usr/bin/env python
"""bench.py — the reference's headline metric on B200: tet-steps/s (and pixels/s) of the per-pixel
ray pass at 2400 x 1800 (BASELINE.json `metric`, configs[2]: synthetic 8M-tet grid, 2400 x 1800,
--alpha_limit 3.0 -X 0.5, plus the Roche lobe and sphere the reference always renders).

    python bench.py [--gpus N] [--steps K] [--warmup W]              # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]   # the reference's OpenMP path

A "step" is one view: rotate -> BVH refit -> solid mask -> entry + tet walk (-> band gather at N > 1).
  value   tet-steps/s with everything resident in HBM, output left on the device (CUDA events on the
          launching stream, max over ranks).
  e2e     the same metric through the C-ABI call a user makes (c5_render) with a pinned HOST output
          buffer: view parameters go host->device, the image device->host, inside the timed region.
  roofline  the walk kernel: algorithmic bytes (72 B per tet-step + 16 B per pixel, SURVEY.md §8d /
          DESIGN.md) / its CUDA-event time, against the measured HBM copy peak.
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, built from /root/reference in the build
          container) timed on this box's host cores on the same workload (N = 1, rank 0 only).
The oracle is only ever the thing compared against or the baseline — never the thing measured as ours.
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

from course5_b200 import api, synth  # noqa: E402

WORKLOAD = "C3"
BYTES_PER_STEP = 72      # 16 B connectivity + 16 B neighbours + 24 B one new vertex + 16 B (alpha, Q)
BYTES_PER_PIXEL = 16     # {tau, I} doubles written


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_capture():
    """What the committed `ncu --set full` capture of the walk says (profiles/walk_traffic.json), if it exists."""
    path = os.path.join(ROOT, "profiles", "walk_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return {}


def ncu_traffic_per_launch():
    """dram bytes per walk launch from the committed ncu capture, if one exists."""
    return ncu_capture().get("dram_bytes_per_launch")


class ClockSampler(threading.Thread):
    """Polls SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


class Watchdog(threading.Thread):
    """Ends the process if the run stops making progress. A multi-GPU run that deadlocks (a collective
    some rank never joins, a kernel that spins for ever) would otherwise sit on its GPUs until an
    outer limit kills it; this way it fails fast, says in which phase, and frees the devices."""

    def __init__(self, rank: int, limit_s: float):
        super().__init__(daemon=True)
        self.rank, self.limit_s = rank, limit_s
        self.phase, self.t_phase, self.t0 = "start", time.monotonic(), time.monotonic()
        self.verbose = os.environ.get("C5_BENCH_VERBOSE") == "1"

    def tick(self, phase: str, limit_s: float | None = None):
        now = time.monotonic()
        if self.verbose or self.rank == 0:
            print(f"[bench rank {self.rank}] +{now - self.t0:6.1f}s {phase}", file=sys.stderr, flush=True)
        self.phase, self.t_phase = phase, now
        if limit_s is not None:
            self.limit_s = limit_s

    def run(self):
        while True:
            time.sleep(2.0)
            stalled = time.monotonic() - self.t_phase
            if stalled > self.limit_s:
                print(f"[bench rank {self.rank}] WATCHDOG: no progress for {stalled:.0f}s in phase '{self.phase}'; "
                      "exiting (exit code 3)", file=sys.stderr, flush=True)
                os._exit(3)


def workload():
    mesh, view = synth.make_config(WORKLOAD)
    return mesh, view


def reference_solids(D):
    """The reference's Roche lobe + sphere from the host-side generator (bit-identical to the
    reference's own, tests/test_host.py)."""
    from course5_b200 import hostlib
    return hostlib.make_solids(D)


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path on this box's host cores
# ------------------------------------------------------------------------------------------------

def reference_sample(view):
    """Bounded sample of the workload for the CPU arms: same mesh, same flags, same solids, at half
    the linear resolution (1/4 of the pixels) so that a step is seconds, not half a minute."""
    return dict(view, res_x=view["res_x"] // 2, res_y=view["res_y"] // 2)


def time_reference(mesh, view, *, steps, warmup):
    from oracle import refbind
    cores = os.cpu_count() or 1
    threads = min(32, cores)           # MAX_NUMBER_OF_THREADS = 32 (config.hpp:39)
    tet_pts = mesh.tet_points()
    kind = "reference" if os.path.exists(refbind.REF_SO) else "port"
    flags = dict(X=view["X"], Y=view["Y"], D=view["D"], I=view["I"], alpha_limit=view["alpha_limit"])
    times, total_steps = [], 0
    if kind == "reference":
        ref = refbind.Ref()
        for k in range(warmup + steps):
            img = ref.render(tet_pts, mesh.alpha, mesh.q, res_x=view["res_x"], res_y=view["res_y"],
                             threads=threads, solids=1, raw=False, **flags)
            # the reference's own timed region: plane ctor + find_intersections + trace_rays (main.cpp:126-130)
            t = img.timings["ctor"] + img.timings["find"] + img.timings["trace"]
            if k >= warmup:
                times.append(t)
            total_steps = img.total_steps
    else:
        port = refbind.Port()
        threads = cores
        for k in range(warmup + steps):
            img = port.render(tet_pts, mesh.alpha, mesh.q, res_x=view["res_x"], res_y=view["res_y"],
                              threads=threads, X=flags["X"], Y=flags["Y"], I=flags["I"],
                              alpha_limit=flags["alpha_limit"])
            if k >= warmup:
                times.append(img.timings["trace"])
            total_steps = img.total_steps
    t_step = float(np.mean(times))
    return dict(value=total_steps / t_step, seconds_per_step=t_step, tet_steps=total_steps, kind=kind,
                cores=threads, pixels=view["res_x"] * view["res_y"])


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mesh, view = workload()
    sample = reference_sample(view)
    r = time_reference(mesh, sample, steps=args.steps, warmup=args.warmup)
    sample_txt = (f"{WORKLOAD} mesh ({mesh.n_tets} tets) + reference solids, same flags, "
                  f"{sample['res_x']}x{sample['res_y']} (1/4 of the pixels), {r['tet_steps']} tet-steps per step")
    line = {
        "impl": "reference", "metric": "tet_steps_per_sec", "value": r["value"], "unit": "tet-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * r["seconds_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "pixels_per_sec": r["pixels"] / r["seconds_per_step"],
        "config": config_dict(mesh, view, args.gpus),
        "cpu_baseline": {"value": r["value"], "unit": "tet-steps/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": sample_txt},
        "e2e": {"value": r["value"], "unit": "tet-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def config_dict(mesh, view, n_gpus, gather="p2p", lanes=1, name=WORKLOAD):
    return {
        "workload": (f"{name}: synthetic Kuhn-split grid, {mesh.n_tets} tets / {mesh.n_points} points, "
                     f"{view['res_x']}x{view['res_y']}, -X {view['X']} -Y {view['Y']} --alpha_limit "
                     f"{view['alpha_limit']}, reference Roche lobe + sphere as solids"),
        "res_x": view["res_x"], "res_y": view["res_y"], "n_tets": mesh.n_tets,
        "views_in_flight": lanes,
        "grazing_kernel": ("after the pixel kernel (NCCL shares the device)" if n_gpus > 1 else
                           "after the pixel kernel" if os.environ.get("C5_GRAZE_SERIAL") else "beside the pixel kernel (side stream)"),
        "parallelism": "single GPU" if n_gpus == 1 else
                       f"{n_gpus} row bands (time-balanced), mesh replicated, " +
                       ("bands stored into rank 0's image over NVLink peer mappings by the walk kernel, one barrier per view"
                        if gather != "sendrecv" else "one grouped ncclSend/ncclRecv gather-v to rank 0 per view"),
        "l2": "inputs larger than L2 (cell records alone are 64 B x n_tets >> 126 MB); no explicit flush",
    }


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist
    from course5_b200.dist import BandRenderer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun --nproc-per-node {args.gpus}")
    dry = args.dry_run_hostsim   # CPU rehearsal of this function's control flow (tests/test_bench.py); not a measurement
    if not dry and not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU path "
                         "(use --impl reference for the CPU baseline)")
    if dry:
        device = torch.device("cpu")
        lib = api.load_library(os.path.join(ROOT, "tests", "hostsim", "libc5hostsim.so"))
    else:
        torch.cuda.set_device(local_rank)
        device = torch.device("cuda", local_rank)
        lib = None
    dog = Watchdog(rank, limit_s=float(os.environ.get("C5_BENCH_STALL_LIMIT", "150" if world == 1 else "90")))
    dog.start()
    dog.tick("init process group" if world > 1 else "single process")
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line
        if dry:
            dist.init_process_group("gloo")
        else:
            dist.init_process_group("nccl", device_id=device)

    dog.tick("synthetic mesh + solids on the host")
    if dry:
        mesh = synth.kuhn_cube(8, seed=3)
        view = dict(res_x=160, res_y=120, X=0.5, Y=0.0, I=0.0, D=0.0, alpha_limit=3.0)
        solids = None
    else:
        mesh, view = workload()
        solids = reference_solids(view["D"])

    dog.tick("upload (topology, BVH)")
    ctx = api.Context(devices=(0 if dry else local_rank,), lib=lib)
    t0 = time.perf_counter()
    info = ctx.upload_mesh(mesh.points, mesh.tets, mesh.alpha, mesh.q)
    if solids is not None:
        ctx.upload_solids(solids[0], True)
        ctx.upload_solids(solids[1], False)
    upload_s = time.perf_counter() - t0
    v = api.make_view(view["res_x"], view["res_y"], X=view["X"], Y=view["Y"], I=view["I"],
                      alpha_limit=view["alpha_limit"], lib=ctx.lib)
    br = BandRenderer(ctx, device=device, rank=rank, world=world, gather=args.gather, lanes=args.lanes)

    def barrier():
        if world > 1:
            dist.barrier()
        if not dry:
            torch.cuda.synchronize(device)

    class HostClock:   # dry run only: stands in for a CUDA event
        def record(self):
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return 1e3 * (other.t - self.t)

    # ---- value: everything resident, output stays on the device -------------------------------
    # warm-up views also settle the band cuts: tet-steps of the previous view, weighted by the time
    # each band took (a few iterations; a sweep does the same from frame to frame)
    n_warm = max(args.warmup, 3) if world == 1 else max(args.warmup, 6)
    for k in range(n_warm):
        dog.tick(f"warm-up view {k}")
        _, _, bands = br.render(v, rebalance="time")
    dog.tick(f"bands {bands}")
    barrier()
    sampler = ClockSampler(local_rank)
    if not dry:
        sampler.start()
    launches0 = br.kernel_launches()
    ev0, ev1 = (HostClock(), HostClock()) if dry else (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    walk_ms, steps_total, stats_last = [], 0, None
    ev0.record()
    # The views are only enqueued (no host readback between them; statistics come from a second
    # pass): consecutive views alternate between two lanes — the context and a sibling that shares
    # its mesh, each on its own stream — so the tail of one view's walk overlaps the start of the
    # next. N > 1, gather=p2p: the walk stores straight into rank 0's image over NVLink and the
    # barrier of view k is left in flight (three images); gather=sendrecv: the same with one grouped
    # ncclSend/ncclRecv per view.
    dog.tick(f"timed region: {args.steps} views, {br.n_lanes} in flight, gather={br.gather_mode}")
    for _ in range(args.steps):
        _, _, bands = br.render(v, rebalance=False, stats=False, pipeline=True)
    br.finish()
    ev1.record()
    barrier()
    dog.tick("statistics pass")
    clocks = sampler.stop() if not dry else {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    launches = br.kernel_launches() - launches0
    elapsed_ms = max(ev0.elapsed_time(ev1), 1e-6)
    for _ in range(3):
        _, st, bands = br.render(v, rebalance=False)
        walk_ms.append(st["ms_walk"])
        stats_last = st
    barrier()
    band_steps = stats_last["tet_steps"]

    # ---- e2e: the public C-ABI call with a pinned HOST output buffer ---------------------------
    # N = 1: c5_render into a pinned buffer. N > 1: the image is ONE pinned shared-memory segment;
    # every rank's c5_render writes its band in place over its own PCIe link, then one barrier.
    # Wall clock around the calls a user makes; the image is complete in host memory at the end
    # of every step.
    dog.tick("e2e: host image")
    from course5_b200.dist import SharedHostImage
    if os.environ.get("C5_BENCH_FAKE_HANG") == "1" and args.lanes > 1:   # tests/test_bench.py: the fallback path
        time.sleep(10_000)
    old_style = world > 1 and args.e2e == "gather"
    if world == 1 or old_style:
        host_out = torch.empty((view["res_y"], view["res_x"], 2), dtype=torch.float64)
        if not dry:
            host_out = host_out.pin_memory()
        host_np = host_out.numpy()
        shared = None
    else:
        shared = SharedHostImage(ctx, view["res_x"], view["res_y"], rank=rank, world=world)
        host_np = shared.array
    lo, hi = bands[rank]
    ve = api.View.from_buffer_copy(v)
    ve.row_begin, ve.row_end = lo, hi
    def e2e_step():
        if old_style:      # band gather on the devices, then ONE device-to-host copy of the image on rank 0
            img, _, _ = br.render(v, rebalance=False, stats=False)
            if rank == 0:
                host_out.copy_(img, non_blocking=False)
        else:
            ctx.render(ve, out=host_np)
            if world > 1:
                shared.barrier()

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if shared is not None:
        shared.close()

    dog.tick("reduce over ranks")
    # ---- reduce over ranks ----------------------------------------------------------------------
    t = torch.tensor([elapsed_ms, e2e_s * 1e3, float(np.mean(walk_ms))], dtype=torch.float64, device=device)
    s = torch.tensor([band_steps, launches], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    elapsed_ms, e2e_ms, walk_ms_max = (float(x) for x in t.cpu())
    total_steps, total_launches = (int(x) for x in s.cpu())
    # the roofline line describes the walk of the band with the most tet-steps (rank 0's may be empty)
    mine = torch.tensor([float(band_steps), float((bands[rank][1] - bands[rank][0]) * view["res_x"]),
                         float(np.mean(walk_ms))], dtype=torch.float64, device=device)
    per_rank = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, mine)
    else:
        per_rank = [mine]
    per_rank = [p.cpu().tolist() for p in per_rank]
    busiest = max(range(world), key=lambda r: per_rank[r][0])

    if rank == 0:
        pixels = view["res_x"] * view["res_y"]
        ms_per_step = elapsed_ms / args.steps
        value = total_steps / (ms_per_step * 1e-3)
        e2e_value = total_steps / (e2e_ms * 1e-3 / args.steps)
        peak, peak_src = measured_hbm_peak()
        # dominant kernel (tet_walk_fp64 + its grazing-ray kernel) on the busiest band; at N = 1 the whole image
        k_steps, k_pixels, k_ms = per_rank[busiest]
        k_ms = max(k_ms, 1e-9)      # (only the dry run has zero device times)
        achieved = (k_steps * BYTES_PER_STEP + k_pixels * BYTES_PER_PIXEL) / (k_ms * 1e-3) / 1e9
        line = {
            "metric": "tet_steps_per_sec", "value": value, "unit": "tet-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": n_warm, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "pixels_per_sec": pixels / (ms_per_step * 1e-3),
            "tet_steps_per_view": total_steps,
            "config": config_dict(mesh, view, world, br.gather_mode, br.n_lanes, "dry-run cube (no solids)" if dry else WORKLOAD),
            "e2e": {"value": e2e_value, "unit": "tet-steps/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": int(api.C.sizeof(api.View)), "d2h_bytes_per_step": pixels * 16,
                    "api": "c5_render (pinned host buffer)" if world == 1 else
                    "BandRenderer.render + one device-to-host copy on rank 0" if old_style else
                    "c5_render (row band) into one pinned shared-memory host image, one barrier per view"},
            "gpu_launches": total_launches,
            "roofline": {"kernel": "tet_walk_fp64", "rank": busiest, "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "frac_of_8000_GBps_nominal": achieved / 8000.0,
                         "traffic": ncu_traffic_per_launch(),
                         "peak_source": peak_src, "kernel_ms": k_ms,
                         "ncu": {k: ncu_capture().get(k) for k in ("l1_hit_rate_pct", "l2_hit_rate_pct",
                                                                   "l1_data_pipe_wavefronts_pct_of_peak", "source")},
                         "algorithmic_bytes_per_launch": int(k_steps * BYTES_PER_STEP + k_pixels * BYTES_PER_PIXEL)},
            "bands": [list(b) for b in bands],
            "phases_ms": {k: stats_last[k] for k in ("ms_rotate", "ms_bvh", "ms_mask", "ms_walk", "ms_total")},
            "one_off": {"upload_and_topology_s": upload_s, "device_bytes": int(info.device_bytes),
                        "boundary_faces": int(info.n_boundary_faces)},
            "clocks": clocks,
        }
        if os.environ.get("C5_BENCH_ATTEMPTS"):
            line["attempts"] = json.loads(os.environ["C5_BENCH_ATTEMPTS"])   # configurations left before this one
        if dry:
            line["data"] = "DRY RUN on the host-loop test build: control flow only, not a measurement"
        if world == 1 and not args.no_cpu_baseline and not dry:
            dog.tick("cpu_baseline: the reference on the host cores", limit_s=1800.0)
            sample = reference_sample(view)
            r = time_reference(mesh, sample, steps=1, warmup=0)
            line["cpu_baseline"] = {
                "value": r["value"], "unit": "tet-steps/s", "cores": r["cores"], "kind": r["kind"],
                "seconds": r["seconds_per_step"],
                "sample": (f"{WORKLOAD} mesh + reference solids, same flags, {sample['res_x']}x{sample['res_y']} "
                           f"(1/4 of the pixels), {r['tet_steps']} tet-steps, 1 run")}
        print(json.dumps(line), flush=True)
    dog.tick("teardown", limit_s=120.0)
    br.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--lanes", type=int, default=2, choices=[1, 2, 3, 4],
                    help="views in flight per GPU (the context and lanes-1 siblings sharing its mesh, one stream each)")
    ap.add_argument("--gather", choices=["auto", "p2p", "sendrecv"], default="sendrecv",
                    help="N > 1: how bands reach rank 0's image. sendrecv = one grouped ncclSend/ncclRecv per view "
                         "(default: the transport of this round's 8-GPU scaling run); p2p = stored by the walk kernel straight into rank "
                         "0's image over NVLink peer mappings (validated at 2 GPUs, same speed there)")
    ap.add_argument("--dry-run-hostsim", action="store_true",
                    help="rehearse the control flow on CPU with tests/hostsim (gloo, tiny mesh); prints a line marked as a dry run")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--e2e", choices=["shared-host", "gather"], default="shared-host",
                    help="N > 1: e2e through one pinned shared-memory host image written by every rank (default), or "
                         "through the band gather plus one device-to-host copy on rank 0")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("C5_BENCH_CHILD") is None:
        run_with_fallback(args)
    else:
        run_ours(args)


def _die_with_parent():
    """In the child, before exec: if the launcher kills this rank, its measurement process goes too
    (no orphan left on a GPU)."""
    import ctypes
    import signal
    try:
        ctypes.CDLL("libc.so.6", use_errno=True).prctl(1, int(signal.SIGKILL), 0, 0, 0)   # PR_SET_PDEATHSIG
    except Exception:
        pass


def run_with_fallback(args):
    """N > 1: every rank runs the measurement in a CHILD process and falls back to a more conservative
    configuration if the first one does not finish. A multi-process GPU run that deadlocks cannot be
    rescued from inside (the collective never returns); the child's own watchdog ends it, every rank
    sees a non-zero exit code at about the same time, and all of them start the next attempt — on a
    fresh rendezvous port, since the first attempt's store is dead. The parent touches neither CUDA
    nor NCCL, so a killed child leaves the devices free. What was attempted and why it was left is
    recorded in the line that finally prints ("attempts")."""
    import subprocess
    rank = int(os.environ.get("RANK", "0"))
    base_port = int(os.environ.get("MASTER_PORT", "29500"))
    attempts = [dict(gather=args.gather, lanes=args.lanes, e2e=args.e2e)]
    # the fallback is the shape of this round's first 8-GPU run: one view in flight, NCCL gather, one copy to the host
    safe = dict(gather="sendrecv", lanes=1, e2e="gather")
    if attempts[0] != safe:
        attempts.append(safe)
    log = []
    for i, a in enumerate(attempts):
        # a port of its own per attempt, the same on every rank, away from the launcher's (whose next run may
        # well use base_port + 1)
        port = 31000 + (base_port + 17 * (i + 1)) % 2000
        env = dict(os.environ, C5_BENCH_CHILD="1", MASTER_PORT=str(port), TORCHELASTIC_USE_AGENT_STORE="False",
                   C5_BENCH_ATTEMPTS=json.dumps(log))
        cmd = [sys.executable, os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps", str(args.steps),
               "--warmup", str(args.warmup), "--gather", a["gather"], "--lanes", str(a["lanes"]), "--e2e", a["e2e"]]
        if args.dry_run_hostsim:
            cmd.append("--dry-run-hostsim")
        if args.no_cpu_baseline:
            cmd.append("--no-cpu-baseline")
        try:   # stderr passes through; the hard limit is a second line of defence behind the child's watchdog
            p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, text=True, preexec_fn=_die_with_parent,
                               timeout=float(os.environ.get("C5_BENCH_ATTEMPT_LIMIT", "600")))
        except subprocess.TimeoutExpired:
            p = subprocess.CompletedProcess(cmd, returncode=124, stdout="")
        if p.returncode == 0:
            if rank == 0:   # the JSON line only (NCCL prints its version banner on stdout)
                for out_line in p.stdout.splitlines():
                    if out_line.startswith("{"):
                        print(out_line, flush=True)
            return
        log.append(dict(a, exit_code=p.returncode))
        print(f"[bench rank {rank}] attempt {a} ended with exit code {p.returncode}"
              + ("; falling back" if i + 1 < len(attempts) else ""), file=sys.stderr, flush=True)
    raise SystemExit(3)


if __name__ == "__main__":
    main()
