#!/usr/bin/env python
"""bench.py : photons propagated / s on N B200s (BASELINE.json metric), one JSON line on rank 0.

    python bench.py --gpus 1 --steps K --warmup W            # this engine (libphox.so)
    python bench.py --impl reference ...                     # the CPU arm (oracle port, host threads)
    torchrun --nproc-per-node N bench.py --gpus N ...        # one rank per GPU, weak scaling

A "step" is one event: the hot path (genstep -> photons -> bounce loop -> hits) over one batch of
synthetic gensteps.  Default workload = the configuration the north_star target is quoted on, the
8x8 CsI+SiPM scintillation geometry (BASELINE config 3) with 12.5 M photons per GPU
(100 M / 8 GPUs); --workload selects the other configs.

    value   photons/s, whole job, gensteps already resident in HBM, hits left on the device
            (phox_simulate_device), timed with CUDA events on the launch stream, max over ranks
    e2e     same metric through the public host-buffer API (Simulator.simulate_np): gensteps in
            pinned host memory -> H2D -> simulate -> hits D2H, every step
    roofline  HBM roofline of the dominant kernel (k_wf_propagate, ~3/4 of the bounce loop, on the bench
            workload): its algorithmic bytes from the event's own counts (168 B per live photon + 16 B
            per survivor + 24 B per ray its home cell settles) / average launch
            duration, from CUDA events the library records between the kernels on the launch stream (a
            separate pass of <= 3 steps with phox_set_profiling on, right after the timed region).  The path-level figure of
            SURVEY 8(d), 132 + 128 f_hit bytes per photon over the whole bounce loop, is reported
            beside it (path_*)
    cpu_baseline  the CPU oracle (a port, NOT Geant4 and NOT the OptiX build - neither installs
            here) on a bounded sample of the same workload, all host threads
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region through NVML (the same counters
    `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints, B200_PROFILING.md recipe), from a
    thread so that no process is spawned next to the measurement."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index = index
        self.sm, self.reasons, self.power = [], set(), []
        self.stop_flag = threading.Event()
        self.recording = threading.Event()
        self.thread = None
        self.smax = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            # first use of each query can block the driver for a long time on a fresh box: pay that here, outside the timed region
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            (pynvml.nvmlDeviceGetCurrentClocksEventReasons if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons")
             else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons)(self.h)
            pynvml.nvmlDeviceGetPowerUsage(self.h)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                if self.recording.is_set():
                    self.sm.append(sm)
                    self.power.append(pw)
                    for bit, name in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def start(self):
        """start the sampling thread (before the warm-up: the first NVML calls made from a new thread block the
        driver for 50-150 ms, measured, which must not fall into the timed region); samples are kept only between
        begin() and stop()"""
        if self.nv is None:
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def begin(self):
        self.recording.set()

    def stop(self):
        if self.nv is None or self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.stop_flag.set()
        self.thread.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.smax, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "power_w_max": max(self.power) if self.power else None}


def host_threads():
    """host threads this process may run on.  Asked for explicitly: under torch.distributed.run the environment carries
    OMP_NUM_THREADS=1, which would otherwise turn the CPU arm into a single-thread run (SCALE_r01: 41 k photons/s, timeouts)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def oracle_throughput(w, sample_photons, nthreads=0):
    """time the CPU oracle on the first gensteps of the workload holding ~sample_photons photons"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _ref import Oracle
    gs = w["gensteps"]
    num = gs.view(np.uint32)[:, 0, 3].astype(np.int64)
    ip = w["input_photons"]
    if ip is not None:
        n = min(sample_photons, len(ip))
        g = gs[:1].copy(); g.view(np.uint32)[0, 0, 3] = n
        sub, ipn = g, ip[:n]
    else:
        k = int(np.searchsorted(np.cumsum(num), sample_photons)) + 1
        sub, ipn = gs[:k], None
        n = int(num[:k].sum())
    orc = Oracle()
    threads = host_threads() if nthreads <= 0 else nthreads
    t0 = time.perf_counter()
    r = orc.simulate(w["geom"], sub, ipn, max_bounce=w["config"].get("max_bounce", 31), use_boxes=2, nthreads=threads, arrays=False)
    dt = time.perf_counter() - t0
    return dict(value=n / dt, seconds=dt, photons=n, cores=threads, rays=r["nray"], hits=r["nhit"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="phox", choices=["phox", "reference"])
    ap.add_argument("--workload", default="sipm8x8_scint")
    ap.add_argument("--photons", type=int, default=12_500_000, help="photons per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=6_000_000, help="photons of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="diagnosis: do not sample NVML clocks during the timed region")
    ap.add_argument("--max-slot", type=int, default=0, help="photons per launch (0 = library default); an event is sliced at genstep granularity")
    ap.add_argument("--accel", default="bvh", choices=["bvh", "nohome", "brute"], help="bvh = two-level BVH + home cells (default); nohome = BVH alone (A/B); brute = validation loop")
    ap.add_argument("--kernel-mode", default="auto", choices=["auto", "persistent", "wavefront"], help="form of the bounce loop (include/phox.h PHOX_KERNEL_*)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warm = max(args.warmup, 3) if args.impl == "phox" else max(args.warmup, 0)

    from eic_opticks_b200 import workloads

    # ---------------- reference arm: CPU implementation of the path on the host cores ----------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        w = workloads.WORKLOADS[args.workload](num_photon=min(args.photons, args.cpu_sample * 2))
        for _ in range(min(warm, 1)):
            oracle_throughput(w, args.cpu_sample // 4)
        vals = [oracle_throughput(w, args.cpu_sample) for _ in range(max(args.steps, 1))]
        tot_ph = sum(v["photons"] for v in vals); tot_s = sum(v["seconds"] for v in vals)
        value = tot_ph / tot_s
        sample = "%d photons (first gensteps of the %s workload) per step, prim tests culled by box trees over instances and prims" % (vals[0]["photons"], args.workload)
        print(json.dumps({
            "impl": "reference", "metric": "photons propagated/sec", "value": value, "unit": "photons/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / len(vals), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": args.workload, "photons_per_gpu_per_step": args.photons, "cpu_sample_photons_per_step": vals[0]["photons"],
                                                            "max_bounce": w["config"].get("max_bounce", 31), "host_threads": vals[0]["cores"],
                                                            "note": "CPU oracle port of the reference path (Geant4 and the OptiX build cannot be installed here); each step is a bounded sample "
                                                                    "of the GPU arm's workload (its first gensteps), the rate is per photon"},
            "cpu_baseline": {"value": value, "unit": "photons/s", "cores": vals[0]["cores"], "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "photons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return 0

    # ---------------- this engine ------------------------------------------------------------------
    import torch
    import torch.distributed as dist
    import eic_opticks_b200 as ph
    from eic_opticks_b200 import parallel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # weak scaling: the global event has world * photons; rank r takes its contiguous genstep share
    big_file = None
    if args.workload == "pmt_wall_torch" and args.photons > 32_000_000:
        # BASELINE config 4 at its stated size (1 G photons / 8 GPUs = 125 M per rank): the photon-file contents of this rank's
        # photon range are drawn on the device with torch (same disc source: r = R u1, phi = 2 pi u2, straight down, 420 nm, zero
        # flags) - the host Philox restatement of src/torch.cpp needs a minute per rank at this size.  Synthetic either way.
        from eic_opticks_b200 import gensteps as G
        w = workloads.WORKLOADS[args.workload](num_photon=1000)
        hx, hy = w["geom"]["half"]
        gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
        u = torch.rand((args.photons, 2), generator=gen, device=dev, dtype=torch.float32)
        r, phi = (min(hx, hy) - 600.0) * u[:, 0], 6.283185307179586 * u[:, 1]
        big_file = torch.zeros((args.photons, 4, 4), dtype=torch.float32, device=dev)
        big_file[:, 0, 0] = r * torch.cos(phi); big_file[:, 0, 1] = r * torch.sin(phi); big_file[:, 0, 2] = 1500.0
        big_file[:, 1, 2] = -1.0
        big_file[:, 2, 0] = torch.sin(phi); big_file[:, 2, 1] = -torch.cos(phi); big_file[:, 2, 3] = 420.0
        del u, r, phi
        big_host = torch.empty((args.photons, 4, 4), dtype=torch.float32).pin_memory()      # the ONE host copy (8 GB per rank at 125 M photons)
        big_host.copy_(big_file)
        gs_r, ip_r, off_r, cnt_r = G.input_photon_genstep(args.photons), big_host.numpy(), rank * args.photons, args.photons
    elif world > 1 and args.workload in ("pmt_wall_torch", "sphere_leak_torch"):
        # input-photon workloads: every rank makes its own photon-range of the global event (its own Philox seed) instead of the
        # whole array - at BASELINE config 4's size (1 G photons) that array is 64 GB per process
        w = workloads.WORKLOADS[args.workload](num_photon=args.photons, seed=rank)
        gs_r, ip_r, off_r, cnt_r = w["gensteps"], w["input_photons"], rank * args.photons, args.photons
    else:
        w = workloads.WORKLOADS[args.workload](num_photon=args.photons * world)
        gs_r, ip_r, off_r, cnt_r = parallel.shard_event(w["gensteps"], rank, world, w["input_photons"])
    g = w["geom"]
    kmode = {"auto": 0, "persistent": 1, "wavefront": 2}[args.kernel_mode]
    sim = ph.Simulator.Create(g["foundry"], g["bnd"], g["optical"], g["icdf"], device=local_rank, event_mode=ph.MODE_MINIMAL, kernel_mode=kmode, **w["config"])
    if args.max_slot > 0:
        sim.set_config(max_slot=args.max_slot)
    if args.accel != "bvh":
        sim.set_config(accel={"nohome": ph.ACCEL_BVH_NOHOME, "brute": ph.ACCEL_BRUTE}[args.accel])
    stream = torch.cuda.current_stream(dev)
    sim.set_stream(stream.cuda_stream)

    d_gs = torch.from_numpy(gs_r).to(dev)
    d_ip = (big_file if big_file is not None else torch.from_numpy(ip_r).to(dev)) if ip_r is not None else None
    h_gs = torch.from_numpy(gs_r).pin_memory()
    h_ip = (big_host if big_file is not None else torch.from_numpy(ip_r).pin_memory()) if ip_r is not None else None
    # pinned destination of the e2e hits: every photon could be a hit, except at the 125 M-photon size where a quarter is plenty
    # (11.6 % of the PMT-wall photons are detected) and 8 ranks x 8 GB of pinned memory would not be
    h_hits = torch.empty((max(cnt_r if big_file is None else cnt_r // 4, 1), 4, 4), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)     # 256 MB > 126 MB L2

    # N > 1: every event's hits end up on rank 0 (the reference hands hits to one host process).  The gather of event k is posted
    # on a second stream after event k+1 was launched and travels while it runs; staging and receive buffers are allocated once.
    gather = parallel.PipelinedHitGather(max(1024, cnt_r // 4), dev, dst=0) if world > 1 else None

    def step_device(event_id):
        sim.simulate_device(d_gs.data_ptr(), len(gs_r), d_ip.data_ptr() if d_ip is not None else 0, 0 if ip_r is None else len(ip_r), event_id, off_r)
        if gather is not None:
            gather.push(sim)

    def step_e2e(event_id):
        gs_np = h_gs.numpy(); ip_np = h_ip.numpy() if h_ip is not None else None
        sim.simulate_np_into(gs_np, event_id, ip_np, off_r, h_hits.numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, sampler=None):
        st_sum = dict(num_kernel=0, simulate_kernel_seconds=0.0, compact_kernel_seconds=0.0, num_ray=0, num_hit=0, num_launch=0, num_home_ray=0)
        barrier()
        if sampler:
            sampler.begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms = 0.0
        per_step = []
        for k in range(steps):
            flush.fill_(float(k))                      # evict L2 between timed iterations (not timed)
            torch.cuda.synchronize(dev)
            e0.record(stream)
            fn(1 + k)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            per_step.append(e0.elapsed_time(e1))
            ms += per_step[-1]
            st = sim.stats()
            for key in st_sum:
                st_sum[key] += st[key]
        if gather is not None and fn is step_device:
            # the last event's gather is part of the job: post it and wait for it inside the timed region
            t0 = time.perf_counter()
            gather.drain()                             # posts the gather of the last event and synchronises its stream
            drain_ms = 1e3 * (time.perf_counter() - t0)
            ms += drain_ms
            st_sum["gather_drain_ms"] = round(drain_ms, 3)
        barrier()
        clocks = sampler.stop() if sampler else None
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        st_sum["step_ms"] = [round(v, 3) for v in per_step]
        return float(t.item()), st_sum, clocks

    sampler = ClockSampler(local_rank)
    if not args.no_clocks:
        sampler.start()
    for k in range(warm):
        step_device(1 + k)           # same event ids as the timed steps: every buffer reaches its steady-state size here
    for k in range(2):
        step_e2e(1 + k)
    ms_dev, st_dev, clocks = timed(step_device, args.steps, None if args.no_clocks else sampler)
    ms_e2e, st_e2e, _ = timed(step_e2e, args.steps)

    tot = torch.tensor([float(cnt_r)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    photons_per_step = float(tot.item())
    value = photons_per_step * args.steps / (ms_dev * 1e-3)
    e2e = photons_per_step * args.steps / (ms_e2e * 1e-3)

    # per-kernel pass (not part of `value`): CUDA events between the kernels of the bounce loop, on the launch stream
    sim.set_profiling(True)
    prof = dict(trace_kernel_seconds=0.0, propagate_kernel_seconds=0.0, num_trace_launch=0, num_ray=0, num_home_ray=0, simulate_kernel_seconds=0.0)
    prof_photons = 0
    for k in range(min(args.steps, 3)):
        prof_photons += cnt_r
        flush.fill_(float(k)); torch.cuda.synchronize(dev)
        step_device(100 + k)
        torch.cuda.synchronize(dev)
        st = sim.stats()
        for key in prof:
            prof[key] += st[key]
    sim.set_profiling(False)

    if rank == 0:
        peak, peak_kind = load_peaks()
        f_hit = st_dev["num_hit"] / max(1, cnt_r * args.steps)
        bytes_per_photon = 132.0 + 128.0 * f_hit if ip_r is None else 196.0 + 128.0 * f_hit
        loop_s = st_dev["simulate_kernel_seconds"] / max(1, st_dev["num_launch"])
        wave = prof["num_trace_launch"] > 0
        traffic_src, tj = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic_r2b.json")
        if os.path.exists(tpath) and args.workload == "sipm8x8_scint" and args.accel == "bvh":
            with open(tpath) as f:
                tj = json.load(f)
            traffic_src = "profiles/traffic_r2b.json: ncu dram__bytes_read+write of one launch of the kernel / its live photons, x live photons per launch here"
        second = None
        if wave:
            # Two kernels per bounce.  Algorithmic bytes, from the event's own counts:
            #  k_wf_propagate (physics + home-cell pass), per live photon: 104 B read (list entry 4, hit record 32, photon 64 - the draw
            #    count travels in its index word -, home 4) + 64 B written (photon); per survivor 8 B (next list entry + its home) + 32 B
            #    (hit record of the next bounce, when its home cell settles the ray) or 8 B (pending-list entry: list position + list entry, when it does not).
            #    Without home cells: 100 B read + 64 B written per live photon, 4 B per survivor.
            #  k_wf_trace (BVH traversal of the pending rays), per ray: 44 B read (pending entry 4, list entry 4, position/time/direction 32,
            #    home 4) + 32 B written (hit record)
            L = prof["num_trace_launch"]
            live = prof["num_ray"]                                   # live photons summed over the bounces = rays of the event(s)
            surv = max(0, prof["num_ray"] - prof_photons)            # survivors = the live photons of bounces 1 ..
            home = prof["num_home_ray"]
            prop_bytes = (168.0 * live + 16.0 * surv + 24.0 * home) if home > 0 else (164.0 * live + 4.0 * surv)
            trace_rays = live - home
            trace_bytes = 76.0 * trace_rays
            prop_s = prof["propagate_kernel_seconds"] / L
            trace_s = prof["trace_kernel_seconds"] / L
            loop_prof_s = max(1e-12, prof["simulate_kernel_seconds"])
            kp = {"kernel": "k_wf_propagate", "bound": "hbm", "achieved": prop_bytes / L / prop_s / 1e9, "peak": peak, "unit": "GB/s",
                  "kernel_ms": prop_s * 1e3, "algorithmic_bytes_per_photon": prop_bytes / max(1, live), "photons_per_launch": live / L,
                  "kernel_share_of_bounce_loop": prof["propagate_kernel_seconds"] / loop_prof_s,
                  "traffic": (tj["k_wf_propagate"]["dram_bytes_per_photon"] * live / L) if tj else None}
            kt = {"kernel": "k_wf_trace", "bound": "hbm", "achieved": trace_bytes / L / trace_s / 1e9, "peak": peak, "unit": "GB/s",
                  "kernel_ms": trace_s * 1e3, "algorithmic_bytes_per_ray": 76.0, "rays_per_launch": trace_rays / L,
                  "kernel_share_of_bounce_loop": prof["trace_kernel_seconds"] / loop_prof_s,
                  "traffic": (tj["k_wf_trace"]["dram_bytes_per_ray"] * trace_rays / L) if tj else None}
            for k in (kp, kt):
                k["frac"] = k["achieved"] / peak
            dom, second = (kp, kt) if prop_s >= trace_s else (kt, kp)
        else:
            dom = {"kernel": "k_simulate", "bound": "hbm", "achieved": cnt_r * bytes_per_photon / loop_s / 1e9, "peak": peak, "unit": "GB/s",
                   "kernel_ms": loop_s * 1e3, "kernel_share_of_bounce_loop": 1.0, "traffic": None}
            dom["frac"] = dom["achieved"] / peak
        out = {
            "metric": "photons propagated/sec", "value": value, "unit": "photons/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "photons_per_gpu_per_step": cnt_r, "gensteps_per_gpu": int(len(gs_r)), "max_bounce": sim.cfg.max_bounce,
                       "event_mode": "Minimal", "rng_mode": "DEBUG_TAG", "accel": {"bvh": "two-level BVH + home cells", "nohome": "two-level BVH", "brute": "brute force"}[args.accel], "kernel_mode": args.kernel_mode, "max_slot": args.max_slot,
                       "l2": "256 MB flush between timed steps",
                       "sharding": "gensteps partitioned over ranks, absolute photon offsets, every event's hits gathered to rank 0 (NCCL send/recv on a second stream, "
                                   "overlapped with the next event; the last gather is drained inside the timed region)" if world > 1 else "single GPU"},
            "rays_per_s": st_dev["num_ray"] * world / (ms_dev * 1e-3), "bounces_per_photon": st_dev["num_ray"] / max(1, cnt_r * args.steps),
            "hit_fraction": f_hit, "home_ray_fraction": st_dev["num_home_ray"] / max(1, st_dev["num_ray"]), "step_ms": st_dev["step_ms"], "e2e_step_ms": st_e2e["step_ms"],
            "e2e": {"value": e2e, "unit": "photons/s", "h2d_bytes_per_step": int(gs_r.nbytes + (ip_r.nbytes if ip_r is not None else 0)),
                    "d2h_bytes_per_step": int(64 * st_e2e["num_hit"] / max(1, args.steps)), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(st_dev["num_kernel"] + st_e2e["num_kernel"]),
            "roofline": dict(dom, peak_source=peak_kind, traffic_source=traffic_src,
                             rays_per_launch_all=(prof["num_ray"] / max(1, prof["num_trace_launch"])) if wave else None,
                             bounce_loop_ms=loop_s * 1e3, bounce_loop_share_of_step=st_dev["simulate_kernel_seconds"] / (ms_dev * 1e-3),
                             path_algorithmic_bytes_per_photon=bytes_per_photon, path_achieved_gbs=cnt_r * bytes_per_photon / loop_s / 1e9,
                             second_kernel=second,
                             note="k_wf_propagate = physics + home-cell candidate pass (HBM / latency bound); k_wf_trace = BVH traversal of the "
                                  "rays the home cells did not settle (issue / latency bound); ncu traffic and stall breakdown in profiles/"),
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:         # N = 1 only: at N > 1 the other ranks would idle in a barrier behind a CPU loop
            wc = workloads.WORKLOADS[args.workload](num_photon=min(args.photons, args.cpu_sample * 2))
            cb = oracle_throughput(wc, args.cpu_sample)
            out["cpu_baseline"] = {"value": cb["value"], "unit": "photons/s", "cores": cb["cores"], "kind": "port",
                                   "sample": "%d photons of the same workload, CPU oracle (not Geant4), %.1f s" % (cb["photons"], cb["seconds"])}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sim.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
