#!/usr/bin/env python
"""bench.py — SVGF frame throughput (temporal + variance + 5 a-trous levels) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 4k|1080p|720p|8k] [--storage f16|f32]
  python bench.py --impl reference ...      # the reference's own kernels (oracle/_ref), else the CPU oracle

A "step" is one frame of the synthetic camera-pan sequence through svgf_frame.  Default
workload = BASELINE.json configs[2] (3840x2160 sequence, the configuration the <0.5 ms target is quoted on);
with --gpus N every rank filters its own independent 4K stream (weak scaling, no data-path collective —
BASELINE config 5's sharding).  Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {"720p": (1280, 720), "1080p": (1920, 1080), "4k": (3840, 2160), "8k": (7680, 4320)}
# algorithmic bytes per pixel per frame: each pass reads every plane it needs
# once and writes each output once, reference layouts, steady state.
BYTES_PER_PX = {"f16": {"temporal": 98, "variance": 17, "atrous_level": 40, "atrous_hist": 8},
                "f32": {"temporal": 130, "variance": 33, "atrous_level": 56, "atrous_hist": 16}}
IN_BYTES_PER_PX = {"f16": 8 + 8 + 16 + 8, "f32": 8 + 8 + 16 + 16}   # normal + uv + motion + noisy colour
OUT_BYTES_PER_PX = {"f16": 8, "f32": 16}


def ncu_traffic_per_launch(args):
    """Mean dram__bytes_read + dram__bytes_write per a-trous launch from the committed `ncu --set full` capture of this
    command (profiles/atrous_r01final.metrics.csv, five consecutive levels of one 4K fp16 frame), or None for any other
    workload: the capture is evidence for the default configuration only."""
    import csv
    if args.workload != "4k" or args.storage != "f16" or args.levels != 5 or args.flags or args.prefilter or args.reproj:
        return None, None
    path = os.path.join(ROOT, "profiles", "atrous_r01final.metrics.csv")
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", "atrous_r01s8.metrics.csv")
    try:
        rows = {r[0]: r for r in csv.reader(open(path)) if r}
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd, wr = rows["dram__bytes_read.sum"], rows["dram__bytes_write.sum"]
        per = [float(a) * scale[rd[1]] + float(b) * scale[wr[1]] for a, b in zip(rd[2:], wr[2:])]
        return int(sum(per) / len(per)), "profiles/%s (ncu --set full, mean of %d levels)" % (os.path.basename(path), len(per))
    except Exception:
        return None, None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Polls NVML for SM clock and throttle reasons while the timed region runs."""

    def __init__(self, index, period=0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.stop_flag = [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


def dist_setup(n_gpus, need_cuda=True):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        if need_cuda:
            raise RuntimeError("bench.py measures the CUDA path: no CUDA device is visible (there is no CPU implementation to fall back to)")
        return rank, world, local     # --impl reference without a GPU: the scalar port on the host cores
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    return rank, world, local


def barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_ranks(x, world):
    """x of every rank, in rank order (diagnostics: load balance of the band partition)."""
    if world == 1:
        return [x]
    import torch
    import torch.distributed as dist
    t = torch.zeros(world, dtype=torch.float64, device="cuda")
    t[dist.get_rank()] = x
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.tolist()]


def sum_over_ranks(x, world):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


class FrameRing:
    """R pre-generated frames of the pan sequence resident in HBM (procedural generator, CUDA)."""

    def __init__(self, W, H, R, storage, seed, device):
        import torch
        from svgf_b200 import synth
        from svgf_b200.filter import GBuffer
        self.R = R
        cdt = torch.float16 if storage == "f16" else torch.float32
        self.gbuf = [GBuffer(W, H, device) for _ in range(R)]
        self.colour = [torch.empty(H, W, 4, dtype=cdt, device=device) for _ in range(R)]
        for t in range(R):
            synth.frame_device(self.gbuf[t], self.colour[t], t, seed=seed)
        torch.cuda.synchronize()


def run_ours(args, rank, world, local):
    import torch
    from svgf_b200 import SvgfFilter, _lib
    from svgf_b200._lib import SvgfFrameBuffers, SvgfGBuffer
    W, H = WORKLOADS[args.workload]
    dev = torch.device("cuda", local)
    K, Wm = args.steps, args.warmup
    R = min(K + Wm, args.ring)
    ring = FrameRing(W, H, R, args.storage, seed=rank, device=dev)
    f = SvgfFilter(W, H, device=dev, storage=args.storage)
    f.SpatialFilterSteps = args.levels
    f.params.flags = args.flags
    f.params.variance_prefilter = args.prefilter
    f.params.reproj_mode = args.reproj
    lib = f.lib
    stream = torch.cuda.current_stream(dev)
    sptr = C.c_void_p(stream.cuda_stream)

    # per-frame argument structs: gbuf[P] / render[P] point straight at ring slot t, [Q] at slot t-1 (no copies)
    def frame_args(t):
        P = t & 1
        cur, prev = t % R, (t - 1) % R
        g = (SvgfGBuffer * 2)()
        g[P] = ring.gbuf[cur].as_struct()
        g[1 - P] = ring.gbuf[prev].as_struct()
        b = SvgfFrameBuffers()
        b.render[P] = ring.colour[cur].data_ptr()
        b.render[1 - P] = ring.colour[prev].data_ptr()
        for k in range(2):
            b.moments[k] = f.MomentsBuffer[k].data_ptr()
            b.filter[k] = f.FilterBuffer[k].data_ptr()
        b.history = f.HistoryLengthBuffer.data_ptr()
        b.ping_pong = P
        return g, b

    calls = [frame_args(t) for t in range(Wm + K)]
    f.Reset()

    def step(t):
        g, b = calls[t]
        st = lib.svgf_frame(f._ctx, C.byref(f.params), C.byref(g), C.byref(b), sptr)
        if st:
            raise RuntimeError(f"svgf_frame -> {st} (cuda {lib.svgf_last_cuda_error(f._ctx)})")

    for t in range(Wm):
        step(t)
    barrier(world)
    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    launches0 = f.launches
    f.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for t in range(Wm, Wm + K):
        step(t)
    e1.record(stream)
    barrier(world)
    sampler.stop_flag = True
    ms = e0.elapsed_time(e1)
    prof = f.profile_end()
    launches = f.launches - launches0
    sampler.join()
    ms_max = max_over_ranks(ms, world)
    value = world * W * H * K / (ms_max * 1e-3) / 1e9

    # ---- end to end through the host-buffer entry point: pinned host inputs, H2D + frame + D2H per step ----
    from svgf_b200 import synth
    import numpy as np
    n_host = 4
    cdt = torch.float16 if args.storage == "f16" else torch.float32
    host = []
    for t in range(n_host):
        hp = {"normal": torch.empty(H, W, 4, dtype=torch.int16).pin_memory(), "uv": torch.empty(H, W, 4, dtype=torch.int16).pin_memory(),
              "motion": torch.empty(H, W, 4, dtype=torch.float32).pin_memory(), "colour": torch.empty(H, W, 4, dtype=cdt).pin_memory()}
        src = ring.gbuf[t % R]
        synth.frame_device(src, ring.colour[t % R], t, seed=rank)       # regenerate (the timed run consumed the colour)
        hp["normal"].copy_(src.normal); hp["uv"].copy_(src.uv); hp["motion"].copy_(src.motion); hp["colour"].copy_(ring.colour[t % R])
        host.append(hp)
    result = torch.empty(H, W, 4, dtype=cdt).pin_memory()
    torch.cuda.synchronize()
    Ke = max(4, min(K, args.e2e_steps))

    def e2e_step(t, reset=False):
        hp = host[t % n_host]
        f.frame_host(hp["normal"], hp["uv"], hp["motion"], hp["colour"], result=result, reset=reset)

    for t in range(3):
        e2e_step(t, reset=(t == 0))
    barrier(world)
    e0.record(stream)
    for t in range(3, 3 + Ke):
        e2e_step(t)
    e1.record(stream)
    barrier(world)
    e2e_ms = max_over_ranks(e0.elapsed_time(e1), world)
    e2e_value = world * W * H * Ke / (e2e_ms * 1e-3) / 1e9
    checksum = float(result.float().sum())       # the host-side read of the step's result

    if rank != 0:
        return None
    peak, peak_src = measured_peaks()
    traffic, traffic_src = ncu_traffic_per_launch(args)
    bpp = BYTES_PER_PX[args.storage]
    n_levels = args.levels
    at_bytes = (bpp["atrous_level"] * n_levels + (bpp["atrous_hist"] if n_levels else 0)) * W * H
    at_launches = max(1, (launches - 2 * K) // K) if n_levels else 0
    at_ms_per_frame = prof["atrous_ms"] / max(1, prof["frames"])
    frame_bytes = (bpp["temporal"] + bpp["variance"]) * W * H + at_bytes
    achieved = at_bytes / (at_ms_per_frame * 1e-3) / 1e9 if n_levels and at_ms_per_frame > 0 else None
    line = {
        "metric": "svgf_frame_throughput", "value": round(value, 4), "unit": "Gpix/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": round(ms_max / K, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 compute, %s storage" % ("fp16" if args.storage == "f16" else "fp32"), "data": "synthetic",
        "config": {"workload": f"BASELINE configs[2]: {W}x{H} camera-pan sequence, temporal + variance + {n_levels} a-trous levels"
                               + (f"; {world} independent streams, one per GPU" if world > 1 else ""),
                   "width": W, "height": H, "atrous_levels": n_levels, "storage": args.storage, "frames_resident": R,
                   "l2": "every step reads a fresh frame (%.0f MB of inputs > 126 MB L2)" % (IN_BYTES_PER_PX[args.storage] * W * H / 1e6),
                   "params": "reference defaults (history 24, depth 0.8, normal 0.9, phi colour 10, phi normal 128)", "flags": args.flags,
                   "variance_prefilter": args.prefilter, "reproj_mode": args.reproj},
        "e2e": {"value": round(e2e_value, 4), "unit": "Gpix/s", "h2d_bytes_per_step": IN_BYTES_PER_PX[args.storage] * W * H,
                "d2h_bytes_per_step": OUT_BYTES_PER_PX[args.storage] * W * H, "ms_per_step": round(e2e_ms / Ke, 4), "steps": Ke,
                "api": "svgf_frame_host (pinned host buffers; copy-in, kernels and copy-out of consecutive frames overlap on three streams; timed region starts with the pipeline drained)", "result_checksum": checksum},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "a-trous levels (%d launches/frame)" % at_launches,
                     "achieved": round(achieved, 1) if achieved else None, "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                     "bytes_per_launch": int(at_bytes / max(1, at_launches)), "ms_per_launch": round(at_ms_per_frame / max(1, at_launches), 5),
                     "peak_source": peak_src},
        "frame_roofline": {"algorithmic_bytes": int(frame_bytes), "achieved": round(frame_bytes / (ms_max / K * 1e-3) / 1e9, 1),
                           "frac": round(frame_bytes / (ms_max / K * 1e-3) / 1e9 / peak, 4), "unit": "GB/s"},
        "stage_ms_per_frame": {k: round(prof[k] / max(1, prof["frames"]), 5) for k in ("temporal_ms", "variance_ms", "atrous_ms")},
        "clocks": sampler.result(),
    }
    return line


def run_bands(args, rank, world, local):
    """--mode bands (BASELINE configs[3]): ONE frame stream, every frame split into `world` horizontal bands with
    per-level halo rows exchanged between neighbouring ranks (NCCL send/recv over NVLink); strong scaling."""
    import torch
    from svgf_b200 import synth
    from svgf_b200.bands import balanced_bounds, make_gpu_banded_filter, required_apron
    from svgf_b200.filter import GBuffer
    W, H = WORKLOADS[args.workload]
    dev = torch.device("cuda", local)
    K, Wm = args.steps, args.warmup
    R = K + Wm
    # auto (-1): when the bands are tall, let a wide apron absorb every level's halo (one overlapped state exchange per
    # frame, no per-level exchange: the exchanges are latency- and host-bound, redundant rows are cheap); short bands keep
    # the 32-row apron and exchange before levels 3 and 4
    if args.band_exchange_from < 0:
        wide = required_apron(args.levels, args.levels, args.band_max_motion)
        wide = (wide + 7) // 8 * 8
        if H // world >= 4 * wide:
            args.band_exchange_from, args.band_apron = args.levels, max(args.band_apron, wide)
        else:
            args.band_exchange_from = min(3, args.levels)
    L = min(args.band_exchange_from, args.levels)
    cdt = torch.float16 if args.storage == "f16" else torch.float32
    # every rank generates the full frames procedurally on its own GPU and keeps only its local rows (band + aprons)
    full_g, full_c = GBuffer(W, H, dev), torch.empty(H, W, 4, dtype=cdt, device=dev)
    bounds = None
    if args.band_balance and world > 1:
        # equal estimated work per band: background pixels (linear depth 0) are passed through by the a-trous levels
        synth.frame_device(full_g, full_c, 0, seed=0)
        live = (full_g.motion[..., 2] != 0).float().mean(dim=1).cpu().numpy()
        bounds = balanced_bounds(live + args.band_bg_cost * (1.0 - live), world, min_rows=args.band_apron)
    bf = make_gpu_banded_filter(W, H, rank, world, dev, storage=args.storage, levels=args.levels, apron=args.band_apron,
                                exchange_from_level=L, max_motion_rows=args.band_max_motion, overlap_state=bool(args.band_overlap_state),
                                bounds=bounds)
    if args.band_dry_run:
        import svgf_b200.bands as _bands
        _bands.post_exchange = lambda *a_, **k_: []
    f, band = bf.f, bf.band
    Hl = band.local_height
    ring_g = [GBuffer(W, Hl, dev) for _ in range(R)]
    ring_c = [torch.empty(Hl, W, 4, dtype=cdt, device=dev) for _ in range(R)]
    for t in range(R):
        synth.frame_device(full_g, full_c, t, seed=0)
        sl = slice(band.ly0, band.ly1)
        ring_g[t].normal.copy_(full_g.normal[sl]); ring_g[t].uv.copy_(full_g.uv[sl]); ring_g[t].motion.copy_(full_g.motion[sl])
        ring_c[t].copy_(full_c[sl])
    del full_g, full_c
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream(dev)
    f.Reset()

    def step(t):
        P = f.PingPongInx
        f.Framebuffer[P], f.RenderBuffer[P] = ring_g[t], ring_c[t]      # inputs are consumed in place, no copies
        bf.Filter()
        bf.EndFrame()

    for t in range(Wm):
        step(t)
    barrier(world)
    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    launches0 = f.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for t in range(Wm, Wm + K):
        step(t)
    bf.drain()                      # a state exchange posted under the last frame's levels completes inside the timed region
    e1.record(stream)
    barrier(world)
    sampler.stop_flag = True
    ms_max = max_over_ranks(e0.elapsed_time(e1), world)
    # per-rank time: with --band-dry-run (no exchanges, wrong pixels) the ranks are independent and this is each band's own
    # compute + host cost, i.e. the load balance; otherwise neighbours wait for each other and the values converge
    ms_by_rank = all_ranks(e0.elapsed_time(e1) / K, world)
    launches = sum_over_ranks(f.launches - launches0, world)
    all_rows = [int(v) for v in all_ranks(float(Hl), world)]
    sampler.join()
    if rank != 0:
        return None
    peak, peak_src = measured_peaks()
    bpp = BYTES_PER_PX[args.storage]
    frame_bytes = (bpp["temporal"] + bpp["variance"] + bpp["atrous_level"] * args.levels + (bpp["atrous_hist"] if args.levels else 0)) * W * H
    halo_rows = sum(2 << i for i in range(L, args.levels))
    halo_bytes = 2 * (world - 1) * (halo_rows * W * OUT_BYTES_PER_PX[args.storage] + bf.state_apron * W * (OUT_BYTES_PER_PX[args.storage] * 3 // 2 + 1))
    value = W * H * K / (ms_max * 1e-3) / 1e9
    return {
        "metric": "svgf_frame_throughput", "value": round(value, 4), "unit": "Gpix/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": round(ms_max / K, 5), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 compute, %s storage" % ("fp16" if args.storage == "f16" else "fp32"), "data": "synthetic",
        "config": {"workload": f"BASELINE configs[3]: {W}x{H} frames in {world} horizontal band(s), temporal + variance + {args.levels} "
                               f"a-trous levels, halo exchange (NCCL send/recv) before levels >= {L}, state exchange "
                               f"{'overlapped with levels 1..' if args.band_overlap_state else 'at frame start'}", "width": W, "height": H,
                   "atrous_levels": args.levels, "storage": args.storage, "band_rows": band.y1 - band.y0, "apron_rows": bf.apron,
                   "band_bounds": bounds, "exchange_from_level": L, "overlap_state": bool(args.band_overlap_state), "exchanges_per_frame": 1 + args.levels - L, "local_rows_by_rank": all_rows,
                   "ms_per_step_by_rank": [round(v, 4) for v in ms_by_rank], "dry_run_no_exchange": bool(args.band_dry_run),
                   "halo_bytes_per_frame_all_ranks": int(halo_bytes),
                   "l2": "every step reads a fresh frame"},
        "gpu_launches": int(launches),
        "frame_roofline": {"algorithmic_bytes": int(frame_bytes), "achieved": round(frame_bytes / (ms_max / K * 1e-3) / 1e9, 1),
                           "frac_of_n_gpu_peak": round(frame_bytes / (ms_max / K * 1e-3) / 1e9 / (peak * world), 4), "unit": "GB/s",
                           "peak_source": peak_src},
        "clocks": sampler.result(),
    }


def cpu_baseline(args):
    """The scalar oracle on the host cores, on a bounded sample of the workload (rank 0, N = 1)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    from oracle_lib import OracleFilter, oracle
    from svgf_b200 import synth
    W, H = WORKLOADS[args.workload]
    # bounded sample: a horizontal band of the frame (full width) so that ~4 warm + 2 timed frames stay within ~20 s
    Hs = max(64, min(H, int(args.cpu_budget_px / W)))
    y0 = (H - Hs) // 2
    o = OracleFilter(W, Hs, storage=args.storage)
    o.params.atrous_iterations = args.levels
    o.Reset()
    n_warm, n_timed = 4, 2
    t_acc = 0.0
    for t in range(n_warm + n_timed):
        o.set_inputs(synth.frame_host(W, H, t, storage=args.storage, rows=(y0, y0 + Hs)))
        t0 = time.perf_counter()
        o.Filter()
        if t >= n_warm:
            t_acc += time.perf_counter() - t0
        o.EndFrame()
    cores = oracle().svgf_oracle_get_threads()
    return {"value": round(W * Hs * n_timed / t_acc / 1e9, 6), "unit": "Gpix/s", "cores": cores, "kind": "port",
            "sample": f"{n_timed} steady-state frames (after {n_warm} warm-up frames) of the middle {W}x{Hs} band of the {W}x{H} sequence, "
                      f"scalar C++ oracle, OpenMP over rows", "ms_per_frame_sample": round(t_acc / n_timed * 1e3, 1)}


def run_reference(args, rank, world, local):
    """--impl reference: the reference's own Filter.cuh kernels (oracle/_ref, patched for compilation only) on
    this GPU when the .so exists, else the scalar oracle port on the host cores."""
    if rank != 0:
        return None
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    W, H = WORKLOADS[args.workload]
    import numpy as np
    from oracle_lib import RefKernels, RefParams, ref, ref_available
    from svgf_b200 import _lib as L
    # n_gpus is the launch's N (the contract's key); the reference has no multi-GPU path, so one GPU does the work
    base = {"impl": "reference", "metric": "svgf_frame_throughput", "unit": "Gpix/s", "n_gpus": max(1, world), "gpus_used": 1, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic",
            "config": {"workload": f"BASELINE configs[2]: {W}x{H} camera-pan sequence, temporal + variance + {args.levels} a-trous levels",
                       "width": W, "height": H, "atrous_levels": args.levels, "storage": "f16"}}
    import torch as _torch
    if not ref_available() or args.storage != "f16" or not _torch.cuda.is_available():
        cb = cpu_baseline(args)
        base.update({"value": cb["value"], "ms_per_step": cb["ms_per_frame_sample"], "dtype": "f32/f64 compute, fp16 storage",
                     "cpu_baseline": cb, "gpu_launches": 0,
                     "e2e": {"value": cb["value"], "unit": "Gpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return base
    import torch
    from svgf_b200 import synth
    from svgf_b200.filter import GBuffer
    dev = torch.device("cuda", local)
    p = L.default_params()
    p.atrous_iterations = args.levels
    rp = RefParams.from_svgf(p, moments_quirk=0)
    r = RefKernels(W, H)
    K, Wm = args.steps, args.warmup
    R = min(K + Wm, args.ring)
    ring = FrameRing(W, H, R, "f16", seed=0, device=dev)
    D2D = 3
    tot_ms = 0.0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ms1 = C.c_float()
    for t in range(Wm + K):
        P = ref().svgf_ref_get_ping_pong(r.ctx)
        g = ring.gbuf[t % R]
        # the reference's producers write the cudaArrays / RenderBuffer directly; filling them is not filter time
        assert ref().svgf_ref_set_gbuffer(r.ctx, P, g.normal.data_ptr(), g.uv.data_ptr(), g.motion.data_ptr(), D2D) == 0
        assert ref().svgf_ref_set_plane(r.ctx, 0, P, ring.colour[t % R].data_ptr(), D2D) == 0
        assert ref().svgf_ref_time_frames(r.ctx, rp, 1, C.byref(ms1)) == 0
        if t >= Wm:
            tot_ms += ms1.value
    value = W * H * K / (tot_ms * 1e-3) / 1e9
    # end to end with host buffers
    host = []
    for t in range(4):
        g = ring.gbuf[t % R]
        synth.frame_device(g, ring.colour[t % R], t, seed=0)
        host.append({"normal": g.normal.cpu().pin_memory(), "uv": g.uv.cpu().pin_memory(), "motion": g.motion.cpu().pin_memory(),
                     "colour": ring.colour[t % R].cpu().pin_memory()})
    result = torch.empty(H, W, 4, dtype=torch.float16).pin_memory()
    ref().svgf_ref_reset(r.ctx)
    Ke = max(4, min(K, args.e2e_steps))
    torch.cuda.synchronize()
    t_e2e = 0.0
    for t in range(3 + Ke):
        hp = host[t % 4]
        t0 = time.perf_counter()
        assert ref().svgf_ref_frame_host(r.ctx, rp, hp["normal"].data_ptr(), hp["uv"].data_ptr(), hp["motion"].data_ptr(),
                                         hp["colour"].data_ptr(), result.data_ptr(), None) == 0
        torch.cuda.synchronize()
        if t >= 3:
            t_e2e += time.perf_counter() - t0
    e2e_value = W * H * Ke / t_e2e / 1e9
    base.update({
        "value": round(value, 4), "ms_per_step": round(tot_ms / K, 4), "dtype": "f32/f64 compute, fp16 storage",
        "gpu_launches": (2 + args.levels) * K,
        "cpu_baseline": {"value": round(value, 4), "unit": "Gpix/s", "cores": 0, "kind": "reference",
                         "sample": "the reference is a CUDA program with no CPU implementation of this path: its own kernels "
                                   "(src/Filter.cuh:359-624, six by-value signature edits to compile, math untouched) built for sm_100a "
                                   "and launched like src/App.cu:469-514 on this B200; every step of the workload"},
        "e2e": {"value": round(e2e_value, 4), "unit": "Gpix/s", "h2d_bytes_per_step": IN_BYTES_PER_PX["f16"] * W * H,
                "d2h_bytes_per_step": OUT_BYTES_PER_PX["f16"] * W * H, "ms_per_step": round(t_e2e / Ke * 1e3, 4),
                "api": "reference kernels behind host buffers (cudaArray uploads + stages + result download)"},
    })
    r.close()
    return base


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=192)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="4k", choices=sorted(WORKLOADS))
    ap.add_argument("--size", default=None, help="WxH override of --workload (diagnostics; e.g. 7680x2240 = one band of an 8K frame)")
    ap.add_argument("--storage", default="f16", choices=["f16", "f32"])
    ap.add_argument("--levels", type=int, default=5)
    ap.add_argument("--ring", type=int, default=200, help="max distinct frames kept resident in HBM")
    ap.add_argument("--e2e-steps", type=int, default=48)
    ap.add_argument("--cpu-budget-px", type=float, default=1.6e6, help="pixels per frame of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--prefilter", type=int, default=0, help="svgf_params.variance_prefilter (1 = 3x3 Gaussian, not in the reference)")
    ap.add_argument("--reproj", type=int, default=0, help="svgf_params.reproj_mode (1 = bilinear 2x2, not in the reference)")
    ap.add_argument("--flags", type=int, default=0, help="svgf_params.flags for A/B runs (8 = no uniform-normal tile shortcut)")
    ap.add_argument("--band-dry-run", type=int, default=0, help="--mode bands diagnostics: 1 = skip every exchange (pixels near band edges are wrong); shows per-rank load")
    ap.add_argument("--band-balance", type=int, default=1, help="--mode bands: 1 = band heights balanced by estimated work (background rows are cheap), 0 = equal heights")
    ap.add_argument("--band-bg-cost", type=float, default=0.45, help="--mode bands: cost of a background pixel relative to a filtered one")
    ap.add_argument("--band-apron", type=int, default=32, help="--mode bands: apron rows on each side of a band")
    ap.add_argument("--band-exchange-from", type=int, default=-1, help="--mode bands: first a-trous level that exchanges its halo (lower levels recompute it in the apron); -1 = choose from the band height")
    ap.add_argument("--band-max-motion", type=int, default=8, help="--mode bands: vertical reach (rows) of the temporal gather covered by the apron")
    ap.add_argument("--band-overlap-state", type=int, default=1, help="--mode bands: 1 = post the previous-frame state exchange under levels 1..N-1")
    ap.add_argument("--mode", default="streams", choices=["streams", "bands"],
                    help="multi-GPU sharding: independent frame streams per GPU (weak scaling, default) or one frame in "
                         "horizontal bands with per-level halo exchange (strong scaling; BASELINE configs[3], use --workload 8k)")
    args = ap.parse_args()
    if args.size:
        w_, h_ = (int(v) for v in args.size.lower().split("x"))
        WORKLOADS[args.size] = (w_, h_)
        args.workload = args.size
    if args.warmup < 3:
        args.warmup = 3
    if args.workload == "8k":
        args.ring = min(args.ring, 48)

    import __graft_entry__ as g
    if not (os.path.exists(os.path.join(ROOT, "svgf_b200", "libsvgf_b200.so")) and os.path.exists(os.path.join(ROOT, "oracle", "libsvgf_oracle.so"))):
        g.build()
    rank, world, local = dist_setup(args.gpus, need_cuda=(args.impl != "reference"))
    if args.impl == "reference":
        line = run_reference(args, rank, world, local)
    elif args.mode == "bands":
        line = run_bands(args, rank, world, local)
    else:
        line = run_ours(args, rank, world, local)
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
    if rank == 0 and line is not None:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
