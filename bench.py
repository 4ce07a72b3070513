#!/usr/bin/env python
"""bench.py — SVGF frame throughput (temporal + variance + 5 a-trous levels) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 4k|1080p|720p|8k] [--storage f16|f32]
  python bench.py --impl reference ...      # the reference's own kernels (oracle/_ref), else the CPU oracle

A "step" is one frame of the synthetic camera-pan sequence through svgf_frame.  Default workload =
BASELINE.json configs[2] (3840x2160 sequence, the configuration the <0.5 ms target is quoted on); with --gpus N every rank
filters its own independent 4K stream (weak scaling, no data-path collective).  Prints ONE JSON line (rank 0) carrying,
besides the contract's keys: `roofline`, `cpu_baseline`, `parity` (the oracle on a band of the same frames, teacher-forced),
`general_case` (no uniform-normal tile shortcut), `e2e` with its PCIe ceiling, `stage_ms_per_frame` (incl. the TAA resolve),
`streams_1080p` (BASELINE configs[4]: 64 independent 1080p streams over the N GPUs, concurrent CUDA streams per GPU) and,
for N > 1, `bands` (BASELINE configs[3]: 8K frames in N bands through the native NCCL band driver, with a bit-identity check
of the stitched bands against the whole frame filtered on one GPU).
"""
import argparse
import ctypes as C
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {"720p": (1280, 720), "1080p": (1920, 1080), "4k": (3840, 2160), "8k": (7680, 4320)}
# algorithmic bytes per pixel per frame: each pass reads every plane it needs
# once and writes each output once, reference layouts, steady state.
BYTES_PER_PX = {"f16": {"temporal": 98, "variance": 17, "atrous_level": 40, "atrous_hist": 8, "taa": 24},
                "f32": {"temporal": 130, "variance": 33, "atrous_level": 56, "atrous_hist": 16, "taa": 48}}
IN_BYTES_PER_PX = {"f16": 8 + 8 + 16 + 8, "f32": 8 + 8 + 16 + 16}   # normal + uv + motion + noisy colour
OUT_BYTES_PER_PX = {"f16": 8, "f32": 16}
MIN_TIMED_MS = 100.0          # the K-step timed region is repeated until at least this much device time has been measured


def workload_name(W, H, levels, world=1, kind="streams"):
    base = f"BASELINE configs[2]: {W}x{H} camera-pan sequence, temporal + variance + {levels} a-trous levels"
    return base + (f"; {world} independent streams, one per GPU" if world > 1 and kind == "streams" else "")


def base_config(W, H, levels, storage):
    """config keys shared by both arms (the driver compares them)."""
    return {"workload": workload_name(W, H, levels), "width": W, "height": H, "atrous_levels": levels, "storage": storage,
            "params": "reference defaults (history 24, depth 0.8, normal 0.9, phi colour 10, phi normal 128)", "seed": 0}


def ncu_traffic_per_launch(args):
    """Mean dram__bytes_read + dram__bytes_write per a-trous launch from the committed `ncu --set full` capture of this
    command (five consecutive levels of one 4K fp16 frame), or None for any other workload: the capture is evidence for the
    default configuration only."""
    import csv
    if args.workload != "4k" or args.storage != "f16" or args.levels != 5 or args.flags or args.prefilter or args.reproj:
        return None, None
    for name in ("atrous_r02.metrics.csv", "atrous_r01final.metrics.csv"):
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            continue
        try:
            rows = {r[0]: r for r in csv.reader(open(path)) if r}
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd, wr = rows["dram__bytes_read.sum"], rows["dram__bytes_write.sum"]
            per = [float(a) * scale[rd[1]] + float(b) * scale[wr[1]] for a, b in zip(rd[2:], wr[2:])]
            return int(sum(per) / len(per)), "profiles/%s (ncu --set full, mean of %d levels)" % (name, len(per))
        except Exception:
            pass
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


def _reduce(x, world, op):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
    return float(t.item())


def max_over_ranks(x, world):
    return _reduce(x, world, "MAX")


def sum_over_ranks(x, world):
    return _reduce(x, world, "SUM")


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


class stdout_to_stderr:
    """NCCL prints its version banner on stdout at communicator creation when NCCL_DEBUG is VERSION or WARN; this line must
    stay the only thing bench.py writes there.  Redirects file descriptor 1 (C-level writes included) for the block."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


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
        self.seed = seed
        for t in range(R):
            synth.frame_device(self.gbuf[t], self.colour[t], t, seed=seed)
        torch.cuda.synchronize()

    def restore_colour(self, lo, hi):
        """The filter consumes the noisy radiance in place; regenerate frames lo..hi-1 (outside any timed region)."""
        from svgf_b200 import synth
        for t in range(lo, hi):
            synth.frame_device(self.gbuf[t % self.R], self.colour[t % self.R], t, seed=self.seed)


class DeviceSequence:
    """A SvgfFilter stepping through a FrameRing with zero copies: gbuf[P] / render[P] point straight at ring slot t."""

    def __init__(self, W, H, dev, storage, ring, levels, flags=0, prefilter=0, reproj=0):
        from svgf_b200 import SvgfFilter
        from svgf_b200._lib import SvgfFrameBuffers, SvgfGBuffer
        self.f = SvgfFilter(W, H, device=dev, storage=storage)
        self.f.SpatialFilterSteps = levels
        self.f.params.flags, self.f.params.variance_prefilter, self.f.params.reproj_mode = flags, prefilter, reproj
        self.ring, self.R = ring, ring.R
        self._G, self._B = SvgfGBuffer, SvgfFrameBuffers
        self.calls = {}

    def args(self, t):
        if t not in self.calls:
            f, ring, R = self.f, self.ring, self.R
            P = t & 1
            cur, prev = t % R, (t - 1) % R
            g = (self._G * 2)()
            g[P] = ring.gbuf[cur].as_struct()
            g[1 - P] = ring.gbuf[prev].as_struct()
            b = self._B()
            b.render[P] = ring.colour[cur].data_ptr()
            b.render[1 - P] = ring.colour[prev].data_ptr()
            for k in range(2):
                b.moments[k] = f.MomentsBuffer[k].data_ptr()
                b.filter[k] = f.FilterBuffer[k].data_ptr()
            b.history = f.HistoryLengthBuffer.data_ptr()
            b.ping_pong = P
            self.calls[t] = (g, b)
        return self.calls[t]

    def step(self, t, sptr):
        g, b = self.args(t)
        f = self.f
        st = f.lib.svgf_frame(f._ctx, C.byref(f.params), C.byref(g), C.byref(b), sptr)
        if st:
            raise RuntimeError(f"svgf_frame -> {st} (cuda {f.lib.svgf_last_cuda_error(f._ctx)})")


def pcie_ceiling(dev, h2d_bytes, d2h_bytes):
    """Host<->device copy bandwidth of this GPU from pinned memory: each direction alone and both at once.  The ceiling of
    the e2e figure is the larger of the two directions' transfer times at the ALONE rates (a pipelined frame cannot finish
    faster than its bigger copy); the concurrent rates show what full-duplex contention costs."""
    import torch
    n = 256 << 20
    hin, hout = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    din, dout = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def run(do_in, do_out):
        best = [0.0, 0.0]
        for it in range(3):
            torch.cuda.synchronize()
            if do_in:
                with torch.cuda.stream(s1):
                    ev[0].record(s1)
                    for _ in range(3):
                        din.copy_(hin, non_blocking=True)
                    ev[1].record(s1)
            if do_out:
                with torch.cuda.stream(s2):
                    ev[2].record(s2)
                    for _ in range(3):
                        hout.copy_(dout, non_blocking=True)
                    ev[3].record(s2)
            torch.cuda.synchronize()
            if it:
                if do_in:
                    best[0] = max(best[0], 3 * n / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9)
                if do_out:
                    best[1] = max(best[1], 3 * n / (ev[2].elapsed_time(ev[3]) * 1e-3) / 1e9)
        return best

    h2d_alone = run(True, False)[0]
    d2h_alone = run(False, True)[1]
    both = run(True, True)
    step_ms = max(h2d_bytes / (h2d_alone * 1e9), d2h_bytes / (d2h_alone * 1e9)) * 1e3
    return {"h2d_gbs": round(h2d_alone, 2), "d2h_gbs": round(d2h_alone, 2), "h2d_gbs_full_duplex": round(both[0], 2),
            "d2h_gbs_full_duplex": round(both[1], 2), "ms_per_step_at_ceiling": round(step_ms, 4),
            "how": "3 x 256 MiB pinned copies per direction, best of 2 after a warm-up; ceiling = the slower direction's bytes per step at its stand-alone rate"}


def run_streams_1080p(args, rank, world, local):
    """BASELINE configs[4]: 64 independent 1080p streams sharded over the N GPUs (stream s on GPU s mod N), every stream
    with its own context, state and CUDA stream; frames of the resident streams of a GPU are issued round-robin."""
    import torch
    W, H = WORKLOADS["1080p"]
    dev = torch.device("cuda", local)
    total = 64
    mine = [s for s in range(total) if s % world == rank]
    R, Wm, K = 6, 6, max(4, args.stream_frames)
    seqs, streams = [], []
    for s in mine:
        ring = FrameRing(W, H, R, args.storage, seed=s, device=dev)
        seqs.append(DeviceSequence(W, H, dev, args.storage, ring, args.levels))
        streams.append(torch.cuda.Stream(dev))
    for q, cs in zip(seqs, streams):
        with torch.cuda.stream(cs):
            q.f.Reset()
    torch.cuda.synchronize()
    sp = [C.c_void_p(cs.cuda_stream) for cs in streams]

    def run(t0, t1):
        for t in range(t0, t1):
            for q, cs, p in zip(seqs, streams, sp):
                q.step(t, p)

    run(0, Wm)
    torch.cuda.synchronize()
    for q in seqs:      # frames Wm.. reuse ring slots 0..: regenerate the colour planes the warm-up consumed
        q.ring.restore_colour(Wm, Wm + K)
    barrier(world)
    main = torch.cuda.current_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for cs in streams:
        cs.wait_event(e0)
    launches0 = sum(q.f.launches for q in seqs)
    # the ring holds R frames: time the steps in laps of R - 1 frames with the colour planes restored in between
    ms = 0.0
    t = Wm
    done = 0
    while done < K:
        lap = min(R - 1, K - done)
        e0.record(main)
        for cs in streams:
            cs.wait_event(e0)
        run(t, t + lap)
        for cs in streams:
            main.wait_stream(cs)
        e1.record(main)
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
        t += lap
        done += lap
        if done < K:
            for q in seqs:
                q.ring.restore_colour(t, t + min(R - 1, K - done))
            torch.cuda.synchronize()
    launches = sum(q.f.launches for q in seqs) - launches0
    barrier(world)
    ms_max = max_over_ranks(ms, world)
    launches = sum_over_ranks(launches, world)
    value = total * W * H * K / (ms_max * 1e-3) / 1e9
    for q in seqs:
        q.f.close()
    return {"workload": f"BASELINE configs[4]: 64 independent {W}x{H} streams over {world} GPU(s), {len(mine)} resident per GPU on concurrent CUDA streams",
            "value": round(value, 4), "unit": "Gpix/s", "frames_per_stream": K, "ms_per_frame_per_stream_slot": round(ms_max / K / max(1, len(mine)), 5),
            "ms_total": round(ms_max, 3), "streams_per_gpu": len(mine), "gpu_launches": int(launches), "scaling": "strong (64 streams in total)"}


def band_transport(args, world):
    """Measured on B200 (8K frame): N=2 1.739 / 1.752 ms (peer memory / NCCL), N=4 0.961 / 0.973, N=8 0.675 / 0.712."""
    if args.band_transport != "auto":
        return args.band_transport
    return "ipc" if world >= 8 else "nccl"


def run_band_frames(args, rank, world, local, W, H, K, Wm, check_frames, transport=None, fixed_bounds=None):
    """K timed 8K frames in `world` bands through the native band driver + a bit-identity check on rank 0.
    transport: "ipc" (peer-memory pulls over NVLink) or "nccl" (send/recv); fixed_bounds: skip the balancing passes."""
    transport = transport or band_transport(args, world)
    import numpy as np
    import torch
    import torch.distributed as dist
    from svgf_b200 import SvgfFilter, synth
    from svgf_b200.band_driver import BandDriver
    from svgf_b200.bands import balanced_bounds
    from svgf_b200.filter import GBuffer
    dev = torch.device("cuda", local)
    cdt = torch.float16 if args.storage == "f16" else torch.float32
    full_g, full_c = GBuffer(W, H, dev), torch.empty(H, W, 4, dtype=cdt, device=dev)
    bounds = fixed_bounds
    if args.band_balance and world > 1 and fixed_bounds is None:
        # equal estimated work per band: background pixels (linear depth 0) are passed through by the a-trous levels
        synth.frame_device(full_g, full_c, 0, seed=0)
        live = (full_g.motion[..., 2] != 0).float().mean(dim=1).cpu().numpy()
        bounds = balanced_bounds(live + args.band_bg_cost * (1.0 - live), world, min_rows=32)
    def make_driver(bnds):
        with stdout_to_stderr():
            return BandDriver(W, H, rank, world, dev, storage=args.storage, levels=args.levels, bounds=bnds, transport=transport)

    calibration = []
    if args.band_balance >= 2 and world > 1 and fixed_bounds is None:
        # measured balance: a few frames WITHOUT exchanges (SVGF_FLAG_BAND_NO_EXCHANGE: every rank runs at its own speed)
        # give each band's cost per row; the boundaries are moved to equalise the predicted times.  Twice.
        for it in range(2):
            cal = make_driver(bounds)
            cal.Reset()
            cal.params.flags = 512
            sl_c = cal.local_rows()
            g_c = GBuffer(W, cal.Height, dev)
            c_c = torch.empty(cal.Height, W, 4, dtype=cdt, device=dev)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            cur = torch.cuda.current_stream(dev)
            t_cal = 0.0
            for t in range(7):
                synth.frame_device(full_g, full_c, t, seed=0)
                g_c.normal.copy_(full_g.normal[sl_c]); g_c.uv.copy_(full_g.uv[sl_c]); g_c.motion.copy_(full_g.motion[sl_c])
                c_c.copy_(full_c[sl_c])
                P = cal.PingPongInx
                cal.Framebuffer[P].normal.copy_(g_c.normal); cal.Framebuffer[P].uv.copy_(g_c.uv); cal.Framebuffer[P].motion.copy_(g_c.motion)
                cal.RenderBuffer[P].copy_(c_c)
                ev0.record(cur)
                cal.Filter()
                ev1.record(cur)
                cal.EndFrame()
                if t >= 4:                   # steady state (history >= 4); only the filter is timed, not the input generation
                    torch.cuda.synchronize()
                    t_cal += ev0.elapsed_time(ev1) / 3
            t_rank = all_ranks(t_cal, world)
            rows_rank = all_ranks(float(cal.y1 - cal.y0), world)
            cal.close()
            del cal, g_c, c_c
            calibration.append([round(v, 4) for v in t_rank])
            speed = [r / max(t, 1e-6) for r, t in zip(rows_rank, t_rank)]             # band rows per ms
            # time to equalise: sum_i rows_i' = H with rows_i' / speed_i equal  =>  rows_i' = H * speed_i / sum(speed)
            share = [v / sum(speed) for v in speed]
            new_rows = [max(32, int(round(H * v))) for v in share]
            new_rows[-1] += H - sum(new_rows)
            bounds = [0]
            for r in new_rows:
                bounds.append(bounds[-1] + r)
            barrier(world)
    bd = make_driver(bounds)
    sl = bd.local_rows()
    R = Wm + K
    ring_g = [GBuffer(W, bd.Height, dev) for _ in range(R)]
    ring_c = [torch.empty(bd.Height, W, 4, dtype=cdt, device=dev) for _ in range(R)]
    # ---- bit identity of the stitched bands against the whole frame on one GPU (rank 0), first frames after a reset ----
    whole = None
    if rank == 0 and check_frames:
        whole = SvgfFilter(W, H, device=dev, storage=args.storage)
        whole.SpatialFilterSteps = args.levels
        whole.Reset()
    bad = 0
    own_fb = list(bd.Framebuffer)         # Reset() zeroes the driver's G-buffer members: they must never alias ring frames then
    bd.Reset()
    ranges = [None] * world
    if world > 1:
        dist.all_gather_object(ranges, (bd.y0, bd.y1))
    else:
        ranges = [(bd.y0, bd.y1)]
    for t in range(R):
        synth.frame_device(full_g, full_c, t, seed=0)
        ring_g[t].normal.copy_(full_g.normal[sl]); ring_g[t].uv.copy_(full_g.uv[sl]); ring_g[t].motion.copy_(full_g.motion[sl])
        ring_c[t].copy_(full_c[sl])
        if t < check_frames:
            P = bd.PingPongInx
            bd.Framebuffer[P] = ring_g[t]
            bd.RenderBuffer[P].copy_(ring_c[t])
            bd.Filter()
            bd.sync()
            mine = bd.result_band().contiguous()
            if whole is not None:
                whole.Framebuffer[P].normal.copy_(full_g.normal); whole.Framebuffer[P].uv.copy_(full_g.uv); whole.Framebuffer[P].motion.copy_(full_g.motion)
                whole.RenderBuffer[P].copy_(full_c)
                whole.Filter()
                ref = whole.FilterBuffer[0]
                bad += int((mine.view(torch.uint8) != ref[bd.y0:bd.y1].contiguous().view(torch.uint8)).sum())
                for r in range(1, world):
                    y0, y1 = ranges[r]
                    buf = torch.empty((y1 - y0, W, 4), dtype=cdt, device=dev)
                    dist.recv(buf, src=r)
                    bad += int((buf.view(torch.uint8) != ref[y0:y1].contiguous().view(torch.uint8)).sum())
                whole.EndFrame()
            elif world > 1:
                dist.send(mine, dst=0)
            bd.EndFrame()
    if whole is not None:
        whole.close()
        del whole
    del full_g, full_c
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    # ---- timing: a fresh sequence, inputs consumed in place from the resident ring ----
    stream = torch.cuda.current_stream(dev)
    # (round 2, found at N = 4: the check loop above leaves ring frames in bd.Framebuffer, and Reset() then zeroed the G-buffers
    # of warm-up frames 1 and 2 - the history restarted at frame 3 and the first timed frames ran the 7x7 variance estimate on
    # every pixel: +0.12 ms per frame over 24 frames for whichever transport was measured first)
    bd.Framebuffer = own_fb
    bd.Reset()
    bd.params.flags = args.flags          # e.g. 512 = SVGF_FLAG_BAND_NO_EXCHANGE (diagnostics: compute only)
    keep = bd.RenderBuffer

    def step(t):
        P = bd.PingPongInx
        bd.Framebuffer[P] = ring_g[t]
        bd.RenderBuffer[P] = ring_c[t]
        bd.Filter()
        bd.EndFrame()

    for t in range(Wm):
        step(t)
    barrier(world)
    launches0 = bd.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for t in range(Wm, Wm + K):
        step(t)
    bd.sync()                       # the state exchange posted under the last frame's levels completes inside the timed region
    e1.record(stream)
    barrier(world)
    ms = e0.elapsed_time(e1)
    ms_max = max_over_ranks(ms, world)
    ms_by_rank = all_ranks(ms / K, world)
    launches = sum_over_ranks(bd.launches - launches0, world)
    rows = [int(v) for v in all_ranks(float(bd.Height), world)]
    bad_total = int(sum_over_ranks(float(bad), world))
    bd.RenderBuffer = keep
    bd.close()
    value = W * H * K / (ms_max * 1e-3) / 1e9
    bpp = BYTES_PER_PX[args.storage]
    frame_bytes = (bpp["temporal"] + bpp["variance"] + bpp["atrous_level"] * args.levels + (bpp["atrous_hist"] if args.levels else 0)) * W * H
    peak, _ = measured_peaks()
    return {"workload": f"BASELINE configs[3]: {W}x{H} frames in {world} horizontal band(s), native band driver (include/svgf_band.h): "
                        + ("peer-memory transport (CUDA IPC mappings, flag words, a one-thread wait and one pull grid per exchange over NVLink)" if transport == "ipc" else "NCCL send/recv")
                        + " of 16 + 32 halo rows before levels 3 and 4, boundary row blocks first, state exchange under levels 1-4",
            "transport": transport,
            "value": round(value, 4), "unit": "Gpix/s", "ms_per_step": round(ms_max / K, 5), "steps": K, "warmup": Wm, "scaling": "strong",
            "band_bounds": bounds, "band_balance": {0: "equal heights", 1: "background share of the first frame", 2: "measured: two calibration passes without exchanges"}[min(args.band_balance, 2)],
            "calibration_ms_by_rank": calibration or None, "local_rows_by_rank": rows, "ms_per_step_by_rank": [round(v, 4) for v in ms_by_rank],
            "gpu_launches": int(launches),
            "bit_identical_to_one_gpu": (bad_total == 0) if check_frames else None, "checked_frames": check_frames, "mismatching_bytes": bad_total,
            "frac_of_n_gpu_hbm_peak": round(frame_bytes / (ms_max / K * 1e-3) / 1e9 / (peak * world), 4)}


def run_ours(args, rank, world, local):
    import torch
    W, H = WORKLOADS[args.workload]
    dev = torch.device("cuda", local)
    K, Wm = args.steps, args.warmup
    R = min(K + Wm, args.ring)
    ring = FrameRing(W, H, R, args.storage, seed=rank, device=dev)
    seq = DeviceSequence(W, H, dev, args.storage, ring, args.levels, args.flags, args.prefilter, args.reproj)
    f = seq.f
    stream = torch.cuda.current_stream(dev)
    sptr = C.c_void_p(stream.cuda_stream)
    lap = min(K, R - Wm) if R > Wm else 0
    if lap < K:
        raise RuntimeError(f"--ring {args.ring} holds fewer than warmup + steps = {Wm + K} frames of {W}x{H}")
    f.Reset()
    for t in range(Wm):
        seq.step(t, sptr)
    barrier(world)
    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    launches0 = f.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # EXACTLY K steps per timed region; the region is repeated (same K frames, colour planes regenerated and the sequence
    # re-warmed in between, untimed) until MIN_TIMED_MS of device time has been measured
    total_ms, repeats, prof_acc, frames_acc = 0.0, 0, {"temporal_ms": 0.0, "variance_ms": 0.0, "atrous_ms": 0.0}, 0
    while True:
        f.profile_begin()
        e0.record(stream)
        for t in range(Wm, Wm + K):
            seq.step(t, sptr)
        e1.record(stream)
        barrier(world)
        total_ms += e0.elapsed_time(e1)
        prof = f.profile_end()
        for k in prof_acc:
            prof_acc[k] += prof[k]
        frames_acc += prof["frames"]
        repeats += 1
        if max_over_ranks(total_ms, world) >= MIN_TIMED_MS or repeats >= args.max_repeats:
            break
        ring.restore_colour(0, Wm + K)
        f.Reset()
        for t in range(Wm):
            seq.step(t, sptr)
        barrier(world)
    launches = f.launches - launches0
    sampler.stop_flag = True
    sampler.join()
    ms_max = max_over_ranks(total_ms, world)
    n_timed = K * repeats
    value = world * W * H * n_timed / (ms_max * 1e-3) / 1e9
    ms_per_step = ms_max / n_timed
    stage = {k: round(v / max(1, frames_acc), 5) for k, v in prof_acc.items()}

    # ---- TAA + sRGB resolve (the step after the path, reference src/App.cu:516-522) on the filtered frames, timed alone ----
    f.TAA()
    torch.cuda.synchronize()
    n_taa = 32
    e0.record(stream)
    for _ in range(n_taa):
        f.TAA()
        f.EndFrame()
    e1.record(stream)
    torch.cuda.synchronize()
    taa_ms = e0.elapsed_time(e1) / n_taa
    stage["taa_ms"] = round(taa_ms, 5)

    # ---- general case: the same K frames with the uniform-normal tile shortcut off (a curved / normal-mapped scene) ----
    general = None
    if not (args.flags & 8) and not args.skip_extras:
        ring.restore_colour(0, Wm + K)
        gseq = DeviceSequence(W, H, dev, args.storage, ring, args.levels, args.flags | 8, args.prefilter, args.reproj)
        gseq.f.Reset()
        for t in range(Wm):
            gseq.step(t, sptr)
        barrier(world)
        e0.record(stream)
        for t in range(Wm, Wm + K):
            gseq.step(t, sptr)
        e1.record(stream)
        barrier(world)
        gms = max_over_ranks(e0.elapsed_time(e1), world) / K
        general = {"ms_per_step": round(gms, 5), "value": round(world * W * H / (gms * 1e-3) / 1e9, 4), "unit": "Gpix/s",
                   "what": "SVGF_FLAG_NO_UNIFORM_TILES: every tap evaluates the normal weight (no planar-tile shortcut)"}
        gseq.f.close()

    # ---- end to end through the host-buffer entry point: pinned host inputs, H2D + frame + D2H per step ----
    n_host = 4
    cdt = torch.float16 if args.storage == "f16" else torch.float32
    host = []
    ring.restore_colour(0, n_host)
    for t in range(n_host):
        hp = {"normal": torch.empty(H, W, 4, dtype=torch.int16).pin_memory(), "uv": torch.empty(H, W, 4, dtype=torch.int16).pin_memory(),
              "motion": torch.empty(H, W, 4, dtype=torch.float32).pin_memory(), "colour": torch.empty(H, W, 4, dtype=cdt).pin_memory()}
        src = ring.gbuf[t % R]
        hp["normal"].copy_(src.normal); hp["uv"].copy_(src.uv); hp["motion"].copy_(src.motion); hp["colour"].copy_(ring.colour[t % R])
        host.append(hp)
    result = torch.empty(H, W, 4, dtype=cdt).pin_memory()
    torch.cuda.synchronize()
    Ke = max(4, min(K, args.e2e_steps)) if args.e2e_steps > 0 else 0      # --e2e-steps 0: profiling runs only

    def e2e_step(t, reset=False):
        hp = host[t % n_host]
        f.frame_host(hp["normal"], hp["uv"], hp["motion"], hp["colour"], result=result, reset=reset)

    for t in range(3 if Ke else 0):
        e2e_step(t, reset=(t == 0))
    barrier(world)
    e0.record(stream)
    for t in range(3, 3 + Ke):
        e2e_step(t)
    e1.record(stream)
    barrier(world)
    e2e_ms = max(max_over_ranks(e0.elapsed_time(e1), world), 1e-9)
    e2e_value = world * W * H * Ke / (e2e_ms * 1e-3) / 1e9
    checksum = float(result.float().sum()) if Ke else None      # the host-side read of the step's result
    h2d, d2h = IN_BYTES_PER_PX[args.storage] * W * H, OUT_BYTES_PER_PX[args.storage] * W * H
    ceiling = None
    if not args.skip_extras:
        barrier(world)          # all ranks measure their link at the same time: that is the condition the e2e number ran under
        ceiling = pcie_ceiling(dev, h2d, d2h)
        ceiling["ms_per_step_at_ceiling"] = round(max_over_ranks(ceiling["ms_per_step_at_ceiling"], world), 4)
    del host, result
    f.close()
    del ring, seq
    torch.cuda.empty_cache()

    streams_rec = bands_rec = None
    if not args.skip_extras:
        streams_rec = run_streams_1080p(args, rank, world, local)
        torch.cuda.empty_cache()
        if world > 1:
            bw, bh = WORKLOADS["8k"]
            bands_rec = run_band_frames(args, rank, world, local, bw, bh, K=args.band_frames, Wm=6, check_frames=3)
            # the same frames, same band heights, over the other transport (timing only)
            other = "nccl" if band_transport(args, world) == "ipc" else "ipc"
            bands_other = run_band_frames(args, rank, world, local, bw, bh, K=args.band_frames, Wm=6, check_frames=1, transport=other,
                                          fixed_bounds=bands_rec["band_bounds"])

    if rank != 0:
        return None
    peak, peak_src = measured_peaks()
    traffic, traffic_src = ncu_traffic_per_launch(args)
    bpp = BYTES_PER_PX[args.storage]
    n_levels = args.levels
    at_bytes = (bpp["atrous_level"] * n_levels + (bpp["atrous_hist"] if n_levels else 0)) * W * H
    at_launches = n_levels
    at_ms = stage["atrous_ms"]
    frame_bytes = (bpp["temporal"] + bpp["variance"]) * W * H + at_bytes
    achieved = at_bytes / (at_ms * 1e-3) / 1e9 if n_levels and at_ms > 0 else None
    cfg = base_config(W, H, n_levels, args.storage)
    cfg["workload"] = workload_name(W, H, n_levels, world)
    cfg.update({"frames_resident": R, "l2": "every step reads a fresh frame (%.0f MB of inputs > 126 MB L2)" % (IN_BYTES_PER_PX[args.storage] * W * H / 1e6),
                "flags": args.flags, "variance_prefilter": args.prefilter, "reproj_mode": args.reproj,
                "timed_region": f"{K} steps x {repeats} repeat(s) = {n_timed} frames, {ms_max:.1f} ms of device time"})
    line = {
        "metric": "svgf_frame_throughput", "value": round(value, 4), "unit": "Gpix/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 compute, %s storage" % ("fp16" if args.storage == "f16" else "fp32"), "data": "synthetic",
        "config": cfg,
        "e2e": {"value": round(e2e_value, 4), "unit": "Gpix/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": round(e2e_ms / max(1, Ke), 4), "steps": Ke,
                "api": "svgf_frame_host (pinned host buffers; copy-in, kernels and copy-out of consecutive frames overlap on three streams; timed region starts with the pipeline drained)",
                "result_checksum": checksum, "pcie_ceiling": ceiling,
                "frac_of_pcie_ceiling": round(ceiling["ms_per_step_at_ceiling"] / (e2e_ms / Ke), 4) if (ceiling and Ke) else None},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "a-trous levels (%d launches/frame)" % at_launches,
                     "achieved": round(achieved, 1) if achieved else None, "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                     "bytes_per_launch": int(at_bytes / max(1, at_launches)), "ms_per_launch": round(at_ms / max(1, at_launches), 5),
                     "peak_source": peak_src,
                     "note": "the level is bound by the FP32 pipe / register-file operand bandwidth, not by HBM (DESIGN.md section 6); frac is algorithmic bytes over the HBM peak as the contract asks"},
        "frame_roofline": {"algorithmic_bytes": int(frame_bytes), "achieved": round(frame_bytes / (ms_per_step * 1e-3) / 1e9, 1),
                           "frac": round(frame_bytes / (ms_per_step * 1e-3) / 1e9 / peak, 4), "unit": "GB/s"},
        "stage_ms_per_frame": stage,
        "taa": {"ms_per_frame": round(taa_ms, 5), "frac_of_hbm_peak": round(bpp["taa"] * W * H / (taa_ms * 1e-3) / 1e9 / peak, 4),
                "bytes_per_px": bpp["taa"], "what": "svgf_taa (reference src/Filter.cuh:288-357), not part of `value`"},
        "general_case": general,
        "clocks": sampler.result(),
    }
    if streams_rec:
        line["streams_1080p"] = streams_rec
    if bands_rec:
        line["bands"] = bands_rec
        line["bands_" + bands_other["transport"]] = {k: bands_other[k] for k in ("transport", "value", "unit", "ms_per_step", "steps", "ms_per_step_by_rank",
                                                                                   "bit_identical_to_one_gpu", "checked_frames", "gpu_launches")}
    return line


def run_bands(args, rank, world, local):
    """--mode bands (BASELINE configs[3]): ONE frame stream, every frame split into `world` horizontal bands (native band
    driver, include/svgf_band.h); strong scaling."""
    W, H = WORKLOADS[args.workload]
    rec = run_band_frames(args, rank, world, local, W, H, K=args.steps, Wm=max(4, args.warmup), check_frames=args.band_check_frames)
    if rank != 0:
        return None
    cfg = base_config(W, H, args.levels, args.storage)
    cfg["workload"] = rec.pop("workload")
    for k in ("transport", "band_bounds", "band_balance", "calibration_ms_by_rank", "local_rows_by_rank", "ms_per_step_by_rank", "bit_identical_to_one_gpu", "checked_frames", "mismatching_bytes"):
        cfg[k] = rec.pop(k)
    return {"metric": "svgf_frame_throughput", "value": rec["value"], "unit": "Gpix/s", "n_gpus": world, "steps": rec["steps"], "warmup": rec["warmup"],
            "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 compute, %s storage" % ("fp16" if args.storage == "f16" else "fp32"), "data": "synthetic", "config": cfg,
            "gpu_launches": rec["gpu_launches"], "frame_roofline": {"frac_of_n_gpu_peak": rec["frac_of_n_gpu_hbm_peak"]}}


def oracle_band_sample(args, with_parity):
    """The scalar oracle on the host cores, on a bounded sample of the workload (rank 0, N = 1): the middle band of the
    frames, full width.  with_parity: the CUDA path filters the same band from the same state (teacher-forced: every
    frame starts from the oracle's buffers) and the worst differences over the timed frames are reported."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    from oracle_lib import OracleFilter, oracle, reference_defaults
    from svgf_b200 import synth
    W, H = WORKLOADS[args.workload]
    Hs = max(64, min(H, int(args.cpu_budget_px / W)))
    y0 = (H - Hs) // 2
    o = OracleFilter(W, Hs, storage=args.storage, params=reference_defaults(args.levels))
    o.Reset()
    f = None
    if with_parity:
        import torch
        from common import half_ulp_diff
        from gpu_util import load_state_from_oracle, npy
        from svgf_b200 import SvgfFilter
        f = SvgfFilter(W, Hs, storage=args.storage)
        f.SpatialFilterSteps = args.levels
        f.Reset()
    n_warm, n_timed = 4, 2
    t_acc = 0.0
    par = {"max_rel_err_rgb": 0.0, "max_rel_err_variance": 0.0, "max_abs_err": 0.0, "history_mismatches": 0, "moments_mismatches": 0,
           "max_fp16_ulps": 0, "fraction_differing": 0.0, "frames_compared": 0}
    for t in range(n_warm + n_timed):
        o.set_inputs(synth.frame_host(W, H, t, storage=args.storage, rows=(y0, y0 + Hs)))
        if f is not None:
            load_state_from_oracle(f, o)
        t0 = time.perf_counter()
        o.Filter()
        if t >= n_warm:
            t_acc += time.perf_counter() - t0
        if f is not None:
            f.Filter()
            P = o.PingPongInx
            got, want = npy(f.FilterBuffer[0]), o.FilterBuffer[0]
            d = np.abs(got.astype(np.float64) - want.astype(np.float64))
            w64 = np.abs(want.astype(np.float64))
            par["max_rel_err_rgb"] = max(par["max_rel_err_rgb"], float((d[..., :3] / np.maximum(w64[..., :3], 1e-2)).max()))
            par["max_rel_err_variance"] = max(par["max_rel_err_variance"], float((d[..., 3] / np.maximum(w64[..., 3], 2.5e-3)).max()))
            par["max_abs_err"] = max(par["max_abs_err"], float(d.max()))
            par["history_mismatches"] += int((npy(f.HistoryLengthBuffer) != o.HistoryLengthBuffer).sum())
            par["moments_mismatches"] += int((npy(f.MomentsBuffer[P]).view(np.uint8) != o.MomentsBuffer[P].view(np.uint8)).sum())
            if args.storage == "f16":
                u = half_ulp_diff(got, want)
                par["max_fp16_ulps"] = max(par["max_fp16_ulps"], int(u.max()))
                par["fraction_differing"] = max(par["fraction_differing"], float((u > 0).mean()))
            else:
                par["fraction_differing"] = max(par["fraction_differing"], float((d > 0).mean()))
            par["frames_compared"] += 1
        o.EndFrame()
    if f is not None:
        f.close()
    cores = oracle().svgf_oracle_get_threads()
    cb = {"value": round(W * Hs * n_timed / t_acc / 1e9, 6), "unit": "Gpix/s", "cores": cores, "kind": "port",
          "sample": f"{n_timed} steady-state frames (after {n_warm} warm-up frames) of the middle {W}x{Hs} band of the {W}x{H} sequence, "
                    f"scalar C++ oracle, OpenMP over rows", "ms_per_frame_sample": round(t_acc / n_timed * 1e3, 1)}
    if f is None:
        return cb, None
    par.update({"vs": "oracle/svgf_oracle.cpp (scalar restatement of src/Filter.cuh:359-624)",
                "sample": f"frames 0..{n_warm + n_timed - 1} of the middle {W}x{Hs} band, every frame started from the oracle's state (teacher-forced), "
                          f"result after {args.levels} a-trous levels",
                "bar": ("fp16 storage: every value within 2 fp16 ulps or 1e-4 absolute" if args.storage == "f16" else "fp32 storage: max relative error 1e-4 (floors 1e-2 radiance, 2.5e-3 variance)"),
                "full_report": "profiles/parity_r02.json (tools/parity_report.py: configs 1-3, both storages, teacher-forced and free-running)"})
    for k in ("max_rel_err_rgb", "max_rel_err_variance", "max_abs_err", "fraction_differing"):
        par[k] = float(f"{par[k]:.3e}")
    return cb, par


def run_reference(args, rank, world, local):
    """--impl reference: the reference's own Filter.cuh kernels (oracle/_ref, patched for compilation only) on
    this GPU when the .so exists, else the scalar oracle port on the host cores.  Nothing of the product is on this path:
    parameters are the reference's literals, inputs come from the generator library (libsvgf_synth.so) only."""
    if rank != 0:
        return None
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    W, H = WORKLOADS[args.workload]
    from oracle_lib import RefKernels, RefParams, ref, ref_available
    # n_gpus is the launch's N (the contract's key); the reference has no multi-GPU path, so one GPU does the work
    base = {"impl": "reference", "metric": "svgf_frame_throughput", "unit": "Gpix/s", "n_gpus": max(1, world), "gpus_used": 1, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic",
            "config": base_config(W, H, args.levels, "f16")}
    import torch as _torch
    if not ref_available() or args.storage != "f16" or not _torch.cuda.is_available():
        cb, _ = oracle_band_sample(args, with_parity=False)
        base.update({"value": cb["value"], "ms_per_step": cb["ms_per_frame_sample"], "dtype": "f32/f64 compute, fp16 storage",
                     "cpu_baseline": cb, "gpu_launches": 0,
                     "e2e": {"value": cb["value"], "unit": "Gpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return base
    import torch
    dev = torch.device("cuda", local)
    # src/App.h:109-114, SpatialFilterSteps = BASELINE's level count
    rp = RefParams(args.levels, 0.8, 0.9, 24, 10.0, 128.0, 0)
    r = RefKernels(W, H)
    K, Wm = args.steps, args.warmup
    R = min(K + Wm, args.ring)
    ring = FrameRing(W, H, R, "f16", seed=0, device=dev)
    D2D = 3
    tot_ms = 0.0
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
    ring.restore_colour(0, 4)
    for t in range(4):
        g = ring.gbuf[t % R]
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
                "api": "reference kernels behind host buffers (cudaArray uploads + stages + result download, synchronous like the reference's frame loop)"},
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
    ap.add_argument("--ring", type=int, default=200, help="max distinct frames kept resident in HBM (must hold warmup + steps)")
    ap.add_argument("--max-repeats", type=int, default=12, help="upper bound on repeats of the K-step timed region (each repeat re-warms the sequence)")
    ap.add_argument("--e2e-steps", type=int, default=48)
    ap.add_argument("--cpu-budget-px", type=float, default=1.6e6, help="pixels per frame of the CPU-baseline / parity sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-extras", action="store_true", help="only the headline measurement and e2e (no general case, PCIe ceiling, 1080p streams, bands)")
    ap.add_argument("--stream-frames", type=int, default=20, help="timed frames per stream of the 64 x 1080p record")
    ap.add_argument("--band-frames", type=int, default=24, help="timed 8K frames of the `bands` record (N > 1)")
    ap.add_argument("--band-check-frames", type=int, default=3, help="--mode bands: frames compared bit for bit with the whole frame on rank 0")
    ap.add_argument("--prefilter", type=int, default=0, help="svgf_params.variance_prefilter (1 = 3x3 Gaussian, not in the reference)")
    ap.add_argument("--reproj", type=int, default=0, help="svgf_params.reproj_mode (1 = bilinear 2x2, not in the reference)")
    ap.add_argument("--flags", type=int, default=0, help="svgf_params.flags for A/B runs (8 = no uniform-normal tile shortcut, 32 = no staged levels)")
    ap.add_argument("--band-transport", default="ipc", choices=["auto", "ipc", "nccl"],
                    help="bands: halo transport of the native driver (ipc = peer memory; auto: peer memory from 8 ranks on, NCCL below)")
    ap.add_argument("--band-balance", type=int, default=2, help="bands: 0 = equal heights, 1 = heights balanced by the background share of the first frame, 2 = 1 + two measured calibration passes")
    ap.add_argument("--band-bg-cost", type=float, default=0.45, help="bands: cost of a background pixel relative to a filtered one")
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
    W, H = WORKLOADS[args.workload]
    if args.impl == "ours" and args.mode == "streams":
        # the resident ring must hold warmup + steps distinct frames (an 8K frame is 1.3 GB): bound K by what fits ~100 GB
        fit = int(100e9 // (IN_BYTES_PER_PX[args.storage] * W * H)) - args.warmup
        if args.steps > fit:
            args.steps = max(4, fit)
        args.ring = max(args.ring, args.warmup + args.steps) if args.warmup + args.steps <= fit + args.warmup else args.ring

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
            cb, par = oracle_band_sample(args, with_parity=True)
            line["cpu_baseline"] = cb
            line["parity"] = par
    if rank == 0 and line is not None:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
