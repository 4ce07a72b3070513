"""One large frame in horizontal bands, one band per rank, through the native band driver (include/svgf_band.h,
csrc/svgf_band.cu): the driver owns an NCCL communicator and a side stream, runs each exchanging level's boundary row
blocks first and overlaps the halo exchange with the interior.  ``svgf_b200.bands`` is the general, level-by-level
fallback on torch.distributed (any level count, any backend; also what the CPU tests drive).

``BandDriver`` mirrors ``SvgfFilter`` for the LOCAL image of this rank (band + aprons): same member names, same call order
(fill Framebuffer[P] / RenderBuffer[P] rows [ly0, ly1) of the frame, ``Filter()``, read ``result_band()``, ``EndFrame()``).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import SvgfError, SvgfFrameBuffers, SvgfGBuffer
from .filter import GBuffer


def broadcast_unique_id(rank, src=0, group=None, device=None):
    """ncclGetUniqueId on `src`, broadcast to every rank over torch.distributed; returns the 128 bytes."""
    import torch.distributed as dist
    lib = _lib.lib()
    buf = (C.c_ubyte * 128)()
    if rank == src:
        st = lib.svgf_band_unique_id(buf)
        if st != _lib.SVGF_OK:
            raise SvgfError(st, "svgf_band_unique_id")
    t = torch.tensor(list(buf), dtype=torch.uint8)
    if dist.get_backend(group) == "nccl":
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    dist.broadcast(t, src=src, group=group)
    return bytes(t.cpu().tolist())


class BandDriver:
    def __init__(self, width, height, rank, world, device, storage="f16", levels=5, bounds=None, unique_id=None, group=None,
                 _handle=None, transport="nccl"):
        """transport: "nccl" (send/recv on a driver-owned communicator) or "ipc" (peer-memory pulls over NVLink: CUDA IPC
        handles exchanged once through torch.distributed, any backend; no NCCL on the data path)."""
        self.lib = _lib.lib()
        self.device = torch.device("cuda", device) if not isinstance(device, torch.device) else device
        self.rank, self.world, self.Width, self.FullHeight = rank, world, int(width), int(height)
        self.transport = transport
        if _handle is not None:                                # a member of a BandGroup: the group created the native band
            self._h = _handle
        elif transport == "ipc":
            import torch.distributed as dist
            rb = (C.c_int32 * (world + 1))(*bounds) if bounds is not None else None
            self._h = C.c_void_p()
            st = self.lib.svgf_band_create_ipc(C.byref(self._h), self.device.index or 0, rank, world, self.Width, self.FullHeight,
                                               {"f16": _lib.SVGF_STORE_F16, "f32": _lib.SVGF_STORE_F32}[storage], rb)
            if st != _lib.SVGF_OK:
                raise SvgfError(st, "svgf_band_create_ipc")
            if world > 1:
                blob = (C.c_ubyte * 640)()
                self._check(self.lib.svgf_band_ipc_export(self._h, blob), "svgf_band_ipc_export")
                blobs = [None] * world
                dist.all_gather_object(blobs, bytes(blob), group=group)
                up = (C.c_ubyte * 640)(*blobs[rank - 1]) if rank > 0 else None
                down = (C.c_ubyte * 640)(*blobs[rank + 1]) if rank + 1 < world else None
                self._check(self.lib.svgf_band_ipc_connect(self._h, up, down), "svgf_band_ipc_connect")
        else:
            if world > 1 and unique_id is None:
                unique_id = broadcast_unique_id(rank, group=group, device=self.device)
            uid = (C.c_ubyte * 128)(*unique_id) if unique_id is not None else None
            rb = (C.c_int32 * (world + 1))(*bounds) if bounds is not None else None
            self._h = C.c_void_p()
            st = self.lib.svgf_band_create(C.byref(self._h), self.device.index or 0, rank, world, self.Width, self.FullHeight,
                                           {"f16": _lib.SVGF_STORE_F16, "f32": _lib.SVGF_STORE_F32}[storage], uid, rb)
            if st != _lib.SVGF_OK:
                raise SvgfError(st, "svgf_band_create")
        rows = (C.c_int32 * 4)()
        self.lib.svgf_band_rows(self._h, C.byref(rows))
        self.y0, self.y1, self.ly0, self.ly1 = (int(v) for v in rows)
        self.Height = self.ly1 - self.ly0                      # rows of the local image
        H, W, dev = self.Height, self.Width, self.device
        cdt = torch.float16 if storage == "f16" else torch.float32
        self.Framebuffer = [GBuffer(W, H, dev), GBuffer(W, H, dev)]
        self.RenderBuffer = [torch.zeros(H, W, 4, dtype=cdt, device=dev) for _ in range(2)]
        self.MomentsBuffer = [torch.zeros(H, W, 2, dtype=cdt, device=dev) for _ in range(2)]
        self.FilterBuffer = [torch.zeros(H, W, 4, dtype=cdt, device=dev) for _ in range(2)]
        self.HistoryLengthBuffer = torch.zeros(H, W, dtype=torch.uint8, device=dev)
        self.PingPongInx = 0
        self.params = _lib.default_params()
        self.params.atrous_iterations = levels

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, st, where):
        if st != _lib.SVGF_OK:
            raise SvgfError(st, where, self.lib.svgf_band_last_error(self._h))

    def _bufs(self):
        b = SvgfFrameBuffers()
        for k in range(2):
            b.render[k] = self.RenderBuffer[k].data_ptr()
            b.moments[k] = self.MomentsBuffer[k].data_ptr()
            b.filter[k] = self.FilterBuffer[k].data_ptr()
        b.history = self.HistoryLengthBuffer.data_ptr()
        b.ping_pong = self.PingPongInx
        return b

    @property
    def launches(self):
        return int(self.lib.svgf_band_launch_count(self._h))

    def local_rows(self):
        """slice of FRAME rows held by the local image"""
        return slice(self.ly0, self.ly1)

    def Reset(self):
        """svgf_band_reset + zeroed G-buffer / filter members (whatever they refer to at the time of the call: see
        SvgfFilter.Reset - a harness that aliases them to its input ring must restore them before resetting)."""
        b = self._bufs()
        self._check(self.lib.svgf_band_reset(self._h, C.byref(b), self._stream()), "svgf_band_reset")
        for g in self.Framebuffer:
            g.zero_()
        for f in self.FilterBuffer:
            f.zero_()
        self.PingPongInx = 0

    def Filter(self):
        b = self._bufs()
        g = (SvgfGBuffer * 2)(self.Framebuffer[0].as_struct(), self.Framebuffer[1].as_struct())
        self._check(self.lib.svgf_band_frame(self._h, C.byref(self.params), C.byref(g), C.byref(b), self._stream()), "svgf_band_frame")

    def EndFrame(self):
        self.PingPongInx = 1 - self.PingPongInx

    def sync(self):
        self._check(self.lib.svgf_band_sync(self._h, self._stream()), "svgf_band_sync")

    def result_band(self):
        """The owned rows of the result (a view into FilterBuffer[0])."""
        return self.FilterBuffer[0][self.y0 - self.ly0:self.y1 - self.ly0]

    def close(self):
        if getattr(self, "_h", None):
            self.lib.svgf_band_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BandGroup:
    """All bands of one frame inside this process (svgf_band_create_group / svgf_band_group_frame): band g on devices[g] -
    several GPUs driven by one thread, or every band on the same GPU, which is how the band schedule (plan, row-block ranges,
    halo and state exchanges) is tested bit for bit on a one-GPU box.  Halos travel by cudaMemcpyPeerAsync, not NCCL.
    ``bands[g]`` is a BandDriver for band g's local image (fill its Framebuffer / RenderBuffer rows, read result_band());
    ``Filter()`` / ``EndFrame()`` / ``Reset()`` act on all bands."""

    def __init__(self, width, height, devices, storage="f16", levels=5, bounds=None):
        self.lib = _lib.lib()
        world = len(devices)
        devs = [torch.device("cuda", d) if not isinstance(d, torch.device) else d for d in devices]
        hs = (C.c_void_p * world)()
        dv = (C.c_int32 * world)(*[d.index or 0 for d in devs])
        rb = (C.c_int32 * (world + 1))(*bounds) if bounds is not None else None
        st = self.lib.svgf_band_create_group(hs, dv, world, int(width), int(height),
                                             {"f16": _lib.SVGF_STORE_F16, "f32": _lib.SVGF_STORE_F32}[storage], rb)
        if st != _lib.SVGF_OK:
            raise SvgfError(st, "svgf_band_create_group")
        self.world = world
        self.bands = [BandDriver(width, height, g, world, devs[g], storage=storage, levels=levels, _handle=C.c_void_p(hs[g]))
                      for g in range(world)]
        self.params = _lib.default_params()
        self.params.atrous_iterations = levels
        self.streams = None          # optional: one torch.cuda.Stream per band (default: each device's current stream)

    def Reset(self):
        for b in self.bands:
            b.Reset()

    def Filter(self):
        if self.streams is None:
            return self._filter([b._stream() for b in self.bands])
        for b, s in zip(self.bands, self.streams):
            s.wait_stream(torch.cuda.current_stream(b.device))
        self._filter([C.c_void_p(s.cuda_stream) for s in self.streams])
        for b, s in zip(self.bands, self.streams):
            torch.cuda.current_stream(b.device).wait_stream(s)

    def _filter(self, stream_ptrs):
        w = self.world
        hs = (C.c_void_p * w)(*[b._h for b in self.bands])
        g = (SvgfGBuffer * (2 * w))()
        bufs = (SvgfFrameBuffers * w)()
        streams = (C.c_void_p * w)()
        for i, b in enumerate(self.bands):
            g[2 * i], g[2 * i + 1] = b.Framebuffer[0].as_struct(), b.Framebuffer[1].as_struct()
            bufs[i] = b._bufs()
            streams[i] = stream_ptrs[i]
        st = self.lib.svgf_band_group_frame(hs, w, C.byref(self.params), g, bufs, streams)
        if st != _lib.SVGF_OK:
            raise SvgfError(st, "svgf_band_group_frame", max(self.lib.svgf_band_last_error(b._h) for b in self.bands))

    def EndFrame(self):
        for b in self.bands:
            b.EndFrame()

    def sync(self):
        for b in self.bands:
            b.sync()

    def close(self):
        for b in self.bands:
            b.close()
