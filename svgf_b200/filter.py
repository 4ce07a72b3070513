"""Host-side mirror of the filter part of the reference's ``application`` class.

The reference has no plugin API for this path: the boundary is three private methods of ``application``
(``TemporalFilter``, ``FilterMoments``, ``WaveletFilter``; reference src/App.cu:469-514) called in order from
``Render()`` (src/App.cu:552-556) over buffers allocated in ``ResizeRenderTextures()`` (src/App.cu:742-778),
with the ping-pong index flipped in ``EndFrame()`` (src/App.cu:366-375).  ``SvgfFilter`` keeps those names,
buffer roles, tunables (src/App.h:109-114) and call order, and forwards every stage to the C ABI of
include/svgf.h.  PyTorch is used only to own device memory and streams.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import SvgfError, SvgfFrameBuffers, SvgfGBuffer


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class GBuffer:
    """One ``framebuffer`` of the reference (4 attachments, src/App.cu:746-752) as dense device planes.

    position : float32 [H, W, 4] or None (never read by the filter)
    normal   : int16   [H, W, 4]  fp16 bit patterns (normal xyz, material)
    uv       : int16   [H, W, 4]  fp16 bit patterns (barycentrics, instance index)
    motion   : float32 [H, W, 4]  (motion px xy, linear depth, depth derivative)
    """

    def __init__(self, width, height, device, with_position=False):
        z = dict(device=device)
        self.position = torch.zeros(height, width, 4, dtype=torch.float32, **z) if with_position else None
        self.normal = torch.zeros(height, width, 4, dtype=torch.int16, **z)
        self.uv = torch.zeros(height, width, 4, dtype=torch.int16, **z)
        self.motion = torch.zeros(height, width, 4, dtype=torch.float32, **z)

    def as_struct(self):
        g = SvgfGBuffer()
        g.position_id = self.position.data_ptr() if self.position is not None else None
        g.normal_mat = self.normal.data_ptr()
        g.uv_inst = self.uv.data_ptr()
        g.motion_depth = self.motion.data_ptr()
        g.position_pitch = g.normal_pitch = g.uv_pitch = g.motion_pitch = 0
        return g

    def zero_(self):
        for t in (self.position, self.normal, self.uv, self.motion):
            if t is not None:
                t.zero_()


class TextureGBuffer:
    """A ``framebuffer`` as the reference's filter receives it (src/App.cu:473-475): three cudaTextureObject_t over
    cudaArrays (src/CudaUtil.h:68-99).  Filled from a linear ``GBuffer`` (the stand-in for the rasteriser's GL
    attachments); ``as_struct`` hands the texture objects to the C ABI with SVGF_PITCH_TEXTURE."""

    def __init__(self, width, height, device):
        self.device = device
        self._h = C.c_void_p()
        with torch.cuda.device(device):
            rc = _lib.synth_lib().svgf_synth_texgbuf_create(width, height, C.byref(self._h))
        if rc:
            raise RuntimeError(f"svgf_synth_texgbuf_create failed ({rc})")
        tex = (C.c_ulonglong * 3)()
        _lib.synth_lib().svgf_synth_texgbuf_objects(self._h, C.byref(tex))
        self.normal_tex, self.uv_tex, self.motion_tex = int(tex[0]), int(tex[1]), int(tex[2])

    def upload(self, gbuf):
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = _lib.synth_lib().svgf_synth_texgbuf_upload(self._h, _ptr(gbuf.normal), _ptr(gbuf.uv), _ptr(gbuf.motion), C.c_void_p(stream))
        if rc:
            raise RuntimeError(f"svgf_synth_texgbuf_upload failed ({rc})")

    def as_struct(self):
        g = SvgfGBuffer()
        g.position_id, g.position_pitch = None, 0
        g.normal_mat, g.normal_pitch = self.normal_tex, _lib.SVGF_PITCH_TEXTURE
        g.uv_inst, g.uv_pitch = self.uv_tex, _lib.SVGF_PITCH_TEXTURE
        g.motion_depth, g.motion_pitch = self.motion_tex, _lib.SVGF_PITCH_TEXTURE
        return g

    def zero_(self):
        pass

    def close(self):
        if self._h:
            _lib.synth_lib().svgf_synth_texgbuf_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SvgfFilter:
    """The SVGF stage of the reference's frame loop, driven through libsvgf_b200.so.

    Member names follow reference src/App.h:109-141.  Typical frame (reference src/App.cu:548-556, 366-375)::

        f.Framebuffer[f.PingPongInx]  <- this frame's G-buffer       (Rasterize)
        f.RenderBuffer[f.PingPongInx] <- this frame's noisy radiance (Trace)
        f.Filter()                    # TemporalFilter + FilterMoments + WaveletFilter, fused where profitable
        result = f.FilterBuffer[0]
        f.EndFrame()
    """

    def __init__(self, width, height, device=0, storage="f16"):
        if not torch.cuda.is_available():
            raise RuntimeError("svgf_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.lib = _lib.lib()
        self.Width, self.Height = int(width), int(height)
        self.device = torch.device("cuda", device) if not isinstance(device, torch.device) else device
        self.storage = {"f16": _lib.SVGF_STORE_F16, "f32": _lib.SVGF_STORE_F32}[storage]
        cdt = torch.float16 if storage == "f16" else torch.float32
        H, W, dev = self.Height, self.Width, self.device
        # ResizeRenderTextures(), reference src/App.cu:742-778
        self.Framebuffer = [GBuffer(W, H, dev), GBuffer(W, H, dev)]
        self.RenderBuffer = [torch.zeros(H, W, 4, dtype=cdt, device=dev) for _ in range(2)]
        self.MomentsBuffer = [torch.zeros(H, W, 2, dtype=cdt, device=dev) for _ in range(2)]
        self.FilterBuffer = [torch.zeros(H, W, 4, dtype=cdt, device=dev) for _ in range(2)]
        self.HistoryLengthBuffer = torch.zeros(H, W, dtype=torch.uint8, device=dev)
        self.TAABuffer = None     # allocated by the first TAA() call
        self.PingPongInx = 0
        # tunables, reference src/App.h:109-114 (SpatialFilterSteps is 3 there; BASELINE.json measures 5)
        self.params = _lib.default_params()
        ctx = C.c_void_p()
        st = self.lib.svgf_create(C.byref(ctx), self.device.index or 0, W, H, self.storage)
        if st != _lib.SVGF_OK:
            raise SvgfError(st, "svgf_create")
        self._ctx = ctx

    # -- tunables under the reference's member names ------------------------------------------------------
    def _tunable(field):  # noqa: N805
        def get(self):
            return getattr(self.params, field)

        def set_(self, v):
            setattr(self.params, field, v)
        return property(get, set_)

    SpatialFilterSteps = _tunable("atrous_iterations")
    DepthThreshold = _tunable("depth_threshold")
    NormalThreshold = _tunable("normal_threshold")
    HistoryLength = _tunable("history_cap")
    PhiColour = _tunable("phi_colour")
    PhiNormal = _tunable("phi_normal")
    del _tunable

    # -- plumbing ----------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, st, where):
        if st != _lib.SVGF_OK:
            raise SvgfError(st, where, self.lib.svgf_last_cuda_error(self._ctx))

    def _bufs(self):
        b = SvgfFrameBuffers()
        for k in range(2):
            b.render[k] = self.RenderBuffer[k].data_ptr()
            b.moments[k] = self.MomentsBuffer[k].data_ptr()
            b.filter[k] = self.FilterBuffer[k].data_ptr()
        b.history = self.HistoryLengthBuffer.data_ptr()
        b.ping_pong = self.PingPongInx
        return b

    def close(self):
        if getattr(self, "_ctx", None):
            self.lib.svgf_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(self.lib.svgf_launch_count(self._ctx))

    # -- the reference's stage methods ----------------------------------------------------------------------
    def Reset(self):
        """First-frame state (undefined in the reference; D12): zero history/colour/moments, all reprojection fails.
        Also zeroes whatever ``Framebuffer[0..1]`` and ``FilterBuffer[0..1]`` refer to AT THE TIME OF THE CALL: a caller that
        has pointed them at its own input planes (zero-copy stepping) must point them back first, or lose those inputs."""
        b = self._bufs()
        self._check(self.lib.svgf_reset(self._ctx, C.byref(b), self._stream()), "svgf_reset")
        for g in self.Framebuffer:
            g.zero_()
        for f in self.FilterBuffer:
            f.zero_()
        self.PingPongInx = 0

    def TemporalFilter(self):
        """application::TemporalFilter(), reference src/App.cu:469-478."""
        P, Q = self.PingPongInx, 1 - self.PingPongInx
        gc, gp = self.Framebuffer[P].as_struct(), self.Framebuffer[Q].as_struct()
        st = self.lib.svgf_temporal(self._ctx, C.byref(self.params), C.byref(gc), C.byref(gp),
                                    _ptr(self.RenderBuffer[Q]), _ptr(self.RenderBuffer[P]),
                                    _ptr(self.HistoryLengthBuffer), _ptr(self.MomentsBuffer[P]),
                                    _ptr(self.MomentsBuffer[Q]), self._stream())
        self._check(st, "svgf_temporal")

    def FilterMoments(self, moments_index=None):
        """application::FilterMoments(), reference src/App.cu:480-489.  The reference always passes
        MomentsBuffer[0] (src/App.cu:484); the default here is the current frame's plane (D4) —
        pass moments_index=0 for the reference's behaviour."""
        P = self.PingPongInx
        m = self.MomentsBuffer[P if moments_index is None else moments_index]
        gc = self.Framebuffer[P].as_struct()
        st = self.lib.svgf_variance(self._ctx, C.byref(self.params), C.byref(gc), _ptr(self.RenderBuffer[P]), _ptr(m),
                                    _ptr(self.HistoryLengthBuffer), _ptr(self.FilterBuffer[0]), self._stream())
        self._check(st, "svgf_variance")

    def WaveletFilter(self):
        """application::WaveletFilter(), reference src/App.cu:491-514: SpatialFilterSteps levels ping-ponging
        FilterBuffer[0]/[1], level 0 also feeding RenderBuffer[PingPongInx]; result in FilterBuffer[0]."""
        P = self.PingPongInx
        gc = self.Framebuffer[P].as_struct()
        res = C.c_void_p()
        st = self.lib.svgf_atrous(self._ctx, C.byref(self.params), C.byref(gc), _ptr(self.FilterBuffer[0]),
                                  _ptr(self.FilterBuffer[1]), _ptr(self.RenderBuffer[P]), 0,
                                  int(self.params.atrous_iterations), C.byref(res), self._stream())
        self._check(st, "svgf_atrous")
        if res.value == self.FilterBuffer[1].data_ptr():  # odd step count, reference src/App.cu:510-513
            self.FilterBuffer[0].copy_(self.FilterBuffer[1])

    def Filter(self):
        """The three stages of Render() (reference src/App.cu:552-556) as one svgf_frame call."""
        b = self._bufs()
        g = (SvgfGBuffer * 2)(self.Framebuffer[0].as_struct(), self.Framebuffer[1].as_struct())
        self._check(self.lib.svgf_frame(self._ctx, C.byref(self.params), C.byref(g), C.byref(b), self._stream()), "svgf_frame")

    def TAA(self):
        """application::TAA(), reference src/App.cu:516-522: FilterBuffer[0] -> TAABuffer[PingPongInx], history =
        TAABuffer[1 - PingPongInx] (the previous frame's resolve).  The reference resolves into FilterBuffer[1] and reads its
        history from that same plane while writing it (D13); two planes swapped per frame make that read well-defined."""
        if self.TAABuffer is None:
            self.TAABuffer = [torch.zeros_like(self.FilterBuffer[0]) for _ in range(2)]
        P = self.PingPongInx
        st = self.lib.svgf_taa(self._ctx, _ptr(self.FilterBuffer[0]), _ptr(self.TAABuffer[1 - P]), _ptr(self.TAABuffer[P]), self._stream())
        self._check(st, "svgf_taa")
        return self.TAABuffer[P]

    def EndFrame(self):
        """application::EndFrame()'s filter-relevant line, reference src/App.cu:374."""
        self.PingPongInx = 1 - self.PingPongInx

    # -- benchmarking helpers ----------------------------------------------------------------------------------
    def profile_begin(self):
        self._check(self.lib.svgf_profile_begin(self._ctx), "svgf_profile_begin")

    def profile_end(self):
        ms = (C.c_double * 3)()
        n = C.c_int()
        self._check(self.lib.svgf_profile_end(self._ctx, C.byref(ms), C.byref(n)), "svgf_profile_end")
        return {"temporal_ms": ms[0], "variance_ms": ms[1], "atrous_ms": ms[2], "frames": n.value}

    def last_dispatch(self):
        """Kernel family (svgf_dispatch_family) that ran each a-trous level of the most recent Filter / WaveletFilter call."""
        fam = (C.c_int32 * 16)()
        n = self.lib.svgf_last_dispatch(self._ctx, fam, 16)
        return [int(fam[i]) for i in range(min(n, 16))]

    def invalidate_guide(self):
        self.lib.svgf_invalidate_guide(self._ctx)

    def frame_host(self, normal, uv, motion, colour, result=None, history_out=None, reset=False):
        """svgf_frame_host: one frame from HOST (ideally pinned) tensors/arrays to a HOST result."""
        st = self.lib.svgf_frame_host(self._ctx, C.byref(self.params), _ptr(normal), _ptr(uv), _ptr(motion), _ptr(colour),
                                      _ptr(result), _ptr(history_out), 1 if reset else 0, self._stream())
        self._check(st, "svgf_frame_host")
