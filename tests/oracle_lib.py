"""ctypes access to the TEST-ONLY checkers: oracle/libsvgf_oracle.so (scalar C++ restatement) and, when
present, oracle/_ref/libsvgf_refkernels.so (the reference's own kernels).  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only."""
import ctypes as C
import os

import numpy as np

from svgf_b200._lib import SvgfFrameBuffers, SvgfGBuffer, SvgfParams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_PATH = os.path.join(ROOT, "oracle", "libsvgf_oracle.so")
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libsvgf_refkernels.so")

_o = None
_r = None


def oracle():
    global _o
    if _o is None:
        o = C.CDLL(ORACLE_PATH)
        o.svgf_oracle_f2h.restype = C.c_uint16
        o.svgf_oracle_f2h.argtypes = [C.c_float]
        o.svgf_oracle_h2f.restype = C.c_float
        o.svgf_oracle_h2f.argtypes = [C.c_uint16]
        o.svgf_oracle_set_threads.argtypes = [C.c_int]
        o.svgf_oracle_get_threads.restype = C.c_int
        P, G = C.POINTER(SvgfParams), C.POINTER(SvgfGBuffer)
        v = C.c_void_p
        o.svgf_oracle_temporal.argtypes = [P, C.c_int, C.c_int, C.c_int, G, G, v, v, v, v, v, v]
        o.svgf_oracle_variance.argtypes = [P, C.c_int, C.c_int, C.c_int, G, v, v, v, v]
        o.svgf_oracle_atrous_level.argtypes = [P, C.c_int, C.c_int, C.c_int, G, v, v, v, C.c_int]
        o.svgf_oracle_frame.argtypes = [P, C.c_int, C.c_int, C.c_int, C.POINTER(SvgfGBuffer * 2), C.POINTER(SvgfFrameBuffers)]
        o.svgf_oracle_taa.argtypes = [C.c_int, C.c_int, C.c_int, v, v, v]
        o.svgf_oracle_demodulate.argtypes = [C.c_int, C.c_int, C.c_int, v, v]
        o.svgf_oracle_remodulate.argtypes = [C.c_int, C.c_int, C.c_int, v, v, v]
        _o = o
    return _o


def reference_defaults(levels=5):
    """svgf_params filled LITERALLY from the reference's members (src/App.h:109-114; SpatialFilterSteps is 3 there, BASELINE.json
    measures 5) with the added knobs neutral - without loading the product library (bench.py's reference arm must not map it)."""
    p = SvgfParams()
    p.history_cap, p.depth_threshold, p.normal_threshold, p.phi_colour, p.phi_normal = 24, 0.8, 0.9, 10.0, 128.0
    p.atrous_iterations = levels
    p.phi_depth, p.alpha_min, p.moments_alpha_min = 1.0, 0.0, 0.0
    p.mesh_id_mode = p.reproj_mode = p.variance_prefilter = p.depth_test_mode = 0
    p.flags = 0
    return p


def np_gbuf(normal, uv, motion):
    g = SvgfGBuffer()
    g.position_id = None
    g.normal_mat = normal.ctypes.data
    g.uv_inst = uv.ctypes.data
    g.motion_depth = motion.ctypes.data
    g.position_pitch = g.normal_pitch = g.uv_pitch = g.motion_pitch = 0
    return g


def _chk(rc, where):
    if rc != 0:
        raise RuntimeError(f"{where} -> status {rc}")


class OracleFilter:
    """numpy twin of svgf_b200.filter.SvgfFilter running the scalar oracle: same member names, same call order."""

    def __init__(self, width, height, storage="f16", params=None):
        self.Width, self.Height = width, height
        self.storage = 0 if storage == "f16" else 1
        cdt = np.float16 if storage == "f16" else np.float32
        H, W = height, width
        self.normal = [np.zeros((H, W, 4), np.uint16) for _ in range(2)]
        self.uv = [np.zeros((H, W, 4), np.uint16) for _ in range(2)]
        self.motion = [np.zeros((H, W, 4), np.float32) for _ in range(2)]
        self.RenderBuffer = [np.zeros((H, W, 4), cdt) for _ in range(2)]
        self.MomentsBuffer = [np.zeros((H, W, 2), cdt) for _ in range(2)]
        self.FilterBuffer = [np.zeros((H, W, 4), cdt) for _ in range(2)]
        self.HistoryLengthBuffer = np.zeros((H, W), np.uint8)
        self.PingPongInx = 0
        self.params = params if params is not None else reference_defaults()

    def gbuf(self, k):
        return np_gbuf(self.normal[k], self.uv[k], self.motion[k])

    def set_inputs(self, planes):
        P = self.PingPongInx
        self.normal[P][...] = planes["normal"]
        self.uv[P][...] = planes["uv"]
        self.motion[P][...] = planes["motion"]
        self.RenderBuffer[P][...] = planes["colour"]

    def Reset(self):
        for lst in (self.normal, self.uv, self.motion, self.RenderBuffer, self.MomentsBuffer, self.FilterBuffer):
            for a in lst:
                a[...] = 0
        self.HistoryLengthBuffer[...] = 0
        self.PingPongInx = 0

    def TemporalFilter(self):
        P, Q = self.PingPongInx, 1 - self.PingPongInx
        hprev = self.HistoryLengthBuffer.copy()
        gc, gp = self.gbuf(P), self.gbuf(Q)
        _chk(oracle().svgf_oracle_temporal(C.byref(self.params), self.Width, self.Height, self.storage, C.byref(gc), C.byref(gp),
                                           self.RenderBuffer[Q].ctypes.data, self.RenderBuffer[P].ctypes.data,
                                           hprev.ctypes.data, self.HistoryLengthBuffer.ctypes.data,
                                           self.MomentsBuffer[P].ctypes.data, self.MomentsBuffer[Q].ctypes.data), "oracle temporal")

    def FilterMoments(self, moments_index=None):
        P = self.PingPongInx
        m = self.MomentsBuffer[P if moments_index is None else moments_index]
        gc = self.gbuf(P)
        _chk(oracle().svgf_oracle_variance(C.byref(self.params), self.Width, self.Height, self.storage, C.byref(gc),
                                           self.RenderBuffer[P].ctypes.data, m.ctypes.data,
                                           self.HistoryLengthBuffer.ctypes.data, self.FilterBuffer[0].ctypes.data), "oracle variance")

    def WaveletFilter(self):
        P = self.PingPongInx
        gc = self.gbuf(P)
        pp = 0
        for i in range(self.params.atrous_iterations):
            _chk(oracle().svgf_oracle_atrous_level(C.byref(self.params), self.Width, self.Height, self.storage, C.byref(gc),
                                                   self.FilterBuffer[pp].ctypes.data, self.FilterBuffer[1 - pp].ctypes.data,
                                                   self.RenderBuffer[P].ctypes.data, i), "oracle atrous")
            pp = 1 - pp
        if self.params.atrous_iterations % 2:
            self.FilterBuffer[0][...] = self.FilterBuffer[1]

    def Filter(self):
        b = SvgfFrameBuffers()
        for k in range(2):
            b.render[k] = self.RenderBuffer[k].ctypes.data
            b.moments[k] = self.MomentsBuffer[k].ctypes.data
            b.filter[k] = self.FilterBuffer[k].ctypes.data
        b.history = self.HistoryLengthBuffer.ctypes.data
        b.ping_pong = self.PingPongInx
        g = (SvgfGBuffer * 2)(self.gbuf(0), self.gbuf(1))
        _chk(oracle().svgf_oracle_frame(C.byref(self.params), self.Width, self.Height, self.storage, C.byref(g), C.byref(b)),
             "oracle frame")

    def TAA(self):
        """svgf_oracle_taa: FilterBuffer[0] -> TAABuffer[PingPongInx], history = TAABuffer[1 - PingPongInx] (snapshot, D13)."""
        if getattr(self, "TAABuffer", None) is None:
            self.TAABuffer = [np.zeros_like(self.FilterBuffer[0]) for _ in range(2)]
        P = self.PingPongInx
        _chk(oracle().svgf_oracle_taa(self.Width, self.Height, self.storage, self.FilterBuffer[0].ctypes.data,
                                      self.TAABuffer[1 - P].ctypes.data, self.TAABuffer[P].ctypes.data), "oracle taa")
        return self.TAABuffer[P]

    def EndFrame(self):
        self.PingPongInx = 1 - self.PingPongInx


def oracle_taa(filtered, history, storage):
    """One TAA + sRGB resolve by the oracle on numpy planes; returns the output plane."""
    H, W = filtered.shape[:2]
    out = np.zeros_like(filtered)
    _chk(oracle().svgf_oracle_taa(W, H, 0 if storage == "f16" else 1, np.ascontiguousarray(filtered).ctypes.data,
                                  np.ascontiguousarray(history).ctypes.data, out.ctypes.data), "oracle taa")
    return out


# ---- the reference's own kernels (GPU only) ---------------------------------------------------------------
class RefParams(C.Structure):
    _fields_ = [("SpatialFilterSteps", C.c_int), ("DepthThreshold", C.c_float), ("NormalThreshold", C.c_float),
                ("HistoryLength", C.c_int), ("PhiColour", C.c_float), ("PhiNormal", C.c_float), ("moments_quirk", C.c_int)]

    @classmethod
    def from_svgf(cls, p, moments_quirk=0):
        return cls(p.atrous_iterations, p.depth_threshold, p.normal_threshold, p.history_cap, p.phi_colour, p.phi_normal,
                   moments_quirk)


def ref_available():
    return os.path.exists(REF_PATH)


def ref():
    global _r
    if _r is None:
        r = C.CDLL(REF_PATH)
        v, RP = C.c_void_p, C.POINTER(RefParams)
        r.svgf_ref_create.argtypes = [C.POINTER(v), C.c_int, C.c_int]
        r.svgf_ref_destroy.argtypes = [v]
        r.svgf_ref_reset.argtypes = [v]
        r.svgf_ref_set_gbuffer.argtypes = [v, C.c_int, v, v, v, C.c_int]
        r.svgf_ref_set_plane.argtypes = [v, C.c_int, C.c_int, v, C.c_int]
        r.svgf_ref_get_plane.argtypes = [v, C.c_int, C.c_int, v, C.c_int]
        r.svgf_ref_set_ping_pong.argtypes = [v, C.c_int]
        r.svgf_ref_get_ping_pong.argtypes = [v]
        for n in ("svgf_ref_temporal", "svgf_ref_variance", "svgf_ref_wavelet"):
            getattr(r, n).argtypes = [v, RP]
        r.svgf_ref_atrous_level.argtypes = [v, RP, C.c_int]
        if hasattr(r, "svgf_ref_taa"):
            r.svgf_ref_taa.argtypes = [v]
        r.svgf_ref_frame.argtypes = [v, RP, C.c_int]
        r.svgf_ref_frame_host.argtypes = [v, RP, v, v, v, v, v, v]
        r.svgf_ref_time_frames.argtypes = [v, RP, C.c_int, C.POINTER(C.c_float)]
        _r = r
    return _r


H2D, D2H, D2D = 1, 2, 3
PLANE_RENDER, PLANE_MOMENTS, PLANE_FILTER, PLANE_HISTORY = 0, 1, 2, 3


class RefKernels:
    """The reference's kernels behind numpy in/out (fp16 storage only — that is all the reference has)."""

    def __init__(self, width, height):
        self.W, self.H = width, height
        self.ctx = C.c_void_p()
        _chk(ref().svgf_ref_create(C.byref(self.ctx), width, height), "svgf_ref_create")
        _chk(ref().svgf_ref_reset(self.ctx), "svgf_ref_reset")

    def close(self):
        if self.ctx:
            ref().svgf_ref_destroy(self.ctx)
            self.ctx = None

    def set_gbuffer(self, slot, normal, uv, motion):
        _chk(ref().svgf_ref_set_gbuffer(self.ctx, slot, normal.ctypes.data, uv.ctypes.data, motion.ctypes.data, H2D), "set_gbuffer")

    def set_plane(self, which, slot, arr):
        arr = np.ascontiguousarray(arr)
        _chk(ref().svgf_ref_set_plane(self.ctx, which, slot, arr.ctypes.data, H2D), "set_plane")

    def get_plane(self, which, slot):
        shape, dt = {0: ((self.H, self.W, 4), np.float16), 1: ((self.H, self.W, 2), np.float16),
                     2: ((self.H, self.W, 4), np.float16), 3: ((self.H, self.W), np.uint8)}[which]
        out = np.empty(shape, dt)
        _chk(ref().svgf_ref_get_plane(self.ctx, which, slot, out.ctypes.data, D2H), "get_plane")
        return out

    def load_state(self, of):
        """Copy an OracleFilter's complete state (both G-buffers, all planes, ping-pong) into the device buffers."""
        for k in range(2):
            self.set_gbuffer(k, of.normal[k], of.uv[k], of.motion[k])
            self.set_plane(PLANE_RENDER, k, of.RenderBuffer[k])
            self.set_plane(PLANE_MOMENTS, k, of.MomentsBuffer[k])
            self.set_plane(PLANE_FILTER, k, of.FilterBuffer[k])
        self.set_plane(PLANE_HISTORY, 0, of.HistoryLengthBuffer)
        ref().svgf_ref_set_ping_pong(self.ctx, of.PingPongInx)
