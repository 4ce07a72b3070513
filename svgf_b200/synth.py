"""Procedural G-buffer + 1-spp radiance sequences (libsvgf_synth.so): stands in for the reference's
rasteriser and path tracer, which need OpenGL / OptiX.  The same per-pixel function runs as a
CUDA kernel (``frame_device``) or an OpenMP loop (``frame_host``) and produces identical bits."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import SynthCfg

PAN_PX, VERT_PX, HALF_PERIOD = 3.25, 0.5, 32   # BASELINE config 2: 64-frame pan with a reversal at frame 32


def make_cfg(width, height, frame, seed=0, storage="f16", pan_px=PAN_PX, vert_px=VERT_PX, half_period=HALF_PERIOD):
    return SynthCfg(width, height, seed, frame, pan_px, vert_px, half_period, 0 if storage == "f16" else 1)


def frame_host(width, height, frame, seed=0, storage="f16", with_position=False, threads=0, rows=None, **kw):
    """Returns dict of numpy planes: normal/uv (uint16 [H,W,4]), motion (float32 [H,W,4]),
    colour (float16|float32 [H,W,4]) and optionally position.  rows=(y0, y1) generates only that band of the
    frame (planes are then [y1-y0, W, 4])."""
    cfg = make_cfg(width, height, frame, seed, storage, **kw)
    y0, y1 = rows if rows is not None else (0, height)
    h = y1 - y0
    out = {
        "normal": np.empty((h, width, 4), np.uint16),
        "uv": np.empty((h, width, 4), np.uint16),
        "motion": np.empty((h, width, 4), np.float32),
        "colour": np.empty((h, width, 4), np.float16 if storage == "f16" else np.float32),
    }
    pos = np.empty((h, width, 4), np.float32) if with_position else None
    if with_position:
        out["position"] = pos
    rc = _lib.synth_lib().svgf_synth_rows_host(
        C.byref(cfg), y0, y1, pos.ctypes.data if with_position else None, out["normal"].ctypes.data, out["uv"].ctypes.data,
        out["motion"].ctypes.data, out["colour"].ctypes.data, threads)
    if rc:
        raise RuntimeError(f"svgf_synth_rows_host failed ({rc})")
    return out


def frame_device(gbuf, colour, frame, seed=0, **kw):
    """Fill a ``svgf_b200.filter.GBuffer`` and a colour plane (torch CUDA tensors) for frame `frame`, on the
    current stream."""
    import torch
    h, w = colour.shape[0], colour.shape[1]
    storage = "f16" if colour.dtype == torch.float16 else "f32"
    cfg = make_cfg(w, h, frame, seed, storage, **kw)
    stream = torch.cuda.current_stream(colour.device).cuda_stream
    with torch.cuda.device(colour.device):
        rc = _lib.synth_lib().svgf_synth_frame_device(
            C.byref(cfg), gbuf.position.data_ptr() if gbuf.position is not None else None, gbuf.normal.data_ptr(),
            gbuf.uv.data_ptr(), gbuf.motion.data_ptr(), colour.data_ptr(), C.c_void_p(stream))
    if rc:
        raise RuntimeError(f"svgf_synth_frame_device failed (cudaError {rc})")
