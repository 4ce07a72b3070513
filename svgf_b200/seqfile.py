"""``.svgfseq`` - a raw dump of filter inputs and outputs, frame by frame.

The reference has no file format for its filter stage: the G-buffer lives in GL textures and the radiance in a CUDA
buffer (reference src/App.cu:742-778), so nothing it produces can be carried to another machine.  This container lets
the procedural generator, the scalar checker and the CUDA path exchange EXACT inputs and outputs: every plane is stored
in the texel layout the C ABI takes (include/svgf.h svgf_gbuffer / svgf_frame_buffers), little-endian, dense
``y * W + x``.

    header   64 bytes: magic "SVGFSEQ1", u32 width, u32 height, u32 storage (0 = fp16, 1 = fp32), u32 plane mask,
                       u32 frame count (patched on close), 36 reserved zero bytes
    frame    the planes selected by the mask, in PLANES order, each W*H texels

Planes: the four G-buffer attachments and the noisy radiance are INPUTS; result / history / moments / colour_history are
what a filter wrote for that frame (FilterBuffer[0], HistoryLengthBuffer, MomentsBuffer[P], RenderBuffer[P])."""
import struct

import numpy as np

MAGIC = b"SVGFSEQ1"
HEADER = struct.Struct("<8sIIIII36x")
# name -> (channels, dtype for fp16 storage, dtype for fp32 storage)
PLANES = [
    ("position", 4, np.float32, np.float32),
    ("normal", 4, np.uint16, np.uint16),
    ("uv", 4, np.uint16, np.uint16),
    ("motion", 4, np.float32, np.float32),
    ("colour", 4, np.float16, np.float32),
    ("result", 4, np.float16, np.float32),
    ("history", 1, np.uint8, np.uint8),
    ("moments", 2, np.float16, np.float32),
    ("colour_history", 4, np.float16, np.float32),
]
INPUT_PLANES = ("normal", "uv", "motion", "colour")
OUTPUT_PLANES = ("result", "history", "moments", "colour_history")


def _mask(names):
    known = [p[0] for p in PLANES]
    for n in names:
        if n not in known:
            raise ValueError(f"unknown plane {n!r}")
    return sum(1 << i for i, p in enumerate(PLANES) if p[0] in names)


def _layout(mask, storage):
    return [(name, ch, np.dtype(d32 if storage else d16)) for i, (name, ch, d16, d32) in enumerate(PLANES) if mask >> i & 1]


class SeqWriter:
    def __init__(self, path, width, height, storage="f16", planes=INPUT_PLANES):
        self.W, self.H, self.storage = int(width), int(height), {"f16": 0, "f32": 1}[storage]
        self.mask = _mask(planes)
        self.layout = _layout(self.mask, self.storage)
        self.frames = 0
        self.f = open(path, "wb")
        self.f.write(HEADER.pack(MAGIC, self.W, self.H, self.storage, self.mask, 0))

    def write(self, planes):
        """planes: dict name -> array of shape [H, W, C] ([H, W] for history) with the plane's dtype (same-size integer views
        are accepted for the fp16 planes)."""
        for name, ch, dt in self.layout:
            a = np.ascontiguousarray(planes[name])
            if a.dtype != dt:
                if a.dtype.itemsize != dt.itemsize:
                    raise ValueError(f"plane {name}: dtype {a.dtype} is not {dt}")
                a = a.view(dt)
            if a.size != self.W * self.H * ch:
                raise ValueError(f"plane {name}: {a.shape} is not {self.H}x{self.W}x{ch}")
            self.f.write(a.tobytes())
        self.frames += 1

    def close(self):
        if self.f:
            self.f.seek(0)
            self.f.write(HEADER.pack(MAGIC, self.W, self.H, self.storage, self.mask, self.frames))
            self.f.close()
            self.f = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class SeqReader:
    def __init__(self, path):
        self.f = open(path, "rb")
        head = self.f.read(HEADER.size)
        if len(head) != HEADER.size:
            raise ValueError("not a .svgfseq file (short header)")
        magic, self.W, self.H, self.storage_id, self.mask, self.frames = HEADER.unpack(head)
        if magic != MAGIC or self.storage_id not in (0, 1) or self.W == 0 or self.H == 0:
            raise ValueError("not a .svgfseq file")
        self.storage = "f32" if self.storage_id else "f16"
        self.layout = _layout(self.mask, self.storage_id)
        self.frame_bytes = sum(self.W * self.H * ch * dt.itemsize for _, ch, dt in self.layout)
        self.f.seek(0, 2)
        if self.f.tell() != HEADER.size + self.frames * self.frame_bytes:
            raise ValueError("truncated .svgfseq file")

    @property
    def planes(self):
        return [name for name, _, _ in self.layout]

    def read(self, index):
        if not 0 <= index < self.frames:
            raise IndexError(index)
        self.f.seek(HEADER.size + index * self.frame_bytes)
        out = {}
        for name, ch, dt in self.layout:
            n = self.W * self.H * ch
            a = np.frombuffer(self.f.read(n * dt.itemsize), dtype=dt)
            out[name] = a.reshape(self.H, self.W) if name == "history" else a.reshape(self.H, self.W, ch)
        return out

    def __iter__(self):
        return (self.read(i) for i in range(self.frames))

    def close(self):
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
