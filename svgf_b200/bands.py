"""One large frame split into horizontal bands, one band per rank (BASELINE config 4).

Rank g owns image rows [y0, y1) and holds a LOCAL image of rows [y0 - A, y1 + A) clipped to the frame (A = apron =
32 rows = the widest single-level halo, 2 * 2^4).  The path needs no all-reduce; the only exchange is between
vertical neighbours:

  * a-trous level i reads +-2*2^i rows of the previous level's output (reference src/Filter.cuh:571-576): before
    level i every rank sends its top / bottom 2*2^i BAND rows of that buffer to the neighbour above / below, into
    the neighbour's apron (NCCL send/recv over NVLink on GPUs; gloo in the CPU tests);
  * the temporal pass gathers previous-frame texels at the motion-vector target (src/Filter.cuh:225-258): at the
    start of a frame the aprons of the previous-frame state (colour history, moments, history lengths) are refreshed
    the same way (`state_apron` rows, covering vertical motion of up to that many pixels per frame);
  * the variance pass's 7x7 window (src/Filter.cuh:465) and every level's own halo are covered by running each stage
    over the whole local image: rows of the apron that lie closer than the stage's reach to the artificial border
    come out wrong, but they are never read for a band row before the next exchange overwrites them.

The class is backend-agnostic: it drives any object with the SvgfFilter interface (svgf_b200.filter.SvgfFilter on
a GPU; the CPU checker behind torch CPU views in the world_size-2 gloo test) through the three
callables below, so the partition / exchange logic is identical in both.
"""
from dataclasses import dataclass

import torch
import torch.distributed as dist

APRON = 32          # rows; = halo of the last of 5 levels (2 * 2^4)
MAX_LEVELS = 5      # levels beyond 5 would need 2 * 2^level > APRON rows


@dataclass
class Band:
    rank: int
    world: int
    H: int
    y0: int
    y1: int           # owned rows [y0, y1)
    ly0: int
    ly1: int          # local image rows [ly0, ly1)

    @property
    def local_height(self):
        return self.ly1 - self.ly0

    def loc(self, gy):
        """local row index of global row gy"""
        return gy - self.ly0


def band_of(H, world, rank, apron=APRON):
    base, rem = divmod(H, world)
    y0 = rank * base + min(rank, rem)
    y1 = y0 + base + (1 if rank < rem else 0)
    return Band(rank, world, H, y0, y1, max(0, y0 - apron), min(H, y1 + apron))


def check_partition(H, world, levels, apron=APRON):
    if levels > MAX_LEVELS or 2 * (1 << max(levels - 1, 0)) > apron:
        raise ValueError(f"{levels} a-trous levels need a {2 << (levels - 1)}-row apron (> {apron})")
    if world > 1 and H // world < apron:
        raise ValueError(f"bands of {H // world} rows are shorter than the {apron}-row halo: use fewer ranks")


def exchange_rows(band, planes, rows, group=None):
    """Refresh `rows` apron rows on each side of the band in every tensor of `planes` (local images, dim 0 = rows)
    with the neighbours' band rows.  One batched send/recv per call; a no-op on one rank."""
    if band.world == 1 or rows <= 0:
        return
    ops, keep = [], []
    up, down = band.rank - 1, band.rank + 1
    for t in planes:
        if up >= 0:
            send = t[band.loc(band.y0):band.loc(band.y0 + rows)]              # my top band rows -> upper neighbour's bottom apron
            recv = t[band.loc(band.y0 - rows):band.loc(band.y0)]              # upper neighbour's bottom band rows -> my top apron
            ops += [dist.P2POp(dist.isend, send, up, group), dist.P2POp(dist.irecv, recv, up, group)]
            keep += [send, recv]
        if down < band.world:
            send = t[band.loc(band.y1 - rows):band.loc(band.y1)]
            recv = t[band.loc(band.y1):band.loc(band.y1 + rows)]
            ops += [dist.P2POp(dist.isend, send, down, group), dist.P2POp(dist.irecv, recv, down, group)]
            keep += [send, recv]
    for w in dist.batch_isend_irecv(ops):
        w.wait()


class BandedFilter:
    """Drives one rank's band of a frame through temporal + variance + N a-trous levels with per-level halo exchange.

    backend : object with the SvgfFilter interface sized (W, band.local_height): RenderBuffer / MomentsBuffer /
              FilterBuffer / HistoryLengthBuffer / PingPongInx / params, and
    ops     : dict of callables  'temporal_variance'(backend) -> runs the temporal and variance passes over the local
              image, leaving the variance output in the returned buffer;  'atrous_level'(backend, level, src, dst);
              'as_tensor'(buffer) -> torch tensor view of a backend buffer (identity for SvgfFilter).
    """

    def __init__(self, backend, band, ops, levels=5, state_apron=APRON, group=None):
        check_partition(band.H, band.world, levels)
        self.f, self.band, self.ops, self.levels, self.group = backend, band, ops, levels, group
        self.state_apron = min(state_apron, APRON)
        self.frame = 0

    def _t(self, buf):
        return self.ops["as_tensor"](buf)

    def Filter(self):
        f, b = self.f, self.band
        P, Q = f.PingPongInx, 1 - f.PingPongInx
        if self.frame > 0:      # previous-frame state in the aprons (the reset frame has none)
            exchange_rows(b, [self._t(f.RenderBuffer[Q]), self._t(f.MomentsBuffer[Q]), self._t(f.HistoryLengthBuffer)],
                          self.state_apron, self.group)
        src = self.ops["temporal_variance"](f)                    # -> FilterBuffer[k] holding the variance pass's output
        k = 0 if src is f.FilterBuffer[0] else 1
        for level in range(self.levels):
            exchange_rows(b, [self._t(f.FilterBuffer[k])], 2 << level, self.group)
            self.ops["atrous_level"](f, level, f.FilterBuffer[k], f.FilterBuffer[1 - k])
            k = 1 - k
        self.result_index = k
        self.frame += 1
        return f.FilterBuffer[k]

    def result_band(self):
        """The owned rows of the final result (a view into the local image)."""
        t = self._t(self.f.FilterBuffer[self.result_index])
        return t[self.band.loc(self.band.y0):self.band.loc(self.band.y1)]

    def EndFrame(self):
        self.f.EndFrame()


# ---- GPU backend glue: svgf_b200.filter.SvgfFilter through the C ABI -------------------------------------------------
def _gpu_temporal_variance(f):
    """svgf_frame with zero a-trous levels = the fused temporal + variance passes (variance output in FilterBuffer[0])."""
    n = f.params.atrous_iterations
    f.params.atrous_iterations = 0
    try:
        f.Filter()
    finally:
        f.params.atrous_iterations = n
    return f.FilterBuffer[0]


def _gpu_atrous_level(f, level, src, dst):
    import ctypes as C
    from ._lib import SVGF_OK, SvgfError
    P = f.PingPongInx
    g = f.Framebuffer[P].as_struct()
    res = C.c_void_p()
    st = f.lib.svgf_atrous(f._ctx, C.byref(f.params), C.byref(g), C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()),
                           C.c_void_p(f.RenderBuffer[P].data_ptr()), level, 1, C.byref(res), f._stream())
    if st != SVGF_OK:
        raise SvgfError(st, "svgf_atrous", f.lib.svgf_last_cuda_error(f._ctx))
    assert res.value == dst.data_ptr()


GPU_OPS = {"temporal_variance": _gpu_temporal_variance, "atrous_level": _gpu_atrous_level, "as_tensor": lambda t: t}


def make_gpu_banded_filter(W, H, rank, world, device, storage="f16", levels=5, group=None):
    from .filter import SvgfFilter
    band = band_of(H, world, rank)
    f = SvgfFilter(W, band.local_height, device=device, storage=storage)
    f.SpatialFilterSteps = levels
    return BandedFilter(f, band, GPU_OPS, levels=levels, group=group)
