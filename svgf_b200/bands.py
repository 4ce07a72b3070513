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

Two refinements cut the exposed communication (measured in DESIGN.md section 7): levels below `exchange_from_level` skip
their exchange and let a wider apron absorb their halo (redundant rows instead of a latency-bound send/recv), and the
previous-frame state exchange can be posted right after level 0 so that it runs under the remaining levels.

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


def band_of(H, world, rank, apron=APRON, bounds=None):
    """Band of `rank`: equal row counts, or the rows [bounds[rank], bounds[rank + 1]) of balanced_bounds()."""
    if bounds is not None:
        if len(bounds) != world + 1 or bounds[0] != 0 or bounds[-1] != H or any(b1 <= b0 for b0, b1 in zip(bounds, bounds[1:])):
            raise ValueError("bounds must be world + 1 increasing row indices from 0 to H")
        y0, y1 = int(bounds[rank]), int(bounds[rank + 1])
    else:
        base, rem = divmod(H, world)
        y0 = rank * base + min(rank, rem)
        y1 = y0 + base + (1 if rank < rem else 0)
    return Band(rank, world, H, y0, y1, max(0, y0 - apron), min(H, y1 + apron))


def balanced_bounds(row_cost, world, min_rows=APRON):
    """Band boundaries with (nearly) equal summed `row_cost` per band instead of equal row counts.

    The a-trous levels skip background pixels (reference src/Filter.cuh:554-558) and take a cheaper path on uniform
    tiles, so rows cost very different amounts; with equal-height bands the slowest rank sets the frame time.  row_cost
    is any per-row estimate every rank can compute identically (bench.py: fraction of non-background pixels of the
    first frame, background weighted 0.45 - measured with `bench.py --band-dry-run 1` on 8 B200s).  Every band keeps at least `min_rows` rows (it must be able to feed its
    neighbours' aprons)."""
    import numpy as np
    c = np.maximum(np.asarray(row_cost, dtype=np.float64), 0.0) + 1e-9
    H = c.shape[0]
    if world * min_rows > H:
        raise ValueError(f"{world} bands of at least {min_rows} rows do not fit {H} rows")
    cum = np.concatenate([[0.0], np.cumsum(c)])
    bounds = [0]
    for k in range(1, world):
        y = int(np.argmin(np.abs(cum - cum[-1] * k / world)))      # boundary whose cumulative cost is nearest the k-th share
        y = max(y, bounds[-1] + min_rows)              # keep the band above tall enough ...
        y = min(y, H - (world - k) * min_rows)         # ... and leave room for the ones below
        bounds.append(y)
    bounds.append(H)
    return bounds


def required_apron(levels, exchange_from_level=0, max_motion_rows=0):
    """Apron rows a band needs when a-trous levels < exchange_from_level run WITHOUT a halo exchange (their halo is
    computed redundantly in the apron) and the others exchange 2 * 2^level rows first.  A stage's output is right only
    `reach` rows further in from the artificial border than its input was: temporal `max_motion_rows` (the gather at the
    motion-vector target), variance 3 (7x7 window), level i 2 * 2^i."""
    L = min(max(exchange_from_level, 0), levels)
    redundant = (3 + max_motion_rows + sum(2 << i for i in range(L))) if L > 0 else 0
    exchanged = max([2 << i for i in range(L, levels)], default=0)
    return max(redundant, exchanged)


def check_partition(H, world, levels, apron=APRON, exchange_from_level=0, max_motion_rows=0):
    need = required_apron(levels, exchange_from_level, max_motion_rows)
    if need > apron:
        raise ValueError(f"{levels} a-trous levels (halo exchange from level {exchange_from_level}) need a {need}-row apron (> {apron})")
    if world > 1 and H // world < apron:
        raise ValueError(f"bands of {H // world} rows are shorter than the {apron}-row halo: use fewer ranks")


def check_band(band, apron):
    if band.world > 1 and band.y1 - band.y0 < apron:
        raise ValueError(f"band of {band.y1 - band.y0} rows is shorter than the {apron}-row halo")


def build_exchange_ops(band, planes, rows, group=None):
    """The batched send/recv list that refreshes `rows` apron rows on each side of the band for every tensor of `planes`
    (local images, dim 0 = rows) with the neighbours' band rows.  The list only refers to tensor views, so it can be built
    once per (planes, rows) and posted every frame."""
    ops = []
    up, down = band.rank - 1, band.rank + 1
    for t in planes:
        if up >= 0:
            send = t[band.loc(band.y0):band.loc(band.y0 + rows)]              # my top band rows -> upper neighbour's bottom apron
            recv = t[band.loc(band.y0 - rows):band.loc(band.y0)]              # upper neighbour's bottom band rows -> my top apron
            ops += [dist.P2POp(dist.isend, send, up, group), dist.P2POp(dist.irecv, recv, up, group)]
        if down < band.world:
            send = t[band.loc(band.y1 - rows):band.loc(band.y1)]
            recv = t[band.loc(band.y1):band.loc(band.y1 + rows)]
            ops += [dist.P2POp(dist.isend, send, down, group), dist.P2POp(dist.irecv, recv, down, group)]
    return ops


def post_exchange(band, planes, rows, group=None, cache=None):
    """Start the exchange and return the pending work handles ([] on one rank); finish_exchange() completes it.  One
    batched send/recv.  `cache` (a dict owned by the caller) keeps the op lists across frames: building 12 tensor views
    and P2POps per call is host time that a 0.5 ms band frame notices."""
    if band.world == 1 or rows <= 0:
        return []
    if cache is None:
        ops = build_exchange_ops(band, planes, rows, group)
    else:
        key = (tuple(t.data_ptr() for t in planes), rows)
        ops = cache.get(key)
        if ops is None:
            ops = cache[key] = build_exchange_ops(band, planes, rows, group)
    return [(w, ops) for w in dist.batch_isend_irecv(ops)]


def finish_exchange(pending):
    for w, _ in pending:
        w.wait()


def exchange_rows(band, planes, rows, group=None, cache=None):
    """Refresh `rows` apron rows on each side of the band in every tensor of `planes` (local images, dim 0 = rows)
    with the neighbours' band rows.  One batched send/recv per call; a no-op on one rank."""
    finish_exchange(post_exchange(band, planes, rows, group, cache))


class BandedFilter:
    """Drives one rank's band of a frame through temporal + variance + N a-trous levels with per-level halo exchange.

    backend : object with the SvgfFilter interface sized (W, band.local_height): RenderBuffer / MomentsBuffer /
              FilterBuffer / HistoryLengthBuffer / PingPongInx / params, and
    ops     : dict of callables  'temporal_variance'(backend) -> runs the temporal and variance passes over the local
              image, leaving the variance output in the returned buffer;  'atrous_level'(backend, level, src, dst);
              'as_tensor'(buffer) -> torch tensor view of a backend buffer (identity for SvgfFilter).
    """

    def __init__(self, backend, band, ops, levels=5, state_apron=None, group=None, exchange_from_level=0, max_motion_rows=0,
                 overlap_state=False):
        """exchange_from_level : a-trous levels below it skip the halo exchange and rely on the apron instead (0 = exchange
                              before every level; `levels` = no per-level exchange at all, one state exchange per frame);
        max_motion_rows     : vertical reach of the temporal gather the apron has to cover in that mode;
        overlap_state       : post the previous-frame state exchange of frame t+1 right after level 0 of frame t (the
                              three planes are final by then) so that it runs under levels 1..N-1."""
        self.apron = max(band.y0 - band.ly0, band.ly1 - band.y1) if band.world > 1 else APRON
        check_partition(band.H, band.world, levels, self.apron if band.world > 1 else 1 << 30, exchange_from_level, max_motion_rows)
        check_band(band, self.apron)
        self.f, self.band, self.ops, self.levels, self.group = backend, band, ops, levels, group
        self.exchange_from_level = exchange_from_level
        self.state_apron = self.apron if state_apron is None else min(state_apron, self.apron)
        self.overlap_state = overlap_state
        self._pending_state = None
        self._op_cache = {}
        self.frame = 0

    def _t(self, buf):
        return self.ops["as_tensor"](buf)

    def _state_planes(self, idx):
        f = self.f
        return [self._t(f.RenderBuffer[idx]), self._t(f.MomentsBuffer[idx]), self._t(f.HistoryLengthBuffer)]

    def Filter(self):
        f, b = self.f, self.band
        P, Q = f.PingPongInx, 1 - f.PingPongInx
        if self._pending_state is not None:     # posted under the previous frame's levels 1..N-1
            finish_exchange(self._pending_state)
            self._pending_state = None
        elif self.frame > 0:                    # previous-frame state in the aprons (the reset frame has none)
            exchange_rows(b, self._state_planes(Q), self.state_apron, self.group, self._op_cache)
        src = self.ops["temporal_variance"](f)                    # -> FilterBuffer[k] holding the variance pass's output
        k = 0 if src is f.FilterBuffer[0] else 1
        level = 0
        while level < self.levels:
            if level >= self.exchange_from_level:
                exchange_rows(b, [self._t(f.FilterBuffer[k])], 2 << level, self.group, self._op_cache)
            # levels that need no exchange in between go out as ONE backend call (fewer host round trips per frame);
            # the group ends after level 0 when the state exchange is to be posted there
            n = 1
            while (level + n < self.levels and level + n < self.exchange_from_level and not (self.overlap_state and level == 0)):
                n += 1
            if "atrous_levels" in self.ops:
                k = self.ops["atrous_levels"](f, level, n, k)
            else:
                for lv in range(level, level + n):
                    self.ops["atrous_level"](f, lv, f.FilterBuffer[k], f.FilterBuffer[1 - k])
                    k = 1 - k
            if level == 0 and self.overlap_state:
                # colour history (written by level 0), moments and history lengths of THIS frame are final: they are the
                # next frame's previous-frame state
                self._pending_state = post_exchange(b, self._state_planes(P), self.state_apron, self.group, self._op_cache)
            level += n
        if self.levels == 0 and self.overlap_state:
            self._pending_state = post_exchange(b, self._state_planes(P), self.state_apron, self.group, self._op_cache)
        self.result_index = k
        self.frame += 1
        return f.FilterBuffer[k]

    def drain(self):
        """Complete a posted state exchange (call before reading state planes from outside or tearing down)."""
        if self._pending_state is not None:
            finish_exchange(self._pending_state)
            self._pending_state = None

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


def _gpu_atrous_levels(f, first, n, k):
    """Levels first .. first+n-1 in one svgf_atrous call, ping-ponging FilterBuffer[k] -> FilterBuffer[1-k] -> ...;
    returns the index of the buffer holding the result."""
    import ctypes as C
    from ._lib import SVGF_OK, SvgfError
    P = f.PingPongInx
    g = f.Framebuffer[P].as_struct()
    res = C.c_void_p()
    src, dst = f.FilterBuffer[k], f.FilterBuffer[1 - k]
    st = f.lib.svgf_atrous(f._ctx, C.byref(f.params), C.byref(g), C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()),
                           C.c_void_p(f.RenderBuffer[P].data_ptr()), first, n, C.byref(res), f._stream())
    if st != SVGF_OK:
        raise SvgfError(st, "svgf_atrous", f.lib.svgf_last_cuda_error(f._ctx))
    return 0 if res.value == f.FilterBuffer[0].data_ptr() else 1


GPU_OPS = {"temporal_variance": _gpu_temporal_variance, "atrous_level": _gpu_atrous_level, "atrous_levels": _gpu_atrous_levels,
           "as_tensor": lambda t: t}


def make_gpu_banded_filter(W, H, rank, world, device, storage="f16", levels=5, group=None, apron=APRON, exchange_from_level=0,
                           max_motion_rows=0, overlap_state=False, bounds=None):
    from .filter import SvgfFilter
    band = band_of(H, world, rank, apron, bounds)
    f = SvgfFilter(W, band.local_height, device=device, storage=storage)
    f.SpatialFilterSteps = levels
    return BandedFilter(f, band, GPU_OPS, levels=levels, group=group, exchange_from_level=exchange_from_level,
                        max_motion_rows=max_motion_rows, overlap_state=overlap_state)
