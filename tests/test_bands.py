"""Band partition of one frame across ranks (svgf_b200/bands.py, BASELINE config 4).

CPU: world_size-2 gloo run with the scalar oracle as the per-band backend — the stitched bands must be BIT-IDENTICAL
to the oracle on the whole frame (every band pixel sees exactly the inputs it would see in the full image once the
per-level halo rows and the previous-frame state aprons have been exchanged).  GPU (-m gpu): the same through the C
ABI on one device, two bands exchanged in-process, against the unpartitioned CUDA path."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle_lib import OracleFilter, oracle
from svgf_b200 import synth
from svgf_b200.bands import APRON, BandedFilter, balanced_bounds, band_of, check_partition, required_apron

W, H, FRAMES, LEVELS = 64, 96, 3, 5
# (image height, apron rows, first level that exchanges its halo, state exchange posted under levels 1..N-1)
MODES = {
    "exchange_every_level": (96, APRON, 0, False),
    "levels_0_2_redundant_state_overlapped": (96, APRON, 3, True),
    "no_level_exchange_state_overlapped": (160, 76, 5, True),
    "unequal_bands": (96, APRON, 0, False),
}
BOUNDS = {"unequal_bands": [0, 60, 96]}
MAX_MOTION_ROWS = 8


def test_band_geometry():
    for world in (1, 2, 3, 4, 8):
        rows = []
        for r in range(world):
            b = band_of(4320, world, r)
            assert b.ly0 == max(0, b.y0 - APRON) and b.ly1 == min(4320, b.y1 + APRON)
            rows += list(range(b.y0, b.y1))
        assert rows == list(range(4320))                      # bands tile the frame exactly once
    check_partition(4320, 8, 5)
    with pytest.raises(ValueError):
        check_partition(96, 8, 5)                             # 12-row bands cannot feed a 32-row halo
    with pytest.raises(ValueError):
        check_partition(4320, 2, 6)                           # a sixth level needs a 64-row apron
    # halo of the levels that do not exchange + variance window + motion reach, or the widest exchanged halo
    assert required_apron(5, 0) == 32 and required_apron(5, 3, 8) == 32 and required_apron(5, 5, 8) == 3 + 8 + 62
    with pytest.raises(ValueError):
        check_partition(4320, 8, 5, apron=32, exchange_from_level=5, max_motion_rows=8)
    check_partition(4320, 8, 5, apron=76, exchange_from_level=5, max_motion_rows=8)


def test_balanced_bounds():
    cost = np.concatenate([np.full(750, 0.2), np.ones(4320 - 750)])          # cheap background rows on top
    for world in (2, 4, 8):
        b = balanced_bounds(cost, world, min_rows=80)
        assert b[0] == 0 and b[-1] == 4320 and len(b) == world + 1
        work = [cost[b[k]:b[k + 1]].sum() for k in range(world)]
        assert max(work) / (cost.sum() / world) < 1.02                         # within 2 % of a perfect split
        assert min(b1 - b0 for b0, b1 in zip(b, b[1:])) >= 80
        assert b[1] > 4320 // world                                            # the top band is taller than an equal split
    assert balanced_bounds(np.ones(96), 2) == [0, 48, 96]
    assert balanced_bounds(np.r_[np.zeros(90), np.ones(6)], 2, min_rows=32) == [0, 64, 96]   # min_rows wins over balance
    with pytest.raises(ValueError):
        balanced_bounds(np.ones(96), 4, min_rows=32)
    with pytest.raises(ValueError):
        band_of(96, 2, 0, bounds=[0, 96])


# ---- oracle backend for BandedFilter -------------------------------------------------------------------------------
def _o_temporal_variance(o):
    o.TemporalFilter()
    o.FilterMoments()
    return o.FilterBuffer[0]


def _o_atrous_level(o, level, src, dst):
    g = o.gbuf(o.PingPongInx)
    rc = oracle().svgf_oracle_atrous_level(C.byref(o.params), o.Width, o.Height, o.storage, C.byref(g), src.ctypes.data,
                                           dst.ctypes.data, o.RenderBuffer[o.PingPongInx].ctypes.data, level)
    assert rc == 0


def _as_tensor(a):
    return torch.from_numpy(a.reshape(a.shape[0], -1).view(np.uint8))      # rows x bytes, shares memory


ORACLE_OPS = {"temporal_variance": _o_temporal_variance, "atrous_level": _o_atrous_level, "as_tensor": _as_tensor}


def _frames(H=H):
    # vertical motion of ~2.5 px/frame so that reprojection crosses the band boundary
    return [synth.frame_host(W, H, t, vert_px=2.5) for t in range(FRAMES)]


def _full_oracle(H=H):
    o = OracleFilter(W, H, storage="f16")
    o.params.atrous_iterations = LEVELS
    o.Reset()
    outs = []
    for planes in _frames(H):
        o.set_inputs(planes)
        o.Filter()
        outs.append((o.FilterBuffer[0].copy(), o.HistoryLengthBuffer.copy(), o.RenderBuffer[o.PingPongInx].copy()))
        o.EndFrame()
    return outs


def _worker(rank, world, port, tmp, mode="exchange_every_level"):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Hm, apron, from_level, overlap = MODES[mode]
        band = band_of(Hm, world, rank, apron, BOUNDS.get(mode))
        o = OracleFilter(W, band.local_height, storage="f16")
        o.params.atrous_iterations = LEVELS
        o.Reset()
        bf = BandedFilter(o, band, ORACLE_OPS, levels=LEVELS, exchange_from_level=from_level, max_motion_rows=MAX_MOTION_ROWS,
                          overlap_state=overlap)
        for t, planes in enumerate(_frames(Hm)):
            o.set_inputs({k: v[band.ly0:band.ly1] for k, v in planes.items()})
            res = bf.Filter()
            P = o.PingPongInx
            sl = slice(band.loc(band.y0), band.loc(band.y1))
            np.savez(os.path.join(tmp, f"r{rank}_f{t}.npz"), result=res[sl], history=o.HistoryLengthBuffer[sl], colour=o.RenderBuffer[P][sl])
            bf.EndFrame()
        bf.drain()
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("world", [2])
def test_bands_with_halo_exchange_equal_the_whole_frame_gloo(world, mode, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), mode), nprocs=world, join=True)
    want = _full_oracle(MODES[mode][0])
    for t in range(FRAMES):
        parts = [np.load(tmp_path / f"r{r}_f{t}.npz") for r in range(world)]
        for key, idx in (("result", 0), ("history", 1), ("colour", 2)):
            got = np.concatenate([p[key] for p in parts], axis=0)
            assert np.array_equal(got.view(np.uint8), want[t][idx].view(np.uint8)), f"frame {t}: {key} differs from the unpartitioned oracle"


# ---- GPU: two bands on one device, exchanged in-process, against the unpartitioned CUDA path ------------------------
@pytest.mark.gpu
def test_bands_on_gpu_equal_the_whole_frame():
    from svgf_b200 import SvgfFilter
    from svgf_b200.bands import GPU_OPS
    Wg, Hg, world = 256, 192, 2
    frames = [synth.frame_host(Wg, Hg, t, vert_px=2.5) for t in range(FRAMES)]
    full = SvgfFilter(Wg, Hg, storage="f16")
    full.Reset()
    bands = [band_of(Hg, world, r) for r in range(world)]
    parts = [SvgfFilter(Wg, b.local_height, storage="f16") for b in bands]
    for p in parts:
        p.Reset()

    def up(f, planes):
        P = f.PingPongInx
        f.Framebuffer[P].normal.copy_(torch.from_numpy(planes["normal"].view(np.int16)))
        f.Framebuffer[P].uv.copy_(torch.from_numpy(planes["uv"].view(np.int16)))
        f.Framebuffer[P].motion.copy_(torch.from_numpy(planes["motion"]))
        f.RenderBuffer[P].copy_(torch.from_numpy(planes["colour"]))

    def swap(ts, rows):
        # the exchange of svgf_b200.bands.exchange_rows for two bands living in one process
        a, b = bands
        for ta, tb in ts:
            ta[a.loc(a.y1):a.loc(a.y1 + rows)].copy_(tb[b.loc(b.y0):b.loc(b.y0 + rows)])
            tb[b.loc(b.y0 - rows):b.loc(b.y0)].copy_(ta[a.loc(a.y1 - rows):a.loc(a.y1)])

    for t, planes in enumerate(frames):
        up(full, planes)
        full.Filter()
        for f, b in zip(parts, bands):
            up(f, {k: v[b.ly0:b.ly1] for k, v in planes.items()})
        fa, fb = parts
        Q = 1 - fa.PingPongInx
        if t > 0:
            swap([(fa.RenderBuffer[Q], fb.RenderBuffer[Q]), (fa.MomentsBuffer[Q], fb.MomentsBuffer[Q]),
                  (fa.HistoryLengthBuffer, fb.HistoryLengthBuffer)], APRON)
        for f in parts:
            GPU_OPS["temporal_variance"](f)
        k = 0
        for level in range(LEVELS):
            swap([(fa.FilterBuffer[k], fb.FilterBuffer[k])], 2 << level)
            for f in parts:
                GPU_OPS["atrous_level"](f, level, f.FilterBuffer[k], f.FilterBuffer[1 - k])
            k = 1 - k
        torch.cuda.synchronize()
        got = torch.cat([f.FilterBuffer[k][b.loc(b.y0):b.loc(b.y1)] for f, b in zip(parts, bands)]).cpu().numpy()
        want = full.FilterBuffer[0].cpu().numpy()
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), f"frame {t}: banded result differs from the whole-frame CUDA path"
        hist = torch.cat([f.HistoryLengthBuffer[b.loc(b.y0):b.loc(b.y1)] for f, b in zip(parts, bands)]).cpu().numpy()
        assert np.array_equal(hist, full.HistoryLengthBuffer.cpu().numpy())
        full.EndFrame()
        for f in parts:
            f.EndFrame()


@pytest.mark.gpu
@pytest.mark.parametrize("from_level,overlap", [(0, False), (3, True), (5, True)])
def test_banded_driver_on_one_rank_equals_svgf_frame(from_level, overlap):
    """world = 1: no exchange, but the driver's own call sequence (svgf_frame with zero levels, then the a-trous levels in
    groups through svgf_atrous) must reproduce one svgf_frame call bit for bit, in every grouping."""
    from svgf_b200 import SvgfFilter
    from svgf_b200.bands import make_gpu_banded_filter
    Wg, Hg = 256, 192
    full = SvgfFilter(Wg, Hg, storage="f16")
    full.Reset()
    bf = make_gpu_banded_filter(Wg, Hg, 0, 1, torch.device("cuda", 0), exchange_from_level=from_level, overlap_state=overlap)
    bf.f.Reset()
    for t in range(FRAMES):
        planes = synth.frame_host(Wg, Hg, t, vert_px=2.5)
        for f in (full, bf.f):
            P = f.PingPongInx
            f.Framebuffer[P].normal.copy_(torch.from_numpy(planes["normal"].view(np.int16)))
            f.Framebuffer[P].uv.copy_(torch.from_numpy(planes["uv"].view(np.int16)))
            f.Framebuffer[P].motion.copy_(torch.from_numpy(planes["motion"]))
            f.RenderBuffer[P].copy_(torch.from_numpy(planes["colour"]))
        full.Filter()
        res = bf.Filter()
        torch.cuda.synchronize()
        assert torch.equal(res.view(torch.uint8), full.FilterBuffer[0].view(torch.uint8)), f"frame {t}"
        assert torch.equal(bf.f.HistoryLengthBuffer, full.HistoryLengthBuffer)
        assert torch.equal(bf.f.RenderBuffer[P].view(torch.uint8), full.RenderBuffer[P].view(torch.uint8))
        full.EndFrame(); bf.EndFrame()
    bf.drain()
