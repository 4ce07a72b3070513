"""The band schedule on ONE GPU: every band of the frame lives in this process (svgf_band_create_group; halos by device copies
instead of NCCL) and runs the same plan, kernels and row-block ranges as the NCCL driver.  The stitched bands must equal the
whole frame filtered in one piece, bit for bit - result, colour history, moments, history lengths - for 2 to 8 bands, equal
and unequal heights, 2 to 5 levels, both storage formats.  (tests/test_band_driver.py runs the NCCL transport when the box
has two or more GPUs; bench.py --gpus N records the same check.)"""
import pytest

pytestmark = pytest.mark.gpu


def _run(W, H, world, levels, storage, bounds=None, frames=5, own_streams=False):
    import torch
    from svgf_b200 import SvgfFilter, synth
    from svgf_b200.band_driver import BandGroup
    from svgf_b200.filter import GBuffer
    dev = torch.device("cuda", 0)
    cdt = torch.float16 if storage == "f16" else torch.float32
    grp = BandGroup(W, H, [dev] * world, storage=storage, levels=levels, bounds=bounds)
    if own_streams:
        grp.streams = [torch.cuda.Stream(dev) for _ in range(world)]
    whole = SvgfFilter(W, H, device=dev, storage=storage)
    whole.SpatialFilterSteps = levels
    whole.Reset()
    grp.Reset()
    full_g, full_c = GBuffer(W, H, dev), torch.empty(H, W, 4, dtype=cdt, device=dev)
    bad = {}
    assert [b.y0 for b in grp.bands] == ([0] + [b.y1 for b in grp.bands[:-1]]) and grp.bands[-1].y1 == H
    for t in range(frames):
        synth.frame_device(full_g, full_c, t, seed=0)
        P = whole.PingPongInx
        for b in grp.bands:
            sl = b.local_rows()
            b.Framebuffer[P].normal.copy_(full_g.normal[sl]); b.Framebuffer[P].uv.copy_(full_g.uv[sl]); b.Framebuffer[P].motion.copy_(full_g.motion[sl])
            b.RenderBuffer[P].copy_(full_c[sl])
        whole.Framebuffer[P].normal.copy_(full_g.normal); whole.Framebuffer[P].uv.copy_(full_g.uv); whole.Framebuffer[P].motion.copy_(full_g.motion)
        whole.RenderBuffer[P].copy_(full_c)
        grp.Filter()
        grp.sync()
        whole.Filter()
        ref = {"result": whole.FilterBuffer[0], "colour_history": whole.RenderBuffer[P], "moments": whole.MomentsBuffer[P],
               "history": whole.HistoryLengthBuffer}
        for b in grp.bands:
            own = slice(b.y0 - b.ly0, b.y1 - b.ly0)
            mine = {"result": b.FilterBuffer[0][own], "colour_history": b.RenderBuffer[P][own], "moments": b.MomentsBuffer[P][own],
                    "history": b.HistoryLengthBuffer[own]}
            for name, plane in mine.items():
                n = int((plane.contiguous().view(torch.uint8) != ref[name][b.y0:b.y1].contiguous().view(torch.uint8)).sum())
                if n:
                    bad[f"frame{t}.band{b.rank}.{name}"] = n
        whole.EndFrame()
        grp.EndFrame()
    torch.cuda.synchronize()
    launches = [b.launches for b in grp.bands]
    grp.close()
    assert not bad, bad
    return launches


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("storage", ["f16", "f32"])
def test_bands_on_one_gpu_equal_the_whole_frame(world, storage):
    _run(1920, 1080, world, 5, storage)


@pytest.mark.parametrize("levels", [2, 3, 4])
def test_fewer_levels(levels):
    _run(1280, 720, 3, levels, "f16")


def test_unequal_heights_and_a_ragged_frame():
    _run(1030, 420, 3, 5, "f32", bounds=[0, 100, 290, 420])
    _run(1030, 420, 4, 5, "f16", bounds=[0, 64, 200, 333, 420])


def test_each_band_on_its_own_stream():
    _run(1920, 1080, 4, 5, "f16", own_streams=True)


def test_an_8k_frame_in_eight_bands():
    """the shape bench.py --gpus 8 runs (7680 x 4320, 540-row bands): interior bands launch both boundary strips in one grid"""
    _run(7680, 4320, 8, 5, "f16", frames=3)


def test_group_members_refuse_the_per_rank_entry_point():
    import ctypes as C
    import torch
    from svgf_b200 import _lib
    from svgf_b200.band_driver import BandGroup
    grp = BandGroup(512, 256, [torch.device("cuda", 0)] * 2)
    b = grp.bands[0]
    bufs = b._bufs()
    g = (_lib.SvgfGBuffer * 2)(b.Framebuffer[0].as_struct(), b.Framebuffer[1].as_struct())
    st = _lib.lib().svgf_band_frame(b._h, C.byref(grp.params), C.byref(g), C.byref(bufs), b._stream())
    assert st == _lib.SVGF_INVALID_ARG
    grp.close()
