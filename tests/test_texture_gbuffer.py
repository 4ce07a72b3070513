"""-m gpu: G-buffers handed over as the reference hands them - cudaTextureObject_t over cudaArrays (src/App.cu:473-475,
src/CudaUtil.h:68-99) - through the same svgf_gbuffer struct (SVGF_PITCH_TEXTURE).  No GL is needed for that: the arrays
are cudaMallocArray allocations like oracle/ref_harness.cu makes for the reference's own kernels.  Every output must be
BIT-identical to the run on linear planes: the texels are the same, only the fetch path differs."""
import ctypes as C

import numpy as np
import pytest
import torch

from gpu_util import npy, upload_inputs
from svgf_b200 import SvgfFilter, _lib, synth
from svgf_b200.filter import GBuffer, TextureGBuffer

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("storage", ["f16", "f32"])
@pytest.mark.parametrize("size", [(258, 129), (641, 363)])
@pytest.mark.parametrize("reproj", [0, 1])
def test_texture_backed_gbuffers_give_bit_identical_frames(size, storage, reproj):
    W, H = size
    dev = torch.device("cuda", 0)
    a, b = SvgfFilter(W, H, storage=storage), SvgfFilter(W, H, storage=storage)
    a.params.reproj_mode = b.params.reproj_mode = reproj
    a.Reset(); b.Reset()
    tex = [TextureGBuffer(W, H, dev), TextureGBuffer(W, H, dev)]
    lin = [GBuffer(W, H, dev), GBuffer(W, H, dev)]
    for t in tex:
        t.upload(lin[0])                       # both slots start as the zeroed G-buffer of a reset
    b.Framebuffer = tex
    for t in range(5):
        planes = synth.frame_host(W, H, t, storage=storage)
        upload_inputs(a, planes)
        P = b.PingPongInx
        lin[P].normal.copy_(torch.from_numpy(planes["normal"].view(np.int16)))
        lin[P].uv.copy_(torch.from_numpy(planes["uv"].view(np.int16)))
        lin[P].motion.copy_(torch.from_numpy(planes["motion"]))
        tex[P].upload(lin[P])
        b.RenderBuffer[P].copy_(torch.from_numpy(planes["colour"]))
        a.Filter(); b.Filter()
        assert torch.equal(a.HistoryLengthBuffer, b.HistoryLengthBuffer), f"frame {t}"
        for x, y in ((a.FilterBuffer[0], b.FilterBuffer[0]), (a.RenderBuffer[P], b.RenderBuffer[P]), (a.MomentsBuffer[P], b.MomentsBuffer[P])):
            assert torch.equal(x.view(torch.uint8), y.view(torch.uint8)), f"frame {t}"
        a.EndFrame(); b.EndFrame()


def test_stage_entry_points_take_textures_and_mixed_planes():
    # svgf_temporal / svgf_variance / svgf_atrous one by one, with only SOME planes as textures
    W, H = 130, 67
    dev = torch.device("cuda", 0)
    a, b = SvgfFilter(W, H), SvgfFilter(W, H)
    a.Reset(); b.Reset()
    tex = TextureGBuffer(W, H, dev)
    for t in range(2):
        planes = synth.frame_host(W, H, t)
        upload_inputs(a, planes); upload_inputs(b, planes)
        P = a.PingPongInx
        tex.upload(b.Framebuffer[P])
        a.TemporalFilter(); a.FilterMoments(); a.WaveletFilter()
        # b: the current G-buffer has its motion and normal planes as textures, uv as linear memory
        g = b.Framebuffer[P].as_struct()
        tg = tex.as_struct()
        g.motion_depth, g.motion_pitch = tg.motion_depth, tg.motion_pitch
        g.normal_mat, g.normal_pitch = tg.normal_mat, tg.normal_pitch
        gp = b.Framebuffer[1 - P].as_struct()
        lib, ctx, s = b.lib, b._ctx, b._stream()
        v = lambda t_: C.c_void_p(t_.data_ptr())
        assert lib.svgf_temporal(ctx, C.byref(b.params), C.byref(g), C.byref(gp), v(b.RenderBuffer[1 - P]), v(b.RenderBuffer[P]),
                                 v(b.HistoryLengthBuffer), v(b.MomentsBuffer[P]), v(b.MomentsBuffer[1 - P]), s) == 0
        assert lib.svgf_variance(ctx, C.byref(b.params), C.byref(g), v(b.RenderBuffer[P]), v(b.MomentsBuffer[P]), v(b.HistoryLengthBuffer),
                                 v(b.FilterBuffer[0]), s) == 0
        res = C.c_void_p()
        assert lib.svgf_atrous(ctx, C.byref(b.params), C.byref(g), v(b.FilterBuffer[0]), v(b.FilterBuffer[1]), v(b.RenderBuffer[P]), 0, 5,
                               C.byref(res), s) == 0
        out = b.FilterBuffer[0] if res.value == b.FilterBuffer[0].data_ptr() else b.FilterBuffer[1]
        assert torch.equal(a.FilterBuffer[0].view(torch.uint8), out.view(torch.uint8)), f"frame {t}"
        assert torch.equal(a.HistoryLengthBuffer, b.HistoryLengthBuffer)
        if res.value != b.FilterBuffer[0].data_ptr():
            b.FilterBuffer[0].copy_(b.FilterBuffer[1])
        a.EndFrame(); b.EndFrame()
