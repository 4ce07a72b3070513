"""-m gpu: the uniform-normal tile shortcut of the packed a-trous kernel (svgf_kernels_packed.cuh).

A tile whose staged texels all carry one normal vector evaluates the normal weight once instead of per tap.  The
claim is that this never changes a bit of the output; it is pinned two ways: against the oracle on scenes made of
large planar regions (where most tiles take the shortcut), and against the same kernel with the shortcut disabled
(SVGF_FLAG_NO_UNIFORM_TILES) on those scenes and on the procedural benchmark scene."""
import ctypes as C

import numpy as np
import pytest
import torch

from common import assert_close
from gpu_util import load_state_from_oracle, npy, upload_inputs
from oracle_lib import OracleFilter, oracle
from svgf_b200 import SvgfFilter, _lib, synth

pytestmark = pytest.mark.gpu


def planar_scene(rng, W, H, storage):
    """Two planar regions (distinct fp16 unit normals) split by a slanted line, a background strip on top, smooth
    per-pixel depth, white-noise colour with 20 % zero variances."""
    yy, xx = np.mgrid[0:H, 0:W]
    left = xx + 0.3 * yy < 0.55 * W
    na = np.array([0.0, 1.0, 0.0]); nb = np.array([0.6, 0.0, 0.8])
    normal = np.zeros((H, W, 4), np.float16)
    normal[left, :3] = na.astype(np.float16)
    normal[~left, :3] = nb.astype(np.float16)
    normal[..., 3] = 2
    uv = np.zeros((H, W, 4), np.float16)
    uv[..., 3] = np.where(left, 1, 2)
    motion = np.zeros((H, W, 4), np.float32)
    motion[..., 2] = (4.0 + 0.01 * xx + 0.02 * yy + np.where(left, 0.0, 3.0)).astype(np.float32)
    motion[..., 3] = 0.02
    bg = yy < H // 8
    normal[bg] = 0; uv[bg] = 0; motion[bg] = 0
    cdt = np.float16 if storage == "f16" else np.float32
    colour = rng.uniform(0, 1.1, size=(H, W, 4)).astype(cdt)
    colour[..., 3] = (rng.uniform(0, 0.05, size=(H, W)) * (rng.uniform(size=(H, W)) < 0.8)).astype(cdt)
    return {"normal": normal.view(np.uint16), "uv": uv.view(np.uint16), "motion": motion, "colour": colour}


def _one_level(f, level, flags):
    P = f.PingPongInx
    f.params.flags = flags
    res = C.c_void_p()
    gs = f.Framebuffer[P].as_struct()
    st = f.lib.svgf_atrous(f._ctx, C.byref(f.params), C.byref(gs), C.c_void_p(f.FilterBuffer[0].data_ptr()),
                           C.c_void_p(f.FilterBuffer[1].data_ptr()), C.c_void_p(f.RenderBuffer[P].data_ptr()), level, 1,
                           C.byref(res), f._stream())
    assert st == 0
    return f.FilterBuffer[1].clone(), f.RenderBuffer[P].clone()


@pytest.mark.parametrize("storage", ["f16", "f32"])
@pytest.mark.parametrize("level", [0, 1, 2, 3, 4])
def test_planar_scene_every_level_against_the_oracle_and_against_the_general_form(storage, level):
    W, H = 900, 420
    rng = np.random.default_rng(100 + level)
    planes = planar_scene(rng, W, H, storage)
    of = OracleFilter(W, H, storage=storage)
    of.set_inputs(planes)
    of.FilterBuffer[0][...] = planes["colour"]
    P = of.PingPongInx
    g = of.gbuf(P)
    out = np.zeros_like(of.FilterBuffer[0])
    hc = of.RenderBuffer[P].copy()
    assert oracle().svgf_oracle_atrous_level(C.byref(of.params), W, H, of.storage, C.byref(g), of.FilterBuffer[0].ctypes.data,
                                             out.ctypes.data, hc.ctypes.data, level) == 0
    f = SvgfFilter(W, H, storage=storage)
    load_state_from_oracle(f, of)
    got, got_h = _one_level(f, level, 0)
    assert_close(npy(got), out, storage, f"planar scene, level {level}")
    assert_close(npy(got_h), hc, storage, f"planar scene, colour history after level {level}")
    load_state_from_oracle(f, of)
    ref, ref_h = _one_level(f, level, _lib.SVGF_FLAG_NO_UNIFORM_TILES)
    assert torch.equal(got.view(torch.uint8), ref.view(torch.uint8)), "shortcut changed output bits"
    assert torch.equal(got_h.view(torch.uint8), ref_h.view(torch.uint8))


@pytest.mark.parametrize("storage", ["f16", "f32"])
@pytest.mark.parametrize("base_flags", [0, _lib.SVGF_FLAG_NO_STAGED_LEVELS])
def test_benchmark_scene_sequence_is_bit_identical_with_and_without_the_shortcut(storage, base_flags):
    # both the staged run (lattice kernel: uniformity from the segment map) and the level-by-level path (packed kernel:
    # uniformity from the staged texels) form the exponent with the same single FMA as the general form
    W, H, N = 1280, 720, 6
    a, b = SvgfFilter(W, H, storage=storage), SvgfFilter(W, H, storage=storage)
    a.params.flags = base_flags
    b.params.flags = base_flags | _lib.SVGF_FLAG_NO_UNIFORM_TILES
    a.Reset(); b.Reset()
    for t in range(N):
        planes = synth.frame_host(W, H, t, storage=storage)
        upload_inputs(a, planes); upload_inputs(b, planes)
        a.Filter(); b.Filter()
        P = a.PingPongInx
        assert torch.equal(a.FilterBuffer[0].view(torch.uint8), b.FilterBuffer[0].view(torch.uint8)), f"frame {t}: result differs"
        assert torch.equal(a.RenderBuffer[P].view(torch.uint8), b.RenderBuffer[P].view(torch.uint8)), f"frame {t}: history differs"
        a.EndFrame(); b.EndFrame()
