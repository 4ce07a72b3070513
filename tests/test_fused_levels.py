"""-m gpu: a-trous levels 0 and 1 as one launch (svgf_kernels_fused.cuh, SVGF_FLAG_FUSE_LEVELS_01) - the
"two-level-fused" variant of BASELINE configs[2].

The fused kernel keeps the level-0 result in shared memory, rounded through the storage format like a plane written by
one launch and read by the next, so its output must be BIT-identical to two single-level launches: final plane and the
colour history written by level 0, in both storage modes, at sizes that are not multiples of the 64 x 24 tile, through
svgf_atrous and through svgf_frame (where the fused launch changes the ping-pong parity)."""
import ctypes as C

import numpy as np
import pytest
import torch

from common import assert_close
from gpu_util import load_state_from_oracle, npy, upload_inputs
from oracle_lib import OracleFilter, oracle
from svgf_b200 import SvgfFilter, _lib, synth
from test_uniform_tiles import planar_scene

pytestmark = pytest.mark.gpu


def _levels(f, n, flags):
    P = f.PingPongInx
    f.params.flags = flags
    res = C.c_void_p()
    gs = f.Framebuffer[P].as_struct()
    st = f.lib.svgf_atrous(f._ctx, C.byref(f.params), C.byref(gs), C.c_void_p(f.FilterBuffer[0].data_ptr()),
                           C.c_void_p(f.FilterBuffer[1].data_ptr()), C.c_void_p(f.RenderBuffer[P].data_ptr()), 0, n,
                           C.byref(res), f._stream())
    assert st == 0
    out = f.FilterBuffer[0] if res.value == f.FilterBuffer[0].data_ptr() else f.FilterBuffer[1]
    assert res.value == out.data_ptr()
    return out.clone(), f.RenderBuffer[P].clone()


@pytest.mark.parametrize("storage", ["f16", "f32"])
@pytest.mark.parametrize("size", [(900, 420), (64, 24), (130, 50), (258, 97)])
def test_two_levels_fused_equal_two_launches_and_the_oracle(storage, size):
    W, H = size
    rng = np.random.default_rng(7 + W)
    planes = planar_scene(rng, W, H, storage)
    of = OracleFilter(W, H, storage=storage)
    of.set_inputs(planes)
    of.FilterBuffer[0][...] = planes["colour"]
    P = of.PingPongInx
    g = of.gbuf(P)
    f = SvgfFilter(W, H, storage=storage)
    load_state_from_oracle(f, of)
    l0 = f.launches
    ref, ref_h = _levels(f, 2, _lib.SVGF_FLAG_NO_STAGED_LEVELS)    # two single-level launches of the packed kernel
    load_state_from_oracle(f, of)
    l1 = f.launches
    got, got_h = _levels(f, 2, _lib.SVGF_FLAG_FUSE_LEVELS_01)
    assert (l1 - l0) - (f.launches - l1) == 1, "levels 0 and 1 did not go out as one launch"
    assert torch.equal(got.view(torch.uint8), ref.view(torch.uint8)), "fusion changed output bits"
    assert torch.equal(got_h.view(torch.uint8), ref_h.view(torch.uint8)), "fusion changed the colour history"
    # against the oracle, teacher-forced: level 1 of the oracle runs on the level-0 plane the GPU produced (chaining two
    # oracle levels would compare two free-running paths through weights that are ill-conditioned where the variance is 0)
    load_state_from_oracle(f, of)
    mid, _ = _levels(f, 1, 0)
    mid_np = np.ascontiguousarray(npy(mid)).astype(of.FilterBuffer[0].dtype)
    lvl1 = np.zeros_like(of.FilterBuffer[0])
    hc = of.RenderBuffer[P].copy()
    assert oracle().svgf_oracle_atrous_level(C.byref(of.params), W, H, of.storage, C.byref(g), mid_np.ctypes.data, lvl1.ctypes.data,
                                             hc.ctypes.data, 1) == 0
    assert_close(npy(got), lvl1, storage, "fused levels 0+1 vs oracle level 1 on the GPU's level-0 plane")
    lvl0 = np.zeros_like(of.FilterBuffer[0])
    assert oracle().svgf_oracle_atrous_level(C.byref(of.params), W, H, of.storage, C.byref(g), of.FilterBuffer[0].ctypes.data,
                                             lvl0.ctypes.data, hc.ctypes.data, 0) == 0
    assert_close(npy(got_h), hc, storage, "colour history of the fused launch vs oracle")


@pytest.mark.parametrize("storage", ["f16", "f32"])
@pytest.mark.parametrize("levels", [2, 5])
def test_sequence_through_svgf_frame_is_bit_identical_fused_and_unfused(storage, levels):
    W, H, N = 1280, 720, 5
    a, b = SvgfFilter(W, H, storage=storage), SvgfFilter(W, H, storage=storage)
    a.SpatialFilterSteps = b.SpatialFilterSteps = levels
    a.params.flags = _lib.SVGF_FLAG_NO_STAGED_LEVELS
    b.params.flags = _lib.SVGF_FLAG_FUSE_LEVELS_01
    a.Reset(); b.Reset()
    for t in range(N):
        planes = synth.frame_host(W, H, t, storage=storage)
        upload_inputs(a, planes); upload_inputs(b, planes)
        la, lb = a.launches, b.launches
        a.Filter(); b.Filter()
        assert (a.launches - la) - (b.launches - lb) == 1
        P = a.PingPongInx
        assert torch.equal(a.FilterBuffer[0].view(torch.uint8), b.FilterBuffer[0].view(torch.uint8)), f"frame {t}: result differs"
        assert torch.equal(a.RenderBuffer[P].view(torch.uint8), b.RenderBuffer[P].view(torch.uint8)), f"frame {t}: history differs"
        a.EndFrame(); b.EndFrame()


def test_no_level_fusion_flag_wins():
    W, H = 256, 96
    f = SvgfFilter(W, H)
    f.Reset()
    upload_inputs(f, synth.frame_host(W, H, 0))
    f.params.flags = _lib.SVGF_FLAG_FUSE_LEVELS_01 | _lib.SVGF_FLAG_NO_LEVEL_FUSION
    g = SvgfFilter(W, H)
    g.Reset()
    upload_inputs(g, synth.frame_host(W, H, 0))
    lf, lg = f.launches, g.launches
    f.Filter(); g.Filter()
    assert f.launches - lf == g.launches - lg
