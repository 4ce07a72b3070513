"""-m gpu: SVGF_VARIANCE_PREFILTER_GAUSS3 - the SVGF paper's 3x3 Gaussian blur of the variance that scales the luminance
edge-stopping term of every a-trous level (include/svgf.h; the reference leaves it out, src/Filter.cuh:547,562).

The blur is a pre-pass (svgf_kernels_basic.cuh variance_gauss3_kernel: vertical taps per lane, horizontal taps by warp
shuffle) read by the level kernel for the centre pixel.  Checked per level against the scalar oracle with the same
switch, in both storage modes, through the packed kernel (even width) and the per-pixel kernel (odd width), and over a
short sequence through svgf_frame."""
import ctypes as C

import numpy as np
import pytest
import torch

from common import assert_close
from gpu_util import load_state_from_oracle, npy, upload_inputs
from oracle_lib import OracleFilter, oracle
from svgf_b200 import SvgfFilter, _lib, synth
from test_parity_sequence import run_sequence
from test_uniform_tiles import planar_scene

pytestmark = pytest.mark.gpu
GAUSS3 = 1


def _one_level(f, level):
    P = f.PingPongInx
    res = C.c_void_p()
    gs = f.Framebuffer[P].as_struct()
    st = f.lib.svgf_atrous(f._ctx, C.byref(f.params), C.byref(gs), C.c_void_p(f.FilterBuffer[0].data_ptr()),
                           C.c_void_p(f.FilterBuffer[1].data_ptr()), C.c_void_p(f.RenderBuffer[P].data_ptr()), level, 1,
                           C.byref(res), f._stream())
    assert st == 0
    return f.FilterBuffer[1].clone(), f.RenderBuffer[P].clone()


@pytest.mark.parametrize("storage", ["f16", "f32"])
@pytest.mark.parametrize("size", [(900, 420), (333, 97)])     # packed kernel / per-pixel kernel (odd width)
@pytest.mark.parametrize("level", [0, 1, 2, 3, 4])
def test_every_level_against_the_oracle(storage, size, level):
    W, H = size
    rng = np.random.default_rng(300 + level + W)
    planes = planar_scene(rng, W, H, storage)
    # blocky variance so that the blurred value differs strongly from the centre's own
    blk = (rng.uniform(size=(H // 3 + 1, W // 3 + 1)) < 0.4).repeat(3, 0).repeat(3, 1)[:H, :W]
    planes["colour"][..., 3] = np.where(blk, 0.6, planes["colour"][..., 3]).astype(planes["colour"].dtype)
    outs = {}
    for mode in (0, GAUSS3):
        of = OracleFilter(W, H, storage=storage)
        of.params.variance_prefilter = mode
        of.set_inputs(planes)
        of.FilterBuffer[0][...] = planes["colour"]
        P = of.PingPongInx
        g = of.gbuf(P)
        want = np.zeros_like(of.FilterBuffer[0])
        hc = of.RenderBuffer[P].copy()
        assert oracle().svgf_oracle_atrous_level(C.byref(of.params), W, H, of.storage, C.byref(g), of.FilterBuffer[0].ctypes.data,
                                                 want.ctypes.data, hc.ctypes.data, level) == 0
        f = SvgfFilter(W, H, storage=storage)
        f.params.variance_prefilter = mode
        load_state_from_oracle(f, of)
        got, got_h = _one_level(f, level)
        assert_close(npy(got), want, storage, f"prefilter {mode}, level {level}")
        assert_close(npy(got_h), hc, storage, f"prefilter {mode}, colour history after level {level}")
        outs[mode] = want.astype(np.float32)
    assert np.abs(outs[0] - outs[GAUSS3]).max() > 1e-2, "the prefilter changed nothing on this scene"


@pytest.mark.parametrize("storage", ["f16", "f32"])
def test_sequence_through_svgf_frame(storage):
    """Teacher-forced pan sequence at the bars of tests/test_parity_sequence.py, five prefiltered levels per frame."""
    run_sequence(640, 360, 6, storage, teacher_forced=True, variance_prefilter=GAUSS3)


def test_fused_levels_flag_is_ignored_with_the_prefilter():
    W, H = 256, 96
    a, b = SvgfFilter(W, H), SvgfFilter(W, H)
    for f in (a, b):
        f.params.variance_prefilter = GAUSS3
        f.Reset()
        upload_inputs(f, synth.frame_host(W, H, 0))
    b.params.flags = _lib.SVGF_FLAG_FUSE_LEVELS_01
    la, lb = a.launches, b.launches
    a.Filter(); b.Filter()
    assert a.launches - la == b.launches - lb
    assert torch.equal(a.FilterBuffer[0].view(torch.uint8), b.FilterBuffer[0].view(torch.uint8))
