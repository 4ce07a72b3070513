"""The remaining switches of the SVGF paper the reference leaves out or leaves commented out:
the RELATIVE depth test of the reprojection (src/Filter.cuh:241, the commented-out line) and albedo demodulation
(README.md:14,172-174).  Both default off; each has an oracle switch, and - integer / element-wise IEEE work - is compared
bit for bit."""
import ctypes as C

import numpy as np
import pytest

from common import decode_gbuf, random_scene
from oracle_lib import OracleFilter, oracle
import oracle_py
from svgf_b200 import _lib


def _pair_state(W, H, storage, seed):
    rng = np.random.default_rng(seed)
    of = OracleFilter(W, H, storage=storage)
    cur = random_scene(rng, W, H, storage=storage, max_motion=2.5)
    prev = random_scene(rng, W, H, storage=storage)
    keep = rng.uniform(size=(H, W)) < 0.7
    for k in ("normal", "uv"):
        prev[k][keep] = cur[k][keep]
    # previous depths scattered around the current ones so that both forms of the test accept some and reject some
    prev["motion"][keep, 2] = cur["motion"][keep, 2] + rng.normal(scale=0.05, size=int(keep.sum())).astype(np.float32)
    P, Q = 0, 1
    of.set_inputs(cur)
    of.normal[Q][...] = prev["normal"]; of.uv[Q][...] = prev["uv"]; of.motion[Q][...] = prev["motion"]
    cdt = of.RenderBuffer[0].dtype
    of.RenderBuffer[Q][...] = rng.uniform(0, 1.2, size=(H, W, 4)).astype(cdt)
    of.MomentsBuffer[Q][...] = rng.uniform(0, 1, size=(H, W, 2)).astype(cdt)
    of.HistoryLengthBuffer[...] = rng.integers(0, 30, size=(H, W)).astype(np.uint8)
    return of, cur, prev


def test_oracle_relative_depth_test_against_the_python_restatement():
    W, H = 24, 17
    of, cur, prev = _pair_state(W, H, "f32", 3)
    of.params.depth_test_mode = _lib.SVGF_DEPTH_TEST_RELATIVE
    of.params.depth_threshold = 0.8
    h0 = of.HistoryLengthBuffer.copy()
    col_prev, mom_prev, col_cur = of.RenderBuffer[1].copy(), of.MomentsBuffer[1].copy(), of.RenderBuffer[0].copy()
    of.TemporalFilter()
    want_c, want_h, want_m = oracle_py.temporal(of.params, decode_gbuf(cur), decode_gbuf(prev), col_prev, col_cur, h0, mom_prev)
    assert np.array_equal(of.HistoryLengthBuffer, want_h)
    # the switch matters on this input: the absolute test decides differently for some pixels
    of2, _, _ = _pair_state(W, H, "f32", 3)
    of2.TemporalFilter()
    assert (of2.HistoryLengthBuffer != of.HistoryLengthBuffer).any()


def test_oracle_demodulate_remodulate_known_answers():
    W, H = 5, 3
    alb = np.full((H, W, 4), 0.5, np.float32); alb[0, 0, :3] = 0.0
    col = np.full((H, W, 4), 0.25, np.float32); col[..., 3] = 0.125
    c = col.copy()
    assert oracle().svgf_oracle_demodulate(W, H, 1, alb.ctypes.data, c.ctypes.data) == 0
    assert np.all(c[1:, :, :3] == 0.5) and c[0, 0, 0] == np.float32(0.25) / np.float32(1e-3) and np.all(c[..., 3] == 0.125)
    out = np.zeros_like(c)
    assert oracle().svgf_oracle_remodulate(W, H, 1, alb.ctypes.data, c.ctypes.data, out.ctypes.data) == 0
    assert np.all(out[1:, :, :3] == 0.25) and np.all(out[..., 3] == 0.125)


@pytest.mark.gpu
@pytest.mark.parametrize("storage", ["f16", "f32"])
@pytest.mark.parametrize("reproj", [0, 1])
def test_relative_depth_test_is_bit_exact_on_the_gpu(storage, reproj):
    from gpu_util import load_state_from_oracle, npy
    from svgf_b200 import SvgfFilter
    W, H = 130, 67
    of, _, _ = _pair_state(W, H, storage, 11)
    of.params.depth_test_mode = _lib.SVGF_DEPTH_TEST_RELATIVE
    of.params.reproj_mode = reproj
    f = SvgfFilter(W, H, storage=storage)
    f.params.depth_test_mode = _lib.SVGF_DEPTH_TEST_RELATIVE
    f.params.reproj_mode = reproj
    load_state_from_oracle(f, of)
    of.TemporalFilter(); f.TemporalFilter()
    assert np.array_equal(npy(f.HistoryLengthBuffer), of.HistoryLengthBuffer)
    assert np.array_equal(npy(f.MomentsBuffer[0]).view(np.uint8), of.MomentsBuffer[0].view(np.uint8))
    assert np.array_equal(npy(f.RenderBuffer[0]).view(np.uint8), of.RenderBuffer[0].view(np.uint8))
    f.params.depth_test_mode = 7
    with pytest.raises(_lib.SvgfError):
        f.TemporalFilter()


@pytest.mark.gpu
@pytest.mark.parametrize("storage", ["f16", "f32"])
def test_albedo_demodulation_round_trip_is_bit_exact_on_the_gpu(storage):
    import torch
    from gpu_util import npy
    from svgf_b200 import SvgfFilter
    W, H = 257, 129
    rng = np.random.default_rng(5)
    cdt = np.float16 if storage == "f16" else np.float32
    alb = rng.uniform(0.0, 1.0, size=(H, W, 4)).astype(cdt); alb[rng.uniform(size=(H, W)) < 0.05, :3] = 0
    col = rng.uniform(0.0, 1.0, size=(H, W, 4)).astype(cdt)
    f = SvgfFilter(W, H, storage=storage)
    st = 0 if storage == "f16" else 1
    a_d, c_d = torch.from_numpy(alb).cuda(), torch.from_numpy(col).cuda()
    v = lambda t: C.c_void_p(t.data_ptr())
    assert f.lib.svgf_demodulate(f._ctx, v(a_d), v(c_d), f._stream()) == 0
    want = col.copy()
    assert oracle().svgf_oracle_demodulate(W, H, st, alb.ctypes.data, want.ctypes.data) == 0
    assert np.array_equal(npy(c_d).view(np.uint8), want.view(np.uint8))
    out_d = torch.empty_like(c_d)
    assert f.lib.svgf_remodulate(f._ctx, v(a_d), v(c_d), v(out_d), f._stream()) == 0
    want2 = np.zeros_like(want)
    assert oracle().svgf_oracle_remodulate(W, H, st, alb.ctypes.data, want.ctypes.data, want2.ctypes.data) == 0
    assert np.array_equal(npy(out_d).view(np.uint8), want2.view(np.uint8))
    assert f.lib.svgf_demodulate(f._ctx, v(a_d), v(a_d), f._stream()) == _lib.SVGF_INVALID_ARG
