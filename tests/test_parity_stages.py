"""-m gpu: every stage of the CUDA path against the scalar oracle, called through the C ABI on the same
seeded inputs — including ragged / tiny images, pitched G-buffers and both storage modes."""
import ctypes as C

import numpy as np
import pytest
import torch

from common import assert_close, random_scene
from gpu_util import load_state_from_oracle, npy
from oracle_lib import OracleFilter
from svgf_b200 import SvgfFilter, _lib

pytestmark = pytest.mark.gpu

SIZES = [(1, 1), (5, 3), (31, 9), (64, 64), (130, 67), (257, 129)]


def make_pair(W, H, storage, seed, motion=2.5):
    """An oracle filter in a random mid-sequence state and a CUDA filter holding the same state."""
    rng = np.random.default_rng(seed)
    of = OracleFilter(W, H, storage=storage)
    cur = random_scene(rng, W, H, storage=storage, max_motion=motion)
    prev = random_scene(rng, W, H, storage=storage)
    keep = rng.uniform(size=(H, W)) < 0.7       # ~70 % of the pixels can reproject
    for k in ("normal", "uv"):
        prev[k][keep] = cur[k][keep]
    prev["motion"][keep, 2:] = cur["motion"][keep, 2:]
    of.PingPongInx = int(rng.integers(0, 2))
    P, Q = of.PingPongInx, 1 - of.PingPongInx
    of.set_inputs(cur)
    of.normal[Q][...] = prev["normal"]; of.uv[Q][...] = prev["uv"]; of.motion[Q][...] = prev["motion"]
    cdt = of.RenderBuffer[0].dtype
    of.RenderBuffer[Q][...] = rng.uniform(0, 1.2, size=(H, W, 4)).astype(cdt)
    of.MomentsBuffer[Q][...] = rng.uniform(0, 1, size=(H, W, 2)).astype(cdt)
    of.HistoryLengthBuffer[...] = rng.integers(0, 30, size=(H, W)).astype(np.uint8)
    f = SvgfFilter(W, H, storage=storage)
    load_state_from_oracle(f, of)
    return of, f


@pytest.mark.parametrize("storage", ["f16", "f32"])
@pytest.mark.parametrize("size", SIZES)
def test_temporal_is_bit_exact(size, storage):
    W, H = size
    of, f = make_pair(W, H, storage, seed=W * 1000 + H)
    of.TemporalFilter()
    f.TemporalFilter()
    P = of.PingPongInx
    assert np.array_equal(npy(f.HistoryLengthBuffer), of.HistoryLengthBuffer)        # integer work: bit-exact
    # the temporal pass is written without FMA contraction: same roundings as the oracle
    assert np.array_equal(npy(f.RenderBuffer[P]).view(np.uint8), of.RenderBuffer[P].view(np.uint8))
    assert np.array_equal(npy(f.MomentsBuffer[P]).view(np.uint8), of.MomentsBuffer[P].view(np.uint8))


@pytest.mark.parametrize("storage", ["f16", "f32"])
@pytest.mark.parametrize("size", SIZES)
def test_variance(size, storage):
    W, H = size
    of, f = make_pair(W, H, storage, seed=W * 77 + H)
    of.TemporalFilter(); f.TemporalFilter()
    # force a mix of short and long histories
    rng = np.random.default_rng(5)
    of.HistoryLengthBuffer[...] = rng.integers(1, 8, size=(H, W)).astype(np.uint8)
    f.HistoryLengthBuffer.copy_(torch.from_numpy(of.HistoryLengthBuffer))
    of.FilterMoments(); f.FilterMoments()
    assert_close(npy(f.FilterBuffer[0]), of.FilterBuffer[0], storage, f"variance {size}")


@pytest.mark.parametrize("storage", ["f16", "f32"])
@pytest.mark.parametrize("size", SIZES)
@pytest.mark.parametrize("levels", [1, 2, 5])
def test_atrous(size, storage, levels):
    W, H = size
    of, f = make_pair(W, H, storage, seed=W * 13 + H + levels)
    cdt = of.FilterBuffer[0].dtype
    rng = np.random.default_rng(9)
    start = rng.uniform(0, 1.1, size=(H, W, 4)).astype(cdt)
    start[..., 3] = (rng.uniform(0, 0.05, size=(H, W)) * (rng.uniform(size=(H, W)) < 0.8)).astype(cdt)  # 20 % exact zeros
    of.FilterBuffer[0][...] = start
    f.FilterBuffer[0].copy_(torch.from_numpy(start))
    of.params.atrous_iterations = f.params.atrous_iterations = levels
    of.WaveletFilter(); f.WaveletFilter()
    P = of.PingPongInx
    assert_close(npy(f.FilterBuffer[0]), of.FilterBuffer[0], storage, f"atrous x{levels} {size}", max_flips=0.05)
    assert_close(npy(f.RenderBuffer[P]), of.RenderBuffer[P], storage, f"colour history {size}")


@pytest.mark.parametrize("storage", ["f16", "f32"])
def test_frame_equals_stage_by_stage(storage):
    W, H = 200, 120
    of, f = make_pair(W, H, storage, seed=42)
    _, f2 = make_pair(W, H, storage, seed=42)
    f.TemporalFilter(); f.FilterMoments(); f.WaveletFilter()
    f2.Filter()
    P = f.PingPongInx
    for a, b in ((f.FilterBuffer[0], f2.FilterBuffer[0]), (f.RenderBuffer[P], f2.RenderBuffer[P]),
                 (f.MomentsBuffer[P], f2.MomentsBuffer[P]), (f.HistoryLengthBuffer, f2.HistoryLengthBuffer)):
        assert torch.equal(a, b)


def test_pitched_gbuffer_planes():
    W, H = 70, 33
    of, f = make_pair(W, H, "f16", seed=3)
    lib = _lib.lib()
    pad = 5                                  # texels of row padding
    P, Q = f.PingPongInx, 1 - f.PingPongInx
    planes = {}

    def pitched(t):
        big = torch.full((H, W + pad, 4), -7, dtype=t.dtype, device=t.device)
        big[:, :W] = t
        return big

    structs = []
    for k in (P, Q):
        g = _lib.SvgfGBuffer()
        n, u, m = pitched(f.Framebuffer[k].normal), pitched(f.Framebuffer[k].uv), pitched(f.Framebuffer[k].motion)
        planes[k] = (n, u, m)
        g.normal_mat, g.normal_pitch = n.data_ptr(), (W + pad) * 8
        g.uv_inst, g.uv_pitch = u.data_ptr(), (W + pad) * 8
        g.motion_depth, g.motion_pitch = m.data_ptr(), (W + pad) * 16
        structs.append(g)
    of.TemporalFilter()
    st = lib.svgf_temporal(f._ctx, C.byref(f.params), C.byref(structs[0]), C.byref(structs[1]),
                           C.c_void_p(f.RenderBuffer[Q].data_ptr()), C.c_void_p(f.RenderBuffer[P].data_ptr()),
                           C.c_void_p(f.HistoryLengthBuffer.data_ptr()), C.c_void_p(f.MomentsBuffer[P].data_ptr()),
                           C.c_void_p(f.MomentsBuffer[Q].data_ptr()), f._stream())
    assert st == 0
    assert np.array_equal(npy(f.HistoryLengthBuffer), of.HistoryLengthBuffer)
    assert np.array_equal(npy(f.RenderBuffer[P]).view(np.uint8), of.RenderBuffer[P].view(np.uint8))


def test_argument_validation_through_the_abi():
    f = SvgfFilter(32, 16)
    f.params.history_cap = 0
    with pytest.raises(_lib.SvgfError) as e:
        f.Filter()
    assert e.value.status == _lib.SVGF_INVALID_ARG
    f.params.history_cap = 256
    with pytest.raises(_lib.SvgfError):
        f.TemporalFilter()
    f.params.history_cap = 24
    f.params.reproj_mode = 1
    with pytest.raises(_lib.SvgfError) as e:
        f.Filter()
    assert e.value.status == _lib.SVGF_UNSUPPORTED
    f.params.reproj_mode = 0
    f.params.atrous_iterations = 11
    with pytest.raises(_lib.SvgfError):
        f.Filter()
    f.params.atrous_iterations = 0
    f.Filter()                                   # N = 0 is legal (GUI range 0..10)
    ctx = C.c_void_p()
    assert _lib.lib().svgf_create(C.byref(ctx), 0, 0, 16, 0) == _lib.SVGF_INVALID_ARG
    assert _lib.lib().svgf_create(C.byref(ctx), 99, 16, 16, 0) == _lib.SVGF_CUDA_ERROR
    assert _lib.lib().svgf_create(C.byref(ctx), 0, 16, 16, 7) == _lib.SVGF_INVALID_ARG
