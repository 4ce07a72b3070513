"""-m gpu: every stage of the CUDA path against the scalar oracle, called through the C ABI on the same
seeded inputs — including ragged / tiny images, pitched G-buffers and both storage modes."""
import ctypes as C

import numpy as np
import pytest
import torch

from common import F32_FLOOR_RADIANCE, F32_REL_TOL, assert_close, random_scene, rel_err
from gpu_util import load_state_from_oracle, npy
from oracle_lib import OracleFilter, oracle
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
    got, want = npy(f.FilterBuffer[0]), of.FilterBuffer[0]
    if storage == "f32":
        # variance = (M2 - M1^2) * 4/h cancels: with these random moments M2 ~ 0.5 and 4/h up to 4, one fp32 ulp
        # of M2 is 2.4e-7 of variance; relative error is taken against max(|var|, 0.05)
        assert rel_err(got[..., :3], want[..., :3], F32_FLOOR_RADIANCE) <= F32_REL_TOL
        assert rel_err(got[..., 3], want[..., 3], 5e-2) <= F32_REL_TOL
    else:
        # fp16 variance values near the cancellation point flip more often; bound the size of the flips
        assert_close(got, want, storage, f"variance {size}", max_flips=0.10)


def _atrous_inputs(of, f, rng, smooth):
    H, W = of.Height, of.Width
    cdt = of.FilterBuffer[0].dtype
    if smooth:
        # smooth shading + noise whose variance channel is consistent with the noise (what the filter is for)
        yy, xx = np.mgrid[0:H, 0:W]
        base = 0.45 + 0.3 * np.sin(xx / 17.0)[..., None] * np.cos(yy / 11.0)[..., None] * np.array([1.0, 0.8, 0.6])
        sigma = 0.05
        start = np.empty((H, W, 4), np.float64)
        start[..., :3] = base + rng.normal(scale=sigma, size=(H, W, 3))
        start[..., 3] = sigma * sigma * rng.uniform(0.5, 1.5, size=(H, W))
        start = start.astype(cdt)
    else:
        start = rng.uniform(0, 1.1, size=(H, W, 4)).astype(cdt)      # white noise, some > 1 (clamp)
        start[..., 3] = (rng.uniform(0, 0.05, size=(H, W)) * (rng.uniform(size=(H, W)) < 0.8)).astype(cdt)  # 20 % exact zeros
    of.FilterBuffer[0][...] = start
    f.FilterBuffer[0].copy_(torch.from_numpy(start))


@pytest.mark.parametrize("storage", ["f16", "f32"])
@pytest.mark.parametrize("size", SIZES)
@pytest.mark.parametrize("level", [0, 1, 2, 3, 4])
def test_atrous_single_level(size, storage, level):
    # every level on its own from identical inputs (white-noise colours, 20 % zero variances: the worst case for
    # the edge-stopping functions): this is the kernel's own error, nothing to amplify it
    W, H = size
    of, f = make_pair(W, H, storage, seed=W * 13 + H + level)
    _atrous_inputs(of, f, np.random.default_rng(9 + level), smooth=False)
    P = of.PingPongInx
    g = of.gbuf(P)
    out = np.zeros_like(of.FilterBuffer[0])
    hc = of.RenderBuffer[P].copy()
    assert oracle().svgf_oracle_atrous_level(C.byref(of.params), W, H, of.storage, C.byref(g), of.FilterBuffer[0].ctypes.data,
                                             out.ctypes.data, hc.ctypes.data, level) == 0
    res = C.c_void_p()
    gs = f.Framebuffer[P].as_struct()
    st = f.lib.svgf_atrous(f._ctx, C.byref(f.params), C.byref(gs), C.c_void_p(f.FilterBuffer[0].data_ptr()),
                           C.c_void_p(f.FilterBuffer[1].data_ptr()), C.c_void_p(f.RenderBuffer[P].data_ptr()), level, 1,
                           C.byref(res), f._stream())
    assert st == 0 and res.value == f.FilterBuffer[1].data_ptr()
    assert_close(npy(f.FilterBuffer[1]), out, storage, f"a-trous level {level} {size}")
    assert_close(npy(f.RenderBuffer[P]), hc, storage, f"colour history after level {level} {size}")


@pytest.mark.parametrize("storage", ["f16", "f32"])
@pytest.mark.parametrize("size", SIZES[2:])
@pytest.mark.parametrize("levels", [2, 5])
def test_atrous_cascade(size, storage, levels):
    # the full cascade through one svgf_atrous call on well-conditioned content
    W, H = size
    of, f = make_pair(W, H, storage, seed=W * 13 + H + levels)
    _atrous_inputs(of, f, np.random.default_rng(21), smooth=True)
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
    f.params.reproj_mode = 2           # 0 = the reference's truncated nearest fetch, 1 = bilinear; nothing else exists
    with pytest.raises(_lib.SvgfError) as e:
        f.Filter()
    assert e.value.status == _lib.SVGF_UNSUPPORTED
    f.params.reproj_mode = 0
    f.params.variance_prefilter = 2
    with pytest.raises(_lib.SvgfError) as e:
        f.Filter()
    assert e.value.status == _lib.SVGF_UNSUPPORTED
    f.params.variance_prefilter = 0
    f.params.atrous_iterations = 11
    with pytest.raises(_lib.SvgfError):
        f.Filter()
    f.params.atrous_iterations = 0
    f.Filter()                                   # N = 0 is legal (GUI range 0..10)
    ctx = C.c_void_p()
    assert _lib.lib().svgf_create(C.byref(ctx), 0, 0, 16, 0) == _lib.SVGF_INVALID_ARG
    assert _lib.lib().svgf_create(C.byref(ctx), 99, 16, 16, 0) == _lib.SVGF_CUDA_ERROR
    assert _lib.lib().svgf_create(C.byref(ctx), 0, 16, 16, 7) == _lib.SVGF_INVALID_ARG


@pytest.mark.parametrize("variant", ["stream", "bulk"])
@pytest.mark.parametrize("storage", ["f16", "f32"])
@pytest.mark.parametrize("level", [0, 1, 2, 3, 4])
def test_atrous_kernel_variants_meet_the_same_bar(variant, storage, level):
    # the A/B kernel families (svgf_params.flags) against the oracle, one level at a time like test_atrous_single_level
    flag = {"stream": _lib.SVGF_FLAG_ATROUS_STREAM, "bulk": _lib.SVGF_FLAG_ATROUS_BULK}[variant]
    fam = {"stream": _lib.SVGF_FAMILY_STREAM, "bulk": _lib.SVGF_FAMILY_BULK}[variant]
    for W, H in [(130, 67), (258, 129)]:
        of, f = make_pair(W, H, storage, seed=W * 13 + H + level)
        _atrous_inputs(of, f, np.random.default_rng(9 + level), smooth=False)
        P = of.PingPongInx
        g = of.gbuf(P)
        out = np.zeros_like(of.FilterBuffer[0])
        hc = of.RenderBuffer[P].copy()
        assert oracle().svgf_oracle_atrous_level(C.byref(of.params), W, H, of.storage, C.byref(g), of.FilterBuffer[0].ctypes.data,
                                                 out.ctypes.data, hc.ctypes.data, level) == 0
        f.params.flags = flag
        res = C.c_void_p()
        gs = f.Framebuffer[P].as_struct()
        st = f.lib.svgf_atrous(f._ctx, C.byref(f.params), C.byref(gs), C.c_void_p(f.FilterBuffer[0].data_ptr()),
                               C.c_void_p(f.FilterBuffer[1].data_ptr()), C.c_void_p(f.RenderBuffer[P].data_ptr()), level, 1,
                               C.byref(res), f._stream())
        assert st == 0 and f.last_dispatch()[level] == fam
        assert_close(npy(f.FilterBuffer[1]), out, storage, f"{variant} a-trous level {level} {W}x{H}")
        assert_close(npy(f.RenderBuffer[P]), hc, storage, f"{variant} colour history after level {level}")
