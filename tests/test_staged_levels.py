"""-m gpu: the staged a-trous run — first level by the packed kernel writing pre-transformed lattice planes, every later
level by the TMA-staged lattice kernel (svgf_kernels_lattice.cuh), the default for two or more consecutive levels.

Pinned against the oracle teacher-forced per level (the lattice level's input is the storage-format plane the GPU itself
produced one level earlier, so nothing but that level's own error is measured) at sizes that exercise every tile phase of
dilations 2..16, ragged right / bottom tiles and images wider than 3840; BIT-identical to the level-by-level path (same
arithmetic, and the lattice planes hold exactly what a storage-format round trip produces) - so every parity statement
made about the packed kernel transfers; and for the plumbing: dispatch report, uniform-tile shortcut, dependent launch."""
import ctypes as C

import numpy as np
import pytest
import torch

from common import assert_close, random_scene
from gpu_util import load_state_from_oracle, npy, upload_inputs
from oracle_lib import OracleFilter, oracle
from svgf_b200 import SvgfFilter, _lib, synth
from test_parity_stages import _atrous_inputs, make_pair
from test_uniform_tiles import planar_scene

pytestmark = pytest.mark.gpu


def _run_levels(f, first, n, flags=0):
    P = f.PingPongInx
    f.params.flags = flags
    res = C.c_void_p()
    gs = f.Framebuffer[P].as_struct()
    st = f.lib.svgf_atrous(f._ctx, C.byref(f.params), C.byref(gs), C.c_void_p(f.FilterBuffer[0].data_ptr()),
                           C.c_void_p(f.FilterBuffer[1].data_ptr()), C.c_void_p(f.RenderBuffer[P].data_ptr()), first, n,
                           C.byref(res), f._stream())
    assert st == 0, f"svgf_atrous -> {st} (cuda {f.lib.svgf_last_cuda_error(f._ctx)})"
    out = f.FilterBuffer[0] if res.value == f.FilterBuffer[0].data_ptr() else f.FilterBuffer[1]
    assert res.value == out.data_ptr()
    return out.clone()


def _oracle_level(of, src, level):
    P = of.PingPongInx
    g = of.gbuf(P)
    out = np.zeros_like(of.FilterBuffer[0])
    hc = of.RenderBuffer[P].copy()
    src = np.ascontiguousarray(src).astype(of.FilterBuffer[0].dtype)
    assert oracle().svgf_oracle_atrous_level(C.byref(of.params), of.Width, of.Height, of.storage, C.byref(g), src.ctypes.data,
                                             out.ctypes.data, hc.ctypes.data, level) == 0
    return out


# (W, H): ragged in both directions; 3850 is wider than the 4K benchmark frame (31 tile columns, the last one 10 pixels);
# 76 rows leave every dilation-16 row phase with a ragged last tile
SIZES = [(130, 67), (258, 129), (1030, 210), (3850, 76)]


@pytest.mark.parametrize("storage", ["f16", "f32"])
@pytest.mark.parametrize("size", SIZES)
@pytest.mark.parametrize("level", [1, 2, 3, 4])
@pytest.mark.parametrize("scene", ["random", "planar"])
def test_lattice_level_against_the_oracle_teacher_forced(size, storage, level, scene):
    W, H = size
    if scene == "random":
        of, f = make_pair(W, H, storage, seed=W * 7 + H + level)
        _atrous_inputs(of, f, np.random.default_rng(31 + level), smooth=False)
    else:   # large planar regions: most tiles take the uniform-normal shortcut, background rows are wildcards
        of = OracleFilter(W, H, storage=storage)
        planes = planar_scene(np.random.default_rng(200 + level), W, H, storage)
        of.set_inputs(planes)
        of.FilterBuffer[0][...] = planes["colour"]
        f = SvgfFilter(W, H, storage=storage)
        load_state_from_oracle(f, of)
    start = f.FilterBuffer[0].clone()
    # level - 1 alone (packed kernel, storage format out): the teacher for level `level`
    mid = _run_levels(f, level - 1, 1, _lib.SVGF_FLAG_NO_STAGED_LEVELS)
    # the packed kernel itself at this size (ragged tiles, > 3840 wide) against the oracle, same inputs
    assert_close(npy(mid), _oracle_level(of, npy(start), level - 1), storage, f"packed level {level - 1} {size} {scene}")
    want = _oracle_level(of, npy(mid), level)
    # the same two levels as one staged run
    f.FilterBuffer[0].copy_(start)
    got = _run_levels(f, level - 1, 2)
    assert f.last_dispatch()[level - 1:level + 1] == [_lib.SVGF_FAMILY_PACKED_STAGED, _lib.SVGF_FAMILY_LATTICE]
    assert_close(npy(got), want, storage, f"staged level {level} {size} {scene}")
    # and without the uniform-tile shortcut (general form of every tap)
    f.FilterBuffer[0].copy_(start)
    got_general = _run_levels(f, level - 1, 2, _lib.SVGF_FLAG_NO_UNIFORM_TILES)
    assert torch.equal(got_general.view(torch.uint8), got.view(torch.uint8)), "the uniform-tile shortcut changed output bits"
    # and the staged run is the level-by-level run, bit for bit
    f.FilterBuffer[0].copy_(start)
    ref2 = _run_levels(f, level - 1, 2, _lib.SVGF_FLAG_NO_STAGED_LEVELS)
    assert torch.equal(ref2.view(torch.uint8), got.view(torch.uint8)), "staged run differs from two single-level launches"


@pytest.mark.parametrize("storage", ["f16", "f32"])
@pytest.mark.parametrize("size", [(258, 129), (1030, 210)])
@pytest.mark.parametrize("first,n", [(0, 2), (0, 3), (0, 5), (1, 4), (2, 3), (3, 2)])
def test_staged_run_matches_the_level_by_level_path(size, storage, first, n):
    W, H = size
    of, f = make_pair(W, H, storage, seed=W + 3 * H + first)
    _atrous_inputs(of, f, np.random.default_rng(77), smooth=True)
    P = f.PingPongInx
    start, hist0 = f.FilterBuffer[0].clone(), f.RenderBuffer[P].clone()
    ref = _run_levels(f, first, n, _lib.SVGF_FLAG_NO_STAGED_LEVELS)
    ref_hist = f.RenderBuffer[P].clone()
    f.FilterBuffer[0].copy_(start); f.RenderBuffer[P].copy_(hist0)
    got = _run_levels(f, first, n)
    fam = f.last_dispatch()
    assert fam[first] == _lib.SVGF_FAMILY_PACKED_STAGED and all(x == _lib.SVGF_FAMILY_LATTICE for x in fam[first + 1:first + n])
    assert torch.equal(got.view(torch.uint8), ref.view(torch.uint8)), f"staged {first}+{n} vs level by level {size}"
    assert torch.equal(f.RenderBuffer[P].view(torch.uint8), ref_hist.view(torch.uint8))


@pytest.mark.parametrize("storage", ["f16", "f32"])
def test_frame_sequence_staged_and_level_by_level_agree(storage):
    W, H, N = 1280, 720, 6
    a, b = SvgfFilter(W, H, storage=storage), SvgfFilter(W, H, storage=storage)
    b.params.flags = _lib.SVGF_FLAG_NO_STAGED_LEVELS
    a.Reset(); b.Reset()
    for t in range(N):       # free-running: both filters feed on their own outputs
        planes = synth.frame_host(W, H, t, storage=storage)
        upload_inputs(a, planes); upload_inputs(b, planes)
        a.Filter(); b.Filter()
        assert a.last_dispatch() == [_lib.SVGF_FAMILY_PACKED_STAGED] + [_lib.SVGF_FAMILY_LATTICE] * 4
        assert b.last_dispatch() == [_lib.SVGF_FAMILY_PACKED] * 5
        P = a.PingPongInx
        assert torch.equal(a.HistoryLengthBuffer, b.HistoryLengthBuffer)
        assert torch.equal(a.MomentsBuffer[P].view(torch.uint8), b.MomentsBuffer[P].view(torch.uint8))
        assert torch.equal(a.RenderBuffer[P].view(torch.uint8), b.RenderBuffer[P].view(torch.uint8)), "colour history (level 0) differs"
        assert torch.equal(a.FilterBuffer[0].view(torch.uint8), b.FilterBuffer[0].view(torch.uint8)), f"frame {t}: result differs"
        a.EndFrame(); b.EndFrame()


def test_dependent_launch_flag_changes_no_bit():
    W, H = 1030, 210
    of, f = make_pair(W, H, "f16", seed=5)
    _atrous_inputs(of, f, np.random.default_rng(3), smooth=False)
    start = f.FilterBuffer[0].clone()
    a = _run_levels(f, 0, 5)
    f.FilterBuffer[0].copy_(start)
    b = _run_levels(f, 0, 5, _lib.SVGF_FLAG_NO_DEPENDENT_LAUNCH)
    assert torch.equal(a.view(torch.uint8), b.view(torch.uint8))


def test_dispatch_report_makes_the_fallbacks_visible():
    W, H = 256, 96
    f = SvgfFilter(W, H)
    f.Reset()
    upload_inputs(f, synth.frame_host(W, H, 0))
    f.SpatialFilterSteps = 7                     # levels 5 and 6 have no tiled kernel
    f.Filter()
    assert f.last_dispatch() == [_lib.SVGF_FAMILY_PACKED] * 5 + [_lib.SVGF_FAMILY_BASIC] * 2
    f.SpatialFilterSteps = 5
    f.PhiNormal = 16.0                           # below the series range of the tiled kernels
    f.EndFrame(); upload_inputs(f, synth.frame_host(W, H, 1)); f.Filter()
    assert f.last_dispatch() == [_lib.SVGF_FAMILY_BASIC] * 5
    f.PhiNormal = 128.0
    f.EndFrame(); upload_inputs(f, synth.frame_host(W, H, 2)); f.Filter()
    assert f.last_dispatch() == [_lib.SVGF_FAMILY_PACKED_STAGED] + [_lib.SVGF_FAMILY_LATTICE] * 4
    f.PhiNormal = 64.0                           # Taylor series instantiation
    f.EndFrame(); upload_inputs(f, synth.frame_host(W, H, 3)); f.Filter()
    assert f.last_dispatch() == [_lib.SVGF_FAMILY_PACKED_STAGED] + [_lib.SVGF_FAMILY_LATTICE] * 4
    g = SvgfFilter(255, 96)                      # odd width, 8-byte texels: neither pixel pairs nor 16-byte row segments
    g.Reset()
    upload_inputs(g, synth.frame_host(255, 96, 0))
    g.Filter()
    assert g.last_dispatch() == [_lib.SVGF_FAMILY_BASIC] * 5
    h = SvgfFilter(255, 96, storage="f32")       # odd width, 16-byte texels: the bulk-copy kernel
    h.Reset()
    upload_inputs(h, synth.frame_host(255, 96, 0, storage="f32"))
    h.Filter()
    assert h.last_dispatch() == [_lib.SVGF_FAMILY_BULK] * 5


@pytest.mark.parametrize("storage", ["f16", "f32"])
def test_taylor_series_instantiation_of_the_staged_run(storage):
    # 32 <= phi_normal < 100 runs the five-term Taylor form of both kernels
    W, H = 258, 129
    of, f = make_pair(W, H, storage, seed=11)
    of.params.phi_normal = f.params.phi_normal = 48.0
    _atrous_inputs(of, f, np.random.default_rng(5), smooth=False)
    start = f.FilterBuffer[0].clone()
    for level in (1, 4):
        f.FilterBuffer[0].copy_(start)
        mid = _run_levels(f, level - 1, 1, _lib.SVGF_FLAG_NO_STAGED_LEVELS)
        want = _oracle_level(of, npy(mid), level)
        f.FilterBuffer[0].copy_(start)
        got = _run_levels(f, level - 1, 2)
        assert_close(npy(got), want, storage, f"staged level {level}, phi_normal 48")
