"""-m gpu: SVGF_REPROJ_BILINEAR - the SVGF paper's 2x2 bilinear history fetch with per-texel consistency tests
(include/svgf.h; the reference fetches one texel with C truncation, src/Filter.cuh:231-232).

History lengths and moments must be BIT-exact against the scalar oracle with the same switch (the kernel runs the
oracle's FP32 operations un-contracted in the same order, so even the rounded mean history length agrees); colour at the
per-stage bar.  Checked on random scenes with fractional motion, through svgf_temporal and through svgf_frame (with and
without the cached previous-frame guide plane), and on the camera-pan sequence."""
import numpy as np
import pytest
import torch

from common import assert_close, random_scene
from gpu_util import load_state_from_oracle, npy
from oracle_lib import OracleFilter
from svgf_b200 import SvgfFilter, _lib
from test_parity_sequence import run_sequence

pytestmark = pytest.mark.gpu
BILINEAR = 1


def _scene_pair(rng, W, H, storage):
    cur = random_scene(rng, W, H, storage=storage)
    prev = random_scene(rng, W, H, storage=storage)
    keep = rng.uniform(size=(H, W)) < 0.75
    for k in ("normal", "uv"):
        prev[k][keep] = cur[k][keep]
    prev["motion"][keep, 2:] = cur["motion"][keep, 2:]
    cur["motion"][..., :2] = rng.uniform(-4, 4, size=(H, W, 2)).astype(np.float32)
    cur["motion"][rng.uniform(size=(H, W)) < 0.1, :2] = np.float32(1.5)        # exact .5 fractions: equal weights, rounding ties
    cur["motion"][rng.uniform(size=(H, W)) < 0.02, 0] = np.float32(np.nan)     # NaN motion: the reprojection fails
    cur["motion"][rng.uniform(size=(H, W)) < 0.02, 1] = np.float32(3e9)        # far outside
    return cur, prev


@pytest.mark.parametrize("storage", ["f16", "f32"])
@pytest.mark.parametrize("size", [(333, 97), (640, 360)])
def test_temporal_stage_against_the_oracle(storage, size):
    W, H = size
    rng = np.random.default_rng(40 + W)
    cur, prev = _scene_pair(rng, W, H, storage)
    cdt = np.float16 if storage == "f16" else np.float32
    o = OracleFilter(W, H, storage=storage)
    o.params.reproj_mode = BILINEAR
    o.params.history_cap = 31
    o.PingPongInx = 0
    o.set_inputs(cur)
    o.normal[1][...] = prev["normal"]; o.uv[1][...] = prev["uv"]; o.motion[1][...] = prev["motion"]
    o.RenderBuffer[1][...] = rng.uniform(0, 1.2, size=(H, W, 4)).astype(cdt)
    o.MomentsBuffer[1][...] = rng.uniform(0, 1, size=(H, W, 2)).astype(cdt)
    o.HistoryLengthBuffer[...] = rng.integers(0, 40, size=(H, W)).astype(np.uint8)
    f = SvgfFilter(W, H, storage=storage)
    f.params.reproj_mode = BILINEAR
    f.params.history_cap = 31
    load_state_from_oracle(f, o)
    f.TemporalFilter(); o.TemporalFilter()
    torch.cuda.synchronize()
    assert np.array_equal(npy(f.HistoryLengthBuffer), o.HistoryLengthBuffer), "history lengths differ"
    assert (o.HistoryLengthBuffer > 1).mean() > 0.3
    assert np.array_equal(npy(f.MomentsBuffer[0]).view(np.uint8), o.MomentsBuffer[0].view(np.uint8)), "moments differ"
    assert np.array_equal(npy(f.RenderBuffer[0]).view(np.uint8), o.RenderBuffer[0].view(np.uint8)), "accumulated colour / variance differ"


@pytest.mark.parametrize("storage", ["f16", "f32"])
def test_pan_sequence_teacher_forced(storage):
    run_sequence(640, 360, 8, storage, teacher_forced=True, reproj_mode=BILINEAR)


@pytest.mark.parametrize("flags", [0, _lib.SVGF_FLAG_NO_GUIDE_CACHE])
def test_pan_sequence_free_running_history_and_moments_stay_exact(flags):
    """Free-running: history and moments never pass through an ill-conditioned weight, so they stay bit-exact over a
    sequence (run_sequence asserts both every frame); with and without the cached previous-frame guide plane."""
    run_sequence(480, 270, 10, "f32", teacher_forced=False, reproj_mode=BILINEAR, flags=flags)
