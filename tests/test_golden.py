"""CPU-only: pins the scalar oracle to golden vectors produced by the REFERENCE'S OWN kernels.

tests/golden/reference_kernels_*.npz were generated on a B200 by tests/golden/make_golden.py from
oracle/_ref/libsvgf_refkernels.so (= /root/reference/src/Filter.cuh compiled by oracle/Makefile).  The
reference itself has no tests or fixtures for this path; these files are the pins that travel.
Tolerance: the oracle uses glibc's exp/pow in FP64/FP32 where the reference used CUDA's, and gcc does not contract
FMAs, so a stored fp16 may differ by an ulp: every value within 2 fp16 ulps (or 1e-4 absolute), history bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest

from common import f16_errors
from oracle_lib import OracleFilter, oracle
from svgf_b200 import _lib

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    p = os.path.join(GOLDEN, name)
    if not os.path.exists(p):
        pytest.skip(f"{name} not generated yet (tests/golden/make_golden.py needs a B200)")
    return np.load(p)


def check(got, want, what, max_flips=0.02):
    e = f16_errors(got, want)
    assert e["violations"] == 0 and e["flip_fraction"] <= max_flips, f"{what}: {e}"


def test_oracle_matches_reference_kernels_static_sequence():
    g = _load("reference_kernels_static_sequence.npz")
    W, H, frames = int(g["W"]), int(g["H"]), int(g["frames"])
    o = OracleFilter(W, H, storage="f16")
    o.params.mesh_id_mode = _lib.SVGF_MESH_ID_REFERENCE_VACUOUS
    o.params.atrous_iterations = 5
    o.Reset()
    for t in range(frames):
        o.set_inputs({k: g[f"f{t}_in_{k}"] for k in ("normal", "uv", "motion", "colour")})
        P = o.PingPongInx
        o.Filter()
        assert np.array_equal(g[f"f{t}_history"], o.HistoryLengthBuffer), f"frame {t}: history"
        check(o.MomentsBuffer[P], g[f"f{t}_moments"], f"frame {t} moments")
        check(o.FilterBuffer[0], g[f"f{t}_result"], f"frame {t} result", max_flips=0.05)
        check(o.RenderBuffer[P], g[f"f{t}_colour_history"], f"frame {t} colour history", max_flips=0.05)
        o.EndFrame()


def test_generated_inputs_are_reproducible():
    # the committed inputs are exactly what the in-tree generator produces (so the fixtures can be regenerated)
    from svgf_b200 import synth
    g = _load("reference_kernels_static_sequence.npz")
    W, H = int(g["W"]), int(g["H"])
    for t in (0, int(g["frames"]) - 1):
        planes = synth.frame_host(W, H, t, pan_px=0.0, vert_px=0.0)
        for k in ("normal", "uv", "motion", "colour"):
            assert np.array_equal(planes[k].view(np.uint8), g[f"f{t}_in_{k}"].view(np.uint8)), (t, k)


@pytest.mark.parametrize("seed", [0, 1])
def test_oracle_matches_reference_kernels_stage_by_stage(seed):
    g = _load(f"reference_kernels_stages_seed{seed}.npz")
    W, H = int(g["W"]), int(g["H"])
    o = OracleFilter(W, H, storage="f16")
    o.params.mesh_id_mode = _lib.SVGF_MESH_ID_REFERENCE_VACUOUS
    o.PingPongInx = 1
    for k in range(2):
        o.normal[k][...] = g[f"in_normal{k}"]; o.uv[k][...] = g[f"in_uv{k}"]; o.motion[k][...] = g[f"in_motion{k}"]
        o.RenderBuffer[k][...] = g[f"in_render{k}"]; o.MomentsBuffer[k][...] = g[f"in_moments{k}"]
    o.HistoryLengthBuffer[...] = g["in_history"]
    # temporal (src/Filter.cuh:359-404)
    o.TemporalFilter()
    assert np.array_equal(o.HistoryLengthBuffer, g["temporal_history"])
    check(o.RenderBuffer[1], g["temporal_colour"], "temporal colour")
    check(o.MomentsBuffer[1], g["temporal_moments"], "temporal moments")
    # variance (src/Filter.cuh:430-525), intended and MomentsBuffer[0] (src/App.cu:484) plane choices
    assert np.array_equal(o.RenderBuffer[1].view(np.uint16), g["variance_in_colour"].view(np.uint16))
    o.HistoryLengthBuffer[...] = g["variance_in_history"]
    o.FilterMoments()
    check(o.FilterBuffer[0], g["variance_out_quirk0"], "variance")
    o.FilterMoments(moments_index=0)
    check(o.FilterBuffer[0], g["variance_out_quirk1"], "variance, MomentsBuffer[0] quirk")
    # a-trous levels (src/Filter.cuh:527-624)
    gb = o.gbuf(1)
    for level in range(5):
        src = np.ascontiguousarray(g[f"atrous{level}_in"])
        out = np.zeros_like(src)
        hc = np.ascontiguousarray(g["atrous0_render_in"]).copy() if level == 0 else np.zeros_like(src)
        assert oracle().svgf_oracle_atrous_level(C.byref(o.params), W, H, 0, C.byref(gb), src.ctypes.data, out.ctypes.data,
                                                 hc.ctypes.data, level) == 0
        check(out, g[f"atrous{level}_out"], f"a-trous level {level}", max_flips=0.05)
        if level == 0:
            check(hc, g["atrous0_colour_history"], "level-0 colour history", max_flips=0.05)
