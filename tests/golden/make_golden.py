#!/usr/bin/env python
"""Generates tests/golden/reference_kernels_*.npz: outputs of the REFERENCE'S OWN kernels
(filter::TemporalFilter / FilterMoments / FilterKernel, compiled unmodified-in-math from
/root/reference/src/Filter.cuh into oracle/_ref/libsvgf_refkernels.so by oracle/Makefile) on small
deterministic inputs.  The reference ships no golden vectors of its own, so these are the pins
the CPU-only test suite (tests/test_golden.py, -m "not gpu") holds the scalar oracle to.

Must run where a B200 is visible (the kernels are sm_100a):
    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'      then copy the .npz files here.

Cases (all in the reference's fp16 layout, mesh-id test vacuous as on hardware, D2):
  static_sequence : 6 frames of the procedural scene with a static camera (zero motion => the reference's in-place
                    history race, D3, cannot occur and every buffer is well defined).  The reference state is
                    re-synchronised to the oracle's after every frame so that 1-ulp differences do not compound;
                    stored per frame: inputs, and the reference's history / moments / result / colour history.
  stages          : random mid-sequence state with DIFFERENT current and previous G-buffers (motion zeroed):
                    temporal, variance (both moments-plane choices, D4) and each a-trous level 0..4 on its own.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import random_scene  # noqa: E402
from oracle_lib import (PLANE_FILTER, PLANE_HISTORY, PLANE_MOMENTS, PLANE_RENDER, OracleFilter, RefKernels, RefParams,  # noqa: E402
                        oracle, ref)
from svgf_b200 import _lib, synth  # noqa: E402


def static_sequence(W=96, H=64, frames=6):
    o = OracleFilter(W, H, storage="f16")
    o.params.mesh_id_mode = _lib.SVGF_MESH_ID_REFERENCE_VACUOUS
    o.params.atrous_iterations = 5
    rp = RefParams.from_svgf(o.params, moments_quirk=0)
    r = RefKernels(W, H)
    o.Reset()
    out = {"W": W, "H": H, "frames": frames}
    for t in range(frames):
        planes = synth.frame_host(W, H, t, pan_px=0.0, vert_px=0.0)
        o.set_inputs(planes)
        P = o.PingPongInx
        r.set_gbuffer(P, planes["normal"], planes["uv"], planes["motion"])
        r.set_plane(PLANE_RENDER, P, planes["colour"])
        o.Filter()
        assert ref().svgf_ref_frame(r.ctx, rp, 0) == 0
        for k in ("normal", "uv", "motion", "colour"):
            out[f"f{t}_in_{k}"] = planes[k]
        out[f"f{t}_history"] = r.get_plane(PLANE_HISTORY, 0)
        out[f"f{t}_moments"] = r.get_plane(PLANE_MOMENTS, P)
        out[f"f{t}_result"] = r.get_plane(PLANE_FILTER, 0)
        out[f"f{t}_colour_history"] = r.get_plane(PLANE_RENDER, P)
        o.EndFrame()
        ref().svgf_ref_set_ping_pong(r.ctx, o.PingPongInx)
        r.load_state(o)      # the NEXT frame starts from the oracle's state (recorded implicitly: the oracle is deterministic)
    r.close()
    return out


def stages(seed=0, W=72, H=48):
    rng = np.random.default_rng(seed)
    o = OracleFilter(W, H, storage="f16")
    o.params.mesh_id_mode = _lib.SVGF_MESH_ID_REFERENCE_VACUOUS
    cur, prev = random_scene(rng, W, H), random_scene(rng, W, H)
    keep = rng.uniform(size=(H, W)) < 0.7
    for k in ("normal", "uv"):
        prev[k][keep] = cur[k][keep]
    prev["motion"][keep, 2:] = cur["motion"][keep, 2:]
    prev["motion"][..., 2] += (rng.uniform(-1.6, 1.6, size=(H, W)) * (prev["motion"][..., 2] > 0)).astype(np.float32)
    cur["motion"][..., :2] = 0
    o.PingPongInx = 1
    o.set_inputs(cur)
    o.normal[0][...] = prev["normal"]; o.uv[0][...] = prev["uv"]; o.motion[0][...] = prev["motion"]
    o.RenderBuffer[0][...] = rng.uniform(0, 1.2, size=(H, W, 4)).astype(np.float16)
    o.MomentsBuffer[0][...] = rng.uniform(0, 1, size=(H, W, 2)).astype(np.float16)
    o.HistoryLengthBuffer[...] = rng.integers(0, 30, size=(H, W)).astype(np.uint8)
    out = {"W": W, "H": H, "seed": seed}
    for k in range(2):
        out[f"in_normal{k}"], out[f"in_uv{k}"], out[f"in_motion{k}"] = o.normal[k].copy(), o.uv[k].copy(), o.motion[k].copy()
        out[f"in_render{k}"], out[f"in_moments{k}"] = o.RenderBuffer[k].copy(), o.MomentsBuffer[k].copy()
    out["in_history"] = o.HistoryLengthBuffer.copy()
    rp = RefParams.from_svgf(o.params, moments_quirk=0)
    r = RefKernels(W, H)
    r.load_state(o)
    # temporal
    assert ref().svgf_ref_temporal(r.ctx, rp) == 0
    out["temporal_history"] = r.get_plane(PLANE_HISTORY, 0)
    out["temporal_colour"] = r.get_plane(PLANE_RENDER, 1)
    out["temporal_moments"] = r.get_plane(PLANE_MOMENTS, 1)
    # variance: starts from the ORACLE's temporal output with a fresh short history (both moments-plane choices)
    o.TemporalFilter()
    o.HistoryLengthBuffer[...] = rng.integers(1, 8, size=(H, W)).astype(np.uint8)
    out["variance_in_history"] = o.HistoryLengthBuffer.copy()
    out["variance_in_colour"] = o.RenderBuffer[1].copy()
    out["variance_in_moments1"] = o.MomentsBuffer[1].copy()
    for quirk in (0, 1):
        r.load_state(o)
        assert ref().svgf_ref_variance(r.ctx, RefParams.from_svgf(o.params, moments_quirk=quirk)) == 0
        out[f"variance_out_quirk{quirk}"] = r.get_plane(PLANE_FILTER, 0)
    # a-trous levels, each from the oracle's previous level
    o.FilterMoments()
    g = o.gbuf(1)
    for level in range(5):
        out[f"atrous{level}_in"] = o.FilterBuffer[0].copy()
        r.load_state(o)
        assert ref().svgf_ref_atrous_level(r.ctx, rp, level) == 0
        out[f"atrous{level}_out"] = r.get_plane(PLANE_FILTER, 1)
        if level == 0:
            out["atrous0_render_in"] = o.RenderBuffer[1].copy()
            out["atrous0_colour_history"] = r.get_plane(PLANE_RENDER, 1)
        nxt = np.zeros_like(o.FilterBuffer[0])
        hc = o.RenderBuffer[1].copy()
        assert oracle().svgf_oracle_atrous_level(C.byref(o.params), W, H, 0, C.byref(g), o.FilterBuffer[0].ctypes.data,
                                                 nxt.ctypes.data, hc.ctypes.data, level) == 0
        o.FilterBuffer[0][...] = nxt
    r.close()
    return out


if __name__ == "__main__":
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.dirname(os.path.abspath(__file__))
    os.makedirs(dst, exist_ok=True)
    np.savez_compressed(os.path.join(dst, "reference_kernels_static_sequence.npz"), **static_sequence())
    for seed in (0, 1):
        np.savez_compressed(os.path.join(dst, f"reference_kernels_stages_seed{seed}.npz"), **stages(seed))
    print("golden vectors written to", dst)
