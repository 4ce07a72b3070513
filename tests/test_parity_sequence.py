"""-m gpu: whole-frame parity over camera-pan sequences (BASELINE configs 1 and 2) and size-independent
properties at the full benchmark resolution."""
import numpy as np
import pytest
import torch

from common import (F16_ABS_FLOOR, F16_MAX_ULPS, F32_FLOOR_RADIANCE, F32_FLOOR_VARIANCE, F32_REL_TOL, half_ulp_diff)
from gpu_util import load_state_from_oracle, npy, upload_inputs
from oracle_lib import OracleFilter
from svgf_b200 import SvgfFilter, synth

pytestmark = pytest.mark.gpu


# How parity over a SEQUENCE is judged (see DESIGN.md "Parity"):
#  * integer state (history lengths) and the moments plane are bit-exact over the free-running sequence;
#  * teacher-forced: every frame starts from the oracle's state; the frame's outputs (5 levels deep) must meet
#    the per-stage tolerance of tests/common.py — the kernels' own error — except for at most TF_OUTLIERS of the
#    values (ten per million), which sit on zero-variance pixels where one level's 1e-7 difference is already
#    amplified by the next levels of the SAME frame; those stay below TF_F32_MAX / TF_F16_MAX_ULPS.
#  * free-running: the CUDA path feeds on its own outputs for all frames.  The reference math is
#    ill-conditioned wherever the accumulated variance is ~0 (e.g. a pixel whose 1-spp samples were all black:
#    phi_l = PhiColour*sqrt(1e-10) = 1e-4, so a 1e-7 difference in a neighbour's luminance moves a weight by
#    1e-3, and one fp16 ulp moves it by a factor of e) — there, ANY two implementations drift apart, including
#    the reference's own kernels versus the oracle (tests/test_reference_kernels.py measures that yardstick).
#    The bar is therefore on the distribution: at most FREE_F32_OUTLIERS of the values above 1e-4 relative and
#    none above FREE_F32_MAX; fp16: at most FREE_F16_FLIPS differing at all, FREE_F16_OUTLIERS by more than 2 ulps.
TF_OUTLIERS, TF_F32_MAX, TF_F16_MAX_ULPS = 1e-5, 5e-3, 128
FREE_F32_OUTLIERS, FREE_F32_MAX = 1e-3, 2e-2
FREE_F16_FLIPS, FREE_F16_OUTLIERS = 0.03, 3e-3


def _frame_stats(f, o, storage):
    P = o.PingPongInx
    out = {}
    for name, got, want in (("result", f.FilterBuffer[0], o.FilterBuffer[0]), ("colour history", f.RenderBuffer[P], o.RenderBuffer[P])):
        g, w = npy(got), want
        if storage == "f32":
            d = np.abs(g.astype(np.float64) - w.astype(np.float64))
            floor = np.array([F32_FLOOR_RADIANCE] * 3 + [F32_FLOOR_VARIANCE])
            r = d / np.maximum(np.abs(w.astype(np.float64)), floor)
            out[name] = {"max": float(r.max()), "outliers": float((r > F32_REL_TOL).mean())}
        else:
            u = half_ulp_diff(g, w)
            absd = np.abs(g.astype(np.float64) - w.astype(np.float64))
            out[name] = {"max_ulps": int(u.max()), "flips": float((u > 0).mean()),
                         "outliers": float(((u > F16_MAX_ULPS) & (absd > F16_ABS_FLOOR)).mean()), "max_abs": float(absd.max())}
    return out


def run_sequence(W, H, frames, storage, check_every=1, seed=0, steps=5, teacher_forced=False, variance_prefilter=0, reproj_mode=0, flags=0):
    f = SvgfFilter(W, H, storage=storage)
    o = OracleFilter(W, H, storage=storage)
    f.params.variance_prefilter = o.params.variance_prefilter = variance_prefilter
    f.params.reproj_mode = o.params.reproj_mode = reproj_mode
    f.params.flags = flags
    f.SpatialFilterSteps = steps
    o.params.atrous_iterations = steps
    f.Reset(); o.Reset()
    worst = {}
    for t in range(frames):
        planes = synth.frame_host(W, H, t, seed=seed, storage=storage)
        o.set_inputs(planes)
        if teacher_forced and t > 0:
            load_state_from_oracle(f, o)
        else:
            upload_inputs(f, planes)
        f.Filter(); o.Filter()
        if t % check_every == 0 or t == frames - 1:
            P = o.PingPongInx
            assert np.array_equal(npy(f.HistoryLengthBuffer), o.HistoryLengthBuffer), f"frame {t}: history lengths differ"
            assert np.array_equal(npy(f.MomentsBuffer[P]).view(np.uint8), o.MomentsBuffer[P].view(np.uint8)), f"frame {t}: moments differ"
            st = _frame_stats(f, o, storage)
            for name, e in st.items():
                if teacher_forced:
                    if storage == "f32":
                        assert e["outliers"] <= TF_OUTLIERS and e["max"] <= TF_F32_MAX, f"frame {t} {name}: {e}"
                    else:
                        assert e["outliers"] <= TF_OUTLIERS and e["max_ulps"] <= TF_F16_MAX_ULPS and e["flips"] <= 0.02, f"frame {t} {name}: {e}"
                else:
                    if storage == "f32":
                        assert e["outliers"] <= FREE_F32_OUTLIERS and e["max"] <= FREE_F32_MAX, f"frame {t} {name}: {e}"
                    else:
                        assert e["flips"] <= FREE_F16_FLIPS and e["outliers"] <= FREE_F16_OUTLIERS, f"frame {t} {name}: {e}"
                for k, v in e.items():
                    worst[f"{name}.{k}"] = max(worst.get(f"{name}.{k}", 0), v)
        f.EndFrame(); o.EndFrame()
    print(f"\n[{W}x{H} x{frames} {storage} {'teacher-forced' if teacher_forced else 'free-running'}] worst vs oracle: {worst}")
    return worst


@pytest.mark.parametrize("storage", ["f16", "f32"])
def test_config1_720p_four_frames(storage):
    # BASELINE config 1: 1280x720, reset + 3 more frames (covers the h<4 path and the first h>=4 frame)
    run_sequence(1280, 720, 4, storage, teacher_forced=True)


@pytest.mark.parametrize("storage", ["f16", "f32"])
def test_pan_sequence_64_frames_teacher_forced(storage):
    run_sequence(480, 270, 64, storage, teacher_forced=True)


@pytest.mark.parametrize("storage", ["f16", "f32"])
def test_pan_sequence_64_frames_free_running(storage):
    run_sequence(480, 270, 64, storage)


def test_config2_1080p_64_frames_fp32():
    # BASELINE config 2, free-running over the whole 64-frame pan: history and moments bit-exact on every
    # checked frame, radiance/variance distribution bar; checked every 4th frame to bound the oracle's CPU time
    # (the oracle still runs every frame).
    run_sequence(1920, 1080, 64, "f32", check_every=4)


def test_config2_1080p_fp16_reference_layout():
    run_sequence(1920, 1080, 8, "f16", check_every=1)


@pytest.mark.parametrize("steps", [0, 1, 3, 4])
def test_other_level_counts(steps):
    run_sequence(320, 200, 5, "f32", steps=steps, teacher_forced=True)


# ---- size-independent properties at the benchmark resolution (no oracle needed) ----------------------------
W4K, H4K = 3840, 2160


def _run_device_sequence(f, frames, seed=0):
    f.Reset()
    for t in range(frames):
        P = f.PingPongInx
        synth.frame_device(f.Framebuffer[P], f.RenderBuffer[P], t, seed=seed)
        f.Filter()
        f.EndFrame()
    torch.cuda.synchronize()


def test_4k_is_deterministic_and_history_saturates():
    f1, f2 = SvgfFilter(W4K, H4K), SvgfFilter(W4K, H4K)
    f1.HistoryLength = f2.HistoryLength = 6
    _run_device_sequence(f1, 8)
    _run_device_sequence(f2, 8)
    assert torch.equal(f1.FilterBuffer[0], f2.FilterBuffer[0])
    assert torch.equal(f1.HistoryLengthBuffer, f2.HistoryLengthBuffer)
    h = f1.HistoryLengthBuffer
    assert int(h.max()) == 6 and int(h.min()) == 1
    assert float((h == 6).float().mean()) > 0.6
    out = f1.FilterBuffer[0].float()
    assert torch.isfinite(out).all() and float(out[..., :3].min()) >= 0 and float(out[..., :3].max()) <= 1.0


def test_4k_flat_field_is_a_fixed_point():
    # constant radiance over one plane with a static camera: every stage must return the constant
    f = SvgfFilter(W4K, H4K)
    f.Reset()
    for t in range(3):
        P = f.PingPongInx
        g = f.Framebuffer[P]
        g.normal.zero_(); g.normal[..., 2] = 0x3C00          # (0, 0, 1) in fp16 bits
        g.uv.zero_(); g.uv[..., 3] = 0x4000                  # instance 2
        g.motion.zero_(); g.motion[..., 2] = 7.0; g.motion[..., 3] = 0.01
        f.RenderBuffer[P][..., 0] = 0.25; f.RenderBuffer[P][..., 1] = 0.5; f.RenderBuffer[P][..., 2] = 0.75
        f.RenderBuffer[P][..., 3] = 1.0
        f.Filter()
        out = f.FilterBuffer[0].float()
        want = torch.tensor([0.25, 0.5, 0.75], device=out.device)
        assert float((out[..., :3] - want).abs().max()) <= 2.5e-4      # half an fp16 ulp at 0.75
        assert int(f.HistoryLengthBuffer.min()) == t + 1 == int(f.HistoryLengthBuffer.max())
        f.EndFrame()


def test_4k_all_background_outputs_zero():
    f = SvgfFilter(W4K, H4K)
    f.Reset()
    f.RenderBuffer[0][...] = 0.5
    f.Filter()
    assert int(f.FilterBuffer[0].float().abs().max()) == 0          # D7
    assert int(f.HistoryLengthBuffer.max()) == 1
