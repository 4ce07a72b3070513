"""-m gpu: whole-frame parity over camera-pan sequences (BASELINE configs 1 and 2) and size-independent
properties at the full benchmark resolution."""
import numpy as np
import pytest
import torch

from common import assert_close, f16_errors, f32_errors
from gpu_util import npy, upload_inputs
from oracle_lib import OracleFilter
from svgf_b200 import SvgfFilter, synth

pytestmark = pytest.mark.gpu


def run_sequence(W, H, frames, storage, check_every=1, seed=0, steps=5):
    f = SvgfFilter(W, H, storage=storage)
    o = OracleFilter(W, H, storage=storage)
    f.SpatialFilterSteps = steps
    o.params.atrous_iterations = steps
    f.Reset(); o.Reset()
    worst = {}
    for t in range(frames):
        planes = synth.frame_host(W, H, t, seed=seed, storage=storage)
        o.set_inputs(planes)
        upload_inputs(f, planes)
        f.Filter(); o.Filter()
        if t % check_every == 0 or t == frames - 1:
            P = o.PingPongInx
            assert np.array_equal(npy(f.HistoryLengthBuffer), o.HistoryLengthBuffer), f"frame {t}: history lengths differ"
            for name, got, want in (("result", f.FilterBuffer[0], o.FilterBuffer[0]),
                                    ("colour history", f.RenderBuffer[P], o.RenderBuffer[P]),
                                    ("moments", f.MomentsBuffer[P], o.MomentsBuffer[P])):
                e = assert_close(npy(got), want, storage, f"frame {t} {name}", **({"max_flips": 0.05} if storage == "f16" else {}))
                for k, v in e.items():
                    worst[f"{name}.{k}"] = max(worst.get(f"{name}.{k}", 0), v)
        f.EndFrame(); o.EndFrame()
    print(f"\n[{W}x{H} x{frames} {storage}] worst errors vs oracle: {worst}")
    return worst


@pytest.mark.parametrize("storage", ["f16", "f32"])
def test_config1_720p_four_frames(storage):
    # BASELINE config 1: 1280x720, reset + 3 more frames (covers the h<4 path and the first h>=4 frame)
    run_sequence(1280, 720, 4, storage)


@pytest.mark.parametrize("storage", ["f16", "f32"])
def test_pan_sequence_64_frames_small(storage):
    run_sequence(480, 270, 64, storage)


def test_config2_1080p_64_frames_fp32():
    # BASELINE config 2 at the north_star bar: <= 1e-4 relative in FP32 after 5 levels over 64 frames,
    # history bit-exact; checked every 4th frame (state errors would persist) to bound the oracle's CPU time.
    run_sequence(1920, 1080, 64, "f32", check_every=4)


def test_config2_1080p_fp16_reference_layout():
    run_sequence(1920, 1080, 8, "f16", check_every=1)


@pytest.mark.parametrize("steps", [0, 1, 3, 4])
def test_other_level_counts(steps):
    run_sequence(320, 200, 5, "f32", steps=steps)


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
