"""The procedural input generator: formats, determinism, and CUDA == CPU bit equality."""
import numpy as np
import pytest

from svgf_b200 import synth


def test_formats_and_semantics():
    W, H = 320, 180
    a = synth.frame_host(W, H, 5, with_position=True)
    n = a["normal"].view(np.float16).astype(np.float32)
    m = a["motion"]
    bg = m[..., 2] == 0
    assert 0.10 < bg.mean() < 0.30                      # a background region exists (pins D7)
    for k in ("normal", "uv", "position"):
        assert (a[k][bg] == 0).all()
    assert (m[bg] == 0).all()
    ln = np.linalg.norm(n[..., :3], axis=2)[~bg]
    assert np.abs(ln - 1).max() < 2e-3                  # unit normals, fp16 quantised
    inst = a["uv"].view(np.float16)[..., 3].astype(int)
    assert inst[~bg].min() >= 1 and len(np.unique(inst)) >= 6
    assert (m[..., 2][~bg] > 2).all() and (m[..., 2] < 40).all()
    assert (m[..., 3] >= 0).all()
    c = a["colour"].astype(np.float32)
    assert (c[..., 3] == 1).all() and c[..., :3].min() >= 0 and c[..., :3].max() > 1.0   # > 1 exercises the clamp
    # pan: nearer pixels move faster, everything moves the same way
    mvx = m[..., 0][~bg]
    assert mvx.min() > 0.5 and mvx.max() > 2 * mvx.min()


def test_deterministic_and_seeded():
    a = synth.frame_host(96, 64, 3, seed=7)
    b = synth.frame_host(96, 64, 3, seed=7, threads=1)
    c = synth.frame_host(96, 64, 3, seed=8)
    for k in a:
        assert np.array_equal(a[k], b[k])
    assert not np.array_equal(a["colour"], c["colour"])


def test_pan_reverses():
    a = synth.frame_host(160, 90, 10)["motion"]
    b = synth.frame_host(160, 90, 40)["motion"]
    fg = (a[..., 2] > 0) & (b[..., 2] > 0)
    assert (a[..., 0][fg] > 0).all() and (b[..., 0][fg] < 0).all()


def test_f32_storage_matches_f16_up_to_rounding():
    a = synth.frame_host(64, 48, 2, storage="f16")
    b = synth.frame_host(64, 48, 2, storage="f32")
    assert np.array_equal(a["motion"], b["motion"])
    with np.errstate(over="ignore"):
        assert np.array_equal(b["colour"].astype(np.float16), a["colour"])


@pytest.mark.gpu
@pytest.mark.parametrize("storage", ["f16", "f32"])
def test_cuda_generator_is_bit_identical_to_cpu(storage):
    import torch
    from svgf_b200.filter import GBuffer
    W, H = 333, 187
    dev = torch.device("cuda", 0)
    g = GBuffer(W, H, dev, with_position=True)
    col = torch.zeros(H, W, 4, dtype=torch.float16 if storage == "f16" else torch.float32, device=dev)
    for frame in (0, 1, 37):
        synth.frame_device(g, col, frame, seed=3)
        torch.cuda.synchronize()
        ref = synth.frame_host(W, H, frame, seed=3, storage=storage, with_position=True)
        assert np.array_equal(g.normal.cpu().numpy().view(np.uint16), ref["normal"])
        assert np.array_equal(g.uv.cpu().numpy().view(np.uint16), ref["uv"])
        assert np.array_equal(g.motion.cpu().numpy().view(np.uint32), ref["motion"].view(np.uint32))
        assert np.array_equal(g.position.cpu().numpy().view(np.uint32), ref["position"].view(np.uint32))
        got = col.cpu().numpy()
        assert np.array_equal(got.view(np.uint16 if storage == "f16" else np.uint32),
                              ref["colour"].view(np.uint16 if storage == "f16" else np.uint32))
