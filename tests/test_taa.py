"""TAA + sRGB resolve (reference src/Filter.cuh:288-357, launched at src/App.cu:516-522): the oracle's restatement
against hand-derived values (CPU), the CUDA kernel against the oracle (-m gpu), and the oracle against the reference's OWN
kernel on histories that are fixed points of it - the only inputs for which the reference kernel, which reads the plane
it writes (D13), has a deterministic output."""
import ctypes as C

import numpy as np
import pytest

from common import assert_close, half_ulp_diff
from oracle_lib import OracleFilter, oracle_taa


def _scene(rng, W, H, storage):
    cdt = np.float16 if storage == "f16" else np.float32
    yy, xx = np.mgrid[0:H, 0:W]
    base = 0.35 + 0.3 * np.sin(xx / 9.0)[..., None] * np.cos(yy / 7.0)[..., None] * np.array([1.0, 0.7, 0.5])
    img = np.empty((H, W, 4), np.float64)
    img[..., :3] = base + rng.normal(scale=0.08, size=(H, W, 3))
    img[..., 3] = rng.uniform(0, 0.01, size=(H, W))
    img[: H // 6] = 0.0                      # a black band (background): exercises the sRGB linear branch and sqrt(0)
    img[-3:, :, :3] = 1.3                    # values above 1: the [0,1] clamp of imageLoad
    return img.astype(cdt)


def srgb(c):
    c = np.float32(c)
    return np.float32(12.92) * c if c <= np.float32(0.0031308) else np.float32(1.055) * np.float32(c) ** np.float32(1 / 2.4) - np.float32(0.055)


def test_oracle_taa_flat_field_known_answers():
    # flat grey field, history alpha 0 (first frame): mix rate 0 -> the blend keeps the (black) history, the clamp pulls it
    # up to the neighbourhood box = the field itself, so the output is sRGB(field) with alpha 1 - up to the PAL-YUV round
    # trip, whose published 5-digit matrices are inverses of each other only to ~5e-6
    W, H = 9, 7
    f = np.zeros((H, W, 4), np.float32)
    f[..., :3] = 0.25
    out = oracle_taa(f, np.zeros_like(f), "f32")
    assert np.allclose(out[..., :3], srgb(0.25), rtol=0, atol=5e-6) and np.all(out[..., 3] == 1.0)
    # history = the resolved field itself (alpha 1 -> rate 0.5): sqrt(mix(h^2, c^2, .5)) lies above the flat box, the clamp
    # returns the field again: the resolve is idempotent on a flat field
    out2 = oracle_taa(f, out, "f32")
    assert np.allclose(out2[..., :3], srgb(0.25), rtol=0, atol=5e-6)
    # black field: sRGB linear branch, exact zeros, alpha still 1
    z = np.zeros((H, W, 4), np.float32)
    outz = oracle_taa(z, z, "f32")
    assert np.all(outz[..., :3] == 0.0) and np.all(outz[..., 3] == 1.0)


def test_oracle_taa_samples_the_floor_texel_one_pixel_up_left():
    # textureSample (src/Filter.cuh:116-131) returns the floor texel of uv * (size - 1) with uv = (x, y) / size: the centre
    # tap of pixel (x, y) is texel (x - 1, y - 1) for x, y >= 1.  A single bright texel therefore shows up shifted.
    W, H = 16, 12
    f = np.zeros((H, W, 4), np.float32)
    f[5, 7, :3] = 0.8
    hist = np.zeros_like(f)
    hist[..., 3] = 1.0                       # steady state: rate 0.5
    out = oracle_taa(f, hist, "f32")
    lit = np.argwhere(out[..., 0] > 0)
    assert (6, 8) in {tuple(p) for p in lit}                      # centre tap of (8, 6) is texel (7, 5)
    assert out[6, 8, 0] == out[..., 0].max()


def test_oracle_taa_rejects_aliasing_planes():
    from oracle_lib import oracle
    f = np.zeros((4, 4, 4), np.float32)
    assert oracle().svgf_oracle_taa(4, 4, 1, f.ctypes.data, f.ctypes.data, f.ctypes.data) != 0


@pytest.mark.gpu
@pytest.mark.parametrize("storage", ["f16", "f32"])
@pytest.mark.parametrize("size", [(1, 1), (5, 3), (33, 17), (130, 67), (641, 363)])
def test_taa_kernel_against_the_oracle_over_a_short_sequence(size, storage):
    import torch
    from gpu_util import npy
    from svgf_b200 import SvgfFilter
    W, H = size
    rng = np.random.default_rng(W * 31 + H)
    f = SvgfFilter(W, H, storage=storage)
    o = OracleFilter(W, H, storage=storage)
    for t in range(4):
        img = _scene(rng, W, H, storage)
        o.FilterBuffer[0][...] = img
        f.FilterBuffer[0].copy_(torch.from_numpy(img))
        want = o.TAA()
        got = npy(f.TAA())
        # teacher-forced: the kernel's history is the oracle's, so that one resolve is compared, not an accumulated drift
        if storage == "f32":
            # north_star's bar: 1e-4.  The kernel takes pow(x, 1/2.4) as ex2(lg2(x) / 2.4) and sqrt as MUFU.SQRT (2^-21 each)
            # and contracts the YUV dot products; outputs are in [0, 1]
            assert np.abs(got.astype(np.float64) - want).max() <= 1e-4, f"frame {t}"
        else:
            # the resolve is discontinuous where a decoded component is within rounding of 0 (negative -> NaN -> the whole pixel
            # is blacked out, src/Filter.cuh:351): libm's powf(x, 2) vs x * x can still put a handful of pixels on the other side
            from common import f16_errors
            e = f16_errors(got, want)
            assert e["violations"] <= max(2, int(2e-5 * got.size)) and e["flip_fraction"] <= 0.02, f"TAA frame {t} {size}: {e}"
        assert np.all(got[..., 3] == 1.0)
        f.TAABuffer[f.PingPongInx].copy_(torch.from_numpy(want))
        f.EndFrame(); o.EndFrame()


@pytest.mark.gpu
def test_taa_rejects_in_place_history():
    from svgf_b200 import SvgfFilter, _lib
    f = SvgfFilter(16, 16)
    buf = f.FilterBuffer[1]
    st = f.lib.svgf_taa(f._ctx, C.c_void_p(f.FilterBuffer[0].data_ptr()), C.c_void_p(buf.data_ptr()), C.c_void_p(buf.data_ptr()), f._stream())
    assert st == _lib.SVGF_INVALID_ARG


@pytest.mark.gpu
def test_oracle_taa_against_the_reference_kernel_on_fixed_point_histories():
    """The reference kernel reads texel (x-1, y-1) of the plane it writes.  Iterating it on a static input drives that plane
    to a fixed point (every texel reproduces itself); from then on it does not matter whether a thread sees its neighbour's
    old or new value, the launch is deterministic, and its output must equal the oracle's for history = that plane."""
    from oracle_lib import PLANE_FILTER, RefKernels, ref, ref_available
    if not ref_available() or not hasattr(ref(), "svgf_ref_taa"):
        pytest.skip("oracle/_ref/libsvgf_refkernels.so (with svgf_ref_taa) not built")
    W, H = 200, 120
    rng = np.random.default_rng(3)
    img = _scene(rng, W, H, "f16")
    r = RefKernels(W, H)
    r.set_plane(PLANE_FILTER, 0, img)
    r.set_plane(PLANE_FILTER, 1, np.zeros_like(img))
    prev = None
    for it in range(200):
        assert ref().svgf_ref_taa(r.ctx) == 0
        cur = r.get_plane(PLANE_FILTER, 1)
        if prev is not None and np.array_equal(cur.view(np.uint16), prev.view(np.uint16)):
            break
        prev = cur
    fixed = np.array_equal(cur.view(np.uint16), prev.view(np.uint16))
    assert fixed, "the reference kernel did not reach a fixed point in 200 launches"
    again = oracle_taa(img, cur, "f16")
    u = half_ulp_diff(again, cur)
    absd = np.abs(again.astype(np.float64) - cur.astype(np.float64))
    frac_exact = float((u == 0).mean())
    # values beyond 2 ulps: the resolve is ill-conditioned where a decoded component is within rounding of 0 (the PAL-YUV
    # matrices are inverses only to ~5e-6): pow(x, 0.5) turns a 1e-8 difference there into 1e-4, on the negative side into
    # NaN and a blacked-out pixel (src/Filter.cuh:348-351).  The compiled reference contracts its dot products into FMAs and
    # uses CUDA's powf, so such values may differ from the un-contracted restatement; they must be rare, and every one of
    # them must be such a near-zero channel (dark on one side) or a blacked-out pixel.
    far = (u > 2) & (absd > 1e-4)
    black = (again[..., :3] == 0).all(axis=-1) | (cur[..., :3] == 0).all(axis=-1)
    dark = np.minimum(again.astype(np.float32), cur.astype(np.float32)) < 0.03
    r.close()
    assert frac_exact >= 0.98, f"exact fraction {frac_exact}"
    assert far.any(axis=-1).mean() <= 2e-3, f"{int(far.any(axis=-1).sum())} of {far.shape[0] * far.shape[1]} pixels differ by more than 2 ulps"
    unexplained = far & ~dark & ~black[..., None]
    assert unexplained.sum() == 0, f"{int(unexplained.sum())} values differ by more than 2 ulps away from the discontinuity: " \
                                   f"{again[unexplained][:4]} vs {cur[unexplained][:4]}"
