"""Known-answer tests that pin the scalar oracle (oracle/svgf_oracle.cpp) on CPU: fp16 codec against numpy,
the alpha table, hand-derivable micro cases and agreement with an independent pure-Python
restatement on small random scenes."""
import ctypes as C

import numpy as np
import pytest

import oracle_py
from common import decode_gbuf, random_scene, rel_err
from oracle_lib import OracleFilter, oracle
from svgf_b200._lib import default_params


def test_half_decode_all_patterns():
    o = oracle()
    bits = np.arange(65536, dtype=np.uint16)
    want = bits.view(np.float16).astype(np.float32)
    got = np.array([o.svgf_oracle_h2f(int(b)) for b in bits], np.float32)
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan)
    assert np.array_equal(got[~nan].view(np.uint32), want[~nan].view(np.uint32))


def test_half_encode_rne():
    o = oracle()
    rng = np.random.default_rng(1)
    vals = np.concatenate([
        rng.uniform(-2, 2, 20000), rng.uniform(-70000, 70000, 5000), 10.0 ** rng.uniform(-9, -3, 5000),
        np.array([0.0, -0.0, 65504, 65519.99, 65520, 65536, 1e9, -1e9, 2.0 ** -24, 2.0 ** -25, 2.0 ** -25 * 1.0001, 2.0 ** -14,
                  5.96e-8, 6.1e-5, np.inf, -np.inf]),
        # exact ties between two halves
        (np.arange(1024, 2048) + 0.5) / 1024.0,
    ]).astype(np.float32)
    got = np.array([o.svgf_oracle_f2h(float(v)) for v in vals], np.uint16)
    with np.errstate(over="ignore"):
        want = vals.astype(np.float16).view(np.uint16)
    assert np.array_equal(got, want)
    assert o.svgf_oracle_f2h(float("nan")) & 0x7c00 == 0x7c00 and o.svgf_oracle_f2h(float("nan")) & 0x3ff


def test_alpha_table_fp32_divide_equals_reference_fp64_divide():
    # reference src/Filter.cuh:381: Alpha = 1.0 / HistoryLength (FP64) stored to float.  The kernels use one
    # IEEE fp32 divide; identical for every h a uint8 can hold.
    for h in range(1, 256):
        assert np.float32(1.0 / h) == np.float32(1.0) / np.float32(h)


def _mk(W, H, storage="f32", params=None):
    return OracleFilter(W, H, storage=storage, params=params)


def test_flat_field_is_a_fixed_point_of_atrous():
    W, H = 12, 9
    of = _mk(W, H)
    n = np.zeros((H, W, 4), np.float16); n[..., 2] = 1.0
    of.normal[0][...] = n.view(np.uint16)
    of.motion[0][..., 2] = 5.0
    of.motion[0][..., 3] = 0.01
    of.FilterBuffer[0][...] = np.array([0.25, 0.5, 0.75, 0.04], np.float32)
    of.params.atrous_iterations = 5
    of.WaveletFilter()
    out = of.FilterBuffer[0]
    assert np.allclose(out[..., :3], [0.25, 0.5, 0.75], rtol=0, atol=2e-7)
    # variance shrinks: sum(w^2 v)/sum(w)^2 < v for more than one tap
    assert (out[..., 3] < 0.04).all() and (out[..., 3] > 0).all()
    # level-0 output was written to the colour history
    assert np.abs(of.RenderBuffer[0][..., :3] - [0.25, 0.5, 0.75]).max() < 2e-7


def test_atrous_single_tap_weights_by_hand():
    # 5x1 image, step 1: centre pixel x=2 sees taps at distance 1 and 2 with kernel 2/3 and 1/6
    # (KernelWeights {1, 2/3, 1/6}, reference src/Filter.cuh:540), equal depth/normal, variance v, luminance
    # difference dl: w = k * exp(-dl / (phi * sqrt(v + 1e-10)))
    W, H = 5, 1
    of = _mk(W, H)
    n = np.zeros((H, W, 4), np.float16); n[..., 0] = 1.0
    of.normal[0][...] = n.view(np.uint16)
    of.motion[0][..., 2] = 2.0
    of.motion[0][..., 3] = 0.5
    v = 0.09
    g = np.array([0.1, 0.2, 0.5, 0.3, 0.9], np.float32)
    of.FilterBuffer[0][0, :, 0] = g; of.FilterBuffer[0][0, :, 1] = g; of.FilterBuffer[0][0, :, 2] = g
    of.FilterBuffer[0][..., 3] = v
    of.params.atrous_iterations = 1
    of.params.phi_colour = 4.0
    of.WaveletFilter()
    phi = 4.0 * np.sqrt(np.float64(np.float32(1e-10) + np.float32(v)))
    k = {1: np.float32(2 / 3), 2: np.float32(1 / 6)}
    S, A, V = 1.0, float(g[2]), v
    for x in range(5):
        if x == 2:
            continue
        w = float(k[abs(x - 2)]) * np.exp(-abs(float(g[2]) - float(g[x])) / phi)   # luminance of grey = value (coefficients sum to 1)
        S += w; A += w * float(g[x]); V += w * w * v
    got = of.FilterBuffer[0][0, 2]
    assert abs(got[0] - A / S) < 1e-6
    assert abs(got[3] - V / (S * S)) < 1e-6


def test_depth_edge_blocks_filtering():
    # two half planes 100 units apart with a tiny depth derivative: cross-edge weights underflow to zero
    W, H = 10, 6
    of = _mk(W, H)
    n = np.zeros((H, W, 4), np.float16); n[..., 2] = 1.0
    of.normal[0][...] = n.view(np.uint16)
    of.motion[0][..., 2] = 5.0
    of.motion[0][:, 5:, 2] = 105.0
    of.motion[0][..., 3] = 1e-3
    of.FilterBuffer[0][:, :5, :3] = 0.2
    of.FilterBuffer[0][:, 5:, :3] = 0.8
    of.FilterBuffer[0][..., 3] = 1.0
    of.params.atrous_iterations = 3
    of.WaveletFilter()
    out = of.FilterBuffer[0]
    assert np.abs(out[:, :5, :3] - 0.2).max() < 1e-6 and np.abs(out[:, 5:, :3] - 0.8).max() < 1e-6


def test_background_pixels():
    # D7: background (depth 0) never accumulates history, the variance pass outputs zeros for it (all 49
    # weights are zero because the zero normal gives pow(0, phiN) = 0), a-trous passes it through and does
    # not write the colour history.
    W, H = 8, 8
    of = _mk(W, H, storage="f16")
    of.Reset()
    of.RenderBuffer[0][...] = np.float16(0.5)
    of.Filter()
    assert (of.HistoryLengthBuffer == 1).all()
    assert (of.FilterBuffer[0] == 0).all()
    # temporal output (cur colour, variance 0) is still in RenderBuffer[P]: a-trous level 0 did not overwrite it
    assert np.allclose(of.RenderBuffer[0][..., :3].astype(np.float32), 0.5) and (of.RenderBuffer[0][..., 3] == 0).all()
    of.EndFrame()
    of.RenderBuffer[1][...] = np.float16(0.25)
    of.Filter()
    assert (of.HistoryLengthBuffer == 1).all()      # zero normals: dot = 0 < 0.9 forever


def test_history_accumulates_and_saturates_and_alpha():
    W, H = 6, 5
    of = _mk(W, H, storage="f32")
    of.params.history_cap = 5
    of.params.atrous_iterations = 0
    n = np.zeros((H, W, 4), np.float16); n[..., 1] = 1.0
    vals = [0.9, 0.1, 0.5, 0.3, 0.7, 0.2, 0.6, 0.4]
    of.Reset()
    mean = 0.0
    for t, v in enumerate(vals):
        P = of.PingPongInx
        of.normal[P][...] = n.view(np.uint16)
        of.motion[P][..., 2] = 4.0
        of.uv[P][..., 3] = np.float16(3).view(np.uint16)
        of.RenderBuffer[P][..., :3] = v
        of.RenderBuffer[P][..., 3] = 1.0
        of.Filter()
        h = min(5, t + 1)
        assert (of.HistoryLengthBuffer == h).all()
        alpha = np.float32(1.0 / h)
        mean = v if t == 0 else np.float32(np.float32(mean * (np.float32(1) - alpha)) + np.float32(np.float32(v) * alpha))
        assert np.allclose(of.RenderBuffer[P][..., 0], mean, rtol=0, atol=1e-7), t
        of.EndFrame()


def test_history_snapshot_semantics_under_motion():
    # D3: every pixel fetches history from its left neighbour (motion +/-1): with snapshot semantics all
    # interior pixels see the PREVIOUS frame's value regardless of evaluation order.
    W, H = 9, 3
    of = _mk(W, H, storage="f32")
    of.params.atrous_iterations = 0
    n = np.zeros((H, W, 4), np.float16); n[..., 2] = 1.0
    of.Reset()
    for t in range(4):
        P = of.PingPongInx
        of.normal[P][...] = n.view(np.uint16)
        of.motion[P][..., 0] = -1.7 if t else 0.0   # truncates to -1
        of.motion[P][..., 2] = 4.0
        of.RenderBuffer[P][...] = 0.5
        of.Filter()
        of.EndFrame()
    # column x has reprojected successfully min(x, 3) times after the reset frame
    want = np.minimum(np.arange(W), 3) + 1
    assert np.array_equal(of.HistoryLengthBuffer, np.broadcast_to(want.astype(np.uint8), (H, W)))


def test_motion_truncates_toward_zero():
    W, H = 7, 1
    of = _mk(W, H, storage="f32")
    of.params.atrous_iterations = 0
    n = np.zeros((H, W, 4), np.float16); n[..., 2] = 1.0
    for k in range(2):
        of.normal[k][...] = n.view(np.uint16)
        of.motion[k][..., 2] = 4.0
    of.HistoryLengthBuffer[0] = np.arange(W) + 10
    of.RenderBuffer[1][0, :, 0] = np.arange(W) / 10.0      # previous colour identifies the source pixel
    of.PingPongInx = 0
    mv = np.array([0.99, -0.99, 1.0, -1.0, 1.9, -2.5, 0.0], np.float32)
    of.motion[0][0, :, 0] = mv
    of.params.history_cap = 255
    of.TemporalFilter()
    src = np.arange(W) + mv.astype(np.int32)                # C truncation, reference src/Filter.cuh:232
    ok = (src >= 0) & (src < W)
    want = np.where(ok, np.arange(W)[np.clip(src, 0, W - 1)] + 10 + 1, 1)
    assert np.array_equal(of.HistoryLengthBuffer[0], want.astype(np.uint8))


def test_odd_level_count_lands_in_filter0():
    rng = np.random.default_rng(3)
    W, H = 16, 11
    pl = random_scene(rng, W, H, storage="f32")
    outs = []
    for N in (3, 4):
        of = _mk(W, H)
        of.params.atrous_iterations = N
        of.set_inputs(pl)
        of.HistoryLengthBuffer[...] = 9
        of.FilterBuffer[0][...] = pl["colour"]
        of.WaveletFilter()
        outs.append(of.FilterBuffer[0].copy())
    assert not np.array_equal(outs[0], outs[1])
    # N = 3 result equals three explicit levels
    of = _mk(W, H)
    of.set_inputs(pl)
    buf = [pl["colour"].copy(), np.zeros_like(pl["colour"])]
    g = of.gbuf(0)
    for i in range(3):
        oracle().svgf_oracle_atrous_level(C.byref(of.params), W, H, 1, C.byref(g), buf[i & 1].ctypes.data,
                                          buf[1 - (i & 1)].ctypes.data, of.RenderBuffer[0].ctypes.data, i)
    assert np.array_equal(outs[0], buf[1])


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_agrees_with_independent_python_restatement(seed):
    rng = np.random.default_rng(seed)
    W, H = 13, 10
    cur = random_scene(rng, W, H, storage="f32")
    prev = random_scene(rng, W, H, storage="f32")
    # make about half the pixels reprojectable: same surface data, different noise
    keep = rng.uniform(size=(H, W)) < 0.6
    for k in ("normal", "uv"):
        prev[k][keep] = cur[k][keep]
    prev["motion"][keep, 2:] = cur["motion"][keep, 2:]
    cur["motion"][..., :2] = np.where(rng.uniform(size=(H, W, 1)) < 0.7, 0.0, cur["motion"][..., :2])
    p = default_params()
    p.history_cap = 7
    of = _mk(W, H, params=p)
    of.PingPongInx = 0
    of.set_inputs(cur)
    of.normal[1][...] = prev["normal"]; of.uv[1][...] = prev["uv"]; of.motion[1][...] = prev["motion"]
    of.RenderBuffer[1][...] = rng.uniform(0, 1.2, size=(H, W, 4)).astype(np.float32)
    of.MomentsBuffer[1][...] = rng.uniform(0, 1, size=(H, W, 2)).astype(np.float32)
    hist0 = rng.integers(0, 9, size=(H, W)).astype(np.uint8)
    of.HistoryLengthBuffer[...] = hist0
    prev_col, prev_mom = of.RenderBuffer[1].copy(), of.MomentsBuffer[1].copy()

    c_py, h_py, m_py = oracle_py.temporal(p, decode_gbuf(cur), decode_gbuf(prev), prev_col, cur["colour"], hist0, prev_mom)
    of.TemporalFilter()
    assert np.array_equal(of.HistoryLengthBuffer, h_py)
    assert rel_err(of.RenderBuffer[0][..., :3], c_py[..., :3], 1e-3) < 2e-6
    assert np.abs(of.RenderBuffer[0][..., 3] - c_py[..., 3]).max() < 5e-7   # m2 - m1^2 cancels in fp32
    assert rel_err(of.MomentsBuffer[0], m_py, 1e-3) < 5e-6

    v_py = oracle_py.variance(p, decode_gbuf(cur), of.RenderBuffer[0], of.MomentsBuffer[0], of.HistoryLengthBuffer)
    of.FilterMoments()
    assert rel_err(of.FilterBuffer[0][..., :3], v_py[..., :3], 1e-3) < 2e-5
    assert np.abs(of.FilterBuffer[0][..., 3] - v_py[..., 3]).max() < 2e-5

    a_in = of.FilterBuffer[0].copy()
    # keep the centre variance away from 0: with var ~ 0 the luminance phi is 1e-4 and the weights amplify
    # fp32-vs-fp64 luminance rounding (6e-8) by 1e4 — a property of the reference math, exercised on the GPU
    # against the fp32 oracle instead (same operation order), not against this fp64 restatement.
    a_in[..., 3] = np.maximum(a_in[..., 3], 1e-2)
    g = of.gbuf(0)
    for level in (0, 1, 2):
        out = np.zeros_like(a_in)
        hc = of.RenderBuffer[0].copy()
        oracle().svgf_oracle_atrous_level(C.byref(p), W, H, 1, C.byref(g), a_in.ctypes.data, out.ctypes.data, hc.ctypes.data, level)
        o_py, hc_py = oracle_py.atrous(p, decode_gbuf(cur), a_in, level, of.RenderBuffer[0])
        assert rel_err(out[..., :3], o_py[..., :3], 1e-3) < 2e-5, level
        assert rel_err(out[..., 3], o_py[..., 3], 1e-4) < 2e-4, level
        if level == 0:
            assert rel_err(hc, hc_py, 1e-3) < 2e-4


@pytest.mark.parametrize("seed", [3, 4])
def test_variance_prefilter_agrees_with_independent_python_restatement(seed):
    """SVGF_VARIANCE_PREFILTER_GAUSS3 (include/svgf.h; the SVGF paper's 3x3 variance blur, absent from the reference)."""
    rng = np.random.default_rng(seed)
    W, H = 12, 9
    cur = random_scene(rng, W, H, storage="f32")
    p = default_params()
    p.variance_prefilter = 1
    of = _mk(W, H, params=p)
    of.PingPongInx = 0
    of.set_inputs(cur)
    a_in = rng.uniform(0, 1.1, size=(H, W, 4)).astype(np.float32)
    a_in[..., 3] = rng.uniform(1e-2, 0.3, size=(H, W)).astype(np.float32)
    a_in[rng.uniform(size=(H, W)) < 0.3, 3] = 0.9          # strong variance contrast: the blur must matter
    g = of.gbuf(0)
    plain = default_params()
    for level in (0, 1, 2):
        out = np.zeros_like(a_in); out_plain = np.zeros_like(a_in)
        hc = of.RenderBuffer[0].copy()
        assert oracle().svgf_oracle_atrous_level(C.byref(p), W, H, 1, C.byref(g), a_in.ctypes.data, out.ctypes.data, hc.ctypes.data, level) == 0
        assert oracle().svgf_oracle_atrous_level(C.byref(plain), W, H, 1, C.byref(g), a_in.ctypes.data, out_plain.ctypes.data, hc.ctypes.data, level) == 0
        o_py, _ = oracle_py.atrous(p, decode_gbuf(cur), a_in, level, None)
        assert rel_err(out[..., :3], o_py[..., :3], 1e-3) < 2e-5, level
        assert rel_err(out[..., 3], o_py[..., 3], 1e-4) < 2e-4, level
        assert np.abs(out - out_plain).max() > 1e-3, "the prefilter changed nothing"


def test_variance_prefilter_of_a_constant_variance_plane_is_the_identity():
    rng = np.random.default_rng(9)
    W, H = 16, 8
    cur = random_scene(rng, W, H, storage="f16")
    a_in = rng.uniform(0, 1, size=(H, W, 4)).astype(np.float16)
    a_in[..., 3] = np.float16(0.125)                        # exactly representable: the blur returns it exactly
    outs = []
    for mode in (0, 1):
        p = default_params()
        p.variance_prefilter = mode
        of = _mk(W, H, storage="f16", params=p)
        of.set_inputs(cur)
        g = of.gbuf(0)
        out = np.zeros_like(a_in)
        hc = of.RenderBuffer[0].copy()
        assert oracle().svgf_oracle_atrous_level(C.byref(p), W, H, 0, C.byref(g), a_in.ctypes.data, out.ctypes.data, hc.ctypes.data, 1) == 0
        outs.append(out)
    assert np.array_equal(outs[0].view(np.uint16), outs[1].view(np.uint16))


@pytest.mark.parametrize("seed", [5, 6, 7])
def test_bilinear_reprojection_agrees_with_independent_python_restatement(seed):
    """SVGF_REPROJ_BILINEAR (include/svgf.h; the SVGF paper's 2x2 history fetch, absent from the reference)."""
    rng = np.random.default_rng(seed)
    W, H = 13, 10
    cur = random_scene(rng, W, H, storage="f32")
    prev = random_scene(rng, W, H, storage="f32")
    keep = rng.uniform(size=(H, W)) < 0.75
    for k in ("normal", "uv"):
        prev[k][keep] = cur[k][keep]
    prev["motion"][keep, 2:] = cur["motion"][keep, 2:]
    # fractional motion away from .0 / .5 (the fp64 restatement must pick the same texels and the same rounded length)
    cur["motion"][..., :2] = (rng.integers(-2, 3, size=(H, W, 2)) + rng.choice([0.13, 0.31, 0.71, 0.88], size=(H, W, 2))).astype(np.float32)
    p = default_params()
    p.history_cap = 9
    p.reproj_mode = 1
    of = _mk(W, H, params=p)
    of.PingPongInx = 0
    of.set_inputs(cur)
    of.normal[1][...] = prev["normal"]; of.uv[1][...] = prev["uv"]; of.motion[1][...] = prev["motion"]
    of.RenderBuffer[1][...] = rng.uniform(0, 1.2, size=(H, W, 4)).astype(np.float32)
    of.MomentsBuffer[1][...] = rng.uniform(0, 1, size=(H, W, 2)).astype(np.float32)
    hist0 = rng.integers(0, 12, size=(H, W)).astype(np.uint8)
    of.HistoryLengthBuffer[...] = hist0
    prev_col, prev_mom = of.RenderBuffer[1].copy(), of.MomentsBuffer[1].copy()
    c_py, h_py, m_py = oracle_py.temporal(p, decode_gbuf(cur), decode_gbuf(prev), prev_col, cur["colour"], hist0, prev_mom)
    of.TemporalFilter()
    assert np.array_equal(of.HistoryLengthBuffer, h_py)
    assert (h_py > 1).mean() > 0.3, "too few successful reprojections to mean anything"
    assert rel_err(of.RenderBuffer[0][..., :3], c_py[..., :3], 1e-3) < 5e-6
    assert np.abs(of.RenderBuffer[0][..., 3] - c_py[..., 3]).max() < 1e-6
    assert rel_err(of.MomentsBuffer[0], m_py, 1e-3) < 1e-5
    # and it is not the nearest fetch in disguise
    q = default_params(); q.history_cap = 9
    on = _mk(W, H, params=q)
    on.PingPongInx = 0
    on.set_inputs(cur)
    on.normal[1][...] = prev["normal"]; on.uv[1][...] = prev["uv"]; on.motion[1][...] = prev["motion"]
    on.RenderBuffer[1][...] = prev_col; on.MomentsBuffer[1][...] = prev_mom; on.HistoryLengthBuffer[...] = hist0
    on.TemporalFilter()
    assert np.abs(on.RenderBuffer[0] - of.RenderBuffer[0]).max() > 1e-2


def test_bilinear_reprojection_with_integer_motion_is_the_nearest_fetch():
    """Integer motion vectors put all the weight on one texel: both modes must then agree bit for bit."""
    rng = np.random.default_rng(11)
    W, H = 24, 16
    cur = random_scene(rng, W, H, storage="f16")
    cur["motion"][..., :2] = rng.integers(-3, 4, size=(H, W, 2)).astype(np.float32)
    res = []
    for mode in (0, 1):
        p = default_params()
        p.reproj_mode = mode
        of = _mk(W, H, storage="f16", params=p)
        of.PingPongInx = 0
        of.set_inputs(cur)
        of.normal[1][...] = of.normal[0]; of.uv[1][...] = of.uv[0]; of.motion[1][...] = of.motion[0]
        r2 = np.random.default_rng(12)
        of.RenderBuffer[1][...] = r2.uniform(0, 1, size=(H, W, 4)).astype(np.float16)
        of.MomentsBuffer[1][...] = r2.uniform(0, 1, size=(H, W, 2)).astype(np.float16)
        of.HistoryLengthBuffer[...] = r2.integers(0, 30, size=(H, W)).astype(np.uint8)
        of.TemporalFilter()
        res.append((of.RenderBuffer[0].copy(), of.MomentsBuffer[0].copy(), of.HistoryLengthBuffer.copy()))
    for a, b in zip(*res):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_argument_validation():
    of = _mk(4, 4)
    of.params.history_cap = 0
    with pytest.raises(RuntimeError):
        of.Filter()
    of.params.history_cap = 256
    with pytest.raises(RuntimeError):
        of.Filter()
