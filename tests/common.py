"""Shared helpers for the tests: random small scenes, error metrics, numpy<->torch plumbing."""
import numpy as np


def random_scene(rng, W, H, bg_frac=0.15, n_surf=3, max_motion=3.0, storage="f16"):
    """Random but structured inputs for ONE frame: a few depth/normal 'surfaces', some background, smallish
    motion.  Returns planes in the reference texel formats (see include/svgf.h)."""
    surf = rng.integers(0, n_surf, size=(H, W))
    # blocky surfaces so that neighbourhoods mix surfaces
    bx, by = max(1, W // 4), max(1, H // 3)
    surf = surf[::by, ::bx].repeat(by, 0).repeat(bx, 1)[:H, :W]
    if surf.shape != (H, W):
        surf = np.pad(surf, ((0, H - surf.shape[0]), (0, W - surf.shape[1])), mode="edge")
    normals = rng.normal(size=(n_surf, 3))
    normals /= np.linalg.norm(normals, axis=1, keepdims=True)
    n = normals[surf] + rng.normal(scale=0.02, size=(H, W, 3))
    n /= np.linalg.norm(n, axis=2, keepdims=True)
    depth = (3.0 + 4.0 * surf + rng.uniform(0, 0.5, size=(H, W))).astype(np.float32)
    dz = rng.uniform(0.0, 0.2, size=(H, W)).astype(np.float32)
    bg = rng.uniform(size=(H, W)) < bg_frac
    normal = np.zeros((H, W, 4), np.float16)
    normal[..., :3] = n.astype(np.float16)
    normal[..., 3] = (surf % 5).astype(np.float16)
    uv = np.zeros((H, W, 4), np.float16)
    uv[..., :3] = rng.uniform(size=(H, W, 3)).astype(np.float16)
    uv[..., 3] = (surf + 1).astype(np.float16)
    motion = np.zeros((H, W, 4), np.float32)
    motion[..., 0] = rng.uniform(-max_motion, max_motion, size=(H, W))
    motion[..., 1] = rng.uniform(-max_motion, max_motion, size=(H, W))
    motion[..., 2] = depth
    motion[..., 3] = dz
    normal[bg] = 0
    uv[bg] = 0
    motion[bg] = 0
    cdt = np.float16 if storage == "f16" else np.float32
    colour = np.ones((H, W, 4), cdt)
    colour[..., :3] = rng.uniform(0, 1.4, size=(H, W, 3)).astype(cdt)   # some > 1: exercises the clamp (D8)
    return {"normal": normal.view(np.uint16), "uv": uv.view(np.uint16), "motion": motion, "colour": colour}


def decode_gbuf(planes):
    """float views for the pure-Python restatement."""
    return {"normal": planes["normal"].view(np.float16).astype(np.float64),
            "inst": planes["uv"].view(np.float16)[..., 3].astype(np.int64),
            "motion": planes["motion"].astype(np.float64)}


def rel_err(got, want, floor):
    """max over elements of |got - want| / max(|want|, floor)."""
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    return float((np.abs(got - want) / np.maximum(np.abs(want), floor)).max())


def half_ulp_diff(got, want):
    """|difference| in units of fp16 ulps (ordered-integer distance) for two float16 arrays."""
    def key(a):
        b = a.view(np.uint16).astype(np.int32)
        return np.where(b & 0x8000, -(b & 0x7fff), b)
    return np.abs(key(np.ascontiguousarray(got)) - key(np.ascontiguousarray(want)))


# ---- parity tolerances (stated once, used by every GPU parity test) --------------------------------------
# FP32 storage ("parity mode"): north_star's bar is max relative error <= 1e-4.  Relative error is
# |got - want| / max(|want|, floor) with a floor of 1e-2 of the channel's natural range (radiance, moments:
# range [0,1] -> floor 1e-2; variance: the temporal variance is <= 0.25 -> floor 2.5e-3), i.e. an absolute
# error of 1e-6 (2.5e-7 for variance) is the smallest error the bar resolves.
F32_REL_TOL = 1e-4
F32_FLOOR_RADIANCE = 1e-2
F32_FLOOR_VARIANCE = 2.5e-3
# FP16 storage (the reference's layout): every stored value is one fp16 rounding of an fp32 result, so a
# 1e-7 difference before rounding can flip the stored half by one ulp (4.9e-4 relative), and the next stage
# may amplify it slightly.  Bar: every value within 2 fp16 ulps of the oracle (or 1e-4 absolute for values so
# small that 2 ulps < 1e-4), and at most 1% of values differing at all per stage.
F16_MAX_ULPS = 2
F16_ABS_FLOOR = 1e-4
F16_MAX_FLIP_FRACTION = 0.01


def f32_errors(got, want):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    c = got.shape[-1]
    out = {"rgb": rel_err(got[..., :3] if c == 4 else got, want[..., :3] if c == 4 else want, F32_FLOOR_RADIANCE)}
    if c == 4:
        out["var"] = rel_err(got[..., 3], want[..., 3], F32_FLOOR_VARIANCE)
    return out


def assert_f32_close(got, want, what=""):
    e = f32_errors(got, want)
    assert max(e.values()) <= F32_REL_TOL, f"{what}: relative errors {e} exceed {F32_REL_TOL}"
    return e


def f16_errors(got, want):
    got = np.ascontiguousarray(got)
    want = np.ascontiguousarray(want)
    ulps = half_ulp_diff(got, want)
    absd = np.abs(got.astype(np.float64) - want.astype(np.float64))
    bad = (ulps > F16_MAX_ULPS) & (absd > F16_ABS_FLOOR)
    return {"max_ulps": int(ulps.max()), "flip_fraction": float((ulps > 0).mean()), "violations": int(bad.sum()),
            "max_abs": float(absd.max())}


def assert_f16_close(got, want, what="", max_flips=F16_MAX_FLIP_FRACTION):
    e = f16_errors(got, want)
    assert e["violations"] == 0, f"{what}: {e}"
    assert e["flip_fraction"] <= max_flips, f"{what}: {e}"
    return e


def assert_close(got, want, storage, what="", **kw):
    return assert_f32_close(got, want, what) if storage == "f32" else assert_f16_close(got, want, what, **kw)
