#!/usr/bin/env python
"""Whole-frame parity of the CUDA path against the scalar oracle, on disk (runs on the GPU box; ~5 minutes):

  python tools/parity_report.py --out profiles/parity_r02.json

For BASELINE configs 1-3 (1280x720 x 4 frames, 1920x1080 x 64 frames, 3840x2160 x 16 frames), both storage modes, ONE oracle
sequence per case drives two CUDA filters: `teacher_forced` starts every frame from the oracle's buffers (what one frame of
the path adds on its own: temporal + variance + 5 levels), `free_running` feeds on its own outputs for the whole sequence.
Per case: history-length and moments mismatches (bit-exact is the bar), and for the result and the colour history the
worst relative error (floors of tests/common.py), the fraction of values above 1e-4, fp16 ulps and flip fraction.

`fp64_truth`: where the two disagree beyond the bar, who is right?  One 1080p fp32 frame, level by level from IDENTICAL
inputs: the level evaluated in float64 (numpy, reference formulas, src/Filter.cuh:407-427,527-624) is the truth; reported
are |kernel - truth| and |oracle - truth| overall and on the pixels where kernel and oracle differ most.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FLOOR_RGB, FLOOR_VAR, TOL = 1e-2, 2.5e-3, 1e-4


def plane_stats(got, want, storage):
    from common import half_ulp_diff
    g64, w64 = got.astype(np.float64), want.astype(np.float64)
    d = np.abs(g64 - w64)
    floor = np.array([FLOOR_RGB] * 3 + [FLOOR_VAR])
    r = d / np.maximum(np.abs(w64), floor)
    out = {"max_rel_rgb": float(r[..., :3].max()), "max_rel_var": float(r[..., 3].max()), "frac_above_1e-4": float((r > TOL).mean()),
           "max_abs": float(d.max())}
    if storage == "f16":
        u = half_ulp_diff(got, want)
        out["max_ulps"] = int(u.max())
        out["flip_fraction"] = float((u > 0).mean())
        out["frac_above_2ulps"] = float(((u > 2) & (d > 1e-4)).mean())
    return out


def merge_worst(acc, st):
    for k, v in st.items():
        acc[k] = max(acc.get(k, 0), v)


def run_case(W, H, frames, storage, check_every):
    import torch
    from gpu_util import load_state_from_oracle, npy, upload_inputs
    from oracle_lib import OracleFilter
    from svgf_b200 import SvgfFilter, synth
    o = OracleFilter(W, H, storage=storage)
    tf, fr = SvgfFilter(W, H, storage=storage), SvgfFilter(W, H, storage=storage)
    o.Reset(); tf.Reset(); fr.Reset()
    res = {m: {"history_mismatches": 0, "moments_mismatching_bytes": 0, "result": {}, "colour_history": {}, "frames_checked": 0,
               "worst_frame": None} for m in ("teacher_forced", "free_running")}
    worst_key = {"teacher_forced": -1.0, "free_running": -1.0}
    t0 = time.time()
    for t in range(frames):
        planes = synth.frame_host(W, H, t, storage=storage)
        o.set_inputs(planes)
        load_state_from_oracle(tf, o)
        upload_inputs(fr, planes)
        o.Filter(); tf.Filter(); fr.Filter()
        if t % check_every == 0 or t == frames - 1:
            P = o.PingPongInx
            for name, f in (("teacher_forced", tf), ("free_running", fr)):
                r = res[name]
                r["history_mismatches"] += int((npy(f.HistoryLengthBuffer) != o.HistoryLengthBuffer).sum())
                r["moments_mismatching_bytes"] += int((npy(f.MomentsBuffer[P]).view(np.uint8) != o.MomentsBuffer[P].view(np.uint8)).sum())
                st = plane_stats(npy(f.FilterBuffer[0]), o.FilterBuffer[0], storage)
                merge_worst(r["result"], st)
                merge_worst(r["colour_history"], plane_stats(npy(f.RenderBuffer[P]), o.RenderBuffer[P], storage))
                r["frames_checked"] += 1
                key = max(st["max_rel_rgb"], st["max_rel_var"])
                if key > worst_key[name]:
                    worst_key[name], r["worst_frame"] = key, t
        o.EndFrame(); fr.EndFrame()
    torch.cuda.synchronize()
    tf.close(); fr.close()
    return {"width": W, "height": H, "frames": frames, "storage": storage, "checked_every": check_every, "seconds": round(time.time() - t0, 1), **res}


# ---- float64 evaluation of one a-trous level (the truth for `fp64_truth`) -----------------------------------------
def atrous_level_f64(p, planes, inp, level):
    H, W = inp.shape[:2]
    S = 1 << level
    c = np.clip(inp.astype(np.float64), 0.0, 1.0)                                   # imageLoad clamp, :543,:586
    lum = 0.2126 * c[..., 0] + 0.7152 * c[..., 1] + 0.0722 * c[..., 2]
    mot = planes["motion"].astype(np.float64)
    z = np.where(mot[..., 2] == 0.0, 1e30, mot[..., 2])
    dz = np.where(mot[..., 2] == 0.0, 0.0, mot[..., 3])
    n = planes["normal"].view(np.float16).astype(np.float64)[..., :3]
    var = c[..., 3]
    phiL = p.phi_colour * np.sqrt(np.maximum(0.0, np.float64(np.float32(1e-10)) + var))
    phiZ = np.maximum(dz, np.float64(np.float32(1e-6))) * S * p.phi_depth
    KW = [1.0, float(np.float32(2.0 / 3.0)), float(np.float32(1.0 / 6.0))]
    sw = np.ones((H, W))
    acc = c.copy()
    pad = 2 * S

    def shifted(a, dx, dy, fill):
        out = np.full_like(a, fill)
        ys0, ys1 = max(0, -dy), min(H, H - dy)
        xs0, xs1 = max(0, -dx), min(W, W - dx)
        out[ys0:ys1, xs0:xs1] = a[ys0 + dy:ys1 + dy, xs0 + dx:xs1 + dx]
        return out

    inside_ones = np.ones((H, W))
    for yy in range(-2, 3):
        for xx in range(-2, 3):
            if xx == 0 and yy == 0:
                continue
            dx, dy = xx * S, yy * S
            valid = shifted(inside_ones, dx, dy, 0.0)
            cq = shifted(c, dx, dy, 0.0)
            lq, zq, nq = shifted(lum, dx, dy, 0.0), shifted(z, dx, dy, 0.0), shifted(n, dx, dy, 0.0)
            d = np.clip((n * nq).sum(-1), 0.0, 1.0)
            wN = d ** p.phi_normal
            length = np.sqrt(float(xx * xx + yy * yy))
            with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
                wZ = np.where(phiZ * length == 0.0, 0.0, np.abs(z - zq) / (phiZ * length))
                wL = np.abs(lum - lq) / phiL
                w = np.exp(-np.maximum(wL, 0.0) - np.maximum(wZ, 0.0)) * wN * (KW[abs(xx)] * KW[abs(yy)]) * valid
            w = np.nan_to_num(w, nan=0.0)
            sw += w
            acc[..., :3] += w[..., None] * cq[..., :3]
            acc[..., 3] += w * w * cq[..., 3]
    out = acc.copy()
    out[..., :3] = acc[..., :3] / sw[..., None]
    out[..., 3] = acc[..., 3] / (sw * sw)
    bg = z == 1e30
    out[bg] = c[bg]
    return out


def fp64_truth(W, H, frame):
    import torch
    from gpu_util import load_state_from_oracle, npy
    from oracle_lib import OracleFilter, oracle
    from svgf_b200 import SvgfFilter, _lib, synth
    storage = "f32"
    o = OracleFilter(W, H, storage=storage)
    o.Reset()
    for t in range(frame + 1):
        planes = synth.frame_host(W, H, t, storage=storage)
        o.set_inputs(planes)
        if t < frame:
            o.Filter(); o.EndFrame()
    o.TemporalFilter(); o.FilterMoments()
    f = SvgfFilter(W, H, storage=storage)
    P = o.PingPongInx
    g = o.gbuf(P)
    rows = []
    cur = o.FilterBuffer[0].copy()                       # the variance pass's output: the input of level 0 for everyone
    for level in range(5):
        want = np.zeros_like(cur)
        hc = o.RenderBuffer[P].copy()
        assert oracle().svgf_oracle_atrous_level(C.byref(o.params), W, H, o.storage, C.byref(g), cur.ctypes.data, want.ctypes.data,
                                                 hc.ctypes.data, level) == 0
        load_state_from_oracle(f, o)
        f.FilterBuffer[0].copy_(torch.from_numpy(cur))
        f.params.flags = _lib.SVGF_FLAG_NO_STAGED_LEVELS
        res = C.c_void_p()
        gs = f.Framebuffer[P].as_struct()
        assert f.lib.svgf_atrous(f._ctx, C.byref(f.params), C.byref(gs), C.c_void_p(f.FilterBuffer[0].data_ptr()),
                                 C.c_void_p(f.FilterBuffer[1].data_ptr()), C.c_void_p(f.RenderBuffer[P].data_ptr()), level, 1,
                                 C.byref(res), f._stream()) == 0
        got = npy(f.FilterBuffer[1]).astype(np.float64)
        truth = atrous_level_f64(o.params, planes, cur, level)
        floor = np.array([FLOOR_RGB] * 3 + [FLOOR_VAR])
        den = np.maximum(np.abs(truth), floor)
        eg, eo = np.abs(got - truth) / den, np.abs(want.astype(np.float64) - truth) / den
        dgo = np.abs(got - want.astype(np.float64)) / den
        idx = np.argsort(dgo.ravel())[-200:]                 # where kernel and oracle disagree most
        rows.append({"level": level, "kernel_vs_oracle_max_rel": float(dgo.max()), "kernel_vs_truth_max_rel": float(eg.max()),
                     "oracle_vs_truth_max_rel": float(eo.max()), "kernel_vs_truth_frac_above_1e-4": float((eg > TOL).mean()),
                     "oracle_vs_truth_frac_above_1e-4": float((eo > TOL).mean()),
                     "top200_disagreements": {"kernel_vs_truth_median_rel": float(np.median(eg.ravel()[idx])),
                                              "oracle_vs_truth_median_rel": float(np.median(eo.ravel()[idx])),
                                              "kernel_closer_to_truth_fraction": float((eg.ravel()[idx] <= eo.ravel()[idx]).mean())}})
        cur = want                                           # teacher-forced: the next level starts from the oracle's plane
    # the whole cascade: five levels chained in float64 against the oracle's and the kernel's own five levels (one frame,
    # same variance-pass output as input)
    start = o.FilterBuffer[0].copy()
    t = start.astype(np.float64)
    for level in range(5):
        t = atrous_level_f64(o.params, planes, t, level)
    o.WaveletFilter()
    load_state_from_oracle(f, o)
    f.FilterBuffer[0].copy_(torch.from_numpy(start))
    f.params.flags = 0
    f.WaveletFilter()
    got5, want5 = npy(f.FilterBuffer[0]).astype(np.float64), o.FilterBuffer[0].astype(np.float64)
    floor = np.array([FLOOR_RGB] * 3 + [FLOOR_VAR])
    den = np.maximum(np.abs(t), floor)
    eg, eo, dgo = np.abs(got5 - t) / den, np.abs(want5 - t) / den, np.abs(got5 - want5) / den
    idx = np.argsort(dgo.ravel())[-200:]
    cascade = {"kernel_vs_oracle_max_rel": float(dgo.max()), "kernel_vs_oracle_frac_above_1e-4": float((dgo > TOL).mean()),
               "kernel_vs_truth_max_rel": float(eg.max()), "oracle_vs_truth_max_rel": float(eo.max()),
               "kernel_vs_truth_frac_above_1e-4": float((eg > TOL).mean()), "oracle_vs_truth_frac_above_1e-4": float((eo > TOL).mean()),
               "top200_disagreements": {"kernel_vs_truth_median_rel": float(np.median(eg.ravel()[idx])),
                                        "oracle_vs_truth_median_rel": float(np.median(eo.ravel()[idx])),
                                        "kernel_closer_to_truth_fraction": float((eg.ravel()[idx] <= eo.ravel()[idx]).mean())}}
    f.close()
    return {"width": W, "height": H, "frame": frame, "storage": storage,
            "what": "per level from identical inputs: float64 evaluation of the reference formulas = truth; relative error with floors 1e-2 (radiance) / 2.5e-3 (variance)",
            "levels": rows, "five_level_cascade": cascade}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_r02.json"))
    ap.add_argument("--quick", action="store_true", help="small sizes / few frames (smoke run of the tool itself)")
    ap.add_argument("--truth-only", action="store_true", help="only the fp64_truth section, merged into an existing --out file")
    a = ap.parse_args()
    import __graft_entry__ as g
    g.build()
    cases = [(1280, 720, 4, 1), (1920, 1080, 64, 1), (3840, 2160, 16, 1)]
    truth_case = (1920, 1080, 8)
    if a.quick:
        cases, truth_case = [(320, 180, 4, 1), (640, 360, 8, 1)], (320, 180, 3)
    if a.truth_only:
        report = json.load(open(a.out))
        report["fp64_truth"] = fp64_truth(*truth_case)
        print("truth cascade", json.dumps(report["fp64_truth"]["five_level_cascade"]), flush=True)
        json.dump(report, open(a.out, "w"), indent=1)
        return
    report = {"tool": "tools/parity_report.py", "bar": {"history": "bit-exact", "moments": "bit-exact",
                                                         "fp32": "max relative error <= 1e-4 (floors 1e-2 radiance, 2.5e-3 variance)",
                                                         "fp16": "every value within 2 fp16 ulps (or 1e-4 absolute)"},
              "cases": [], "fp64_truth": None}
    for (W, H, n, every) in cases:
        for storage in ("f32", "f16"):
            r = run_case(W, H, n, storage, every)
            report["cases"].append(r)
            print(json.dumps({k: r[k] for k in ("width", "height", "frames", "storage", "seconds")}),
                  "TF", json.dumps(r["teacher_forced"]["result"]), "hist", r["teacher_forced"]["history_mismatches"],
                  "FREE", json.dumps(r["free_running"]["result"]), "hist", r["free_running"]["history_mismatches"], flush=True)
    report["fp64_truth"] = fp64_truth(*truth_case)
    for row in report["fp64_truth"]["levels"]:
        print("truth", json.dumps(row), flush=True)
    print("truth cascade", json.dumps(report["fp64_truth"]["five_level_cascade"]), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(report, open(a.out, "w"), indent=1)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
