#!/usr/bin/env python
"""Error statistics of the CUDA path against the scalar oracle over a pan sequence (runs on the GPU box).
Writes a JSON report; used to choose / document the tolerances in tests/common.py and DESIGN.md."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=360)
    ap.add_argument("--frames", type=int, default=24)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_report.json"))
    a = ap.parse_args()
    import torch
    from common import half_ulp_diff
    from gpu_util import npy, upload_inputs
    from oracle_lib import OracleFilter
    from svgf_b200 import SvgfFilter, synth
    W, H = a.width, a.height
    report = {}
    for storage in ("f32", "f16"):
        f = SvgfFilter(W, H, storage=storage)
        o = OracleFilter(W, H, storage=storage)
        f.Reset(); o.Reset()
        rows = []
        for t in range(a.frames):
            planes = synth.frame_host(W, H, t, storage=storage)
            o.set_inputs(planes); upload_inputs(f, planes)
            f.Filter(); o.Filter()
            P = o.PingPongInx
            row = {"frame": t, "history_mismatch": int((npy(f.HistoryLengthBuffer) != o.HistoryLengthBuffer).sum())}
            for name, got, want in (("result", f.FilterBuffer[0], o.FilterBuffer[0]), ("hist_colour", f.RenderBuffer[P], o.RenderBuffer[P]),
                                    ("moments", f.MomentsBuffer[P], o.MomentsBuffer[P])):
                g, w = npy(got), want
                g64, w64 = g.astype(np.float64), w.astype(np.float64)
                d = np.abs(g64 - w64)
                ent = {"max_abs": float(d.max())}
                chans = {"rgb": slice(0, 3), "var": slice(3, 4)} if g.shape[-1] == 4 else {"m": slice(0, 2)}
                for cn, sl in chans.items():
                    for floor in (1e-1, 1e-2, 2.5e-3, 1e-3, 1e-4):
                        ent[f"{cn}_rel_floor{floor:g}"] = float((d[..., sl] / np.maximum(np.abs(w64[..., sl]), floor)).max())
                    ent[f"{cn}_max_abs"] = float(d[..., sl].max())
                if storage == "f16":
                    u = half_ulp_diff(g, w)
                    ent["ulp_hist"] = np.bincount(np.minimum(u, 8).ravel(), minlength=9).tolist()
                row[name] = ent
            rows.append(row)
            f.EndFrame(); o.EndFrame()
        report[storage] = rows
        last = rows[-1]
        print(storage, "last frame:", json.dumps(last)[:1500])
        worst = {}
        for r in rows:
            for name in ("result", "hist_colour", "moments"):
                for k, v in r[name].items():
                    if isinstance(v, float):
                        worst[f"{name}.{k}"] = max(worst.get(f"{name}.{k}", 0.0), v)
        print(storage, "WORST over sequence:", json.dumps(worst))
        print(storage, "history mismatches:", sum(r["history_mismatch"] for r in rows))
        report[storage + "_worst"] = worst
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(report, open(a.out, "w"))


if __name__ == "__main__":
    main()
