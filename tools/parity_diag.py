#!/usr/bin/env python
"""Stage-by-stage error diagnosis (GPU box): isolated error of every stage (inputs injected from the oracle)
versus the chained error (GPU stages fed by GPU outputs) on the pan sequence, FP32 storage."""
import argparse, ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def stats(got, want, floor):
    d = np.abs(got.astype(np.float64) - want.astype(np.float64)) / np.maximum(np.abs(want.astype(np.float64)), floor)
    return f"max {d.max():.2e} p99.99 {np.quantile(d, 0.9999):.2e} p99 {np.quantile(d, 0.99):.2e} n>1e-4 {(d > 1e-4).sum()}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=640); ap.add_argument("--height", type=int, default=360)
    ap.add_argument("--frames", type=int, default=8); ap.add_argument("--storage", default="f32")
    a = ap.parse_args()
    import torch
    from gpu_util import npy, upload_inputs, load_state_from_oracle
    from oracle_lib import OracleFilter, oracle
    from svgf_b200 import SvgfFilter, synth
    W, H = a.width, a.height
    o = OracleFilter(W, H, storage=a.storage)
    f = SvgfFilter(W, H, storage=a.storage)      # chained
    fi = SvgfFilter(W, H, storage=a.storage)     # isolated: state injected from the oracle before every stage
    o.Reset(); f.Reset()
    for t in range(a.frames):
        planes = synth.frame_host(W, H, t, storage=a.storage)
        o.set_inputs(planes); upload_inputs(f, planes)
        P = o.PingPongInx
        # temporal
        load_state_from_oracle(fi, o)
        if t == 0:
            fi.Reset(); upload_inputs(fi, planes)
        o.TemporalFilter(); f.TemporalFilter(); fi.TemporalFilter()
        print(f"frame {t} temporal: hist mismatch {int((npy(f.HistoryLengthBuffer) != o.HistoryLengthBuffer).sum())} colour-bits-equal "
              f"{np.array_equal(npy(fi.RenderBuffer[P]), o.RenderBuffer[P])} chained-equal {np.array_equal(npy(f.RenderBuffer[P]), o.RenderBuffer[P])}")
        # variance
        load_state_from_oracle(fi, o)
        o.FilterMoments(); f.FilterMoments(); fi.FilterMoments()
        w = o.FilterBuffer[0]
        print(f"   variance isolated rgb {stats(npy(fi.FilterBuffer[0])[..., :3], w[..., :3], 1e-2)} | var {stats(npy(fi.FilterBuffer[0])[..., 3], w[..., 3], 2.5e-3)}")
        print(f"   variance chained  rgb {stats(npy(f.FilterBuffer[0])[..., :3], w[..., :3], 1e-2)} | var {stats(npy(f.FilterBuffer[0])[..., 3], w[..., 3], 2.5e-3)}")
        # levels
        g = o.gbuf(P)
        cur_o = o.FilterBuffer[0].copy()
        cur_g = f.FilterBuffer[0].clone()
        for level in range(5):
            out_o = np.zeros_like(cur_o)
            hc = o.RenderBuffer[P]
            oracle().svgf_oracle_atrous_level(C.byref(o.params), W, H, o.storage, C.byref(g), cur_o.ctypes.data, out_o.ctypes.data,
                                              hc.ctypes.data, level)
            # isolated
            fi.FilterBuffer[0].copy_(torch.from_numpy(cur_o))
            res = C.c_void_p()
            gs = fi.Framebuffer[P].as_struct()
            fi.lib.svgf_atrous(fi._ctx, C.byref(fi.params), C.byref(gs), C.c_void_p(fi.FilterBuffer[0].data_ptr()),
                               C.c_void_p(fi.FilterBuffer[1].data_ptr()), C.c_void_p(fi.RenderBuffer[P].data_ptr()), level, 1, C.byref(res), fi._stream())
            iso = npy(fi.FilterBuffer[1])
            # chained
            f.FilterBuffer[0].copy_(cur_g)
            gs2 = f.Framebuffer[P].as_struct()
            f.lib.svgf_atrous(f._ctx, C.byref(f.params), C.byref(gs2), C.c_void_p(f.FilterBuffer[0].data_ptr()),
                              C.c_void_p(f.FilterBuffer[1].data_ptr()), C.c_void_p(f.RenderBuffer[P].data_ptr()), level, 1, C.byref(res), f._stream())
            cur_g = f.FilterBuffer[1].clone()
            ch = npy(cur_g)
            print(f"   level {level} isolated rgb {stats(iso[..., :3], out_o[..., :3], 1e-2)} | var {stats(iso[..., 3], out_o[..., 3], 2.5e-3)}")
            print(f"   level {level} chained  rgb {stats(ch[..., :3], out_o[..., :3], 1e-2)} | var {stats(ch[..., 3], out_o[..., 3], 2.5e-3)}")
            if level == 4:
                d = np.abs(ch[..., :3].astype(np.float64) - out_o[..., :3]) / np.maximum(np.abs(out_o[..., :3]), 1e-2)
                y, x, c = np.unravel_index(d.argmax(), d.shape)
                print(f"      worst pixel ({x},{y}) ch{c}: got {ch[y, x]} want {out_o[y, x]} hist {o.HistoryLengthBuffer[y, x]} depth {planes['motion'][y, x]}")
            cur_o = out_o
        o.FilterBuffer[0][...] = cur_o
        f.FilterBuffer[0].copy_(cur_g)
        o.EndFrame(); f.EndFrame()


if __name__ == "__main__":
    main()
