#!/usr/bin/env python
"""Work with .svgfseq files (svgf_b200/seqfile.py):

  tools/svgfseq.py gen out.svgfseq --size 640x360 --frames 8 [--storage f16|f32] [--seed 0]
        procedural camera-pan inputs (the benchmark's generator, host twin)
  tools/svgfseq.py filter in.svgfseq out.svgfseq [--levels 5] [--reproj 0|1] [--prefilter 0|1]
        run the inputs through libsvgf_b200.so on cuda:0 (svgf_frame) and write inputs + outputs
  tools/svgfseq.py compare a.svgfseq b.svgfseq
        per-plane comparison of two files (bit-equal / max abs difference)
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from svgf_b200.seqfile import INPUT_PLANES, OUTPUT_PLANES, SeqReader, SeqWriter  # noqa: E402


def gen(a):
    from svgf_b200 import synth
    w, h = (int(v) for v in a.size.lower().split("x"))
    with SeqWriter(a.out, w, h, a.storage, INPUT_PLANES) as wr:
        for t in range(a.frames):
            wr.write(synth.frame_host(w, h, t, seed=a.seed, storage=a.storage))
    print(f"{a.out}: {a.frames} frames of {w}x{h}, {a.storage}")


def filt(a):
    import torch
    from svgf_b200 import SvgfFilter
    with SeqReader(a.inp) as rd, SeqWriter(a.out, rd.W, rd.H, rd.storage, INPUT_PLANES + OUTPUT_PLANES) as wr:
        f = SvgfFilter(rd.W, rd.H, storage=rd.storage)
        f.SpatialFilterSteps = a.levels
        f.params.reproj_mode, f.params.variance_prefilter = a.reproj, a.prefilter
        f.Reset()
        for planes in rd:
            P = f.PingPongInx
            f.Framebuffer[P].normal.copy_(torch.from_numpy(planes["normal"].view(np.int16).copy()))
            f.Framebuffer[P].uv.copy_(torch.from_numpy(planes["uv"].view(np.int16).copy()))
            f.Framebuffer[P].motion.copy_(torch.from_numpy(planes["motion"].copy()))
            f.RenderBuffer[P].copy_(torch.from_numpy(planes["colour"].copy()))
            f.Filter()
            torch.cuda.synchronize()
            out = dict(planes)
            out.update(result=f.FilterBuffer[0].cpu().numpy(), history=f.HistoryLengthBuffer.cpu().numpy(),
                       moments=f.MomentsBuffer[P].cpu().numpy(), colour_history=f.RenderBuffer[P].cpu().numpy())
            wr.write(out)
            f.EndFrame()
    print(f"{a.out}: filtered {a.inp}")


def compare(a):
    worst = 0.0
    with SeqReader(a.a) as ra, SeqReader(a.b) as rb:
        if (ra.W, ra.H, ra.storage) != (rb.W, rb.H, rb.storage):
            sys.exit("geometry or storage differ")
        common = [p for p in ra.planes if p in rb.planes]
        for t in range(min(ra.frames, rb.frames)):
            fa, fb = ra.read(t), rb.read(t)
            for p in common:
                if np.array_equal(fa[p].view(np.uint8), fb[p].view(np.uint8)):
                    continue
                d = float(np.nanmax(np.abs(fa[p].astype(np.float64) - fb[p].astype(np.float64))))
                worst = max(worst, d)
                print(f"frame {t} {p}: max abs difference {d:.3e}")
    print("identical" if worst == 0.0 else f"worst difference {worst:.3e}")
    return worst


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    g = sub.add_parser("gen"); g.add_argument("out"); g.add_argument("--size", default="640x360"); g.add_argument("--frames", type=int, default=8)
    g.add_argument("--storage", default="f16", choices=["f16", "f32"]); g.add_argument("--seed", type=int, default=0); g.set_defaults(fn=gen)
    f = sub.add_parser("filter"); f.add_argument("inp"); f.add_argument("out"); f.add_argument("--levels", type=int, default=5)
    f.add_argument("--reproj", type=int, default=0); f.add_argument("--prefilter", type=int, default=0); f.set_defaults(fn=filt)
    c = sub.add_parser("compare"); c.add_argument("a"); c.add_argument("b"); c.set_defaults(fn=compare)
    a = ap.parse_args()
    a.fn(a)


if __name__ == "__main__":
    main()
