#!/usr/bin/env python
"""Small frames through every kernel family and mode, meant to run under `compute-sanitizer --tool memcheck`:
odd and even widths (per-pixel / packed kernels), fused levels, variance prefilter, bilinear reprojection, both storages."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from gpu_util import upload_inputs  # noqa: E402
from svgf_b200 import SvgfFilter, _lib, synth  # noqa: E402

n = 0
for storage in ("f16", "f32"):
    for (W, H) in ((200, 70), (131, 37)):
        for flags, pre, rep in ((0, 0, 0), (_lib.SVGF_FLAG_FUSE_LEVELS_01, 0, 0), (0, 1, 1), (_lib.SVGF_FLAG_BASIC_KERNELS, 1, 1)):
            f = SvgfFilter(W, H, storage=storage)
            f.params.flags, f.params.variance_prefilter, f.params.reproj_mode = flags, pre, rep
            f.Reset()
            for t in range(3):
                upload_inputs(f, synth.frame_host(W, H, t, storage=storage))
                f.Filter()
                f.EndFrame()
            torch.cuda.synchronize()
            assert torch.isfinite(f.FilterBuffer[0].float()).all()
            n += 1
# round 2: the staged run at sizes that exercise every lattice row phase and the ragged tiles, with and without the
# uniform-normal shortcut / dependent launch / staged levels; TAA; texture-object G-buffers; the band driver's in-process
# group (two-range boundary launches, halo and state copies into the aprons)
for storage in ("f16", "f32"):
    for (W, H) in ((392, 214), (258, 99)):
        for flags in (0, 8, 32, 256):
            f = SvgfFilter(W, H, storage=storage)
            f.params.flags = flags
            f.Reset()
            for t in range(3):
                upload_inputs(f, synth.frame_host(W, H, t, storage=storage))
                f.Filter()
                f.TAA()
                f.EndFrame()
            torch.cuda.synchronize()
            assert torch.isfinite(f.FilterBuffer[0].float()).all()
            n += 1

from svgf_b200.band_driver import BandGroup  # noqa: E402
from svgf_b200.filter import GBuffer  # noqa: E402

dev = torch.device("cuda", 0)
for storage, world, (W, H) in (("f16", 3, (512, 420)), ("f32", 2, (258, 200))):
    cdt = torch.float16 if storage == "f16" else torch.float32
    grp = BandGroup(W, H, [dev] * world, storage=storage, levels=5)
    grp.Reset()
    full_g, full_c = GBuffer(W, H, dev), torch.empty(H, W, 4, dtype=cdt, device=dev)
    for t in range(3):
        synth.frame_device(full_g, full_c, t, seed=0)
        P = grp.bands[0].PingPongInx
        for b in grp.bands:
            sl = b.local_rows()
            b.Framebuffer[P].normal.copy_(full_g.normal[sl]); b.Framebuffer[P].uv.copy_(full_g.uv[sl]); b.Framebuffer[P].motion.copy_(full_g.motion[sl])
            b.RenderBuffer[P].copy_(full_c[sl])
        grp.Filter()
        grp.sync()
        grp.EndFrame()
    torch.cuda.synchronize()
    for b in grp.bands:
        assert torch.isfinite(b.result_band().float()).all()
    grp.close()
    n += 1
print(f"{n} configurations ran")
