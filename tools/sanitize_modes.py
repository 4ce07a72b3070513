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
print(f"{n} configurations ran")
