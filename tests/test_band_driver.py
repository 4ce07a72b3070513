"""The native band driver (include/svgf_band.h): argument validation without a GPU, and - when the box has two or more
GPUs - the stitched bands against the whole frame, bit for bit, over real NCCL ranks (tools/band_check.py under torchrun)."""
import ctypes as C
import json
import os
import subprocess
import sys

import pytest

from svgf_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_band_create_validates_its_partition_before_touching_a_device():
    lib = _lib.lib()
    h = C.c_void_p()
    uid = (C.c_ubyte * 128)()
    assert lib.svgf_band_create(C.byref(h), 0, 0, 2, 256, 256, 0, None, None) == _lib.SVGF_INVALID_ARG        # world > 1 needs an id
    assert lib.svgf_band_create(C.byref(h), 0, 2, 2, 256, 256, 0, uid, None) == _lib.SVGF_INVALID_ARG         # rank out of range
    assert lib.svgf_band_create(C.byref(h), 0, 0, 4, 256, 100, 0, uid, None) == _lib.SVGF_UNSUPPORTED         # 25-row bands < the 32-row halo
    bad = (C.c_int32 * 3)(0, 200, 100)
    assert lib.svgf_band_create(C.byref(h), 0, 0, 2, 256, 100, 0, uid, bad) == _lib.SVGF_INVALID_ARG          # bounds not increasing to H
    short = (C.c_int32 * 3)(0, 240, 256)
    assert lib.svgf_band_create(C.byref(h), 0, 0, 2, 256, 256, 0, uid, short) == _lib.SVGF_UNSUPPORTED        # a 16-row band
    assert not h.value


@pytest.mark.gpu
@pytest.mark.parametrize("size,levels,storage", [((1920, 1080), 5, "f16"), ((1030, 420), 5, "f32"), ((1280, 720), 4, "f16"),
                                                 ((1280, 720), 3, "f16"), ((1280, 720), 2, "f16")])
def test_stitched_bands_equal_the_whole_frame_over_nccl(size, levels, storage):
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs two or more GPUs (bench.py --gpus N records the same check in its JSON line)")
    W, H = size
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(ROOT, "tools", "band_check.py"), "--width", str(W), "--height", str(H),
                        "--levels", str(levels), "--storage", storage], capture_output=True, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stdout[-1500:] + r.stderr[-3000:]
    res = json.loads(lines[-1])
    assert res["bit_identical"], res
    assert r.returncode == 0
