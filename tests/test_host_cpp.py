"""The C++ host mirror (svgf_b200/host/svgf_host.hpp, driven by svgf_host_demo) against the Python mirror: both sit on
the same C ABI, so on the same procedurally generated sequence their final buffers must be bit-identical."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "svgf_b200", "svgf_host_demo")


def fnv1a(b):
    h = 1469598103934665603
    for c in bytes(b):
        h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


def test_host_demo_is_built_and_links_the_c_abi():
    assert os.path.exists(DEMO), "svgf_host_demo was not built (make -C svgf_b200/csrc)"
    out = subprocess.run(["ldd", DEMO], capture_output=True, text=True).stdout
    assert "libsvgf_b200.so" in out and "libsvgf_synth.so" in out


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["--fused", "--staged"])
def test_cpp_host_matches_python_host(mode):
    import torch
    from svgf_b200 import SvgfFilter, synth
    W, H, frames = 320, 184, 6
    r = subprocess.run([DEMO, str(W), str(H), str(frames), mode], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    got = json.loads(r.stdout.strip().splitlines()[-1])
    f = SvgfFilter(W, H, storage="f16")
    f.Reset()
    for t in range(frames):
        P = f.PingPongInx
        synth.frame_device(f.Framebuffer[P], f.RenderBuffer[P], t, seed=0)
        if mode == "--fused":
            f.Filter()
        else:
            f.TemporalFilter(); f.FilterMoments(); f.WaveletFilter()
        f.EndFrame()
    torch.cuda.synchronize()
    res = f.FilterBuffer[0].cpu().numpy().view(np.uint8).tobytes()
    hist = f.HistoryLengthBuffer.cpu().numpy().tobytes()
    assert got["history_fnv1a"] == fnv1a(hist)
    assert got["result_fnv1a"] == fnv1a(res)
