"""The .svgfseq container (svgf_b200/seqfile.py): exact round trips of every plane in the C ABI's
texel layouts on the CPU, and (-m gpu) a file of generated inputs filtered by the CUDA path equals the same frames
filtered from memory, and matches the scalar oracle fed from the file."""
import os
import subprocess
import sys

import numpy as np
import pytest

from svgf_b200 import synth
from svgf_b200.seqfile import HEADER, INPUT_PLANES, OUTPUT_PLANES, SeqReader, SeqWriter

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
TOOL = os.path.join(ROOT, "tools", "svgfseq.py")


@pytest.mark.parametrize("storage", ["f16", "f32"])
def test_round_trip_is_bit_exact(storage, tmp_path):
    W, H, N = 37, 21, 3
    rng = np.random.default_rng(0)
    cdt = np.float16 if storage == "f16" else np.float32
    frames = []
    path = tmp_path / "a.svgfseq"
    with SeqWriter(path, W, H, storage, INPUT_PLANES + OUTPUT_PLANES) as wr:
        for t in range(N):
            p = synth.frame_host(W, H, t, storage=storage)
            p["result"] = rng.uniform(size=(H, W, 4)).astype(cdt)
            p["history"] = rng.integers(0, 255, size=(H, W)).astype(np.uint8)
            p["moments"] = rng.uniform(size=(H, W, 2)).astype(cdt)
            p["colour_history"] = rng.uniform(size=(H, W, 4)).astype(cdt)
            p["colour_history"][0, 0, 0] = np.nan                      # raw bits survive, NaN included
            wr.write(p)
            frames.append(p)
    assert os.path.getsize(path) == HEADER.size + N * W * H * (8 + 8 + 16 + (3 * 4 + 2) * np.dtype(cdt).itemsize + 1)
    with SeqReader(path) as rd:
        assert (rd.W, rd.H, rd.storage, rd.frames) == (W, H, storage, N)
        assert rd.planes == ["normal", "uv", "motion", "colour", "result", "history", "moments", "colour_history"]
        for t, got in enumerate(rd):
            for k, v in frames[t].items():
                assert got[k].dtype == v.dtype and np.array_equal(got[k].view(np.uint8), v.view(np.uint8)), (t, k)
        with pytest.raises(IndexError):
            rd.read(N)


def test_rejects_foreign_and_truncated_files(tmp_path):
    bad = tmp_path / "bad.svgfseq"
    bad.write_bytes(b"not a sequence")
    with pytest.raises(ValueError):
        SeqReader(bad)
    good = tmp_path / "g.svgfseq"
    with SeqWriter(good, 8, 4, "f16") as wr:
        wr.write(synth.frame_host(8, 4, 0))
    data = good.read_bytes()
    (tmp_path / "t.svgfseq").write_bytes(data[:-5])
    with pytest.raises(ValueError):
        SeqReader(tmp_path / "t.svgfseq")
    with pytest.raises(ValueError):
        SeqWriter(tmp_path / "x.svgfseq", 8, 4, "f16", planes=("normal", "albedo"))
    with SeqWriter(tmp_path / "y.svgfseq", 8, 4, "f16") as wr, pytest.raises(ValueError):
        wr.write({k: v[:2] for k, v in synth.frame_host(8, 4, 0).items()})


def test_cli_gen_and_compare(tmp_path):
    a, b = str(tmp_path / "a.svgfseq"), str(tmp_path / "b.svgfseq")
    for out in (a, b):
        subprocess.run([sys.executable, TOOL, "gen", out, "--size", "48x32", "--frames", "2"], check=True, capture_output=True)
    r = subprocess.run([sys.executable, TOOL, "compare", a, b], check=True, capture_output=True, text=True)
    assert "identical" in r.stdout


@pytest.mark.gpu
def test_filtering_a_file_equals_filtering_from_memory_and_matches_the_oracle(tmp_path):
    import torch
    from common import assert_close
    from gpu_util import upload_inputs
    from oracle_lib import OracleFilter
    from svgf_b200 import SvgfFilter
    W, H, N = 320, 180, 4
    src, dst = str(tmp_path / "in.svgfseq"), str(tmp_path / "out.svgfseq")
    subprocess.run([sys.executable, TOOL, "gen", src, "--size", f"{W}x{H}", "--frames", str(N)], check=True, capture_output=True)
    subprocess.run([sys.executable, TOOL, "filter", src, dst], check=True, capture_output=True)
    f = SvgfFilter(W, H)
    o = OracleFilter(W, H)
    f.Reset(); o.Reset()
    with SeqReader(dst) as rd:
        assert rd.frames == N
        for t, rec in enumerate(rd):
            planes = synth.frame_host(W, H, t)
            for k in INPUT_PLANES:
                assert np.array_equal(rec[k].view(np.uint8), planes[k].view(np.uint8))
            upload_inputs(f, planes)
            f.Filter()
            torch.cuda.synchronize()
            P = f.PingPongInx
            assert np.array_equal(rec["result"].view(np.uint8), f.FilterBuffer[0].cpu().numpy().view(np.uint8))
            assert np.array_equal(rec["history"], f.HistoryLengthBuffer.cpu().numpy())
            assert np.array_equal(rec["colour_history"].view(np.uint8), f.RenderBuffer[P].cpu().numpy().view(np.uint8))
            if t == 0:                                 # first frame: no accumulated ill-conditioning, per-stage bar applies
                o.set_inputs({k: rec[k] for k in INPUT_PLANES})
                o.Filter()
                assert np.array_equal(rec["history"], o.HistoryLengthBuffer)
                assert np.array_equal(rec["moments"].view(np.uint8), o.MomentsBuffer[o.PingPongInx].view(np.uint8))
                assert_close(rec["result"], o.FilterBuffer[0], "f16", "frame 0 from the file vs oracle", max_flips=0.03)
            f.EndFrame()
