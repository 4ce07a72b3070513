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


def test_the_other_transports_validate_before_touching_a_device_too():
    lib = _lib.lib()
    h = C.c_void_p()
    assert lib.svgf_band_create_ipc(C.byref(h), 0, 3, 2, 256, 256, 0, None) == _lib.SVGF_INVALID_ARG          # rank out of range
    assert lib.svgf_band_create_ipc(C.byref(h), 0, 0, 4, 256, 100, 0, None) == _lib.SVGF_UNSUPPORTED          # bands < the apron
    assert not h.value
    blob = (C.c_ubyte * 640)()
    assert lib.svgf_band_ipc_export(None, blob) == _lib.SVGF_INVALID_ARG
    assert lib.svgf_band_ipc_connect(None, blob, blob) == _lib.SVGF_INVALID_ARG
    hs = (C.c_void_p * 2)()
    dv = (C.c_int32 * 2)(0, 0)
    assert lib.svgf_band_create_group(hs, dv, 0, 256, 256, 0, None) == _lib.SVGF_INVALID_ARG                  # no bands
    assert lib.svgf_band_create_group(None, dv, 2, 256, 256, 0, None) == _lib.SVGF_INVALID_ARG
    assert lib.svgf_band_create_group(hs, dv, 2, 256, 40, 0, None) == _lib.SVGF_UNSUPPORTED                   # 20-row bands
    assert not hs[0] and not hs[1]
    assert lib.svgf_band_group_frame(None, 2, None, None, None, None) == _lib.SVGF_INVALID_ARG


def _plan(rank, world, lo, hi, rows, levels):
    steps = (_lib.SvgfBandStep * 32)()
    n = _lib.lib().svgf_band_plan(rank, world, lo, hi, rows, levels, steps, 32)
    assert n >= 0
    return [(s.kind, s.level, s.yblock0, s.nyblocks, s.rows, s.yblock1, s.nyblocks1) for s in steps[:n]]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("levels", [2, 3, 4, 5])
@pytest.mark.parametrize("height", [1080, 4320, 333])
def test_band_plan_covers_what_each_level_must_produce(world, levels, height):
    """svgf_band_plan is the schedule svgf_band_frame executes.  For every rank: each level's launches are disjoint row-block
    ranges whose union covers the band plus the rows the non-exchanging levels above it need; a level that feeds an
    exchange has produced the rows its neighbours need BEFORE the exchange is posted; levels >= 3 wait for their halo."""
    LAUNCH, EXCHANGE, WAIT = 0, 1, 2
    base = height // world
    if world > 1 and base < 32:
        pytest.skip("bands shorter than the apron")
    for rank in range(world):
        y0 = rank * base
        y1 = height if rank == world - 1 else y0 + base
        ly0, ly1 = (max(0, y0 - 32), min(height, y1 + 32)) if world > 1 else (y0, y1)
        lo, hi, rows = y0 - ly0, y1 - ly0, ly1 - ly0
        plan = _plan(rank, world, lo, hi, rows, levels)
        assert [lv for _, lv, *_ in plan] == sorted(lv for _, lv, *_ in plan), "levels out of order"
        for l in range(1, levels):
            B = 12 << l
            reach = sum(2 << j for j in range(l + 1, min(levels, 3)))
            need = set(range(max(0, lo - reach), min(rows, hi + reach)))
            mine = [s for s in plan if s[1] == l]
            covered, before_exchange = set(), set()
            seen_exchange = seen_wait = False
            reads_apron = set()          # rows whose 5-tap column reaches a neighbour's rows
            if world > 1 and l >= 3:
                if rank > 0:
                    reads_apron |= set(range(0, lo + (2 << l)))
                if rank + 1 < world:
                    reads_apron |= set(range(hi - (2 << l), rows + B))
            for kind, _, yb0, nyb, r, yb1, nyb1 in mine:
                if kind == WAIT:
                    assert not seen_wait
                    seen_wait = True
                    assert r == 2 << l
                if kind == LAUNCH:
                    blk = set(range(yb0 * B, (yb0 + nyb) * B)) | set(range(yb1 * B, (yb1 + nyb1) * B))
                    assert len(blk) == (nyb + nyb1) * B, "the two ranges of a launch overlap"
                    assert not (blk & covered), f"rank {rank} level {l}: row blocks launched twice"
                    covered |= blk
                    if not seen_wait:
                        assert not (blk & reads_apron), f"rank {rank} level {l}: apron rows read before the halo has arrived"
                    if not seen_exchange:
                        before_exchange |= blk
                elif kind == EXCHANGE:
                    seen_exchange = True
                    assert r == 2 << (l + 1)
                    halo = set()
                    if rank > 0:
                        halo |= set(range(lo, lo + r))
                    if rank + 1 < world:
                        halo |= set(range(hi - r, hi))
                    assert halo <= before_exchange, f"rank {rank} level {l}: exchange posted before its rows were produced"
            assert need <= covered, f"rank {rank} level {l}: rows {sorted(need - covered)[:4]}... never produced"
            assert seen_exchange == (world > 1 and l + 1 < levels and l + 1 >= 3)
            assert seen_wait == (world > 1 and l >= 3)


def test_band_plan_rejects_what_the_driver_does_not_do():
    steps = (_lib.SvgfBandStep * 32)()
    lib = _lib.lib()
    assert lib.svgf_band_plan(0, 2, 0, 500, 532, 6, steps, 32) == -1      # more than 5 levels
    assert lib.svgf_band_plan(0, 2, 0, 500, 532, 1, steps, 32) == -1      # a single level has no staged run
    assert lib.svgf_band_plan(2, 2, 0, 500, 532, 5, steps, 32) == -1
    assert lib.svgf_band_plan(0, 2, 0, 500, 532, 5, steps, 2) == -1       # list too short


def _band_check(n, size, levels, storage, extra, port):
    W, H = size
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tools", "band_check.py"), "--width", str(W), "--height", str(H),
                        "--levels", str(levels), "--storage", storage] + extra, capture_output=True, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stdout[-1500:] + r.stderr[-3000:]
    res = json.loads(lines[-1])
    assert res["bit_identical"], res
    assert r.returncode == 0
    return res


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["nccl", "ipc"])
@pytest.mark.parametrize("size,levels,storage", [((1920, 1080), 5, "f16"), ((1030, 420), 5, "f32"), ((1280, 720), 4, "f16"),
                                                 ((1280, 720), 3, "f16"), ((1280, 720), 2, "f16")])
def test_stitched_bands_equal_the_whole_frame_over_real_ranks(size, levels, storage, transport):
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs two or more GPUs (bench.py --gpus N records the same check in its JSON line)")
    _band_check(n, size, levels, storage, ["--transport", transport], 29517)


@pytest.mark.gpu
@pytest.mark.parametrize("n,size,levels,storage", [(2, (1280, 720), 5, "f16"), (3, (1920, 1080), 5, "f16"), (4, (1030, 420), 5, "f32"),
                                                   (3, (1280, 720), 3, "f16"), (8, (1920, 1080), 5, "f16")])
def test_peer_memory_transport_with_every_rank_on_one_gpu(n, size, levels, storage):
    """n PROCESSES time-slicing cuda:0 (torch.distributed on gloo): real CUDA IPC mappings, real cross-process flag words and
    pull kernels - the peer-memory transport end to end on a one-GPU box.  Stitched bands == whole frame, bit for bit."""
    res = _band_check(n, size, levels, storage, ["--transport", "ipc", "--same-gpu", "--frames", "4"], 29519)
    assert res["transport"] == "ipc" and res["same_gpu"] and res["n_gpus"] == n
