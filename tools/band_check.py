#!/usr/bin/env python
"""Run under torchrun with N >= 2 ranks (one GPU each): filters a short pan sequence in N bands through the native band
driver (include/svgf_band.h) and checks on rank 0 that the stitched bands equal the whole frame filtered on one GPU, bit
for bit - result, colour history, moments and history lengths.  Prints one JSON line; exit code 1 on a mismatch.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/band_check.py

--transport ipc: the peer-memory transport (CUDA IPC handles, flag words, pull kernels; no NCCL on the data path).
--same-gpu: every rank uses cuda:0 and torch.distributed runs on gloo - N processes time-slicing ONE GPU, which is how the
ipc transport (real processes, real IPC mappings, real cross-process flags) is tested on a one-GPU box.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=6)
    ap.add_argument("--levels", type=int, default=5)
    ap.add_argument("--storage", default="f16")
    ap.add_argument("--transport", default="nccl", choices=["nccl", "ipc"])
    ap.add_argument("--same-gpu", action="store_true")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from svgf_b200 import SvgfFilter, synth
    from svgf_b200.band_driver import BandDriver
    from svgf_b200.filter import GBuffer
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    if a.same_gpu:
        local = 0
        assert a.transport == "ipc", "NCCL refuses two ranks on one device"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if a.same_gpu:
        dist.init_process_group(backend="gloo")
    else:
        dist.init_process_group(backend="nccl", device_id=dev)
    on_host = a.same_gpu                       # gloo moves host tensors
    W, H = a.width, a.height
    cdt = torch.float16 if a.storage == "f16" else torch.float32
    bd = BandDriver(W, H, rank, world, dev, storage=a.storage, levels=a.levels, transport=a.transport)
    full_g, full_c = GBuffer(W, H, dev), torch.empty(H, W, 4, dtype=cdt, device=dev)
    whole = SvgfFilter(W, H, device=dev, storage=a.storage) if rank == 0 else None
    if whole is not None:
        whole.SpatialFilterSteps = a.levels
        whole.Reset()
    bd.Reset()
    torch.cuda.synchronize()
    dist.barrier()                             # the first frame's waits are bounded: start together
    bad = {}
    sl = bd.local_rows()
    for t in range(a.frames):
        synth.frame_device(full_g, full_c, t, seed=0)
        P = bd.PingPongInx
        bd.Framebuffer[P].normal.copy_(full_g.normal[sl]); bd.Framebuffer[P].uv.copy_(full_g.uv[sl]); bd.Framebuffer[P].motion.copy_(full_g.motion[sl])
        bd.RenderBuffer[P].copy_(full_c[sl])
        bd.Filter()
        bd.sync()
        planes = {"result": bd.result_band(), "colour_history": bd.RenderBuffer[P][bd.y0 - bd.ly0:bd.y1 - bd.ly0],
                  "moments": bd.MomentsBuffer[P][bd.y0 - bd.ly0:bd.y1 - bd.ly0],
                  "history": bd.HistoryLengthBuffer[bd.y0 - bd.ly0:bd.y1 - bd.ly0]}
        if whole is not None:
            whole.Framebuffer[P].normal.copy_(full_g.normal); whole.Framebuffer[P].uv.copy_(full_g.uv); whole.Framebuffer[P].motion.copy_(full_g.motion)
            whole.RenderBuffer[P].copy_(full_c)
            whole.Filter()
            ref = {"result": whole.FilterBuffer[0], "colour_history": whole.RenderBuffer[P], "moments": whole.MomentsBuffer[P],
                   "history": whole.HistoryLengthBuffer}
        # gather every rank's band rows on rank 0
        bounds = [None] * world
        dist.all_gather_object(bounds, (bd.y0, bd.y1))
        for name, mine in planes.items():
            mine = mine.contiguous()
            if rank == 0:
                n = int((mine.view(torch.uint8) != ref[name][bd.y0:bd.y1].contiguous().view(torch.uint8)).sum())
                for r in range(1, world):
                    y0, y1 = bounds[r]
                    buf = torch.empty((y1 - y0,) + tuple(mine.shape[1:]), dtype=mine.dtype, device="cpu" if on_host else dev)
                    dist.recv(buf, src=r)
                    n += int((buf.to(dev).view(torch.uint8) != ref[name][y0:y1].contiguous().view(torch.uint8)).sum())
                if n:
                    bad[f"frame{t}.{name}"] = n
            else:
                dist.send(mine.cpu() if on_host else mine, dst=0)
        if whole is not None:
            whole.EndFrame()
        bd.EndFrame()
    torch.cuda.synchronize()
    rc = 0
    if rank == 0:
        line = {"check": "band driver vs whole frame", "n_gpus": world, "width": W, "height": H, "frames": a.frames, "levels": a.levels,
                "storage": a.storage, "transport": a.transport, "same_gpu": a.same_gpu, "bit_identical": not bad, "mismatching_bytes": bad}
        print(json.dumps(line), flush=True)
        rc = 0 if not bad else 1
    torch.cuda.synchronize()
    dist.barrier()
    bd.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(rc)


if __name__ == "__main__":
    main()
