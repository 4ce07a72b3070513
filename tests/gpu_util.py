"""Helpers for the -m gpu tests: move numpy planes in and out of a SvgfFilter."""
import numpy as np
import torch


def upload_inputs(f, planes, slot=None):
    P = f.PingPongInx if slot is None else slot
    f.Framebuffer[P].normal.copy_(torch.from_numpy(planes["normal"].view(np.int16)))
    f.Framebuffer[P].uv.copy_(torch.from_numpy(planes["uv"].view(np.int16)))
    f.Framebuffer[P].motion.copy_(torch.from_numpy(planes["motion"]))
    if "colour" in planes:
        f.RenderBuffer[P].copy_(torch.from_numpy(planes["colour"]))


def load_state_from_oracle(f, of):
    """Copy an OracleFilter's full state into a SvgfFilter (both G-buffers, all planes, ping-pong)."""
    for k in range(2):
        f.Framebuffer[k].normal.copy_(torch.from_numpy(of.normal[k].view(np.int16)))
        f.Framebuffer[k].uv.copy_(torch.from_numpy(of.uv[k].view(np.int16)))
        f.Framebuffer[k].motion.copy_(torch.from_numpy(of.motion[k]))
        f.RenderBuffer[k].copy_(torch.from_numpy(of.RenderBuffer[k]))
        f.MomentsBuffer[k].copy_(torch.from_numpy(of.MomentsBuffer[k]))
        f.FilterBuffer[k].copy_(torch.from_numpy(of.FilterBuffer[k]))
    f.HistoryLengthBuffer.copy_(torch.from_numpy(of.HistoryLengthBuffer))
    f.PingPongInx = of.PingPongInx
    f.invalidate_guide()


def npy(t):
    return t.detach().cpu().numpy()
