"""svgf_b200 — B200-native SVGF denoising filter (temporal reprojection + variance estimation + N-level
edge-avoiding a-trous wavelet) behind the C ABI of include/svgf.h; a drop-in for the filter stage of
jacquespillet/SVGF (reference src/App.cu:469-514 -> src/Filter.cuh:359-624)."""
from . import _lib  # noqa: F401
from ._lib import SvgfError, SvgfParams, default_params  # noqa: F401

__all__ = ["SvgfFilter", "GBuffer", "SvgfError", "SvgfParams", "default_params", "synth", "build"]


def __getattr__(name):
    if name in ("SvgfFilter", "GBuffer"):
        from . import filter as _f
        return getattr(_f, name)
    if name == "synth":
        import importlib
        return importlib.import_module(".synth", __name__)
    raise AttributeError(name)


def build(verbose=False):
    """Compile the in-tree libraries (nvcc, sm_100a).  Used by __graft_entry__.build()."""
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run(["make", "-C", os.path.join(here, "csrc")], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("building libsvgf_b200.so / libsvgf_synth.so failed")
