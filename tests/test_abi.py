"""The C-ABI boundary without a GPU: the library loads, exports exactly what include/svgf.h declares, and
fails loudly (no fallback) when there is no device."""
import ctypes as C
import os
import re
import subprocess

import pytest

from svgf_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions(headers=("svgf.h", "svgf_band.h")):
    names = set()
    for h in headers:
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(svgf_[a-z_0-9]+)\s*\(", src))
    return sorted(names)


def test_header_declares_the_north_star_entry_points():
    names = header_functions()
    for n in ("svgf_temporal", "svgf_variance", "svgf_atrous", "svgf_frame", "svgf_create", "svgf_destroy", "svgf_reset"):
        assert n in names


def test_library_exports_every_declared_symbol():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (svgf_[a-z_0-9]+)", out))
    declared = set(header_functions())
    assert declared <= exported, f"declared but not exported: {declared - exported}"
    assert exported <= declared, f"exported but not declared in include/svgf.h: {exported - declared}"


def test_python_binding_table_matches_header():
    assert sorted(n for n, _, _ in _lib.ABI) == header_functions()
    lib = _lib.lib()
    assert lib.svgf_abi_version() == _lib.SVGF_ABI_VERSION


def test_default_params_are_the_reference_defaults():
    p = _lib.default_params()
    # reference src/App.h:109-114 (SpatialFilterSteps is 3 there; BASELINE.json measures 5)
    assert (p.history_cap, p.atrous_iterations) == (24, 5)
    assert (p.depth_threshold, p.normal_threshold, p.phi_colour, p.phi_normal) == pytest.approx((0.8, 0.9, 10.0, 128.0))
    assert (p.phi_depth, p.alpha_min, p.moments_alpha_min) == (1.0, 0.0, 0.0)
    assert (p.mesh_id_mode, p.reproj_mode, p.variance_prefilter, p.flags) == (0, 0, 0, 0)


def test_struct_layouts_match_the_header():
    # sizes a C compiler gives the structs in include/svgf.h
    code = ('#include "svgf.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n", sizeof(svgf_params), '
            'sizeof(svgf_gbuffer), sizeof(svgf_frame_buffers));return 0;}')
    exe = "/tmp/svgf_sizes"
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=code, text=True, check=True)
    sizes = [int(v) for v in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(_lib.SvgfParams), C.sizeof(_lib.SvgfGBuffer), C.sizeof(_lib.SvgfFrameBuffers)]


def test_no_device_means_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ctx = C.c_void_p()
    st = _lib.lib().svgf_create(C.byref(ctx), 0, 64, 64, 0)
    assert st == _lib.SVGF_CUDA_ERROR and not ctx.value
    from svgf_b200 import SvgfFilter
    with pytest.raises(RuntimeError):
        SvgfFilter(64, 64)


def test_product_never_touches_the_oracle():
    # the oracle is test infrastructure: nothing under svgf_b200/ may reference it
    for dirpath, _, files in os.walk(os.path.join(ROOT, "svgf_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                for needle in ("svgf_oracle", "oracle/", "oracle_lib", "oracle_py", "refkernels", "import oracle"):
                    assert needle not in text, f"{fn} references the oracle ({needle})"
