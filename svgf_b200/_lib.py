"""ctypes bindings of the C ABI in include/svgf.h (libsvgf_b200.so) and of the input generator
(libsvgf_synth.so).  The libraries are built in-tree by ``svgf_b200.build``; loading fails loudly when
they are missing — there is no Python or CPU fallback for any stage."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsvgf_b200.so")
SYNTH_PATH = os.path.join(_HERE, "libsvgf_synth.so")

SVGF_OK, SVGF_INVALID_ARG, SVGF_UNSUPPORTED, SVGF_CUDA_ERROR = 0, 1, 2, 3
SVGF_STORE_F16, SVGF_STORE_F32 = 0, 1
SVGF_MESH_ID_INTENDED, SVGF_MESH_ID_REFERENCE_VACUOUS = 0, 1
SVGF_DEPTH_TEST_ABSOLUTE, SVGF_DEPTH_TEST_RELATIVE = 0, 1
SVGF_FLAG_NO_GUIDE_CACHE, SVGF_FLAG_NO_LEVEL_FUSION, SVGF_FLAG_BASIC_KERNELS, SVGF_FLAG_NO_UNIFORM_TILES = 1, 2, 4, 8
SVGF_FLAG_FUSE_LEVELS_01 = 16
SVGF_FLAG_NO_STAGED_LEVELS, SVGF_FLAG_ATROUS_BULK, SVGF_FLAG_ATROUS_STREAM, SVGF_FLAG_NO_DEPENDENT_LAUNCH = 32, 64, 128, 256
SVGF_FLAG_BAND_NO_EXCHANGE = 512
# svgf_dispatch_family (svgf_last_dispatch)
SVGF_FAMILY_BASIC, SVGF_FAMILY_PACKED, SVGF_FAMILY_PACKED_STAGED, SVGF_FAMILY_LATTICE = 1, 2, 3, 4
SVGF_FAMILY_BULK, SVGF_FAMILY_STREAM, SVGF_FAMILY_FUSED01 = 5, 6, 7
SVGF_PITCH_TEXTURE = C.c_size_t(-1).value
SVGF_ABI_VERSION = 2


class SvgfParams(C.Structure):
    """struct svgf_params (include/svgf.h); defaults = reference src/App.h:109-114."""
    _fields_ = [
        ("history_cap", C.c_int32), ("depth_threshold", C.c_float), ("normal_threshold", C.c_float),
        ("phi_colour", C.c_float), ("phi_normal", C.c_float), ("atrous_iterations", C.c_int32),
        ("phi_depth", C.c_float), ("alpha_min", C.c_float), ("moments_alpha_min", C.c_float),
        ("mesh_id_mode", C.c_int32), ("reproj_mode", C.c_int32), ("variance_prefilter", C.c_int32),
        ("flags", C.c_uint32), ("depth_test_mode", C.c_int32),
    ]


class SvgfGBuffer(C.Structure):
    """struct svgf_gbuffer: the reference's cudaFramebuffer (src/App.h:41-44) as pitch-linear planes."""
    _fields_ = [
        ("position_id", C.c_void_p), ("position_pitch", C.c_size_t),
        ("normal_mat", C.c_void_p), ("normal_pitch", C.c_size_t),
        ("uv_inst", C.c_void_p), ("uv_pitch", C.c_size_t),
        ("motion_depth", C.c_void_p), ("motion_pitch", C.c_size_t),
    ]


class SvgfFrameBuffers(C.Structure):
    """struct svgf_frame_buffers: RenderBuffer/MomentsBuffer/FilterBuffer[2] + HistoryLengthBuffer (src/App.h:136-139)."""
    _fields_ = [
        ("render", C.c_void_p * 2), ("moments", C.c_void_p * 2), ("filter", C.c_void_p * 2),
        ("history", C.c_void_p), ("ping_pong", C.c_int32),
    ]


class SynthCfg(C.Structure):
    """struct svgf_synth_cfg (csrc/synth_scene.h)."""
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("seed", C.c_uint32), ("frame", C.c_int32),
        ("pan_px", C.c_float), ("vert_px", C.c_float), ("half_period", C.c_int32), ("storage", C.c_int32),
    ]


# every symbol include/svgf.h declares: (name, restype, argtypes)
ABI = [
    ("svgf_abi_version", C.c_int, []),
    ("svgf_status_string", C.c_char_p, [C.c_int]),
    ("svgf_default_params", None, [C.POINTER(SvgfParams)]),
    ("svgf_create", C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int]),
    ("svgf_destroy", None, [C.c_void_p]),
    ("svgf_reset", C.c_int, [C.c_void_p, C.POINTER(SvgfFrameBuffers), C.c_void_p]),
    ("svgf_temporal", C.c_int, [C.c_void_p, C.POINTER(SvgfParams), C.POINTER(SvgfGBuffer), C.POINTER(SvgfGBuffer),
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("svgf_variance", C.c_int, [C.c_void_p, C.POINTER(SvgfParams), C.POINTER(SvgfGBuffer), C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p]),
    ("svgf_atrous", C.c_int, [C.c_void_p, C.POINTER(SvgfParams), C.POINTER(SvgfGBuffer), C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_void_p]),
    ("svgf_frame", C.c_int, [C.c_void_p, C.POINTER(SvgfParams), C.POINTER(SvgfGBuffer * 2), C.POINTER(SvgfFrameBuffers),
                             C.c_void_p]),
    ("svgf_taa", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("svgf_demodulate", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("svgf_remodulate", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("svgf_invalidate_guide", None, [C.c_void_p]),
    ("svgf_profile_begin", C.c_int, [C.c_void_p]),
    ("svgf_profile_end", C.c_int, [C.c_void_p, C.POINTER(C.c_double * 3), C.POINTER(C.c_int)]),
    ("svgf_last_cuda_error", C.c_int, [C.c_void_p]),
    ("svgf_launch_count", C.c_uint64, [C.c_void_p]),
    ("svgf_last_dispatch", C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_int]),
    ("svgf_frame_host", C.c_int, [C.c_void_p, C.POINTER(SvgfParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
]

class SvgfBandStep(C.Structure):
    """struct svgf_band_step (include/svgf_band.h)."""
    _fields_ = [("kind", C.c_int32), ("level", C.c_int32), ("yblock0", C.c_int32), ("nyblocks", C.c_int32), ("rows", C.c_int32),
                ("yblock1", C.c_int32), ("nyblocks1", C.c_int32)]


# include/svgf_band.h
ABI += [
    ("svgf_band_plan", C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(SvgfBandStep), C.c_int]),
    ("svgf_band_unique_id", C.c_int, [C.c_void_p]),
    ("svgf_band_create", C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                   C.POINTER(C.c_int32)]),
    ("svgf_band_destroy", None, [C.c_void_p]),
    ("svgf_band_rows", None, [C.c_void_p, C.POINTER(C.c_int32 * 4)]),
    ("svgf_band_reset", C.c_int, [C.c_void_p, C.POINTER(SvgfFrameBuffers), C.c_void_p]),
    ("svgf_band_frame", C.c_int, [C.c_void_p, C.POINTER(SvgfParams), C.POINTER(SvgfGBuffer * 2), C.POINTER(SvgfFrameBuffers),
                                  C.c_void_p]),
    ("svgf_band_sync", C.c_int, [C.c_void_p, C.c_void_p]),
    ("svgf_band_create_ipc", C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32)]),
    ("svgf_band_ipc_export", C.c_int, [C.c_void_p, C.c_void_p]),
    ("svgf_band_ipc_connect", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("svgf_band_create_group", C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(C.c_int32)]),
    ("svgf_band_group_frame", C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(SvgfParams), C.POINTER(SvgfGBuffer),
                                         C.POINTER(SvgfFrameBuffers), C.POINTER(C.c_void_p)]),
    ("svgf_band_launch_count", C.c_uint64, [C.c_void_p]),
    ("svgf_band_last_error", C.c_int, [C.c_void_p]),
]

SYNTH_ABI = [
    ("svgf_synth_frame_host", C.c_int, [C.POINTER(SynthCfg)] + [C.c_void_p] * 5 + [C.c_int]),
    ("svgf_synth_rows_host", C.c_int, [C.POINTER(SynthCfg), C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int]),
    ("svgf_synth_frame_device", C.c_int, [C.POINTER(SynthCfg)] + [C.c_void_p] * 6),
    ("svgf_synth_texgbuf_create", C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    ("svgf_synth_texgbuf_destroy", None, [C.c_void_p]),
    ("svgf_synth_texgbuf_upload", C.c_int, [C.c_void_p] * 5),
    ("svgf_synth_texgbuf_objects", None, [C.c_void_p, C.POINTER(C.c_ulonglong * 3)]),
]

_lib = None
_synth = None


def _bind(path, table):
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C svgf_b200/csrc`). svgf_b200 has no fallback implementation.")
    lib = C.CDLL(path)
    for name, res, args in table:
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


def lib():
    """The loaded C-ABI library (libsvgf_b200.so)."""
    global _lib
    if _lib is None:
        _lib = _bind(LIB_PATH, ABI)
        if _lib.svgf_abi_version() != SVGF_ABI_VERSION:
            raise RuntimeError("libsvgf_b200.so ABI version mismatch")
    return _lib


def synth_lib():
    global _synth
    if _synth is None:
        _synth = _bind(SYNTH_PATH, SYNTH_ABI)
    return _synth


class SvgfError(RuntimeError):
    def __init__(self, status, where, cuda_error=0):
        self.status = status
        self.cuda_error = cuda_error
        name = lib().svgf_status_string(status).decode()
        super().__init__(f"{where}: {name}" + (f" (cudaError {cuda_error})" if cuda_error else ""))


def default_params():
    p = SvgfParams()
    lib().svgf_default_params(C.byref(p))
    return p
