"""ctypes binding of libvsb200.so (include/vsb200.h).  Fails loudly when the library is missing: there is no
CPU or PyTorch fallback in this package."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VSB200_LIB selects another build of the same sources (e.g. the experimental -DVSB_PDL variant, see __graft_entry__.py)
LIB_PATH = os.environ.get("VSB200_LIB") or os.path.join(_HERE, "libvsb200.so")

EXPORTS = [
    "vsb_last_error", "vsb_version", "vsb_model_create", "vsb_model_set_tensor", "vsb_model_finalize",
    "vsb_model_destroy", "vsb_embed", "vsb_embedder_forward", "vsb_detect", "vsb_jnd_heatmaps", "vsb_embed_host",
    "vsb_detect_host", "vsb_embed_detect_host", "vsb_frames_host_u8", "vsb_launch_count", "vsb_profile_enable", "vsb_profile_read", "vsb_debug_get_tensor", "vsb_debug_conv",
    "vsb_debug_resample_table",
]

FLAG_CLAMP, FLAG_LOWRES_ATTN, FLAG_NO_ATTENUATION, FLAG_RESIZE_NO_AA = 1, 2, 4, 8
VIDEO_MODES = {"repeat": 0, "alternate": 1, "interpolate": 2}


class ModelDesc(C.Structure):
    _fields_ = [
        ("nbits", C.c_int32), ("hidden", C.c_int32), ("img_size", C.c_int32), ("yuv", C.c_int32),
        ("unet_in_ch", C.c_int32), ("unet_out_ch", C.c_int32), ("unet_levels", C.c_int32), ("unet_z", C.c_int32 * 6),
        ("unet_num_blocks", C.c_int32), ("unet_act", C.c_int32), ("unet_norm", C.c_int32), ("unet_last_tanh", C.c_int32),
        ("ext_depths", C.c_int32 * 4), ("ext_dims", C.c_int32 * 4), ("ext_stem_stride", C.c_int32),
        ("jnd_in_ch", C.c_int32), ("jnd_out_ch", C.c_int32),
    ]


class ConvTest(C.Structure):
    _fields_ = [
        ("loader", C.c_int32), ("B", C.c_int32), ("IH", C.c_int32), ("IW", C.c_int32), ("C0", C.c_int32), ("C1", C.c_int32),
        ("R", C.c_int32), ("S", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32), ("pad_mode", C.c_int32),
        ("N", C.c_int32), ("epi", C.c_int32), ("act", C.c_int32), ("rows_per_sample", C.c_int32), ("block_n", C.c_int32),
        ("src0", C.c_void_p), ("src1", C.c_void_p), ("weights", C.c_void_p),
        ("bias", C.c_void_p), ("resid16", C.c_void_p), ("resid32", C.c_void_p), ("a_scale", C.c_void_p),
        ("ln_w", C.c_void_p), ("ln_b", C.c_void_p),
        ("outc_w", C.c_void_p), ("outc_b", C.c_void_p), ("n_out", C.c_int32),
        ("out16", C.c_void_p), ("out32", C.c_void_p), ("delta", C.c_void_p), ("grn_stats", C.c_void_p),
        ("ld0", C.c_int32), ("ldw", C.c_int32), ("ld_out", C.c_int32),
    ]


_lib = None


def lib():
    """Load the shared library (once).  Raises RuntimeError with build instructions if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  videoseal_b200 has no CPU / PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    L.vsb_last_error.restype = C.c_char_p
    L.vsb_version.restype = C.c_int
    L.vsb_model_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(C.c_void_p)]
    L.vsb_model_set_tensor.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int32]
    L.vsb_model_finalize.argtypes = [C.c_void_p, C.c_int32]
    L.vsb_model_destroy.argtypes = [C.c_void_p]
    L.vsb_model_destroy.restype = None
    L.vsb_embed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                            C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32, C.c_void_p]
    L.vsb_embedder_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    L.vsb_detect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.vsb_jnd_heatmaps.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.vsb_embed_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                 C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32]
    L.vsb_detect_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    L.vsb_embed_detect_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32]
    L.vsb_frames_host_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                     C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32]
    L.vsb_launch_count.argtypes = [C.c_int32]
    L.vsb_launch_count.restype = C.c_int64
    L.vsb_profile_enable.argtypes = [C.c_int32]
    L.vsb_profile_read.argtypes = [C.c_char_p, C.c_int64]
    L.vsb_profile_read.restype = C.c_int64
    L.vsb_debug_get_tensor.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    L.vsb_debug_get_tensor.restype = C.c_int64
    L.vsb_debug_conv.argtypes = [C.POINTER(ConvTest), C.c_void_p]
    L.vsb_debug_resample_table.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        msg = lib().vsb_last_error().decode("utf-8", "replace")
        if rc == -2:
            raise NotImplementedError(msg)
        raise RuntimeError(f"libvsb200 error {rc}: {msg}")
