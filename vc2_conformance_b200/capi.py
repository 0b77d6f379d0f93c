"""
ctypes binding of libvc2b200.so (include/vc2_b200.h).  There is NO fallback: if the CUDA
library has not been built, importing this module's `lib()` raises.
"""

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# VC2B_LIB: developer override (kernel variants built side by side for A/B timing)
LIB_PATH = os.environ.get("VC2B_LIB") or os.path.join(HERE, "_lib", "libvc2b200.so")

MAX_LEVELS = 16
MAX_BANDS = 32

E_BAD_PARAMS = -1
E_UNSUPPORTED = -2
E_SCRATCH_TOO_SMALL = -3

ST_OK = 0
ST_UNEXPECTED_END_OF_STREAM = 1
ST_INVALID_SLICE_Y_LENGTH = 2
ST_INT32_ENVELOPE = 4
ST_COEFF16_RANGE = 5


class Params(C.Structure):
    """struct vc2b_params (include/vc2_b200.h)."""

    _fields_ = [
        ("luma_width", C.c_int32),
        ("luma_height", C.c_int32),
        ("color_diff_width", C.c_int32),
        ("color_diff_height", C.c_int32),
        ("luma_depth", C.c_int32),
        ("color_diff_depth", C.c_int32),
        ("wavelet_index", C.c_int32),
        ("wavelet_index_ho", C.c_int32),
        ("dwt_depth", C.c_int32),
        ("dwt_depth_ho", C.c_int32),
        ("slices_x", C.c_int32),
        ("slices_y", C.c_int32),
        ("is_ld", C.c_int32),
        ("slice_prefix_bytes", C.c_int32),
        ("slice_size_scaler", C.c_int32),
        ("slice_bytes_numerator", C.c_int64),
        ("slice_bytes_denominator", C.c_int64),
        ("quant_matrix", (C.c_int32 * 4) * MAX_LEVELS),
    ]


class Unit(C.Structure):
    """struct vc2b_unit (include/vc2_b200.h): one data unit found by vc2b_index_stream."""

    _fields_ = [
        ("offset", C.c_int64),
        ("next_offset", C.c_int64),
        ("data_offset", C.c_int64),
        ("parse_code", C.c_int32),
        ("flags", C.c_int32),
        ("picture_number", C.c_uint32),
        ("fragment_data_length", C.c_int32),
        ("fragment_slice_count", C.c_int32),
        ("fragment_x_offset", C.c_int32),
        ("fragment_y_offset", C.c_int32),
        ("major_version", C.c_int32),
        ("minor_version", C.c_int32),
        ("profile", C.c_int32),
        ("level", C.c_int32),
        ("base_video_format", C.c_int32),
        ("picture_coding_mode", C.c_int32),
        ("params", Params),
    ]


IDX_NEEDS_BASE_FORMAT = 1
IDX_NEEDS_DEFAULT_QUANT_MATRIX = 2
IDX_NO_SEQUENCE_HEADER = 4
IDX_TRUNCATED = 8
IDX_UNSUPPORTED = 16


class LibraryMissing(ImportError):
    pass


class Vc2bError(RuntimeError):
    def __init__(self, fn, code):
        self.code = code
        names = {-1: "bad parameters", -2: "unsupported parameters", -3: "scratch buffer too small"}
        what = names.get(code, "CUDA error %d" % code)
        RuntimeError.__init__(self, "%s failed: %s" % (fn, what))


_lib = None

_PP = C.POINTER(Params)
_vp = C.c_void_p
_i = C.c_int
_i64 = C.c_int64

# name -> (restype, argtypes); every symbol include/vc2_b200.h declares
SIGNATURES = {
    "vc2b_version": (C.c_char_p, []),
    "vc2b_launch_count": (_i64, []),
    "vc2b_check_params": (_i, [_PP]),
    "vc2b_num_subbands": (_i, [_PP]),
    "vc2b_subband_info": (_i, [_PP, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "vc2b_component_coeff_count": (_i64, [_PP, _i]),
    "vc2b_picture_coeff_count": (_i64, [_PP]),
    "vc2b_padded_dims": (_i, [_PP, _i, _vp, _vp]),
    "vc2b_picture_sample_count": (_i64, [_PP]),
    "vc2b_transform_scratch_bytes": (_i64, [_PP]),
    "vc2b_idwt_max_safe_coeff": (_i64, [_PP]),
    "vc2b_hq_slice_offsets": (_i, [_PP, _vp, _vp, _i, _vp, _vp, _vp]),
    "vc2b_unpack": (_i, [_PP, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "vc2b_unpack16": (_i, [_PP, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vc2b_coeff16_supported": (_i, [_PP]),
    "vc2b_set_unpack_kernel": (_i, [_i]),
    "vc2b_dc_prediction": (_i, [_PP, _vp, _i, _i, _vp]),
    "vc2b_idwt": (_i, [_PP, _vp, _i, _vp, _vp, _i64, _vp]),
    "vc2b_idwt16": (_i, [_PP, _vp, _vp, _i, _vp, _vp, _i64, _vp, _vp]),
    "vc2b_idwt_raw": (_i, [_PP, _i, _vp, _vp, _vp, _i64, _vp]),
    "vc2b_dwt": (_i, [_PP, _vp, _i, _vp, _vp, _i64, _vp]),
    "vc2b_dwt_raw": (_i, [_PP, _i, _vp, _vp, _vp, _i64, _vp]),
    "vc2b_quantize_to_fit": (_i, [_PP, _vp, _i, _i64, _i, _vp, _vp, _vp]),
    "vc2b_slice_bits": (_i, [_PP, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "vc2b_quantize_slices": (_i, [_PP, _vp, _i, _vp, _i, _vp, _vp]),
    "vc2b_quantize_to_fit_lists": (_i, [_vp, _vp, _vp, _i, _i64, _i64, _i, _vp, _vp, _vp]),
    "vc2b_hq_lossless_layout": (_i, [_PP, _vp, _i, _vp, _vp, _vp]),
    "vc2b_pack_slices_lossless": (_i, [_PP, _vp, _i, _vp, _vp, _vp, _vp, _i64, _vp]),
    "vc2b_packed_bytes": (_i64, [_PP, _i64]),
    "vc2b_pack_slices": (_i, [_PP, _vp, _i, _i64, _vp, _vp, _vp, _i64, _vp]),
    "vc2b_inverse_quant": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "vc2b_forward_quant": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "vc2b_clip_offset": (_i, [_vp, _i64, _i, _i, _vp]),
    "vc2b_raw_picture_bytes": (_i64, [_PP]),
    "vc2b_pictures_to_raw": (_i, [_PP, _vp, _i, _vp, _i64, _vp]),
    "vc2b_raw_to_pictures": (_i, [_PP, _vp, _i64, _i, _vp, _vp]),
    "vc2b_picture_compare": (_i, [_PP, _vp, _vp, _i, _vp, _vp]),
    "vc2b_index_stream": (_i64, [_vp, _i64, _vp, _i64]),
}


def lib():
    """Loads libvc2b200.so; raises LibraryMissing (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            "%s not found: build it with `python -m vc2_conformance_b200._build` "
            "(there is no CPU fallback)" % LIB_PATH
        )
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def check(fn, code):
    if code != 0:
        raise Vc2bError(fn, code)
