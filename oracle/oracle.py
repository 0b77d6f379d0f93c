"""
ctypes wrapper around the C oracle (oracle/vc2_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs.  Never imported by the
product package ``vc2_conformance_b200``.
"""

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libvc2oracle.so")

MAX_LEVELS = 16

OK = 0
UNEXPECTED_END_OF_STREAM = 1
INVALID_SLICE_Y_LENGTH = 2
BAD_PARAMS = 3
INT64_ENVELOPE = 4


class Params(C.Structure):
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


def build(force=False):
    """Compile the oracle with gcc (the Makefile recipe) if needed."""
    src = os.path.join(_HERE, "vc2_oracle.c")
    hdr = os.path.join(_HERE, "vc2_oracle.h")
    if (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libvc2oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None

_I64P = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_I32P = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_U16P = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")
_U8P = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    PP = C.POINTER(Params)
    L.vc2o_intlog2.restype = C.c_int32
    L.vc2o_intlog2.argtypes = [C.c_int64]
    for name in ("vc2o_subband_width", "vc2o_subband_height"):
        f = getattr(L, name)
        f.restype = C.c_int32
        f.argtypes = [PP, C.c_int, C.c_int]
    L.vc2o_slice_bytes.restype = C.c_int64
    L.vc2o_slice_bytes.argtypes = [PP, C.c_int, C.c_int]
    L.vc2o_num_subbands.restype = C.c_int32
    L.vc2o_num_subbands.argtypes = [PP]
    L.vc2o_subband_layout.restype = C.c_int64
    L.vc2o_subband_layout.argtypes = [PP, C.c_int, _I32P, _I32P, _I32P, _I32P, _I64P]
    for name in ("vc2o_quant_factor", "vc2o_quant_offset"):
        f = getattr(L, name)
        f.restype = C.c_int64
        f.argtypes = [C.c_int]
    for name in ("vc2o_inverse_quant", "vc2o_forward_quant"):
        f = getattr(L, name)
        f.restype = C.c_int64
        f.argtypes = [C.c_int64, C.c_int]
    L.vc2o_read_sintb_block.restype = C.c_int
    L.vc2o_read_sintb_block.argtypes = [_U8P, C.c_size_t, C.c_uint64, C.c_int64, _I64P, C.c_size_t,
                                        C.POINTER(C.c_uint64)]
    L.vc2o_signed_exp_golomb_length.restype = C.c_int32
    L.vc2o_signed_exp_golomb_length.argtypes = [C.c_int64]
    L.vc2o_transform_data.restype = C.c_int
    L.vc2o_transform_data.argtypes = [PP, _U8P, C.c_size_t, _I64P, _I64P, _I64P, C.c_void_p,
                                      C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_int32)]
    for name in ("vc2o_dc_prediction", "vc2o_apply_dc_prediction"):
        f = getattr(L, name)
        f.restype = None
        f.argtypes = [_I64P, C.c_int32, C.c_int32]
    L.vc2o_filter_bit_shift.restype = C.c_int
    L.vc2o_filter_bit_shift.argtypes = [C.c_int]
    for name in ("vc2o_oned_synthesis", "vc2o_oned_analysis"):
        f = getattr(L, name)
        f.restype = None
        f.argtypes = [_I64P, C.c_int64, C.c_int64, C.c_int]
    L.vc2o_idwt.restype = C.c_int
    L.vc2o_idwt.argtypes = [PP, C.c_int, _I64P, _I64P]
    L.vc2o_decode_component.restype = C.c_int
    L.vc2o_decode_component.argtypes = [PP, C.c_int, _I64P, _U16P]
    L.vc2o_dwt.restype = C.c_int
    L.vc2o_dwt.argtypes = [PP, C.c_int, _I64P, _I64P]
    L.vc2o_encode_component.restype = C.c_int
    L.vc2o_encode_component.argtypes = [PP, C.c_int, _U16P, _I64P]
    L.vc2o_calculate_coeffs_bits.restype = C.c_int64
    L.vc2o_calculate_coeffs_bits.argtypes = [_I64P, C.c_size_t]
    L.vc2o_gather_slice.restype = C.c_int64
    L.vc2o_gather_slice.argtypes = [PP, C.c_int, _I64P, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.vc2o_encode_hq_lossy.restype = C.c_int64
    L.vc2o_encode_hq_lossy.argtypes = [PP, _I64P, _I64P, _I64P, C.c_int64, C.c_int, _U8P, C.c_size_t,
                                       C.c_void_p]
    L.vc2o_encode_hq_fixed.restype = C.c_int64
    L.vc2o_encode_hq_fixed.argtypes = [PP, _I64P, _I64P, _I64P, C.c_int, _U8P, C.c_size_t]
    L.vc2o_encode_ld_lossy.restype = C.c_int64
    L.vc2o_encode_ld_lossy.argtypes = [PP, _I64P, _I64P, _I64P, C.c_int, _U8P, C.c_size_t, C.c_void_p]
    _lib = L
    return L


# ---------------------------------------------------------------------------
# convenience layer
# ---------------------------------------------------------------------------

ORIENT_SLOT = {"LL": 0, "L": 0, "H": 1, "HL": 1, "LH": 2, "HH": 3}


def make_params(
    luma_width,
    luma_height,
    color_diff_width=None,
    color_diff_height=None,
    luma_depth=10,
    color_diff_depth=None,
    wavelet_index=1,
    wavelet_index_ho=None,
    dwt_depth=3,
    dwt_depth_ho=0,
    slices_x=1,
    slices_y=1,
    is_ld=False,
    slice_prefix_bytes=0,
    slice_size_scaler=1,
    slice_bytes_numerator=0,
    slice_bytes_denominator=1,
    quant_matrix=None,
):
    """quant_matrix: {level: {orient: value}} as in the reference's State, or None (all 0)."""
    p = Params()
    p.luma_width = luma_width
    p.luma_height = luma_height
    p.color_diff_width = luma_width if color_diff_width is None else color_diff_width
    p.color_diff_height = luma_height if color_diff_height is None else color_diff_height
    p.luma_depth = luma_depth
    p.color_diff_depth = luma_depth if color_diff_depth is None else color_diff_depth
    p.wavelet_index = wavelet_index
    p.wavelet_index_ho = wavelet_index if wavelet_index_ho is None else wavelet_index_ho
    p.dwt_depth = dwt_depth
    p.dwt_depth_ho = dwt_depth_ho
    p.slices_x = slices_x
    p.slices_y = slices_y
    p.is_ld = 1 if is_ld else 0
    p.slice_prefix_bytes = slice_prefix_bytes
    p.slice_size_scaler = slice_size_scaler
    p.slice_bytes_numerator = slice_bytes_numerator
    p.slice_bytes_denominator = slice_bytes_denominator
    if quant_matrix:
        for level, orients in quant_matrix.items():
            for orient, value in orients.items():
                p.quant_matrix[level][ORIENT_SLOT[orient]] = value
    return p


def layout(p, comp):
    """Returns (level[], orient[], w[], h[], offset[], total) in bitstream order."""
    L = lib()
    nb = L.vc2o_num_subbands(C.byref(p))
    level = np.zeros(nb, np.int32)
    orient = np.zeros(nb, np.int32)
    w = np.zeros(nb, np.int32)
    h = np.zeros(nb, np.int32)
    off = np.zeros(nb, np.int64)
    total = L.vc2o_subband_layout(C.byref(p), comp, level, orient, w, h, off)
    return level, orient, w, h, off, int(total)


def padded_dims(p, comp):
    level, orient, w, h, off, total = layout(p, comp)
    top = p.dwt_depth + p.dwt_depth_ho
    pw = int(w[0]) if top == 0 else 2 * int(w[-1])
    ph = int(h[0]) if p.dwt_depth == 0 else 2 * int(h[-1])
    return pw, ph


def comp_dims(p, comp):
    if comp == 0:
        return p.luma_width, p.luma_height
    return p.color_diff_width, p.color_diff_height


def transform_data(p, data):
    """Returns (status, [cy, c1, c2] int64 flat arrays, qindex[], slice_len[], consumed, bad_slice)."""
    L = lib()
    data = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
    if data.size == 0:
        data = np.zeros(1, np.uint8)
        nbytes = 0
    else:
        nbytes = data.size
    coeffs = [np.zeros(layout(p, c)[5], np.int64) for c in range(3)]
    ns = p.slices_x * p.slices_y
    qindex = np.zeros(ns, np.int32)
    slen = np.zeros(ns, np.int32)
    consumed = C.c_size_t(0)
    bad = C.c_int32(-1)
    st = L.vc2o_transform_data(C.byref(p), data, nbytes, coeffs[0], coeffs[1], coeffs[2],
                               qindex.ctypes.data, slen.ctypes.data, C.byref(consumed), C.byref(bad))
    return st, coeffs, qindex, slen, consumed.value, bad.value


def idwt(p, comp, coeffs):
    pw, ph = padded_dims(p, comp)
    out = np.zeros((ph, pw), np.int64)
    st = lib().vc2o_idwt(C.byref(p), comp, np.ascontiguousarray(coeffs, np.int64), out.reshape(-1))
    assert st == 0, st
    return out


def decode_component(p, comp, coeffs):
    w, h = comp_dims(p, comp)
    out = np.zeros((h, w), np.uint16)
    st = lib().vc2o_decode_component(C.byref(p), comp, np.ascontiguousarray(coeffs, np.int64),
                                     out.reshape(-1))
    assert st == 0, st
    return out


def decode_picture(p, data):
    """transform_data + picture_decode.  Returns (status, [Y, C1, C2] uint16 arrays, consumed)."""
    st, coeffs, qindex, slen, consumed, bad = transform_data(p, data)
    if st != 0:
        return st, None, consumed
    return st, [decode_component(p, c, coeffs[c]) for c in range(3)], consumed


def encode_component(p, comp, pic):
    total = layout(p, comp)[5]
    out = np.zeros(total, np.int64)
    pic = np.ascontiguousarray(pic, np.uint16)
    st = lib().vc2o_encode_component(C.byref(p), comp, pic.reshape(-1), out)
    assert st == 0, st
    return out


def dwt(p, comp, padded_picture):
    total = layout(p, comp)[5]
    out = np.zeros(total, np.int64)
    buf = np.array(padded_picture, dtype=np.int64, order="C", copy=True)
    st = lib().vc2o_dwt(C.byref(p), comp, buf.reshape(-1), out)
    assert st == 0, st
    return out


def dc_prediction(band):
    band = np.array(band, np.int64, order="C", copy=True)
    lib().vc2o_dc_prediction(band.reshape(-1), band.shape[1], band.shape[0])
    return band


def apply_dc_prediction(band):
    band = np.array(band, np.int64, order="C", copy=True)
    lib().vc2o_apply_dc_prediction(band.reshape(-1), band.shape[1], band.shape[0])
    return band


def gather_slice(p, comp, coeffs, sx, sy):
    L = lib()
    n = L.vc2o_gather_slice(C.byref(p), comp, coeffs, sx, sy, None, None)
    out = np.zeros(max(n, 1), np.int64)
    qm = np.zeros(max(n, 1), np.int32)
    L.vc2o_gather_slice(C.byref(p), comp, coeffs, sx, sy, out.ctypes.data, qm.ctypes.data)
    return out[:n], qm[:n]


def safe_lossy_hq_slice_size_scaler(picture_bytes, num_slices):
    """encoder/pictures.py:586 get_safe_lossy_hq_slice_size_scaler."""
    max_slice_bytes = (picture_bytes + num_slices - 1) // num_slices
    return max(1, (max_slice_bytes - 4 + 254) // 255)


def encode_hq_lossy(p, coeffs, picture_bytes, minimum_qindex=0):
    ns = p.slices_x * p.slices_y
    cap = picture_bytes + ns * (p.slice_prefix_bytes + 4 + 3 * p.slice_size_scaler) + 64
    out = np.zeros(cap, np.uint8)
    q = np.zeros(ns, np.int32)
    n = lib().vc2o_encode_hq_lossy(C.byref(p), coeffs[0], coeffs[1], coeffs[2], picture_bytes,
                                   minimum_qindex, out, cap, q.ctypes.data)
    if n < 0:
        raise ValueError("encode_hq_lossy failed (%d)" % n)
    return out[:n].tobytes(), q


def encode_hq_fixed(p, coeffs, qindex=0):
    ns = p.slices_x * p.slices_y
    total = sum(c.size for c in coeffs)
    cap = total * 9 + ns * (p.slice_prefix_bytes + 4 + 3 * p.slice_size_scaler) + 64
    out = np.zeros(cap, np.uint8)
    n = lib().vc2o_encode_hq_fixed(C.byref(p), coeffs[0], coeffs[1], coeffs[2], qindex, out, cap)
    if n < 0:
        raise ValueError("encode_hq_fixed failed (%d)" % n)
    return out[:n].tobytes()


def encode_ld_lossy(p, coeffs, minimum_qindex=0):
    """coeffs: DC prediction NOT yet applied; applied here on copies (LD profile)."""
    ns = p.slices_x * p.slices_y
    coeffs = [c.copy() for c in coeffs]
    for c in range(3):
        level, orient, w, h, off, total = layout(p, c)
        n0 = int(w[0]) * int(h[0])
        band = coeffs[c][:n0].reshape(int(h[0]), int(w[0]))
        coeffs[c][:n0] = apply_dc_prediction(band).reshape(-1)
    cap = (p.slice_bytes_numerator * ns) // p.slice_bytes_denominator + 64
    out = np.zeros(cap, np.uint8)
    q = np.zeros(ns, np.int32)
    n = lib().vc2o_encode_ld_lossy(C.byref(p), coeffs[0], coeffs[1], coeffs[2], minimum_qindex, out,
                                   cap, q.ctypes.data)
    if n < 0:
        raise ValueError("encode_ld_lossy failed (%d)" % n)
    return out[:n].tobytes(), q


def encode_picture(p, pics):
    """picture_encode: [Y, C1, C2] uint16 arrays -> [cy, c1, c2] int64 flat coefficient arrays."""
    return [encode_component(p, c, pics[c]) for c in range(3)]
