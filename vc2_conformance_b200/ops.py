"""
numpy-in / numpy-out wrappers of single C-ABI calls (host -> device -> kernel -> host).
Used by the reference-named functions in ``pseudocode.py`` and by the parity tests; the
throughput path (``device.BatchCodec``) keeps data resident instead.
"""

import ctypes as C

import numpy as np

from . import capi
from .device import _ptr, _torch
from .device import geometry_for as Geometry


def _dev(a, dtype):
    torch = _torch()
    a = np.array(a, dtype=dtype, order="C", copy=True)
    if dtype == np.uint16:
        return torch.from_numpy(a.view(np.int16)).cuda().view(torch.uint16)
    return torch.from_numpy(a).cuda()


def _stream():
    torch = _torch()
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def idwt_component(p, comp, coeffs):
    """idwt (pseudocode/picture_decoding.py:88) of one component: flat int32 coefficients in
    subband order -> int32 array [padded_h, padded_w]."""
    torch = _torch()
    g = Geometry(p)
    assert coeffs.size == g.comp_coeffs[comp]
    d_in = _dev(coeffs, np.int32)
    pw, ph = g.padded[comp]
    d_out = torch.empty(pw * ph, dtype=torch.int32, device="cuda")
    scratch = torch.empty(max(g.scratch_bytes, 16), dtype=torch.uint8, device="cuda")
    capi.check("vc2b_idwt_raw",
               capi.lib().vc2b_idwt_raw(C.byref(p), comp, _ptr(d_in), _ptr(d_out), _ptr(scratch),
                                        scratch.numel(), _stream()))
    return d_out.cpu().numpy().reshape(ph, pw)


def dwt_component(p, comp, padded_picture):
    """dwt (pseudocode/picture_encoding.py:93) of one component: int32 [padded_h, padded_w] ->
    flat int32 coefficients in subband order."""
    torch = _torch()
    g = Geometry(p)
    pw, ph = g.padded[comp]
    assert padded_picture.shape == (ph, pw)
    d_in = _dev(padded_picture, np.int32)
    d_out = torch.empty(g.comp_coeffs[comp], dtype=torch.int32, device="cuda")
    scratch = torch.empty(max(g.scratch_bytes, 16), dtype=torch.uint8, device="cuda")
    capi.check("vc2b_dwt_raw",
               capi.lib().vc2b_dwt_raw(C.byref(p), comp, _ptr(d_in), _ptr(d_out), _ptr(scratch),
                                       scratch.numel(), _stream()))
    return d_out.cpu().numpy()


def inverse_quant(values, quant_indices):
    torch = _torch()
    v = _dev(values, np.int32)
    q = _dev(np.broadcast_to(quant_indices, np.shape(values)), np.int32)
    out = torch.empty_like(v)
    capi.check("vc2b_inverse_quant",
               capi.lib().vc2b_inverse_quant(_ptr(v), _ptr(q), _ptr(out), v.numel(), _stream()))
    return out.cpu().numpy()


def forward_quant(values, quant_indices):
    torch = _torch()
    v = _dev(values, np.int32)
    q = _dev(np.broadcast_to(quant_indices, np.shape(values)), np.int32)
    out = torch.empty_like(v)
    capi.check("vc2b_forward_quant",
               capi.lib().vc2b_forward_quant(_ptr(v), _ptr(q), _ptr(out), v.numel(), _stream()))
    return out.cpu().numpy()


CLIP = 1
ADD_OFFSET = 2
REMOVE_OFFSET = 4


def clip_offset(values, depth, mode):
    v = _dev(values, np.int32)
    capi.check("vc2b_clip_offset", capi.lib().vc2b_clip_offset(_ptr(v), v.numel(), int(depth), int(mode), _stream()))
    return v.cpu().numpy()


def dc_prediction(p, picture_coeffs, inverse=False):
    """dc_prediction (decoder/transform_data_syntax.py:63) or its inverse apply_dc_prediction
    (encoder/pictures.py:202) on the level-0 bands of one picture's flat coefficients."""
    v = _dev(picture_coeffs, np.int32)
    capi.check("vc2b_dc_prediction",
               capi.lib().vc2b_dc_prediction(C.byref(p), _ptr(v), 1, 1 if inverse else 0, _stream()))
    return v.cpu().numpy()


INT64_MAX = (1 << 63) - 1


def quantize_to_fit_lists(coeff_sets, quant_matrix_sets, target_size, align_bits=1, minimum_qindex=0):
    """quantize_to_fit (encoder/pictures.py:295) on explicit lists: returns (qindex, [bits of each set at
    that index], [quantised int32 array per set]).  ``target_size=None``: no search, evaluate at
    ``minimum_qindex`` (quantize_coeffs :251 / calculate_coeffs_bits :217)."""
    torch = _torch()
    nsets = len(coeff_sets)
    offs = np.zeros(nsets + 1, np.int32)
    for i, cs in enumerate(coeff_sets):
        offs[i + 1] = offs[i] + len(cs)
    n = int(offs[nsets])
    values = np.concatenate([np.asarray(cs, np.int32).reshape(-1) for cs in coeff_sets]) if n else np.zeros(0, np.int32)
    qms = np.concatenate([np.asarray(q, np.int32).reshape(-1) for q in quant_matrix_sets]) if n else np.zeros(0, np.int32)
    assert qms.size == values.size
    d_v = _dev(np.append(values, 0), np.int32)
    d_q = _dev(np.append(qms, 0), np.int32)
    d_res = torch.zeros(16, dtype=torch.int32, device="cuda")
    d_out = torch.zeros(n + 1, dtype=torch.int32, device="cuda")
    target = INT64_MAX if target_size is None else int(target_size)
    capi.check("vc2b_quantize_to_fit_lists",
               capi.lib().vc2b_quantize_to_fit_lists(_ptr(d_v), _ptr(d_q), offs.ctypes.data_as(C.c_void_p), nsets,
                                                     target, int(align_bits), int(minimum_qindex), _ptr(d_res),
                                                     _ptr(d_out), _stream()))
    res = d_res.cpu().numpy()
    out = d_out.cpu().numpy()
    return int(res[0]), [int(b) for b in res[1:1 + nsets]], [out[offs[i]:offs[i + 1]] for i in range(nsets)]
