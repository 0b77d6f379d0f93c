"""
Reference-named functions of the picture path, computed by the CUDA kernels.

Same names, argument meaning and in-place behaviour as the functions star-exported by
``vc2_conformance.pseudocode`` (/root/reference/vc2_conformance/pseudocode/__init__.py:96-107)
and the decoder's ``transform_data`` / ``dc_prediction``
(decoder/transform_data_syntax.py:63,116), operating on the reference's containers: ``state``
dictionaries, 2-D nested lists of ``int``, ``{level: {orient: [[int]]}}`` coefficient dictionaries
and ``{"Y","C1","C2","pic_num"}`` pictures.  Converting nested lists to device arrays costs far
more than the kernels; the throughput API is ``device.BatchCodec`` / ``device.SequenceDecoder``.
This module is the strict drop-in layer (see ``install.py``).

Integers are int32 on the device.  Inputs that do not fit raise ``EnvelopeError`` instead of
silently differing from the reference's unbounded arithmetic.
"""

import numpy as np

from . import capi, ops
from .params import Geometry, make_params, params_from_state

__all__ = [
    "EnvelopeError",
    "quant_factor", "quant_offset", "inverse_quant", "forward_quant",
    "idwt", "dwt", "h_synthesis", "vh_synthesis", "h_analysis", "vh_analysis",
    "idwt_pad_removal", "dwt_pad_addition",
    "offset_picture", "offset_component", "clip_picture", "clip_component",
    "remove_offset_picture", "remove_offset_component",
    "picture_decode", "picture_encode", "inverse_wavelet_transform", "forward_wavelet_transform",
    "dc_prediction", "apply_dc_prediction", "transform_data", "DeviceTransform",
    "initialize_fragment_state", "fragment_slice",
]

INT32_MAX = (1 << 31) - 1


class EnvelopeError(OverflowError):
    """A value does not fit the int32 envelope of the CUDA path."""


def _to_i32(values, what):
    a = np.array(values, dtype=object)
    if a.size and (max(int(v) for v in a.reshape(-1)) > INT32_MAX or min(int(v) for v in a.reshape(-1)) < -INT32_MAX):
        raise EnvelopeError("%s outside the int32 envelope" % what)
    return a.astype(np.int32)


# ---- quantisation (pseudocode/quantization.py:18-62) -----------------------------------------

def quant_factor(index):
    """(13.3.2) Host-side integer formula, identical to pseudocode/quantization.py:40."""
    base = 2 ** (index // 4)
    if (index % 4) == 0:
        return 4 * base
    elif (index % 4) == 1:
        return ((503829 * base) + 52958) // 105917
    elif (index % 4) == 2:
        return ((665857 * base) + 58854) // 117708
    else:
        return ((440253 * base) + 32722) // 65444


def quant_offset(index):
    """(13.3.2) pseudocode/quantization.py:54."""
    if index == 0:
        return 1
    elif index == 1:
        return 2
    return (quant_factor(index) + 1) // 2


def inverse_quant(quantized_coeff, quant_index):
    """(13.3.1) pseudocode/quantization.py:18; scalars or array-likes."""
    scalar = np.isscalar(quantized_coeff)
    v = _to_i32(np.atleast_1d(quantized_coeff), "quantized_coeff")
    qi = np.broadcast_to(np.asarray(quant_index, dtype=np.int64), v.shape)
    # the result must fit as well: (m * qf + qo + 2) // 4
    mags = np.abs(v.astype(np.int64)).reshape(-1)
    for m, q in zip(mags, qi.reshape(-1)):
        if m and (int(m) * quant_factor(int(q)) + quant_offset(int(q)) + 2) // 4 > INT32_MAX:
            raise EnvelopeError("inverse_quant result outside the int32 envelope")
    out = ops.inverse_quant(v, qi.astype(np.int32))
    return int(out[0]) if scalar else out


def forward_quant(coeff, quant_index):
    """(13.3.1) pseudocode/quantization.py:30; scalars or array-likes."""
    scalar = np.isscalar(coeff)
    v = _to_i32(np.atleast_1d(coeff), "coeff")
    out = ops.forward_quant(v, np.broadcast_to(np.asarray(quant_index, dtype=np.int32), v.shape))
    return int(out[0]) if scalar else out


# ---- containers -----------------------------------------------------------------------------

def _height(a):
    return len(a)


def _width(a):
    return len(a[0]) if len(a) else 0


def _transform_params(state, w, h, dwt_depth=None, dwt_depth_ho=None):
    """Params for a single component whose PADDED size is w x h."""
    return make_params(
        w, h,
        wavelet_index=state["wavelet_index"],
        wavelet_index_ho=state.get("wavelet_index_ho", state["wavelet_index"]),
        dwt_depth=state["dwt_depth"] if dwt_depth is None else dwt_depth,
        dwt_depth_ho=state.get("dwt_depth_ho", 0) if dwt_depth_ho is None else dwt_depth_ho,
        luma_depth=16,
    )


def _coeff_dict_to_flat(g, comp, coeff_data):
    out = np.zeros(g.comp_coeffs[comp], np.int32)
    for b, band in enumerate(g.bands[comp]):
        arr = _to_i32(coeff_data[band["level"]][g.orient_name(b)], "coefficient")
        assert arr.shape == (band["h"], band["w"]), (arr.shape, band)
        out[band["off"]: band["off"] + band["w"] * band["h"]] = arr.reshape(-1)
    return out


def _flat_to_coeff_dict(g, comp, flat):
    d = {}
    for b, band in enumerate(g.bands[comp]):
        a = flat[band["off"]: band["off"] + band["w"] * band["h"]].reshape(band["h"], band["w"])
        d.setdefault(band["level"], {})[g.orient_name(b)] = a.tolist()
    return d


# ---- synthesis (pseudocode/picture_decoding.py:88-184) --------------------------------------------

def idwt(state, coeff_data):
    """(15.4.1) Inverse DWT of one component; returns the padded picture as nested lists."""
    ho = state.get("dwt_depth_ho", 0)
    dc = coeff_data[0]["LL" if ho == 0 else "L"]
    w = _width(dc) << (state["dwt_depth"] + ho)
    h = _height(dc) << state["dwt_depth"]
    p = _transform_params(state, w, h)
    g = Geometry(p)
    return ops.idwt_component(p, 0, _coeff_dict_to_flat(g, 0, coeff_data)).tolist()


def h_synthesis(state, L_data, H_data):
    """(15.4.2) One horizontal-only synthesis level."""
    p = _transform_params(state, 2 * _width(L_data), _height(L_data), dwt_depth=0, dwt_depth_ho=1)
    flat = np.concatenate([_to_i32(L_data, "L").reshape(-1), _to_i32(H_data, "H").reshape(-1)])
    return ops.idwt_component(p, 0, flat).tolist()


def vh_synthesis(state, LL_data, HL_data, LH_data, HH_data):
    """(15.4.3) One 2-D synthesis level."""
    p = _transform_params(state, 2 * _width(LL_data), 2 * _height(LL_data), dwt_depth=1, dwt_depth_ho=0)
    flat = np.concatenate([_to_i32(a, "subband").reshape(-1) for a in (LL_data, HL_data, LH_data, HH_data)])
    return ops.idwt_component(p, 0, flat).tolist()


def idwt_pad_removal(state, pic, c):
    """(15.4.5) Crops in place (host-side list surgery, as pseudocode/picture_decoding.py:282)."""
    width = state["luma_width"] if c == "Y" else state["color_diff_width"]
    height = state["luma_height"] if c == "Y" else state["color_diff_height"]
    del pic[height:]
    for row in pic:
        del row[width:]


# ---- analysis (pseudocode/picture_encoding.py:93-199) ----------------------------------------------

def dwt(state, picture):
    """Forward DWT of one (already padded) component; returns {level: {orient: [[int]]}}."""
    p = _transform_params(state, _width(picture), _height(picture))
    g = Geometry(p)
    flat = ops.dwt_component(p, 0, _to_i32(picture, "sample"))
    return _flat_to_coeff_dict(g, 0, flat)


def h_analysis(state, data):
    p = _transform_params(state, _width(data), _height(data), dwt_depth=0, dwt_depth_ho=1)
    flat = ops.dwt_component(p, 0, _to_i32(data, "sample"))
    h, w2 = _height(data), _width(data) // 2
    return (flat[: h * w2].reshape(h, w2).tolist(), flat[h * w2:].reshape(h, w2).tolist())


def vh_analysis(state, data):
    p = _transform_params(state, _width(data), _height(data), dwt_depth=1, dwt_depth_ho=0)
    flat = ops.dwt_component(p, 0, _to_i32(data, "sample"))
    h2, w2 = _height(data) // 2, _width(data) // 2
    n = h2 * w2
    return tuple(flat[i * n:(i + 1) * n].reshape(h2, w2).tolist() for i in range(4))


def dwt_pad_addition(state, pic, c):
    """Edge-replicating pad in place (pseudocode/picture_encoding.py:235)."""
    top = state["dwt_depth"] + state.get("dwt_depth_ho", 0)
    w = state["luma_width"] if c == "Y" else state["color_diff_width"]
    h = state["luma_height"] if c == "Y" else state["color_diff_height"]
    sw = 1 << top
    sh = 1 << state["dwt_depth"]
    width = sw * ((w + sw - 1) // sw)
    height = sh * ((h + sh - 1) // sh)
    for row in pic:
        value = row[-1]
        while len(row) < width:
            row.append(value)
    while len(pic) < height:
        pic.append(pic[-1][:])


# ---- offset / clip (pseudocode/picture_decoding.py:301-349, picture_encoding.py:264) -------------------

def _component_depth(state, c):
    return state["luma_depth"] if c == "Y" else state["color_diff_depth"]


def _apply_in_place(comp_data, depth, mode, what):
    if not len(comp_data):
        return
    out = ops.clip_offset(_to_i32(comp_data, what), depth, mode).tolist()
    for row, new in zip(comp_data, out):
        row[:] = new


def offset_component(state, comp_data, c):
    _apply_in_place(comp_data, _component_depth(state, c), ops.ADD_OFFSET, "sample")


def offset_picture(state, current_picture):
    for c in ["Y", "C1", "C2"]:
        offset_component(state, current_picture[c], c)


def clip_component(state, comp_data, c):
    _apply_in_place(comp_data, _component_depth(state, c), ops.CLIP, "sample")


def clip_picture(state, current_picture):
    for c in ["Y", "C1", "C2"]:
        clip_component(state, current_picture[c], c)


def remove_offset_component(state, comp_data, c):
    _apply_in_place(comp_data, _component_depth(state, c), ops.REMOVE_OFFSET, "sample")


def remove_offset_picture(state, current_picture):
    for c in ["Y", "C1", "C2"]:
        remove_offset_component(state, current_picture[c], c)


# ---- DC prediction (decoder/transform_data_syntax.py:63, encoder/pictures.py:202) ----------------------

def _dc(band, inverse):
    if not len(band):
        return
    a = _to_i32(band, "coefficient")
    h, w = a.shape
    p = make_params(w, h, wavelet_index=3, dwt_depth=0, dwt_depth_ho=0)
    flat = np.concatenate([a.reshape(-1)] * 3)
    out = ops.dc_prediction(p, flat, inverse=inverse)[: w * h].reshape(h, w).tolist()
    for row, new in zip(band, out):
        row[:] = new


def dc_prediction(band):
    """(13.4) In place.  A band that comes from a :class:`DeviceTransform` of an LD picture has
    already been predicted on the device (``vc2b_unpack`` applies it): calling this on it, as
    ``fragment_data`` does after the last slice (decoder/fragment_syntax.py:171-181), is a no-op."""
    if isinstance(band, _PredictedBand):
        return
    _dc(band, False)


def apply_dc_prediction(band):
    """Inverse of dc_prediction, in place (encoder/pictures.py:202)."""
    _dc(band, True)


# ---- whole pictures ----------------------------------------------------------------------------

class _PredictedBand(list):
    """Level-0 band of a DeviceTransform: DC prediction already applied on the device."""


class DeviceTransform(dict):
    """``state["y_transform"]``-compatible container produced by :func:`transform_data`.  It keeps
    the dequantised coefficients of all three components on the device (``codec``) so that
    :func:`picture_decode` needs no host round trip, and materialises the reference's
    ``{level: {orient: [[int]]}}`` view lazily when something reads it as a dictionary."""

    def __init__(self, codec, comp):
        dict.__init__(self)
        self.codec = codec
        self.comp = comp
        self._filled = False

    def _fill(self):
        if not self._filled:
            self._filled = True
            g = self.codec.g
            flat = self.codec.coeff_components(0)[self.comp]
            d = _flat_to_coeff_dict(g, self.comp, flat)
            for orient in list(d[0]):
                d[0][orient] = _PredictedBand(d[0][orient])
            dict.update(self, d)

    def __getitem__(self, k):
        self._fill()
        return dict.__getitem__(self, k)

    def __iter__(self):
        self._fill()
        return dict.__iter__(self)

    def __len__(self):
        self._fill()
        return dict.__len__(self)

    def items(self):
        self._fill()
        return dict.items(self)

    def keys(self):
        self._fill()
        return dict.keys(self)

    def values(self):
        self._fill()
        return dict.values(self)


def _state_codec(state, p):
    from .device import BatchCodec

    return BatchCodec(p, 1)


def inverse_wavelet_transform(state):
    """(15.3) Sets current_picture Y/C1/C2 (cropped), as pseudocode/picture_decoding.py:73."""
    yt = state["y_transform"]
    if isinstance(yt, DeviceTransform):
        codec = yt.codec
    else:
        p = params_from_state(state)
        codec = _state_codec(state, p)
        g = codec.g
        flat = np.concatenate([
            _coeff_dict_to_flat(g, c, state[k]) for c, k in enumerate(["y_transform", "c1_transform", "c2_transform"])
        ])
        import torch

        codec.coeffs[: flat.size].copy_(torch.from_numpy(flat))
    codec.idwt_device(1)
    return codec


def picture_decode(state):
    """(15.2) pseudocode/picture_decoding.py:45: IDWT, pad removal, clip, offset, output callback."""
    state["current_picture"] = {}
    state["current_picture"]["pic_num"] = state["picture_number"]
    codec = inverse_wavelet_transform(state)
    planes = codec.picture_planes(0)  # idwt + crop + clip + offset are fused on the device
    for name, plane in zip(["Y", "C1", "C2"], planes):
        state["current_picture"][name] = plane.astype(np.int64).tolist()
    if "_output_picture_callback" in state:
        state["_output_picture_callback"](
            state["current_picture"], state["video_parameters"], state["picture_coding_mode"]
        )


def forward_wavelet_transform(state, current_picture):
    """Pads in place, then transforms (pseudocode/picture_encoding.py:78)."""
    for c in ["Y", "C1", "C2"]:
        dwt_pad_addition(state, current_picture[c], c)
    state["y_transform"] = dwt(state, current_picture["Y"])
    state["c1_transform"] = dwt(state, current_picture["C1"])
    state["c2_transform"] = dwt(state, current_picture["C2"])


def picture_encode(state, current_picture):
    """pseudocode/picture_encoding.py:45; mutates current_picture like the reference."""
    remove_offset_picture(state, current_picture)
    forward_wavelet_transform(state, current_picture)


# ---- transform_data (decoder/transform_data_syntax.py:116) ------------------------------------------

def _max_transform_bytes(p, num_slices):
    if p.is_ld:
        return (num_slices * p.slice_bytes_numerator) // p.slice_bytes_denominator
    return num_slices * (p.slice_prefix_bytes + 4 + 3 * 255 * p.slice_size_scaler)


def transform_data(state):
    """(13.5.2) Reads every slice of the picture from ``state["_file"]`` (the reader must be byte
    aligned, as after transform_parameters' byte_align), unpacks and dequantises on the device
    and leaves the reader exactly after the last slice.  Raises the reference's
    ``UnexpectedEndOfStream`` / ``InvalidSliceYLength`` and runs its per-slice level-constraint
    checks (decoder/transform_data_syntax.py:154,222,250) when that package is importable."""
    from .device import BatchCodec

    p = params_from_state(state)
    codec = BatchCodec(p, 1)
    g = codec.g
    f = state["_file"]
    assert state["next_bit"] == 7, "transform_data must start byte aligned"
    bound = _max_transform_bytes(p, g.num_slices)
    first = state["current_byte"]
    start = f.tell() - (1 if first is not None else 0)  # decoder/io.py:138 tell()
    rest = f.read(max(bound - 1, 0)) if first is not None else b""
    data = (bytes(bytearray([first])) if first is not None else b"") + rest
    codec.decode_streams([data], check=False)  # (also runs the IDWT; the picture is kept for picture_decode)
    st = codec.status_host(1)[0]
    qindex = codec.qindex[: g.num_slices].cpu().numpy()
    errors = _reference_errors()
    if st[0] == capi.ST_INT32_ENVELOPE:
        raise EnvelopeError("picture leaves the int32 envelope")
    # per-slice level constraints, in stream order, up to the failing slice if any
    last = int(st[1]) if st[0] != capi.ST_OK else g.num_slices
    if errors is not None:
        so = codec.slice_offset[: g.num_slices].cpu().numpy().astype(np.int64) & 0xffffffff
        for n in range(min(last + 1, g.num_slices)):
            if st[0] != capi.ST_OK and n == last and st[0] == capi.ST_UNEXPECTED_END_OF_STREAM and p.is_ld:
                break
            errors["assert_level_constraint"](state, "qindex", int(qindex[n]))
            if not p.is_ld and n < last:
                end = int(so[n + 1]) if n + 1 < g.num_slices else int(st[2])
                errors["assert_level_constraint"](state, "total_slice_bytes", end - int(so[n]))
    if st[0] == capi.ST_UNEXPECTED_END_OF_STREAM:
        raise (errors["UnexpectedEndOfStream"]() if errors else EOFError("UnexpectedEndOfStream"))
    if st[0] == capi.ST_INVALID_SLICE_Y_LENGTH:
        n = int(st[1])
        sy, sx = divmod(n, p.slices_x)
        if errors:
            raise errors["InvalidSliceYLength"](int(codec.qindex[n]), 0, sx, sy)
        raise ValueError("InvalidSliceYLength at slice %d" % n)
    consumed = int(st[2])
    # leave the reader after the last slice: current_byte = next byte, next_bit = 7 (decoder/io.py:152)
    f.seek(start + consumed)
    byte = f.read(1)
    state["current_byte"] = bytearray(byte)[0] if len(byte) == 1 else None
    state["next_bit"] = 7
    state["y_transform"] = DeviceTransform(codec, 0)
    state["c1_transform"] = DeviceTransform(codec, 1)
    state["c2_transform"] = DeviceTransform(codec, 2)
    # slice_quantizers of the last slice (decoder/transform_data_syntax.py:255)
    q = int(qindex[-1]) if g.num_slices else 0
    state["quantizer"] = {
        level: {orient: max(q - v, 0) for orient, v in orients.items()}
        for level, orients in state["quant_matrix"].items()
    }


def _reference_errors():
    try:
        from vc2_conformance.decoder.assertions import assert_level_constraint
        from vc2_conformance.decoder.exceptions import InvalidSliceYLength, UnexpectedEndOfStream
    except Exception:
        return None
    return {
        "assert_level_constraint": assert_level_constraint,
        "InvalidSliceYLength": InvalidSliceYLength,
        "UnexpectedEndOfStream": UnexpectedEndOfStream,
    }


# ---- fragmented pictures (decoder/fragment_syntax.py:140-181) ------------------------------------------

class _FragmentedTransform(DeviceTransform):
    """``state["*_transform"]`` of a fragmented picture: slices are buffered as bytes and unpacked on
    the device when the last one has arrived (or when somebody looks at the coefficients)."""

    def __init__(self, shared, comp):
        dict.__init__(self)
        self.shared = shared
        self.comp = comp
        self._filled = False

    @property
    def codec(self):
        return self.shared.finish()

    @codec.setter
    def codec(self, value):  # DeviceTransform.__init__ is bypassed; nothing to set
        pass


class _FragmentBuffer(object):
    def __init__(self, p):
        self.p = p
        self.data = bytearray()
        self._codec = None

    def finish(self):
        from .device import BatchCodec

        if self._codec is None:
            codec = BatchCodec(self.p, 1)
            codec.decode_streams([bytes(self.data)], check=False)
            st = codec.status_host(1)[0]
            if st[0] == capi.ST_INT32_ENVELOPE:
                raise EnvelopeError("picture leaves the int32 envelope")
            if st[0] != capi.ST_OK:
                raise ValueError("fragmented picture incomplete or corrupt (status %d at slice %d)" % (st[0], st[1]))
            self._codec = codec
        return self._codec


def initialize_fragment_state(state):
    """(14.3) decoder/fragment_syntax.py:140: coefficient containers that buffer slice bytes."""
    shared = _FragmentBuffer(params_from_state(state))
    state["y_transform"] = _FragmentedTransform(shared, 0)
    state["c1_transform"] = _FragmentedTransform(shared, 1)
    state["c2_transform"] = _FragmentedTransform(shared, 2)
    state["fragment_slices_received"] = 0
    state["_fragment_slices_remaining"] = state["slices_x"] * state["slices_y"]
    state["fragmented_picture_done"] = False


def _read_bytes(state, n):
    """n whole bytes from the byte-aligned reader (decoder/io.py:152-206 semantics)."""
    assert state["next_bit"] == 7, "slices of a fragment start byte aligned"
    errors = _reference_errors()
    out = bytearray()
    if n == 0:
        return out
    if state["current_byte"] is None:
        raise (errors["UnexpectedEndOfStream"]() if errors else EOFError("UnexpectedEndOfStream"))
    out.append(state["current_byte"])
    rest = state["_file"].read(n)  # n-1 more bytes of this slice + the next current byte
    out += rest[: n - 1]
    if len(rest) < n - 1:
        state["current_byte"] = None
        raise (errors["UnexpectedEndOfStream"]() if errors else EOFError("UnexpectedEndOfStream"))
    state["current_byte"] = bytearray(rest)[n - 1] if len(rest) == n else None
    return out


def fragment_slice(state, sx, sy):
    """Replacement for ``slice`` as called by fragment_data (decoder/fragment_syntax.py:168): consumes
    exactly the bytes of slice (sx, sy) and buffers them; the device unpacks the whole picture once
    every slice is there.  Per-slice level constraints are checked like hq_slice / ld_slice do."""
    shared = state["y_transform"].shared
    errors = _reference_errors()
    if (state["parse_code"] & 0xF8) == 0xC8:  # is_ld
        n = sy * state["slices_x"] + sx
        nbytes = ((n + 1) * state["slice_bytes_numerator"]) // state["slice_bytes_denominator"] - (
            n * state["slice_bytes_numerator"]) // state["slice_bytes_denominator"]
        chunk = _read_bytes(state, nbytes)
        if errors and nbytes:
            errors["assert_level_constraint"](state, "qindex", chunk[0] >> 1)
    else:
        scaler = state["slice_size_scaler"]
        chunk = _read_bytes(state, state["slice_prefix_bytes"] + 1)
        if errors:
            errors["assert_level_constraint"](state, "qindex", chunk[-1])
        for _ in range(3):
            length = _read_bytes(state, 1)
            chunk += length
            chunk += _read_bytes(state, scaler * length[0])
        if errors:
            errors["assert_level_constraint"](state, "total_slice_bytes", len(chunk))
    shared.data += chunk
