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
from .device import geometry_for as Geometry
from .params import make_params, params_from_state

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
    "calculate_coeffs_bits", "quantize_coeffs", "quantize_to_fit", "transform_and_slice_picture",
    "DeviceSliceCoeffs", "make_transform_data_hq_lossy", "make_transform_data_hq_lossless",
    "make_transform_data_ld_lossy",
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


def _raise_eos(errors):
    raise (errors["UnexpectedEndOfStream"]() if errors else EOFError("UnexpectedEndOfStream"))


def _intlog2(n):
    return (n - 1).bit_length()  # pseudocode/vc2_math.py:32


def _ld_slice_bytes(p, n):
    """slice_bytes (pseudocode/slice_sizes.py:95) of LD slice number n."""
    return ((n + 1) * p.slice_bytes_numerator) // p.slice_bytes_denominator - (
        n * p.slice_bytes_numerator) // p.slice_bytes_denominator


def _ld_slice_header(p, data, n):
    """(qindex, slice_y_length, slice_bytes) of LD slice n read from the host bytes (ld_slice,
    decoder/transform_data_syntax.py:145-172), or None where the bytes are not all there."""
    s0 = (n * p.slice_bytes_numerator) // p.slice_bytes_denominator
    nbytes = _ld_slice_bytes(p, n)
    length_bits = _intlog2(8 * nbytes - 7)
    hbytes = (7 + length_bits + 7) // 8
    if s0 + hbytes > len(data):
        return None
    hdr = int.from_bytes(bytes(data[s0:s0 + hbytes]), "big") >> (8 * hbytes - 7 - length_bits)
    return hdr >> length_bits, hdr & ((1 << length_bits) - 1), nbytes


def _max_transform_bytes(p, num_slices):
    if p.is_ld:
        return (num_slices * p.slice_bytes_numerator) // p.slice_bytes_denominator
    return num_slices * (p.slice_prefix_bytes + 4 + 3 * 255 * p.slice_size_scaler)


_window_hint = {}  # parameter bytes -> bytes the last picture of these parameters took


def _unpack_from_reader(state, p, codec):
    """Feeds the picture's slices from the reference's reader state (decoder/io.py) to the device and
    returns (data, status row).  The reader is left exactly after the last slice.

    A seekable file is read in windows (the size of the previous picture of the same parameters plus
    a margin, grown if the device runs out of bytes) and repositioned afterwards; any other file-like
    object is walked slice by slice on the host, which needs no seek."""
    g = codec.g
    f = state["_file"]
    assert state["next_bit"] == 7, "transform_data must start byte aligned"
    first = state["current_byte"]
    if first is None:
        data = b""
        codec.unpack_streams([data])
        return data, codec.status_host(1)[0]
    seekable = False
    try:
        seekable = bool(f.seekable())
    except Exception:
        pass
    if not seekable:
        data = bytearray()
        try:
            for n in range(g.num_slices):
                data += _read_slice_bytes(state, p, n)
        except Exception as e:  # UnexpectedEndOfStream: decode what there is, the device reports the slice
            if type(e).__name__ not in ("UnexpectedEndOfStream", "EOFError"):
                raise
        codec.unpack_streams([bytes(data)])
        return bytes(data), codec.status_host(1)[0]
    bound = _max_transform_bytes(p, g.num_slices)
    key = bytes(bytearray(p))
    start = f.tell() - 1  # decoder/io.py:138 tell(): current_byte is the byte before the file position
    want = bound if p.is_ld else min(bound, max(1 << 16, int(_window_hint.get(key, 0) * 1.25) + 4096))
    data = bytes(bytearray([first]))
    while True:
        more = f.read(want - len(data))
        data += more
        at_eof = len(data) < want
        codec.unpack_streams([data])
        st = codec.status_host(1)[0]
        if st[0] == capi.ST_UNEXPECTED_END_OF_STREAM and not at_eof and want < bound:
            want = min(bound, 4 * want)  # the window was too small, not the stream
            continue
        break
    if st[0] == capi.ST_OK:
        consumed = int(st[2])
        _window_hint[key] = consumed
        # leave the reader after the last slice: current_byte = next byte, next_bit = 7 (decoder/io.py:152)
        f.seek(start + consumed)
        byte = f.read(1)
        state["current_byte"] = bytearray(byte)[0] if len(byte) == 1 else None
        state["next_bit"] = 7
    return data, st


def transform_data(state):
    """(13.5.2) Reads every slice of the picture from ``state["_file"]`` (the reader must be byte
    aligned, as after transform_parameters' byte_align), unpacks and dequantises on the device
    and leaves the reader exactly after the last slice.  Raises the reference's
    ``UnexpectedEndOfStream`` / ``InvalidSliceYLength`` with the reference's arguments and runs its
    per-slice level-constraint checks (decoder/transform_data_syntax.py:154,222,250) when that package
    is importable.  The IDWT is left to :func:`picture_decode`."""
    p = params_from_state(state)
    codec = _state_codec(state, p)
    g = codec.g
    data, st = _unpack_from_reader(state, p, codec)
    errors = _reference_errors()
    if st[0] == capi.ST_INT32_ENVELOPE:
        raise EnvelopeError("picture leaves the int32 envelope")
    qindex = codec.qindex[: g.num_slices].cpu().numpy()
    # per-slice level constraints, in stream order, up to the failing slice if any
    last = int(st[1]) if st[0] != capi.ST_OK else g.num_slices
    if errors is not None:
        check = errors["assert_level_constraint"]
        so = codec.slice_offset[: g.num_slices].cpu().numpy().astype(np.int64) & 0xffffffff
        for n in range(last):
            check(state, "qindex", int(qindex[n]))
            if not p.is_ld:
                end = int(so[n + 1]) if n + 1 < last else _hq_slice_end(p, data, int(so[n]))
                check(state, "total_slice_bytes", end - int(so[n]))
        if last < g.num_slices:
            # the failing slice: its qindex is checked if the reference got as far as reading it
            if p.is_ld:
                s0 = (last * p.slice_bytes_numerator) // p.slice_bytes_denominator
                if s0 < len(data):
                    check(state, "qindex", data[s0] >> 1)
            else:
                off = _hq_slice_end(p, data, int(so[last - 1])) if last > 0 else 0
                if off is not None and off + p.slice_prefix_bytes < len(data):
                    check(state, "qindex", data[off + p.slice_prefix_bytes])
    if st[0] == capi.ST_UNEXPECTED_END_OF_STREAM:
        _raise_eos(errors)
    if st[0] == capi.ST_INVALID_SLICE_Y_LENGTH:
        n = int(st[1])
        sy, sx = divmod(n, p.slices_x)
        _q, slice_y_length, nbytes = _ld_slice_header(p, data, n)
        if errors:
            # InvalidSliceYLength(slice_y_length, slice_bytes, sx, sy): decoder/exceptions.py:1645
            raise errors["InvalidSliceYLength"](slice_y_length, nbytes, sx, sy)
        raise ValueError("InvalidSliceYLength %d (slice of %d bytes) at slice %d" % (slice_y_length, nbytes, n))
    state["y_transform"] = DeviceTransform(codec, 0)
    state["c1_transform"] = DeviceTransform(codec, 1)
    state["c2_transform"] = DeviceTransform(codec, 2)
    # slice_quantizers of the last slice (decoder/transform_data_syntax.py:255)
    q = int(qindex[-1]) if g.num_slices else 0
    state["quantizer"] = {
        level: {orient: max(q - v, 0) for orient, v in orients.items()}
        for level, orients in state["quant_matrix"].items()
    }


def _hq_slice_end(p, data, off):
    """Byte offset just after the HQ slice that starts at ``off`` (hq_slice,
    decoder/transform_data_syntax.py:212-251), or None if its length bytes are not all there."""
    pos = off + p.slice_prefix_bytes + 1
    for _ in range(3):
        if pos >= len(data):
            return None
        pos += 1 + p.slice_size_scaler * data[pos]
    return pos


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
            codec.unpack_streams([bytes(self.data)])
            st = codec.status_host(1)[0]
            if st[0] == capi.ST_INT32_ENVELOPE:
                raise EnvelopeError("picture leaves the int32 envelope")
            if st[0] != capi.ST_OK:
                # truncated streams and bad slice_y_length fields were raised by fragment_slice when it read
                # the slice, as the reference does; what is left is a picture whose slices have not all arrived
                errors = _reference_errors()
                if st[0] == capi.ST_UNEXPECTED_END_OF_STREAM:
                    _raise_eos(errors)
                raise RuntimeError("fragmented picture: device status %d at slice %d" % (st[0], st[1]))
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
        _raise_eos(errors)
    out.append(state["current_byte"])
    rest = state["_file"].read(n)  # n-1 more bytes of this slice + the next current byte
    out += rest[: n - 1]
    if len(rest) < n - 1:
        state["current_byte"] = None
        _raise_eos(errors)
    state["current_byte"] = bytearray(rest)[n - 1] if len(rest) == n else None
    return out


def _read_slice_bytes(state, p, n, check=None):
    """The bytes of slice number n, consumed from the reader as hq_slice / ld_slice would
    (decoder/transform_data_syntax.py:145-251); ``check(name, value)`` receives the level-constraint
    values in the reference's order.  A stream that ends inside the slice raises UnexpectedEndOfStream
    with the bytes read so far appended to nothing (the caller decides what to keep)."""
    if p.is_ld:
        nbytes = _ld_slice_bytes(p, n)
        chunk = _read_bytes(state, nbytes)
        if check and nbytes:
            check("qindex", chunk[0] >> 1)
        return chunk
    chunk = _read_bytes(state, p.slice_prefix_bytes + 1)
    if check:
        check("qindex", chunk[-1])
    for _ in range(3):
        length = _read_bytes(state, 1)
        chunk += length
        chunk += _read_bytes(state, p.slice_size_scaler * length[0])
    if check:
        check("total_slice_bytes", len(chunk))
    return chunk


def fragment_slice(state, sx, sy):
    """Replacement for ``slice`` as called by fragment_data (decoder/fragment_syntax.py:168): consumes
    exactly the bytes of slice (sx, sy) and buffers them; the device unpacks the whole picture once
    every slice is there.  Per-slice level constraints, UnexpectedEndOfStream and InvalidSliceYLength are
    raised here, while the slice is read, like hq_slice / ld_slice do."""
    shared = state["y_transform"].shared
    p = shared.p
    errors = _reference_errors()
    check = (lambda name, value: errors["assert_level_constraint"](state, name, value)) if errors else None
    n = sy * state["slices_x"] + sx
    chunk = _read_slice_bytes(state, p, n, check)
    if p.is_ld:
        # slice_y_length must leave room for itself (decoder/transform_data_syntax.py:166)
        nbytes = len(chunk)
        length_bits = _intlog2(8 * nbytes - 7)
        hbytes = (7 + length_bits + 7) // 8
        hdr = int.from_bytes(bytes(chunk[:hbytes]), "big") >> (8 * hbytes - 7 - length_bits)
        slice_y_length = hdr & ((1 << length_bits) - 1)
        if slice_y_length > 8 * nbytes - 7 - length_bits:
            if errors:
                raise errors["InvalidSliceYLength"](slice_y_length, nbytes, sx, sy)
            raise ValueError("InvalidSliceYLength %d (slice of %d bytes) at slice %d" % (slice_y_length, nbytes, n))
    shared.data += chunk


# ---- encoder: quantisation and slice assembly (encoder/pictures.py:217-792) ---------------------------
#
# Function-level forms (quantize_coeffs, calculate_coeffs_bits, quantize_to_fit) work on the reference's
# Python lists through vc2b_quantize_to_fit_lists.  The whole-picture forms keep the picture's coefficients
# on the device between transform_and_slice_picture and make_transform_data_*: one launch sequence per
# picture instead of one call per slice.

def _i32_list(values, what):
    a = np.asarray(_to_i32(list(values), what), np.int32).reshape(-1) if len(values) else np.zeros(0, np.int32)
    return a


def calculate_coeffs_bits(coeffs):
    """encoder/pictures.py:217: bits of a bounded block holding ``coeffs`` (trailing zeros are free)."""
    v = _i32_list(coeffs, "coefficient")
    if v.size == 0:
        return 0
    _q, bits, _out = ops.quantize_to_fit_lists([v], [np.zeros(v.size, np.int32)], None, 1, 0)
    return bits[0]


def quantize_coeffs(qindex, coeff_values, quant_matrix_values):
    """encoder/pictures.py:251: forward_quant of every value at max(0, qindex - matrix value)."""
    v = _i32_list(coeff_values, "coefficient")
    if v.size == 0:
        return []
    qm = _i32_list(quant_matrix_values, "quantisation matrix value")[: v.size]
    _q, _bits, out = ops.quantize_to_fit_lists([v[: qm.size]], [qm], None, 1, int(qindex))
    return out[0].tolist()


def quantize_to_fit(target_size, coeff_sets, align_bits=1, minimum_qindex=0):
    """encoder/pictures.py:295: smallest qindex >= minimum_qindex at which the sets, each padded to
    ``align_bits``, total at most ``target_size`` bits.  Returns (qindex, [quantised lists])."""
    assert target_size >= 0
    values = [_i32_list(cs[0], "coefficient") for cs in coeff_sets]
    qms = [_i32_list(cs[1], "quantisation matrix value")[: v.size] for cs, v in zip(coeff_sets, values)]
    values = [v[: q.size] for v, q in zip(values, qms)]  # zip() semantics of quantize_coeffs
    qindex, _bits, out = ops.quantize_to_fit_lists(values, qms, int(target_size), int(align_bits),
                                                   int(minimum_qindex))
    return qindex, [o.tolist() for o in out]


def _slice_order(g):
    """Per component: (order, counts, qm) -- ``order`` lists the component's coefficient indices slice by
    slice in bitstream order (level, orientation, then raster inside the slice's rectangle:
    encoder/pictures.py:418-451), ``counts[slice]`` how many each slice holds, ``qm`` the quantisation
    matrix value of every coefficient in that order."""
    cached = getattr(g, "_slice_order", None)
    if cached is not None:
        return cached
    p = g.params
    nx, ny = p.slices_x, p.slices_y
    nb = g.num_subbands
    out = []
    for c in range(3):
        key = np.empty(g.comp_coeffs[c], np.int64)
        qmv = np.empty(g.comp_coeffs[c], np.int32)
        for b, band in enumerate(g.bands[c]):
            w, h, off = band["w"], band["h"], band["off"]
            # the slice holding column x: largest s with (w * s) // nx <= x (slice_left, slice_sizes.py:108)
            sx = ((np.arange(w, dtype=np.int64) + 1) * nx - 1) // w
            sy = ((np.arange(h, dtype=np.int64) + 1) * ny - 1) // h
            key[off: off + w * h] = ((sy[:, None] * nx + sx[None, :]) * nb + b).reshape(-1)
            qmv[off: off + w * h] = p.quant_matrix[band["level"]][band["slot"]]
        order = np.argsort(key, kind="stable")
        counts = np.bincount(key // nb, minlength=nx * ny)
        out.append((order, counts, qmv[order]))
    g._slice_order = out
    return out


class DeviceSliceCoeffs(list):
    """What the device form of :func:`transform_and_slice_picture` returns: the picture's transform
    coefficients stay in ``codec.coeffs`` (DC prediction applied for the low delay profile) for the
    whole-picture ``make_transform_data_*`` replacements.  Anything that reads it as the reference's
    ``[[SliceCoeffs, ...], ...]`` gets that view, built on first access."""

    def __init__(self, codec, is_ld):
        list.__init__(self)
        self.codec = codec
        self.is_ld = is_ld
        self._filled = False

    def _fill(self):
        if self._filled:
            return
        self._filled = True
        from vc2_conformance.encoder.pictures import ComponentCoeffs, SliceCoeffs

        g = self.codec.g
        comps = self.codec.coeff_components(0)
        per = []
        for c, (order, counts, qm) in enumerate(_slice_order(g)):
            vals = comps[c][order]
            ends = np.cumsum(counts)
            per.append([(vals[e - n: e].tolist(), qm[e - n: e].tolist()) for e, n in zip(ends, counts)])
        nx, ny = g.params.slices_x, g.params.slices_y
        for sy in range(ny):
            list.append(self, [SliceCoeffs(*(ComponentCoeffs(*per[c][sy * nx + sx]) for c in range(3)))
                               for sx in range(nx)])

    def __len__(self):
        return self.codec.g.params.slices_y

    def __getitem__(self, i):
        self._fill()
        return list.__getitem__(self, i)

    def __iter__(self):
        self._fill()
        return list.__iter__(self)


def transform_and_slice_picture(codec_features, picture):
    """encoder/pictures.py:350 on the device: offset removal, padding, DWT (picture_encode) and, for the
    low delay profile, apply_dc_prediction; returns a :class:`DeviceSliceCoeffs`."""
    from vc2_conformance.encoder.pictures import get_quantization_marix
    from vc2_conformance.pseudocode.video_parameters import set_coding_parameters
    from vc2_data_tables import Profiles

    from .device import BatchCodec

    state = dict(
        wavelet_index=codec_features["wavelet_index"],
        wavelet_index_ho=codec_features["wavelet_index_ho"],
        dwt_depth=codec_features["dwt_depth"],
        dwt_depth_ho=codec_features["dwt_depth_ho"],
        slices_x=codec_features["slices_x"],
        slices_y=codec_features["slices_y"],
        picture_coding_mode=codec_features["picture_coding_mode"],
    )
    set_coding_parameters(state, codec_features["video_parameters"])
    state["quant_matrix"] = get_quantization_marix(codec_features)
    is_ld = codec_features["profile"] == Profiles.low_delay
    state["parse_code"] = 0xC8 if is_ld else 0xE8
    if is_ld:  # placeholders: the slice sizes are only known to make_transform_data_ld_lossy
        state["slice_bytes_numerator"] = state["slices_x"] * state["slices_y"]
        state["slice_bytes_denominator"] = 1
    p = params_from_state(state)
    codec = BatchCodec(p, 1)
    g = codec.g
    flat = np.empty(g.picture_samples, np.uint16)
    for c, name in enumerate(["Y", "C1", "C2"]):
        w, h = g.dims[c]
        a = np.array(picture[name], dtype=object)
        if a.shape != (h, w):
            raise ValueError("component %s is %r, expected %r" % (name, a.shape, (h, w)))
        if a.size and (min(int(v) for v in a.reshape(-1)) < 0 or max(int(v) for v in a.reshape(-1)) > 0xffff):
            raise EnvelopeError("sample outside the 16-bit envelope of the device pictures")
        flat[g.plane_off[c]: g.plane_off[c] + w * h] = a.astype(np.uint16).reshape(-1)
    codec.upload_pictures(flat[None, :])
    codec.dwt_device(1)
    if is_ld:
        codec.dc_predict_device(1, inverse=True)
    return DeviceSliceCoeffs(codec, is_ld)


def _original(name):
    """The reference's own function of that name (not the rebound one)."""
    from . import install

    return install.original("vc2_conformance.encoder.pictures", name)


def _codec_with(codec, **changes):
    """A view of ``codec`` (same device buffers) for a changed parameter set (slice sizes)."""
    import copy

    from .device import geometry_for
    from .params import copy_params

    view = copy.copy(codec)
    view.p = copy_params(codec.p, **changes)
    view.g = geometry_for(view.p)
    return view


def _per_slice_lists(g, flat_components):
    """[comp][slice] -> list of values in bitstream order."""
    out = []
    for c, (order, counts, _qm) in enumerate(_slice_order(g)):
        vals = flat_components[c][order]
        ends = np.cumsum(counts)
        out.append([vals[e - n: e].tolist() for e, n in zip(ends, counts)])
    return out


def make_transform_data_hq_lossy(picture_bytes, transform_coeffs, minimum_qindex=0, minimum_slice_size_scaler=1):
    """encoder/pictures.py:617 for a :class:`DeviceSliceCoeffs`: one quantize_to_fit launch for every
    slice of the picture; returns (slice_size_scaler, TransformData of HQSlice)."""
    if not isinstance(transform_coeffs, DeviceSliceCoeffs):
        return _original("make_transform_data_hq_lossy")(picture_bytes, transform_coeffs, minimum_qindex,
                                                        minimum_slice_size_scaler)
    from vc2_conformance.bitstream import HQSlice, TransformData
    from vc2_conformance.encoder.exceptions import InsufficientHQPictureBytesError
    from vc2_conformance.encoder.pictures import get_safe_lossy_hq_slice_size_scaler

    codec0 = transform_coeffs.codec
    num_slices = codec0.g.num_slices
    scaler = max(get_safe_lossy_hq_slice_size_scaler(picture_bytes, num_slices), minimum_slice_size_scaler)
    total_coeff_bytes = picture_bytes - num_slices * 4
    if total_coeff_bytes < 0:
        raise InsufficientHQPictureBytesError()
    codec = _codec_with(codec0, slice_size_scaler=scaler)
    g = codec.g
    codec.quantize_device(1, picture_bytes, minimum_qindex)
    qindex = codec.qindex[:num_slices].cpu().numpy()
    bits = codec.slice_bits[: 3 * num_slices].cpu().numpy().reshape(num_slices, 3)
    quantised = codec.quantized_host(1)[0]
    per = _per_slice_lists(g, [quantised[g.comp_off[c]: g.comp_off[c] + g.comp_coeffs[c]] for c in range(3)])
    align = 8 * scaler
    den = num_slices * scaler
    slices = []
    for n in range(num_slices):
        # slice_bytes (slice_sizes.py:95) in units of slice_size_scaler bytes, as at :672-690
        total_length = ((n + 1) * total_coeff_bytes) // den - (n * total_coeff_bytes) // den
        y_length = (int(bits[n, 0]) + align - 1) // align
        c1_length = (int(bits[n, 1]) + align - 1) // align
        slices.append(HQSlice(
            qindex=int(qindex[n]), slice_y_length=y_length, slice_c1_length=c1_length,
            slice_c2_length=total_length - y_length - c1_length,
            y_transform=per[0][n], c1_transform=per[1][n], c2_transform=per[2][n]))
    return scaler, TransformData(hq_slices=slices)


def make_transform_data_hq_lossless(transform_coeffs, minimum_slice_size_scaler=1):
    """encoder/pictures.py:528 for a :class:`DeviceSliceCoeffs`: block sizes on the device, minimal length
    fields, the smallest slice_size_scaler that holds the longest of them."""
    if not isinstance(transform_coeffs, DeviceSliceCoeffs):
        return _original("make_transform_data_hq_lossless")(transform_coeffs, minimum_slice_size_scaler)
    from vc2_conformance.bitstream import HQSlice, TransformData

    codec = transform_coeffs.codec
    g = codec.g
    num_slices = g.num_slices
    scaler, _ps, _offs, _totals = codec.lossless_layout(1, minimum_slice_size_scaler)
    bits = codec.slice_bits[: 3 * num_slices].cpu().numpy().reshape(num_slices, 3)
    per = _per_slice_lists(g, codec.coeff_components(0))
    align = 8 * scaler
    slices = [HQSlice(
        qindex=0,
        slice_y_length=(int(bits[n, 0]) + align - 1) // align,
        slice_c1_length=(int(bits[n, 1]) + align - 1) // align,
        slice_c2_length=(int(bits[n, 2]) + align - 1) // align,
        y_transform=per[0][n], c1_transform=per[1][n], c2_transform=per[2][n]) for n in range(num_slices)]
    return scaler, TransformData(hq_slices=slices)


def make_transform_data_ld_lossy(picture_bytes, transform_coeffs, minimum_qindex=0):
    """encoder/pictures.py:724 for a :class:`DeviceSliceCoeffs`."""
    if not isinstance(transform_coeffs, DeviceSliceCoeffs):
        return _original("make_transform_data_ld_lossy")(picture_bytes, transform_coeffs, minimum_qindex)
    from vc2_conformance.bitstream import LDSlice, TransformData
    from vc2_conformance.encoder.exceptions import InsufficientLDPictureBytesError

    codec0 = transform_coeffs.codec
    num_slices = codec0.g.num_slices
    for n in range(num_slices):  # target_size of every slice (:752-760) must not be negative
        target = 8 * (((n + 1) * picture_bytes) // num_slices - (n * picture_bytes) // num_slices) - 7
        if target - _intlog2(target) < 0:
            raise InsufficientLDPictureBytesError()
    codec = _codec_with(codec0, slice_bytes_numerator=int(picture_bytes), slice_bytes_denominator=num_slices)
    g = codec.g
    codec.quantize_device(1, picture_bytes, minimum_qindex, predict=False)  # DC prediction already applied
    qindex = codec.qindex[:num_slices].cpu().numpy()
    bits = codec.slice_bits[: 3 * num_slices].cpu().numpy().reshape(num_slices, 3)
    quantised = codec.quantized_host(1)[0]
    per = _per_slice_lists(g, [quantised[g.comp_off[c]: g.comp_off[c] + g.comp_coeffs[c]] for c in range(3)])
    slices = []
    for n in range(num_slices):
        c = [v for pair in zip(per[1][n], per[2][n]) for v in pair]  # interleave (:712)
        slices.append(LDSlice(qindex=int(qindex[n]), slice_y_length=int(bits[n, 0]), y_transform=per[0][n],
                              c_transform=c))
    return TransformData(ld_slices=slices)
