"""
Index of a VC-2 byte stream without the Python bit reader.

``index_stream(data)`` walks the data units of a stream in C (``vc2b_index_stream``: parse_info,
sequence_header, picture_header, transform_parameters, fragment_header -- the syntax the reference reads
one bit at a time in decoder/stream.py:194-308, sequence_header.py, picture_syntax.py:67-348 and
fragment_syntax.py:44-140) and returns, for every picture and fragment, the ``capi.Params`` and the byte
offset that :class:`~vc2_conformance_b200.device.BatchCodec` needs.  It performs no conformance checks:
those stay with the reference decoder.
"""

import ctypes as C

from . import capi

PARSE_CODE_SEQUENCE_HEADER = 0x00
PARSE_CODE_END_OF_SEQUENCE = 0x10
PICTURE_PARSE_CODES = (0xC8, 0xE8)
FRAGMENT_PARSE_CODES = (0xCC, 0xEC)


def index_stream(data):
    """``data``: bytes-like object holding the stream.  Returns a list of ``capi.Unit``."""
    buf = (C.c_ubyte * len(data)).from_buffer_copy(bytes(data)) if len(data) else (C.c_ubyte * 1)()
    L = capi.lib()
    n = int(L.vc2b_index_stream(buf, len(data), None, 0))
    if n < 0:
        raise capi.Vc2bError("vc2b_index_stream", n)
    units = (capi.Unit * max(n, 1))()
    m = int(L.vc2b_index_stream(buf, len(data), units, n))
    assert m == n
    return [units[i] for i in range(n)]


def pictures(units):
    """The whole (non-fragmented) pictures of an index: [(picture_number, params, data_offset), ...]."""
    return [(int(u.picture_number), u.params, int(u.data_offset)) for u in units if u.parse_code in PICTURE_PARSE_CODES]


def _slice_sizes(p, data, pos, first_slice, count):
    """Total bytes of ``count`` consecutive slices starting at ``data[pos]`` (slice number ``first_slice``):
    closed form for LD (pseudocode/slice_sizes.py:95), the length bytes for HQ
    (decoder/transform_data_syntax.py:212-251).  Stops at the end of ``data``."""
    start = pos
    for n in range(first_slice, first_slice + count):
        if p.is_ld:
            pos += ((n + 1) * p.slice_bytes_numerator) // p.slice_bytes_denominator - (
                n * p.slice_bytes_numerator) // p.slice_bytes_denominator
        else:
            pos += p.slice_prefix_bytes + 1
            for _ in range(3):
                if pos >= len(data):
                    return len(data) - start
                pos += 1 + p.slice_size_scaler * data[pos]
        if pos >= len(data):
            return len(data) - start
    return pos - start


def picture_payloads(data, units):
    """The transform data of every picture of an indexed stream, whole or FRAGMENTED, ready for
    :meth:`~vc2_conformance_b200.device.BatchCodec.decode_streams`: [(picture_number, params, bytes), ...].
    The slices of a fragmented picture (decoder/fragment_syntax.py:155-181) are gathered from its fragments in
    stream order -- which is slice raster order -- so the device unpacks the picture in one launch."""
    data = bytes(data)
    out = []
    pending = None  # [picture_number, params, bytearray, slices received]
    for k, u in enumerate(units):
        end = int(units[k + 1].offset) if k + 1 < len(units) else len(data)
        if u.parse_code in PICTURE_PARSE_CODES:
            out.append((int(u.picture_number), u.params, data[int(u.data_offset):end]))
        elif u.parse_code in FRAGMENT_PARSE_CODES:
            if u.fragment_slice_count == 0:
                pending = [int(u.picture_number), u.params, bytearray(), 0]
            elif pending is not None:
                p = pending[1]
                first = u.fragment_y_offset * p.slices_x + u.fragment_x_offset
                if first != pending[3]:
                    raise ValueError("fragment at slice %d, expected slice %d" % (first, pending[3]))
                pos = int(u.data_offset)
                size = _slice_sizes(p, data[:end], pos, first, int(u.fragment_slice_count))
                pending[2] += data[pos:pos + size]
                pending[3] += int(u.fragment_slice_count)
                if pending[3] >= p.slices_x * p.slices_y:
                    out.append((pending[0], p, bytes(pending[2])))
                    pending = None
    return out
