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
