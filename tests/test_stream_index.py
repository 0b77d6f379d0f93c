"""vc2b_index_stream against whole streams serialised and deserialised by the unmodified
reference (tools/make_golden_index.py): every data unit offset, parse code, picture number, fragment
header field and transform parameter must equal what the reference's parser read."""

import numpy as np
import pytest

import oracle as O
from golden_utils import load


def _names():
    return [str(n) for n in load("index")["names"]]


@pytest.mark.parametrize("name", _names())
def test_index_matches_reference_parser(name):
    from vc2_conformance_b200 import capi
    from vc2_conformance_b200 import stream_index as SI

    z = load("index")
    stream = z[name + ".stream"].tobytes()
    table, offs, qm, dims = z[name + ".units"], z[name + ".data_offsets"], z[name + ".quant_matrix"], z[name + ".dims"]
    units = SI.index_stream(stream)
    assert len(units) == len(table)
    for u, t, doff in zip(units, table, offs):
        assert (u.offset, u.parse_code) == (int(t[0]), int(t[1]))
        assert u.next_offset == (int(t[0]) + int(t[2]) if t[2] else 0)
        assert u.flags == 0, (name, u.offset, u.flags)
        assert (u.major_version, u.picture_coding_mode) == (int(dims[6]), int(dims[7]))
        is_pic = u.parse_code in SI.PICTURE_PARSE_CODES
        is_frag = u.parse_code in SI.FRAGMENT_PARSE_CODES
        if is_pic or is_frag:
            assert u.picture_number == int(t[3])
            assert (u.fragment_data_length, u.fragment_slice_count, u.fragment_x_offset, u.fragment_y_offset) == \
                tuple(int(v) for v in t[4:8])
            p = u.params
            assert (p.luma_width, p.luma_height, p.color_diff_width, p.color_diff_height, p.luma_depth,
                    p.color_diff_depth) == tuple(int(v) for v in dims[:6])
        if t[8]:  # this unit carries transform parameters (pictures, first fragment of a picture)
            p = u.params
            assert (p.wavelet_index, p.dwt_depth, p.wavelet_index_ho, p.dwt_depth_ho, p.slices_x, p.slices_y) == \
                tuple(int(v) for v in t[9:15])
            if p.is_ld:
                assert (p.slice_bytes_numerator, p.slice_bytes_denominator) == (int(t[15]), int(t[16]))
            else:
                assert (p.slice_prefix_bytes, p.slice_size_scaler) == (int(t[17]), int(t[18]))
            got_qm = np.array([[p.quant_matrix[lv][o] for o in range(4)] for lv in range(capi.MAX_LEVELS)])
            assert (got_qm == qm).all()
        if is_pic:
            assert u.data_offset == int(doff)
        if is_frag and u.fragment_slice_count:
            # the slices of a fragment run up to the next data unit (fragment_data_length counts them in bytes)
            assert u.data_offset + u.fragment_data_length <= (u.next_offset or len(stream))
            assert u.params.slices_x > 0  # carried over from the picture's first fragment


def test_pictures_found_by_the_index_decode_with_the_oracle():
    """The parameters and offsets of the index are enough to decode: the oracle must consume exactly the
    bytes up to the next data unit."""
    from vc2_conformance_b200 import stream_index as SI

    z = load("index")
    for name in _names():
        stream = z[name + ".stream"].tobytes()
        units = SI.index_stream(stream)
        for k, u in enumerate(units):
            if u.parse_code not in SI.PICTURE_PARSE_CODES:
                continue
            po = O.Params()
            for f, _t in u.params._fields_:
                if f != "quant_matrix":
                    setattr(po, f, getattr(u.params, f))
            for lv in range(O.MAX_LEVELS):
                for o in range(4):
                    po.quant_matrix[lv][o] = u.params.quant_matrix[lv][o]
            st, _pic, consumed = O.decode_picture(po, stream[u.data_offset:units[k + 1].offset])
            assert st == 0 and u.data_offset + consumed == units[k + 1].offset, (name, k)


def _oracle_params(p):
    po = O.Params()
    for f, _t in p._fields_:
        if f != "quant_matrix":
            setattr(po, f, getattr(p, f))
    for lv in range(O.MAX_LEVELS):
        for o in range(4):
            po.quant_matrix[lv][o] = p.quant_matrix[lv][o]
    return po


def _default_names():
    return [str(n) for n in load("index_defaults")["names"]]


@pytest.mark.parametrize("name", _default_names())
def test_base_video_formats_and_default_quant_matrices(name):
    """Streams written by the reference encoder that rely on a base video format and on the default
    quantisation matrix (tools/make_golden_index_defaults.py): no NEEDS_* flag, the parameters equal the
    reference decoder's state, and decoding the payloads found by the index gives the reference's pictures."""
    from vc2_conformance_b200 import capi
    from vc2_conformance_b200 import stream_index as SI

    z = load("index_defaults")
    stream = z[name + ".stream"].tobytes()
    units = SI.index_stream(stream)
    want = [int(v) for v in z[name + ".state"]]
    carrying = [u for u in units if u.parse_code in SI.PICTURE_PARSE_CODES + SI.FRAGMENT_PARSE_CODES]
    assert carrying
    for u in units:
        assert u.flags == 0, (name, hex(u.parse_code), u.flags)
    for u in carrying:
        p = u.params
        got = [p.luma_width, p.luma_height, p.color_diff_width, p.color_diff_height, p.luma_depth, p.color_diff_depth,
               p.wavelet_index, p.wavelet_index_ho, p.dwt_depth, p.dwt_depth_ho, p.slices_x, p.slices_y]
        assert got == want, (name, got, want)
        got_qm = np.array([[p.quant_matrix[lv][o] for o in range(4)] for lv in range(capi.MAX_LEVELS)])
        assert (got_qm == z[name + ".quant_matrix"]).all(), name
    payloads = SI.picture_payloads(stream, units)
    assert payloads
    _pn, p, data = payloads[0]
    st, pics, consumed = O.decode_picture(_oracle_params(p), data)
    assert st == 0 and consumed == len(data)
    for c, comp in enumerate(("Y", "C1", "C2")):
        assert (pics[c].astype(np.int64) == z[name + ".picture_" + comp]).all(), (name, comp)


def test_fragmented_pictures_are_gathered_from_the_index():
    """picture_payloads: the slices of a fragmented picture, gathered from its fragments, decode (oracle) to
    the same picture as the whole-picture stream of the same case would."""
    from vc2_conformance_b200 import stream_index as SI

    z = load("index")
    for name in ("hq_fragments", "ld_fragments"):
        stream = z[name + ".stream"].tobytes()
        units = SI.index_stream(stream)
        payloads = SI.picture_payloads(stream, units)
        assert len(payloads) == 2 and [pn for pn, _p, _d in payloads] == [100, 101]
        for _pn, p, data in payloads:
            st, _pics, consumed = O.decode_picture(_oracle_params(p), data)
            assert st == 0 and consumed == len(data), name


def test_hostile_headers_do_not_walk_out_of_the_buffer():
    """slice_prefix_bytes / slice_size_scaler / slice_bytes_* are untrusted exp-Golomb values; with
    next_parse_offset = 0 the indexer walks the slices itself and must neither read outside the buffer nor loop."""
    from vc2_conformance_b200 import capi
    from vc2_conformance_b200 import stream_index as SI

    z = load("index")
    for name in ("hq_v3_asym_picture_no_next_offset", "ld_fields_picture_no_next_offset"):
        stream = bytearray(z[name + ".stream"].tobytes())
        units = SI.index_stream(bytes(stream))
        pic = [u for u in units if u.parse_code in SI.PICTURE_PARSE_CODES][0]
        # corrupt every byte of the picture's header area in turn (transform parameters sit between the
        # parse_info header and the transform data): the call must return, whatever it finds
        for pos in range(int(pic.offset) + 13, int(pic.data_offset)):
            for value in (0x00, 0xff, 0x55):
                bad = bytearray(stream)
                bad[pos] = value
                got = SI.index_stream(bytes(bad))
                assert 1 <= len(got) <= len(units) + 4
                for u in got:
                    assert 0 <= u.offset <= len(bad)
                    assert u.data_offset <= len(bad)


def test_truncated_and_foreign_buffers():
    from vc2_conformance_b200 import capi
    from vc2_conformance_b200 import stream_index as SI

    z = load("index")
    stream = z["hq_v2_picture.stream"].tobytes()
    with pytest.raises(capi.Vc2bError):
        SI.index_stream(b"not a VC-2 stream at all")
    assert SI.index_stream(b"") == []
    cut = SI.index_stream(stream[:50])  # sequence header, then a picture whose header is cut short
    assert len(cut) == 2 and (cut[1].flags & capi.IDX_TRUNCATED)
    # two sequences back to back: the index continues after the end of sequence
    assert len(SI.index_stream(stream + stream)) == 2 * len(SI.index_stream(stream))


@pytest.mark.gpu
def test_pictures_found_by_the_index_decode_on_the_device():
    from vc2_conformance_b200 import stream_index as SI
    from vc2_conformance_b200.device import BatchCodec

    z = load("index")
    for name in _names():
        stream = z[name + ".stream"].tobytes()
        units = SI.index_stream(stream)
        pics = [(k, u) for k, u in enumerate(units) if u.parse_code in SI.PICTURE_PARSE_CODES]
        if not pics:
            continue
        p = pics[0][1].params
        chunks = [stream[u.data_offset:units[k + 1].offset] for k, u in pics]
        codec = BatchCodec(p, len(chunks))
        codec.decode_streams(chunks)
        st = codec.status_host(len(chunks))
        assert (st[:, 0] == 0).all() and (st[:, 2] == np.array([len(c) for c in chunks])).all()
        po = _oracle_params(p)
        for i, c in enumerate(chunks):
            _st, want, _n = O.decode_picture(po, c)
            planes = codec.picture_planes(i)
            for comp in range(3):
                assert (planes[comp] == want[comp]).all(), (name, i, comp)


@pytest.mark.gpu
def test_fragmented_and_default_table_pictures_decode_on_the_device():
    """Fragment slices and table-dependent streams go from the index straight to the device."""
    from vc2_conformance_b200 import stream_index as SI
    from vc2_conformance_b200.device import BatchCodec

    cases = [("index", n) for n in ("hq_fragments", "ld_fragments")] + [("index_defaults", n) for n in _default_names()]
    for gold, name in cases:
        z = load(gold)
        stream = z[name + ".stream"].tobytes()
        payloads = SI.picture_payloads(stream, SI.index_stream(stream))
        p = payloads[0][1]
        chunks = [d for _pn, _p, d in payloads]
        codec = BatchCodec(p, len(chunks))
        codec.decode_streams(chunks)
        st = codec.status_host(len(chunks))
        assert (st[:, 0] == 0).all() and (st[:, 2] == np.array([len(c) for c in chunks])).all(), name
        for i, c in enumerate(chunks):
            _st, want, _n = O.decode_picture(_oracle_params(p), c)
            planes = codec.picture_planes(i)
            for comp in range(3):
                assert (planes[comp] == want[comp]).all(), (name, i, comp)
