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
        po = O.Params()
        for f, _t in p._fields_:
            if f != "quant_matrix":
                setattr(po, f, getattr(p, f))
        for lv in range(O.MAX_LEVELS):
            for o in range(4):
                po.quant_matrix[lv][o] = p.quant_matrix[lv][o]
        for i, c in enumerate(chunks):
            _st, want, _n = O.decode_picture(po, c)
            planes = codec.picture_planes(i)
            for comp in range(3):
                assert (planes[comp] == want[comp]).all(), (name, i, comp)
