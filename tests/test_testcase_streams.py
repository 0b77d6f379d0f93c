"""Streams restating the reference's decoder test cases `slice_padding_data` and `dangling_bounded_block_data`
(test_cases/decoder/pictures.py:645-880): the reference's own per-slice helpers applied to
pictures encoded, serialised and decoded by the unmodified reference (tools/make_golden_testcases.py).
Padding bits of every kind must be ignored; values dangling over the end of a bounded block must decode as
the reference's bounded-block reader decodes them (decoder/io.py:246-252)."""

import numpy as np
import pytest

import oracle as O
from golden_utils import load, params_from_arrays


def _names():
    return [str(n) for n in load("testcases")["names"]]


def _base(name):
    return name.split(".", 1)[0]


def _params_key(z, name):
    """lossless_quantization streams carry their own slice_size_scaler; the others share their picture's."""
    return name + ".params" if name + ".params" in z.files else _base(name) + ".params"


@pytest.mark.parametrize("name", _names())
def test_oracle_decodes_testcase_stream(name):
    z = load("testcases")
    p = params_from_arrays(z[_params_key(z, name)], z[_base(name) + ".quant_matrix"])
    data = z[name + ".data"].tobytes()
    st, out, consumed = O.decode_picture(p, data)
    assert st == 0 and consumed == len(data)
    for c in range(3):
        assert (out[c] == z[name + ".decoded%d" % c]).all(), (name, c)


def test_all_kinds_present():
    names = _names()
    for kind in ("all_zeros", "all_ones", "alternating_1s_and_0s", "alternating_0s_and_1s", "dummy_end_of_sequence",
                 "zero_dangling", "sign_dangling", "stop_and_sign_dangling", "lsb_stop_and_sign_dangling",
                 "lossless_quantization"):
        assert any(kind in n for n in names), kind
    assert any(n.startswith("hq_") for n in names) and any(n.startswith("ld_") for n in names)


@pytest.mark.gpu
@pytest.mark.parametrize("base", sorted({_base(n) for n in _names()}))
def test_device_decodes_testcase_streams(base, unpack_kernel):
    from vc2_conformance_b200.device import BatchCodec
    from vc2_conformance_b200.params import params_from_arrays as dev_params

    z = load("testcases")
    groups = {}
    for n in _names():
        if _base(n) == base:
            groups.setdefault(_params_key(z, n), []).append(n)
    for key, names in groups.items():
        p = dev_params(z[key], z[base + ".quant_matrix"])
        streams = [z[n + ".data"].tobytes() for n in names]
        codec = BatchCodec(p, len(names))
        codec.decode_streams(streams)
        st = codec.status_host(len(names))
        assert (st[:, 0] == 0).all()
        assert (st[:, 2] == np.array([len(s) for s in streams])).all()
        for i, n in enumerate(names):
            coeffs = codec.coeff_components(i)
            planes = codec.picture_planes(i)
            for c in range(3):
                assert (coeffs[c] == z[n + ".dec_coeffs%d" % c]).all(), (n, "coeffs", c)
                assert (planes[c] == z[n + ".decoded%d" % c]).all(), (n, "picture", c)
