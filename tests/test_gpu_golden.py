"""CUDA path (through the C ABI) against the golden vectors generated from the unmodified Python
reference (tools/make_golden.py).  Bit-exact."""

import numpy as np
import pytest

from golden_utils import corrupt_variants, load, picture_case_names

pytestmark = pytest.mark.gpu


def _params(z, name):
    from vc2_conformance_b200.params import params_from_arrays

    return params_from_arrays(z[name + ".params"], z[name + ".quant_matrix"])


@pytest.mark.parametrize("name", picture_case_names())
def test_decode_golden_pictures(name, unpack_kernel):
    from vc2_conformance_b200.device import BatchCodec

    z = load("pictures")
    p = _params(z, name)
    data = z[name + ".data"].tobytes()
    codec = BatchCodec(p, 2)
    codec.decode_streams([data, data])
    st = codec.status_host(2)
    assert (st[:, 0] == 0).all() and (st[:, 2] == len(data)).all()
    ns = p.slices_x * p.slices_y
    assert (codec.qindex[:ns].cpu().numpy() == z[name + ".qindex"]).all()
    for i in range(2):
        coeffs = codec.coeff_components(i)
        planes = codec.picture_planes(i)
        for c in range(3):
            assert (coeffs[c] == z[name + ".dec_coeffs%d" % c]).all(), (name, "coeffs", c)
            assert (planes[c] == z[name + ".decoded%d" % c]).all(), (name, "picture", c)
        assert st[i, 3] == max(int(np.abs(z[name + ".dec_coeffs%d" % c]).max()) for c in range(3)) or p.is_ld


@pytest.mark.parametrize("name", picture_case_names())
def test_decode_corrupt_streams(name, unpack_kernel):
    """Dangling / truncated bounded blocks, garbage lengths, end of stream
    (vc2_conformance/test_cases/decoder/pictures.py:351-472 semantics)."""
    import ctypes as C

    from vc2_conformance_b200 import capi
    from vc2_conformance_b200.device import BatchCodec

    z = load("pictures")
    p = _params(z, name)
    codec = BatchCodec(p, 1)
    for v in corrupt_variants(z, name):
        bdata = z[v + ".data"].tobytes()
        codec.decode_streams([bdata], check=False)
        st = codec.status_host(1)[0]
        exp_status = int(z[v + ".status"][0])
        if st[0] == capi.ST_INT32_ENVELOPE and exp_status == 0:
            # The reference computes in unbounded integers; the CUDA path must flag (not wrap)
            # a picture whose coefficients leave its documented int32 envelope.
            safe = capi.lib().vc2b_idwt_max_safe_coeff(C.byref(p))
            biggest = max(int(np.abs(z[v + ".dec_coeffs%d" % c]).max()) for c in range(3))
            assert biggest > safe, (v, biggest, safe)
            continue
        assert st[0] == exp_status, (v, st)
        if st[0] == 0:
            assert st[2] == int(z[v + ".consumed"][0]), v
            coeffs = codec.coeff_components(0)
            planes = codec.picture_planes(0)
            for c in range(3):
                assert (coeffs[c] == z[v + ".dec_coeffs%d" % c]).all(), (v, "coeffs", c)
                assert (planes[c] == z[v + ".decoded%d" % c]).all(), (v, "picture", c)


def test_transform_golden():
    from vc2_conformance_b200 import ops
    from vc2_conformance_b200.params import make_params

    z = load("transform")
    for k in range(int(z["count"][0])):
        wi, wiho, d, dho, W, H = [int(v) for v in z["case_%d" % k]]
        p = make_params(W, H, wavelet_index=wi, wavelet_index_ho=wiho, dwt_depth=d, dwt_depth_ho=dho)
        got = ops.idwt_component(p, 0, z["coeffs_%d" % k])
        assert (got == z["idwt_%d" % k]).all(), ("idwt", k)
        got = ops.dwt_component(p, 0, z["pic_%d" % k])
        assert (got == z["dwt_%d" % k]).all(), ("dwt", k)


def test_quant_golden():
    from vc2_conformance_b200 import ops

    z = load("quant")
    ok = np.abs(z["inverse_quant"]) < (1 << 31)
    got = ops.inverse_quant(z["vals"][ok], z["qis"][ok])
    assert (got == z["inverse_quant"][ok]).all()
    got = ops.forward_quant(z["vals"], z["qis"])
    assert (got == z["forward_quant"]).all()


def test_dc_prediction_golden():
    from vc2_conformance_b200 import ops
    from vc2_conformance_b200.params import make_params

    z = load("dc_prediction")
    for k in range(int(z["count"][0])):
        band = z["in_%d" % k]
        h, w = band.shape
        p = make_params(w, h, wavelet_index=3, dwt_depth=0, dwt_depth_ho=0)
        flat = np.concatenate([band.reshape(-1)] * 3).astype(np.int32)
        got = ops.dc_prediction(p, flat)
        for c in range(3):
            assert (got[c * w * h:(c + 1) * w * h].reshape(h, w) == z["pred_%d" % k]).all(), k
        got = ops.dc_prediction(p, flat, inverse=True)
        for c in range(3):
            assert (got[c * w * h:(c + 1) * w * h].reshape(h, w) == z["apply_%d" % k]).all(), k


@pytest.mark.parametrize("name", picture_case_names())
def test_encode_golden_pictures(name):
    """picture_encode coefficients, per-slice qindex and the serialised bytes equal the reference's
    (transform_and_slice_picture + make_transform_data_* + serialisation)."""
    from vc2_conformance_b200.device import BatchCodec

    z = load("pictures")
    p = _params(z, name)
    pics = np.concatenate([z[name + ".pic%d" % c].reshape(-1) for c in range(3)])
    codec = BatchCodec(p, 2)
    arg = int(z[name + ".arg"][0])
    lossless = "lossless" in name
    if lossless:
        # make_transform_data_hq_lossless (encoder/pictures.py:528): qindex 0, minimal length fields, the
        # smallest slice_size_scaler that holds them -- the golden parameters carry the reference's choice
        scaler, out = codec.encode_pictures_lossless(np.stack([pics, pics]))
        assert scaler == p.slice_size_scaler
    else:
        out = codec.encode_pictures(np.stack([pics, pics]), arg)
    coeffs = codec.coeff_components(1)
    if not p.is_ld:  # (LD: dc prediction has been applied in place)
        for c in range(3):
            assert (coeffs[c] == z[name + ".enc_coeffs%d" % c]).all(), (name, c)
    ns = p.slices_x * p.slices_y
    assert (codec.qindex[ns:2 * ns].cpu().numpy() == z[name + ".qindex"]).all()
    data = z[name + ".data"].tobytes()
    if p.slice_prefix_bytes:
        # the reference stream carries random prefix bytes; ours are zero: compare decodes
        dec = BatchCodec(p, 1)
        dec.decode_streams([out[0]])
        for c in range(3):
            assert (dec.picture_planes(0)[c] == z[name + ".decoded%d" % c]).all()
    else:
        assert out[0] == data and out[1] == data
