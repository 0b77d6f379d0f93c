"""The C oracle (oracle/vc2_oracle.c) against the golden vectors generated from the unmodified
Python reference (tools/make_golden.py).  CPU only."""

import ctypes as C

import numpy as np
import pytest

import oracle as O
from golden_utils import corrupt_variants, load, params_from_arrays, picture_case_names


def test_quant_tables():
    z = load("quant")
    L = O.lib()
    for i, qf, qo in zip(z["idx"], z["quant_factor"], z["quant_offset"]):
        assert L.vc2o_quant_factor(int(i)) == qf
        assert L.vc2o_quant_offset(int(i)) == qo
    for v, q, iq, fq in zip(z["vals"], z["qis"], z["inverse_quant"], z["forward_quant"]):
        assert L.vc2o_inverse_quant(int(v), int(q)) == iq
        assert L.vc2o_forward_quant(int(v), int(q)) == fq
    for n, r in zip(z["intlog2_n"], z["intlog2"]):
        assert L.vc2o_intlog2(int(n)) == r
    for v, r in zip(z["sint_vals"], z["sint_len"]):
        assert L.vc2o_signed_exp_golomb_length(int(v)) == r


def test_expgolomb_blocks():
    z = load("expgolomb")
    data, meta, out = z["data"], z["meta"], z["out"]
    dpos = opos = 0
    for pre, bits_left, nread, nbytes in meta:
        blob = np.ascontiguousarray(data[dpos:dpos + nbytes])
        exp = out[opos:opos + nread]
        got = np.zeros(nread, np.int64)
        st = O.lib().vc2o_read_sintb_block(blob, int(nbytes), int(pre), int(bits_left), got, int(nread), None)
        assert st == 0
        assert (got == exp).all()
        dpos += nbytes
        opos += nread


def test_lifting():
    z = load("lifting")
    for k in range(int(z["count"][0])):
        a = z["in_%d" % k]
        filt = int(z["filt_%d" % k][0])
        s = a.copy()
        O.lib().vc2o_oned_synthesis(s, 1, s.size, filt)
        assert (s == z["synth_%d" % k]).all()
        an = a.copy()
        O.lib().vc2o_oned_analysis(an, 1, an.size, filt)
        assert (an == z["anal_%d" % k]).all()


def test_transform():
    z = load("transform")
    for k in range(int(z["count"][0])):
        wi, wiho, d, dho, W, H = [int(v) for v in z["case_%d" % k]]
        p = O.make_params(W, H, wavelet_index=wi, wavelet_index_ho=wiho, dwt_depth=d, dwt_depth_ho=dho)
        assert (O.idwt(p, 0, z["coeffs_%d" % k].astype(np.int64)) == z["idwt_%d" % k]).all()
        assert (O.dwt(p, 0, z["pic_%d" % k].astype(np.int64)) == z["dwt_%d" % k]).all()


def test_dc_prediction():
    z = load("dc_prediction")
    for k in range(int(z["count"][0])):
        band = z["in_%d" % k].astype(np.int64)
        assert (O.dc_prediction(band) == z["pred_%d" % k]).all()
        assert (O.apply_dc_prediction(band) == z["apply_%d" % k]).all()
        assert (O.dc_prediction(O.apply_dc_prediction(band)) == band).all()


def test_quantize_to_fit():
    z = load("quantize_to_fit")
    f = O.lib().vc2o_quantize_to_fit
    f.restype = C.c_int
    f.argtypes = [C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
    for k in range(int(z["count"][0])):
        nsets, align, target, minq, qi = [int(v) for v in z["meta_%d" % k]]
        cs = [np.ascontiguousarray(z["c_%d_%d" % (k, s)]) for s in range(nsets)]
        qms = [np.ascontiguousarray(z["qm_%d_%d" % (k, s)]) for s in range(nsets)]
        outs = [np.zeros(c.size, np.int64) for c in cs]
        carr = (C.c_void_p * nsets)(*[c.ctypes.data for c in cs])
        qarr = (C.c_void_p * nsets)(*[q.ctypes.data for q in qms])
        narr = (C.c_size_t * nsets)(*[c.size for c in cs])
        oarr = (C.c_void_p * nsets)(*[o.ctypes.data for o in outs])
        assert f(target, nsets, carr, qarr, narr, align, minq, oarr) == qi
        for s in range(nsets):
            assert (outs[s] == z["q_%d_%d" % (k, s)]).all()
            assert O.lib().vc2o_calculate_coeffs_bits(outs[s], outs[s].size) == int(z["bits_%d_%d" % (k, s)][0])


@pytest.mark.parametrize("name", picture_case_names())
def test_picture_decode(name):
    z = load("pictures")
    p = params_from_arrays(z[name + ".params"], z[name + ".quant_matrix"])
    data = z[name + ".data"].tobytes()
    st, coeffs, qindex, slen, consumed, bad = O.transform_data(p, data)
    assert st == 0 and consumed == len(data)
    assert (qindex == z[name + ".qindex"]).all()
    for c in range(3):
        assert (coeffs[c] == z[name + ".dec_coeffs%d" % c]).all()
        assert (O.decode_component(p, c, coeffs[c]) == z[name + ".decoded%d" % c]).all()
    for v in corrupt_variants(z, name):
        bdata = z[v + ".data"].tobytes()
        st, coeffs, qindex, slen, consumed, bad = O.transform_data(p, bdata)
        assert st == int(z[v + ".status"][0])
        if st == 0:
            assert consumed == int(z[v + ".consumed"][0])
            for c in range(3):
                assert (coeffs[c] == z[v + ".dec_coeffs%d" % c]).all()
                assert (O.decode_component(p, c, coeffs[c]) == z[v + ".decoded%d" % c]).all()


@pytest.mark.parametrize("name", picture_case_names())
def test_picture_encode(name):
    z = load("pictures")
    p = params_from_arrays(z[name + ".params"], z[name + ".quant_matrix"])
    pics = [z[name + ".pic%d" % c] for c in range(3)]
    coeffs = O.encode_picture(p, pics)
    for c in range(3):
        assert (coeffs[c] == z[name + ".enc_coeffs%d" % c]).all()
    data = z[name + ".data"].tobytes()
    arg = int(z[name + ".arg"][0])
    if p.is_ld:
        got, q = O.encode_ld_lossy(p, coeffs)
    elif "lossless" in name:
        got, q = O.encode_hq_fixed(p, coeffs, 0), np.zeros(p.slices_x * p.slices_y, np.int32)
    elif p.slice_prefix_bytes == 0:
        got, q = O.encode_hq_lossy(p, coeffs, arg)
    else:
        pytest.skip("reference stream carries random prefix bytes")
    assert got == data
    assert (q == z[name + ".qindex"]).all()
