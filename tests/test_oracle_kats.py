"""Known-answer values quoted from the reference's OWN test-suite (file:line cited per test,
relative to /root/reference/tests/), checked against the C oracle.  CPU only."""

import numpy as np
import pytest

import oracle as O


def _read(encoded, n=1, bit_offset=0, bits_left=None):
    data = np.frombuffer(encoded, np.uint8).copy()
    out = np.zeros(n, np.int64)
    if bits_left is None:
        bits_left = 8 * data.size
    st = O.lib().vc2o_read_sintb_block(data, data.size, bit_offset, bits_left, out, n, None)
    return st, list(out)


# bitstream/test_bitstream_io.py:618-640 (TestReadSint.test_basic_reading)
@pytest.mark.parametrize(
    "encoded,exp",
    [
        (b"\x80", 0), (b"\x20", 1), (b"\x60", 2), (b"\x08", 3), (b"\x18", 4), (b"\x48", 5),
        (b"\x58", 6), (b"\x02\x00", 7), (b"\x30", -1), (b"\x70", -2), (b"\x0C", -3), (b"\x1C", -4),
        (b"\x4C", -5), (b"\x5C", -6), (b"\x03\x00", -7),
    ],
)
def test_read_sint_vectors(encoded, exp):
    st, out = _read(encoded)
    assert st == 0 and out == [exp]


# bitstream/test_exp_golomb.py:8-66 (lengths: 2*floor(log2(v+1))+1, +1 sign when non-zero)
@pytest.mark.parametrize("v,exp", [(0, 1), (1, 4), (2, 4), (3, 6), (6, 6), (7, 8), (-1, 4), (-7, 8), (255, 18)])
def test_signed_exp_golomb_length(v, exp):
    assert O.lib().vc2o_signed_exp_golomb_length(v) == exp


# vc2_conformance/test_cases/decoder/pictures.py:351-381: a value cut by the end of a bounded
# block is completed with 1-bits (data bits 1, stop, sign=1 => negative); later values are 0.
def test_dangling_value_is_completed_with_ones():
    # "0 0 0" then block ends: pairs (0,0) then follow=0 ... with ones: 0 0 0 [1] [1] => value
    # bits: follow0 data0 follow0 data1(ext) stop(ext) sign(ext): value=(1<<2|0<<1|1)-1=4, neg
    st, out = _read(b"\x00", n=3, bits_left=3)
    assert st == 0 and out == [-4, 0, 0]
    # empty block: everything decodes to zero without consuming bits
    st, out = _read(b"\x00", n=4, bits_left=0)
    assert st == 0 and out == [0, 0, 0, 0]


# decoder/test_decoder_io.py:30-52: reading past the end of the file raises UnexpectedEndOfStream
def test_unexpected_end_of_stream():
    st, out = _read(b"\x00", n=1, bits_left=64)
    assert st == O.UNEXPECTED_END_OF_STREAM


# encoder/test_pictures.py:226-246 (test_calculate_coeffs_bits)
@pytest.mark.parametrize(
    "coeffs,exp",
    [([], 0), ([1], 4), ([3], 6), ([-1], 4), ([1, 3], 10), ([0, 1, 3], 11), ([1, 0, 3, 0], 11),
     ([1, 0, 3, 0, 0, 0], 11), ([0, 0, 0], 0)],
)
def test_calculate_coeffs_bits(coeffs, exp):
    a = np.array(coeffs + [0], np.int64)
    assert O.lib().vc2o_calculate_coeffs_bits(a, len(coeffs)) == exp


# encoder/test_pictures.py:273-286 (test_quantize_coeffs)
@pytest.mark.parametrize(
    "qi,coeffs,qm,exp",
    [(0, [1, 2, 3, 4], [0, 0, 0, 0], [1, 2, 3, 4]), (4, [1, 2, 3, 4], [0, 0, 0, 0], [0, 1, 1, 2]),
     (8, [8, 8, 8], [0, 4, 8], [2, 4, 8])],
)
def test_quantize_coeffs(qi, coeffs, qm, exp):
    got = [O.lib().vc2o_forward_quant(c, max(0, qi - m)) for c, m in zip(coeffs, qm)]
    assert got == exp


# pseudocode/test_quantization.py:7-13
def test_quant_roundtrip_error_bound():
    L = O.lib()
    for n in range(-100, 100):
        assert abs(n - L.vc2o_inverse_quant(L.vc2o_forward_quant(n, 12), 12)) < 8


# SURVEY.md 8(a6) / probe of pseudocode/quantization.py:40-62
def test_first_quant_factors():
    L = O.lib()
    assert [L.vc2o_quant_factor(i) for i in range(12)] == [4, 5, 6, 7, 8, 10, 11, 13, 16, 19, 23, 27]
    assert [L.vc2o_quant_offset(i) for i in range(12)] == [1, 2, 3, 4, 4, 5, 6, 7, 8, 10, 12, 14]


# encoder/test_pictures.py:455-562 (hand-evaluated Haar, horizontal-only depth 1, 3x2 slices)
def _haar_case(is_ld):
    p = O.make_params(12, 4, 6, 4, 8, 8, wavelet_index=3, wavelet_index_ho=3, dwt_depth=0, dwt_depth_ho=1,
                      slices_x=3, slices_y=2, is_ld=is_ld,
                      quant_matrix={0: {"L": 123}, 1: {"H": 321}})
    Y = (128 + np.arange(48).reshape(4, 12)).astype(np.uint16)
    C1 = (128 + np.arange(24).reshape(4, 6)).astype(np.uint16)
    C2 = (128 - np.arange(24).reshape(4, 6)).astype(np.uint16)
    return p, O.encode_picture(p, [Y, C1, C2])


def test_hand_evaluated_haar_slices():
    p, coeffs = _haar_case(False)
    y, qm = O.gather_slice(p, 0, coeffs[0], 0, 0)
    assert list(y) == [1, 3, 13, 15, 1, 1, 1, 1]
    assert list(qm) == [123] * 4 + [321] * 4
    assert list(O.gather_slice(p, 1, coeffs[1], 0, 0)[0]) == [1, 7, 1, 1]
    assert list(O.gather_slice(p, 2, coeffs[2], 0, 0)[0]) == [0, -6, -1, -1]
    assert list(O.gather_slice(p, 0, coeffs[0], 2, 1)[0]) == [33, 35, 45, 47, 1, 1, 1, 1]
    assert list(O.gather_slice(p, 1, coeffs[1], 2, 1)[0]) == [17, 23, 1, 1]
    assert list(O.gather_slice(p, 2, coeffs[2], 2, 1)[0]) == [-16, -22, -1, -1]


# encoder/test_pictures.py:564-606 (test_dc_prediction)
def test_hand_evaluated_dc_prediction():
    p, coeffs = _haar_case(True)
    for c in range(3):
        level, orient, w, h, off, total = O.layout(p, c)
        n0 = int(w[0] * h[0])
        coeffs[c][:n0] = O.apply_dc_prediction(coeffs[c][:n0].reshape(h[0], w[0])).reshape(-1)
    assert list(O.gather_slice(p, 0, coeffs[0], 0, 0)[0]) == [1, 2, 12, 9, 1, 1, 1, 1]
    assert list(O.gather_slice(p, 1, coeffs[1], 0, 0)[0]) == [1, 6, 1, 1]
    assert list(O.gather_slice(p, 2, coeffs[2], 0, 0)[0]) == [0, -6, -1, -1]


# pseudocode/test_picture_encoding.py:150-186 (dwt_pad_addition: replicate last column, then row)
def test_pad_addition_edge_replication():
    p = O.make_params(3, 3, wavelet_index=3, dwt_depth=1, dwt_depth_ho=1, luma_depth=8)
    pic = (128 + np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]])).astype(np.uint16)
    # padded to 4x4: [[1,2,3,3],[4,5,6,6],[7,8,9,9],[7,8,9,9]]; transform then invert => same
    coeffs = O.encode_component(p, 0, pic)
    back = O.idwt(p, 0, coeffs)
    assert (back == np.array([[1, 2, 3, 3], [4, 5, 6, 6], [7, 8, 9, 9], [7, 8, 9, 9]])).all()


# pseudocode/test_picture_encoding.py:113-142 (dwt o idwt identity for 9 depth pairs x 7 filters)
@pytest.mark.parametrize("filt", range(7))
@pytest.mark.parametrize("d,dho", [(0, 0), (1, 0), (2, 0), (0, 1), (0, 2), (1, 1), (2, 1), (1, 2), (2, 2)])
def test_dwt_idwt_roundtrip(filt, d, dho):
    W, H = 32, 16
    p = O.make_params(W, H, wavelet_index=filt, dwt_depth=d, dwt_depth_ho=dho)
    pic = (np.arange(H)[:, None] * 100 + np.arange(W)[None, :]).astype(np.int64)
    assert (O.idwt(p, 0, O.dwt(p, 0, pic)) == pic).all()


# pseudocode/test_vc2_math.py:53-73 (mean rounding) via the DC predictor: mean(a,b,c)=(a+b+c+1)//3
def test_mean_floor_division_on_negatives():
    band = np.array([[-4, 0], [0, 0]], np.int64)
    # y=1,x=1 prediction = mean(left=band[1][0]+pred.., ...) evaluated sequentially
    got = O.dc_prediction(band)
    # row0: [-4, -4]; row1: [-4, mean(-4,-4,-4)=(-12+1)//3=-4]
    assert got.tolist() == [[-4, -4], [-4, -4]]
    band = np.array([[-1, 0], [-1, 0]], np.int64)
    # row0: [-1,-1]; row1: [-2, mean(-2,-1,-1)=(-4+1)//3=-1]
    assert O.dc_prediction(band).tolist() == [[-1, -1], [-2, -1]]


# decoder/test_transform_data_syntax.py:31-62 (slice_y_length bound)
def test_invalid_slice_y_length():
    p = O.make_params(8, 8, wavelet_index=3, dwt_depth=0, dwt_depth_ho=0, slices_x=1, slices_y=1, is_ld=True,
                      slice_bytes_numerator=4, slice_bytes_denominator=1, quant_matrix={0: {"LL": 0}})
    # 32 bits: 7 qindex + 5 length bits (intlog2(25)=5) => 20 left; length=21 is invalid, 20 is ok
    def stream(length):
        v = (0 << 25) | (length << 20)
        return v.to_bytes(4, "big")
    assert O.transform_data(p, stream(21))[0] == O.INVALID_SLICE_Y_LENGTH
    assert O.transform_data(p, stream(20))[0] == O.OK
