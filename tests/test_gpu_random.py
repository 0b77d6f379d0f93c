"""CUDA path against the C oracle on seeded random inputs (sizes the oracle finishes in seconds),
plus size-independent properties at the BASELINE.json sizes."""

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def _mk(**kw):
    from vc2_conformance_b200.params import make_params

    kw.setdefault("quant_matrix", synth.default_quant_matrix(kw.get("dwt_depth", 3), kw.get("dwt_depth_ho", 0)))
    return make_params(**kw)


def _oracle_streams(p, pictures, picture_bytes_list):
    import oracle as O

    po = synth.oracle_params(p)
    streams, decoded = [], []
    for planes, pb in zip(pictures, picture_bytes_list):
        coeffs = O.encode_picture(po, planes)
        if p.is_ld:
            data, q = O.encode_ld_lossy(po, coeffs)
        elif pb is None:
            data = O.encode_hq_fixed(po, coeffs, 0)
        else:
            data, q = O.encode_hq_lossy(po, coeffs, pb)
        st, out, consumed = O.decode_picture(po, data)
        assert st == 0 and consumed == len(data)
        streams.append(data)
        decoded.append(out)
    return streams, decoded


HQ_CASES = [
    # W, H, cw, ch, depth, wi, wiho, d, dho, sx, sy, scaler
    (320, 192, 160, 192, 10, 0, 0, 3, 0, 10, 12, 1),
    (352, 288, 176, 144, 8, 1, 1, 3, 0, 11, 9, 1),
    (256, 128, 256, 128, 12, 5, 6, 2, 3, 2, 8, 2),
    (200, 120, 100, 120, 10, 2, 2, 2, 0, 7, 5, 1),
    (192, 96, 96, 96, 16, 6, 6, 2, 1, 6, 6, 3),
    (128, 64, 64, 64, 10, 4, 3, 1, 1, 4, 4, 1),
]


@pytest.mark.parametrize("case", HQ_CASES)
def test_hq_batch_vs_oracle(case, unpack_kernel):
    from vc2_conformance_b200.device import BatchCodec

    import oracle as O

    W, H, cw, ch, depth, wi, wiho, d, dho, sx, sy, scaler = case
    S = W * H + 2 * cw * ch
    # scaler large enough for the biggest picture_bytes used below (encoder/pictures.py:586)
    scaler = max(scaler, O.safe_lossy_hq_slice_size_scaler(2 * S, sx * sy))
    p = _mk(luma_width=W, luma_height=H, color_diff_width=cw, color_diff_height=ch, luma_depth=depth,
            wavelet_index=wi, wavelet_index_ho=wiho, dwt_depth=d, dwt_depth_ho=dho, slices_x=sx, slices_y=sy,
            slice_size_scaler=scaler)
    S = W * H + 2 * cw * ch
    pics = [synth.noise_picture(p, 0), synth.ramp_picture(p, 1), synth.noise_picture(p, 2),
            synth.ramp_picture(p, 3), synth.noise_picture(p, 4)]
    # ~2 bits/sample, ~6 bits/sample, nearly lossless sizes
    pbs = [S // 4, S // 2, S, S // 3, S * 2]
    streams, decoded = _oracle_streams(p, pics, pbs)
    codec = BatchCodec(p, len(pics))
    codec.decode_streams(streams)
    for i in range(len(pics)):
        planes = codec.picture_planes(i)
        for c in range(3):
            assert (planes[c] == decoded[i][c]).all(), (case, i, c)
    # encoder mirror on the same pictures: bytes identical to the oracle's
    flat = np.stack([synth.flat_picture(pl) for pl in pics])
    for i, pb in enumerate(pbs):
        out = codec.encode_pictures(flat[i:i + 1], pb)
        assert out[0] == streams[i], (case, i)


LD_CASES = [
    (320, 192, 160, 192, 10, 4, 4, 2, 0, 10, 12, 5, 1),
    (128, 128, 64, 64, 8, 1, 1, 3, 0, 4, 8, 3, 1),
    (192, 64, 192, 64, 12, 3, 3, 1, 2, 6, 4, 7, 2),
    (96, 80, 48, 80, 10, 0, 0, 2, 0, 3, 5, 9, 4),
]


@pytest.mark.parametrize("case", LD_CASES)
def test_ld_batch_vs_oracle(case, unpack_kernel):
    from vc2_conformance_b200.device import BatchCodec
    from vc2_conformance_b200.params import copy_params

    W, H, cw, ch, depth, wi, wiho, d, dho, sx, sy, num, den = case
    S = W * H + 2 * cw * ch
    # bytes per picture = num/den of the raw size
    pbytes = (S * depth // 8) * den // (num * 2) * 2
    p = _mk(luma_width=W, luma_height=H, color_diff_width=cw, color_diff_height=ch, luma_depth=depth,
            wavelet_index=wi, wavelet_index_ho=wiho, dwt_depth=d, dwt_depth_ho=dho, slices_x=sx, slices_y=sy,
            is_ld=True, slice_bytes_numerator=pbytes, slice_bytes_denominator=sx * sy)
    pics = [synth.noise_picture(p, 10), synth.ramp_picture(p, 11), synth.ramp_picture(p, 12)]
    streams, decoded = _oracle_streams(p, pics, [None] * 3)
    codec = BatchCodec(p, len(pics))
    codec.decode_streams(streams)
    for i in range(len(pics)):
        planes = codec.picture_planes(i)
        for c in range(3):
            assert (planes[c] == decoded[i][c]).all(), (case, i, c)
    flat = np.stack([synth.flat_picture(pl) for pl in pics])
    out = codec.encode_pictures(flat, 0)
    for i in range(len(pics)):
        assert out[i] == streams[i], (case, i)


@pytest.mark.parametrize("wi", range(7))
@pytest.mark.parametrize("wiho,d,dho", [(None, 3, 0), (1, 2, 1), (6, 1, 2), (5, 0, 2), (0, 2, 2)])
def test_transforms_vs_oracle(wi, wiho, d, dho):
    import oracle as O
    from vc2_conformance_b200 import ops

    wiho = wi if wiho is None else wiho
    W, H = 360, 136  # padded by the transform for most depths
    p = _mk(luma_width=W, luma_height=H, wavelet_index=wi, wavelet_index_ho=wiho, dwt_depth=d, dwt_depth_ho=dho)
    po = synth.oracle_params(p)
    pw, ph = O.padded_dims(po, 0)
    rng = np.random.RandomState(wi * 31 + d * 7 + dho)
    coeffs = rng.randint(-900, 900, pw * ph).astype(np.int32)
    assert (ops.idwt_component(p, 0, coeffs) == O.idwt(po, 0, coeffs.astype(np.int64))).all()
    pic = rng.randint(-2048, 2048, (ph, pw)).astype(np.int32)
    got = ops.dwt_component(p, 0, pic)
    assert (got == O.dwt(po, 0, pic.astype(np.int64))).all()
    # dwt o idwt identity (tests/pseudocode/test_picture_encoding.py:113-142)
    assert (ops.idwt_component(p, 0, got) == pic).all()


def test_elementwise_vs_oracle():
    import oracle as O
    from vc2_conformance_b200 import ops

    rng = np.random.RandomState(5)
    v = rng.randint(-(1 << 17), 1 << 17, 5000).astype(np.int32)
    qi = rng.randint(0, 40, 5000).astype(np.int32)
    L = O.lib()
    exp = np.array([L.vc2o_inverse_quant(int(a), int(b)) for a, b in zip(v, qi)])
    ok = np.abs(exp) < (1 << 31)
    assert (ops.inverse_quant(v, qi)[ok] == exp[ok]).all()
    exp = np.array([L.vc2o_forward_quant(int(a), int(b)) for a, b in zip(v, qi)])
    assert (ops.forward_quant(v, qi) == exp).all()
    x = rng.randint(-3000, 3000, 4000).astype(np.int32)
    assert (ops.clip_offset(x, 10, ops.CLIP | ops.ADD_OFFSET) == np.clip(x, -512, 511) + 512).all()
    assert (ops.clip_offset(x, 10, ops.REMOVE_OFFSET) == x - 512).all()


FULL_CASES = [
    # BASELINE.json configs at full size: (name, params kwargs, picture_bytes or None)
    ("cfg1_1080p_legall_d3_8x8", dict(luma_width=1920, luma_height=1080, color_diff_width=960,
                                      color_diff_height=1080, luma_depth=10, wavelet_index=1, dwt_depth=3,
                                      slices_x=8, slices_y=8, slice_size_scaler=1024)),
    ("cfg1_1080p_legall_d3_60x135", dict(luma_width=1920, luma_height=1080, color_diff_width=960,
                                         color_diff_height=1080, luma_depth=10, wavelet_index=1, dwt_depth=3,
                                         slices_x=60, slices_y=135, slice_size_scaler=8)),
    ("cfg2_2160p_dd97_d4", dict(luma_width=3840, luma_height=2160, color_diff_width=1920,
                                color_diff_height=2160, luma_depth=10, wavelet_index=0, dwt_depth=4,
                                slices_x=120, slices_y=135, slice_size_scaler=16)),
    ("cfg4_1080p_fidelity_daub_444", dict(luma_width=1920, luma_height=1080, luma_depth=12, wavelet_index=5,
                                          wavelet_index_ho=6, dwt_depth=2, dwt_depth_ho=3, slices_x=60,
                                          slices_y=270, slice_size_scaler=8)),
    # BASELINE config 5 (the encoder path): 7680x4320 12-bit 4:4:4, 240x270 slices, both filters SURVEY 8(d) names
    ("cfg5_8k_dd97_d4_444", dict(luma_width=7680, luma_height=4320, luma_depth=12, wavelet_index=0, dwt_depth=4,
                                 slices_x=240, slices_y=270, slice_size_scaler=32)),
    ("cfg5_8k_legall_d3_444", dict(luma_width=7680, luma_height=4320, luma_depth=12, wavelet_index=1, dwt_depth=3,
                                   slices_x=240, slices_y=270, slice_size_scaler=32)),
]


@pytest.mark.parametrize("name,kw", FULL_CASES)
def test_full_size_roundtrip_and_oracle(name, kw):
    """At full size: (1) lossless encode -> decode is the identity on device (round trip through
    DWT, qindex search, packing, slice index, unpack, IDWT); (2) a lossy stream decodes to exactly
    the oracle's picture; (3) the device encoder emits the oracle's bytes."""
    import oracle as O
    from vc2_conformance_b200.device import BatchCodec

    p = _mk(**kw)
    po = synth.oracle_params(p)
    ns = p.slices_x * p.slices_y
    S = p.luma_width * p.luma_height + 2 * p.color_diff_width * p.color_diff_height
    codec = BatchCodec(p, 2, max_stream_bytes=4 * S)
    pics = [synth.noise_picture(p, 0), synth.ramp_picture(p, 1)]
    flat = np.stack([synth.flat_picture(pl) for pl in pics])
    # (1) room for everything at qindex 0
    big = ns * (4 + 255 * p.slice_size_scaler)
    streams = codec.encode_pictures(flat, big)
    assert (codec.qindex[: 2 * ns].cpu().numpy() == 0).all()
    out = codec.decode_streams(streams).cpu().numpy()
    assert (out == flat).all()
    # (2)+(3) lossy at ~2.5 bits/sample against the oracle
    pb = (S * 5 // 16) // 4 * 4
    coeffs = O.encode_picture(po, pics[0])
    data, q = O.encode_hq_lossy(po, coeffs, pb)
    st, dec, consumed = O.decode_picture(po, data)
    assert st == 0
    got = codec.decode_streams([data])
    planes = codec.picture_planes(0)
    for c in range(3):
        assert (planes[c] == dec[c]).all(), (name, c)
    enc = codec.encode_pictures(flat[0:1], pb)
    assert enc[0] == data
    # the encoder stages against the oracle: picture_encode coefficients and the per-slice qindex
    got_coeffs = codec.coeff_components(0)
    for c in range(3):
        assert (got_coeffs[c] == coeffs[c]).all(), (name, "picture_encode", c)
    assert (codec.qindex[:ns].cpu().numpy() == np.asarray(q)).all(), (name, "qindex")
    # forward_quant of the whole picture at the chosen indices (quantize_coeffs, encoder/pictures.py:251):
    # inverse-quantising it must give what the decoder unpacked from the stream
    quantised = codec.quantized_host(1)[0]
    codec.unpack_streams([data])
    unpacked = np.concatenate(codec.coeff_components(0))
    nz = quantised != 0
    assert ((unpacked != 0) == nz).all(), (name, "quantised zero pattern")
    assert (np.sign(unpacked) == np.sign(quantised)).all(), (name, "quantised signs")


@pytest.mark.parametrize("name,kw", [FULL_CASES[1], FULL_CASES[2]])
def test_full_size_lossless_encoder(name, kw):
    """make_transform_data_hq_lossless at full size: minimal length fields, slices of different sizes,
    packed on the device; decoding gives back the pictures and the oracle agrees on the bytes."""
    import oracle as O
    from vc2_conformance_b200.device import BatchCodec
    from vc2_conformance_b200.params import copy_params

    p = _mk(**kw)
    S = p.luma_width * p.luma_height + 2 * p.color_diff_width * p.color_diff_height
    codec = BatchCodec(p, 2, max_stream_bytes=4 * S)
    pics = [synth.noise_picture(p, 7), synth.ramp_picture(p, 8)]
    flat = np.stack([synth.flat_picture(pl) for pl in pics])
    scaler, streams = codec.encode_pictures_lossless(flat)
    ps = copy_params(p, slice_size_scaler=scaler)
    dec = BatchCodec(ps, 2, max_stream_bytes=sum(len(s) for s in streams) + 64)
    out = dec.decode_streams(streams).cpu().numpy()
    assert (out == flat).all()
    # the oracle's fixed-qindex-0 serialisation at the same scaler writes the same bytes
    po = synth.oracle_params(ps)
    assert O.encode_hq_fixed(po, O.encode_picture(po, pics[1]), 0) == streams[1]


def test_full_size_ld_field():
    """BASELINE config 3: 1080i LD field, Haar-with-shift depth 2, 60x135 slices, 4:1."""
    import oracle as O
    from vc2_conformance_b200.device import BatchCodec

    S = 1920 * 540 + 2 * 960 * 540
    pbytes = S * 10 // 8 // 4
    p = _mk(luma_width=1920, luma_height=540, color_diff_width=960, color_diff_height=540, luma_depth=10,
            wavelet_index=4, dwt_depth=2, slices_x=60, slices_y=135, is_ld=True,
            slice_bytes_numerator=pbytes, slice_bytes_denominator=60 * 135)
    po = synth.oracle_params(p)
    pics = [synth.ramp_picture(p, 3), synth.noise_picture(p, 4)]
    streams, decoded = _oracle_streams(p, pics, [None, None])
    codec = BatchCodec(p, 2)
    codec.decode_streams(streams)
    for i in range(2):
        planes = codec.picture_planes(i)
        for c in range(3):
            assert (planes[c] == decoded[i][c]).all(), (i, c)
    flat = np.stack([synth.flat_picture(pl) for pl in pics])
    out = codec.encode_pictures(flat, 0)
    assert out[0] == streams[0] and out[1] == streams[1]


COEFF16_CASES = [
    # W, H, cw, ch, depth, wavelet, dwt_depth, slices_x, slices_y, int16 plane expected
    (1920, 1080, 960, 1080, 10, 0, 3, 60, 135, True),    # several strips across and down
    (1000, 600, 500, 600, 10, 0, 2, 25, 75, False),      # level-1 bands 250 wide: not vector aligned
    (1056, 600, 528, 600, 10, 0, 2, 33, 75, True),       # ragged right / bottom strips
    (1000, 600, 500, 300, 12, 1, 3, 25, 75, False),      # chroma bands 63 wide: not vector aligned
    (3840, 272, 1920, 272, 10, 0, 4, 120, 17, True),     # the bench geometry (32 x 16 slices), padded height
    (480, 1000, 480, 1000, 8, 1, 2, 15, 125, True),      # LeGall
    (512, 256, 256, 256, 10, 2, 2, 16, 32, True),        # Deslauriers-Dubuc (13,7)
    (512, 256, 256, 256, 10, 4, 3, 16, 32, True),        # Haar with shift
    (512, 256, 256, 256, 10, 6, 2, 16, 32, True),        # Daubechies (9,7)
    (456, 130, 232, 66, 10, 0, 2, 19, 11, False),        # uneven slices, padded dimensions
    (640, 480, 320, 480, 10, 0, 2, 4, 2, False),         # large slices: warp-per-slice unpack, int32 plane
]


@pytest.mark.parametrize("case", COEFF16_CASES)
@pytest.mark.parametrize("group", [128, 2])
def test_coeff16_plane(case, group):
    """vc2b_unpack16 + vc2b_idwt16 (high-pass coefficients staged as int16) give the oracle's coefficients and
    pictures; parameter sets the plane does not support quietly use the int32 plane."""
    import oracle as O
    from vc2_conformance_b200.device import BatchCodec

    W, H, cw, ch, depth, wi, d, sx, sy, expect16 = case
    S = W * H + 2 * cw * ch
    p = _mk(luma_width=W, luma_height=H, color_diff_width=cw, color_diff_height=ch, luma_depth=depth,
            wavelet_index=wi, dwt_depth=d, slices_x=sx, slices_y=sy,
            slice_size_scaler=O.safe_lossy_hq_slice_size_scaler(2 * S, sx * sy))
    po = synth.oracle_params(p)
    pics = [synth.noise_picture(p, 40), synth.ramp_picture(p, 41), synth.noise_picture(p, 42)]
    streams, decoded = _oracle_streams(p, pics, [S, S // 2, S // 3])
    codec = BatchCodec(p, len(pics), group=group, coeff16=True)
    assert codec.coeff16 == expect16
    codec.pictures.view(codec.torch.int16).fill_(-1)
    codec.decode_streams(streams)
    for i in range(len(pics)):
        st, comps, _q, _c = O.transform_data(po, streams[i])[:4]
        got = codec.coeff_components(i)
        planes = codec.picture_planes(i)
        for c in range(3):
            assert (got[c] == np.asarray(comps[c]).reshape(-1)).all(), (case, "coefficients", i, c)
            assert (planes[c] == decoded[i][c]).all(), (case, "picture", i, c)


COEFF16_EXTRA = [
    # LD (interleaved chroma block, DC prediction on the int32 level-0 band), several slice sizes
    dict(luma_width=960, luma_height=272, color_diff_width=480, color_diff_height=272, luma_depth=10, wavelet_index=4,
         dwt_depth=2, slices_x=30, slices_y=34, is_ld=1, slice_bytes_numerator=30 * 34 * 90 + 11,
         slice_bytes_denominator=30 * 34),
    dict(luma_width=512, luma_height=256, color_diff_width=512, color_diff_height=256, luma_depth=12, wavelet_index=1,
         dwt_depth=3, slices_x=16, slices_y=16, is_ld=1, slice_bytes_numerator=16 * 16 * 400 + 3,
         slice_bytes_denominator=16 * 16),
    # HQ with slice prefix bytes and a slice_size_scaler above 1
    dict(luma_width=512, luma_height=256, color_diff_width=256, color_diff_height=256, luma_depth=10, wavelet_index=0,
         dwt_depth=3, slices_x=16, slices_y=32, slice_prefix_bytes=3, slice_size_scaler=4),
]


@pytest.mark.parametrize("kw", COEFF16_EXTRA)
def test_coeff16_plane_ld_and_prefix(kw):
    """The int16 coefficient plane on LD pictures and on HQ pictures with prefix bytes / a scaler: coefficients and
    pictures equal the oracle's, and the int16 kernels are the ones that ran."""
    import oracle as O
    from vc2_conformance_b200.device import BatchCodec

    p = _mk(**kw)
    po = synth.oracle_params(p)
    S = p.luma_width * p.luma_height + 2 * p.color_diff_width * p.color_diff_height
    pics = [synth.noise_picture(p, 60), synth.ramp_picture(p, 61), synth.noise_picture(p, 62)]
    streams, decoded = _oracle_streams(p, pics, [S // 2, S // 3, S // 4])
    codec = BatchCodec(p, len(pics), coeff16=True)
    assert codec.coeff16
    host, offs = codec.pack_host_streams(streams)
    codec.decode_host(host, offs, len(pics))
    st = codec.status_host(len(pics))
    assert (st[:, 0] == 0).all(), st   # nothing was redone through the int32 planes
    for i in range(len(pics)):
        comps = O.transform_data(po, streams[i])[1]
        got = codec.coeff_components(i)
        planes = codec.picture_planes(i)
        for c in range(3):
            assert (got[c] == np.asarray(comps[c]).reshape(-1)).all(), ("coefficients", i, c)
            assert (planes[c] == decoded[i][c]).all(), ("picture", i, c)


def test_coeff16_overflow_is_redone_through_the_int32_plane():
    """A lossless 16-bit noise picture has high-pass coefficients beyond int16 (flagged by the unpack), a 16-bit
    ramp has intermediate low-pass bands beyond int16 (flagged by the IDWT level that produces them): both come
    back as VC2B_ST_COEFF16_RANGE and decode_streams / SequenceDecoder.finish decode them again through the
    int32 planes; a picture that fits is left alone."""
    from vc2_conformance_b200 import capi
    from vc2_conformance_b200.device import BatchCodec, SequenceDecoder

    W, H = 512, 256
    p = _mk(luma_width=W, luma_height=H, color_diff_width=W // 2, color_diff_height=H, luma_depth=16,
            wavelet_index=0, dwt_depth=3, slices_x=16, slices_y=32, slice_size_scaler=8)
    rng = np.random.default_rng(7)
    quiet = [(32768 + rng.integers(-400, 400, size=(H, w))).astype(np.uint16) for w in (W, W // 2, W // 2)]
    pics = [synth.noise_picture(p, 50), synth.ramp_picture(p, 51), quiet, synth.noise_picture(p, 53)]
    streams, decoded = _oracle_streams(p, pics, [None] * len(pics))
    codec = BatchCodec(p, len(pics), coeff16=True)
    assert codec.coeff16
    host, offs = codec.pack_host_streams(streams)
    codec.decode_host(host, offs, len(pics))
    st = codec.status_host(len(pics))
    assert list(st[:, 0]) == [capi.ST_COEFF16_RANGE, capi.ST_COEFF16_RANGE, capi.ST_OK, capi.ST_COEFF16_RANGE]
    assert codec.redo_coeff16(st) == [0, 1, 3]
    assert (st[:, 0] == capi.ST_OK).all()
    for i in range(len(pics)):
        planes = codec.picture_planes(i)
        for c in range(3):
            assert (planes[c] == decoded[i][c]).all(), (i, c)

    # the pipelined host-to-host decoder
    torch = codec.torch
    seq = SequenceDecoder(p, chunk=3, nstreams=2, max_chunk_bytes=sum(len(s) for s in streams) + 256)
    chunks = seq.plan(host, offs.numpy())
    out = torch.empty((len(pics), codec.g.picture_samples), dtype=torch.uint16).pin_memory()
    seq.decode(chunks, out)
    seq.synchronize()
    st = seq.finish(chunks, out)
    assert (st[:, 0] == capi.ST_OK).all()
    flat = out.numpy()
    for i in range(len(pics)):
        assert (flat[i] == synth.flat_picture(decoded[i])).all(), i


@pytest.mark.parametrize("mag_bits", [12, 27, 28, 29, 30])
@pytest.mark.parametrize("is_ld", [False, True])
def test_quantiser_search_and_packer_on_large_coefficients(mag_bits, is_ld):
    """quantize_fit_kernel's division-free probe is only valid for |c| < 2^28 (kFastMagLimit): slices with
    larger magnitudes must take the exact path, and pack_slices_kernel's 32-bit code builder must hand
    over to the 64-bit one for long codes.  Coefficients are written straight into the codec's
    coefficient buffer; bytes and qindex must equal the oracle's."""
    import oracle as O
    from vc2_conformance_b200.device import BatchCodec

    if is_ld and mag_bits >= 30:
        pytest.skip("the DC prediction residues of 2^30 magnitudes leave int32 (envelope flag, tested elsewhere)")
    W, H, sx, sy = 64, 32, 4, 4
    S = W * H * 2
    picture_bytes = S // 2
    kw = dict(luma_width=W, luma_height=H, color_diff_width=W // 2, color_diff_height=H, luma_depth=10,
              wavelet_index=1, dwt_depth=2, slices_x=sx, slices_y=sy)
    if is_ld:
        kw.update(is_ld=True, slice_bytes_numerator=picture_bytes, slice_bytes_denominator=sx * sy)
    else:
        kw.update(slice_size_scaler=O.safe_lossy_hq_slice_size_scaler(picture_bytes, sx * sy))
    p = _mk(**kw)
    po = synth.oracle_params(p)
    codec = BatchCodec(p, 2)
    g = codec.g
    rng = np.random.RandomState(mag_bits)
    lim = (1 << mag_bits) - 1
    comps_all = []
    for i in range(2):
        comps = [rng.randint(-lim, lim + 1, n).astype(np.int64) for n in g.comp_coeffs]
        if i == 1:  # mostly small values with a few large ones: some slices fast, some exact
            comps = [np.where(rng.rand(c.size) < 0.02, c, c % 7 - 3) for c in comps]
        comps_all.append(comps)
        flat = np.concatenate(comps).astype(np.int32)
        codec.coeffs[i * g.picture_coeffs:(i + 1) * g.picture_coeffs].copy_(codec.torch.from_numpy(flat))
    codec.quantize_device(2, 0 if is_ld else picture_bytes)
    out, nbytes = codec.pack_device(2, 0 if is_ld else picture_bytes)
    got = out.cpu().numpy()
    qi = codec.qindex[: 2 * sx * sy].cpu().numpy().reshape(2, -1)
    for i in range(2):
        comps = comps_all[i]
        if is_ld:
            # the oracle's LD encoder applies the DC prediction itself; quantize_device did so in place
            want, wq = O.encode_ld_lossy(po, comps)
        else:
            want, wq = O.encode_hq_lossy(po, comps, picture_bytes)
        assert (qi[i] == wq).all(), (mag_bits, is_ld, i)
        assert got[i, :nbytes].tobytes() == want, (mag_bits, is_ld, i)
