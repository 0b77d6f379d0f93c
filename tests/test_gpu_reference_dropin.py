"""The real drop-in: the UNMODIFIED reference (baseline/_ref or /root/reference, found by
oracle/ref_import.py) driven through ``vc2_conformance_b200.install``.

Every test runs the reference's own entry points -- ``encoder.make_sequence``,
``autofill_and_serialise_stream``, ``decoder.parse_stream`` -- once as shipped (pure Python) and once with
``install()`` active (CUDA kernels behind the rebound names) and requires identical results: pictures,
callbacks, serialised bytes, exception types and exception arguments.  Mirrors
/root/reference/tests/encoder/test_pictures.py:1355-1410 (test_lossless / test_lossy) and
tests/decoder/test_transform_data_syntax.py:31-62.
"""

import copy
from io import BytesIO

import numpy as np
import pytest

import ref_import

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_import.available(), reason="no reference tree (baseline/_ref)")]


@pytest.fixture(scope="module")
def ref():
    ref_import.install()
    import vc2_data_tables as T
    from vc2_conformance import bitstream, decoder, encoder
    from vc2_conformance.codec_features import CodecFeatures
    from vc2_conformance.pseudocode.state import State
    from vc2_conformance.pseudocode.video_parameters import VideoParameters

    class R(object):
        pass

    r = R()
    r.T, r.bitstream, r.decoder, r.encoder = T, bitstream, decoder, encoder
    r.CodecFeatures, r.State, r.VideoParameters = CodecFeatures, State, VideoParameters
    return r


@pytest.fixture
def gpu():
    import vc2_conformance_b200.install as I

    yield I
    I.uninstall()


def codec_features(ref, W=64, H=32, cdf=1, depth=8, wi=1, wiho=None, d=2, dho=0, sx=5, sy=7, profile="hq",
                   lossless=False, picture_bytes=500, fragment_slice_count=0, fields=False, qm=None):
    T = ref.T
    exc = (1 << depth) - 1
    wiho = wi if wiho is None else wiho
    if qm is None and (wi, wiho, d, dho) not in T.QUANTISATION_MATRICES:
        # no default matrix for this transform (decoder/picture_syntax.py:331): carry a custom one
        qm = {0: {"L" if dho else "LL": 2}}
        for level in range(1, dho + 1):
            qm[level] = {"H": level % 3}
        for level in range(dho + 1, dho + d + 1):
            qm[level] = {"HL": 1 + level % 2, "LH": level % 3, "HH": 3}
    return ref.CodecFeatures(
        name="basic", level=T.Levels.unconstrained,
        profile=T.Profiles.high_quality if profile == "hq" else T.Profiles.low_delay,
        picture_coding_mode=T.PictureCodingModes.pictures_are_fields if fields else T.PictureCodingModes.pictures_are_frames,
        video_parameters=ref.VideoParameters(
            frame_width=W, frame_height=H, color_diff_format_index=T.ColorDifferenceSamplingFormats(cdf),
            source_sampling=T.SourceSamplingModes.progressive, top_field_first=True, frame_rate_numer=1,
            frame_rate_denom=1, pixel_aspect_ratio_numer=1, pixel_aspect_ratio_denom=1, clean_width=W, clean_height=H,
            left_offset=0, top_offset=0, luma_offset=0, luma_excursion=exc, color_diff_offset=1 << (depth - 1),
            color_diff_excursion=exc, color_primaries_index=T.PresetColorPrimaries.hdtv,
            color_matrix_index=T.PresetColorMatrices.hdtv, transfer_function_index=T.PresetTransferFunctions.tv_gamma),
        wavelet_index=T.WaveletFilters(wi), wavelet_index_ho=T.WaveletFilters(wiho),
        dwt_depth=d, dwt_depth_ho=dho, slices_x=sx, slices_y=sy, fragment_slice_count=fragment_slice_count,
        lossless=lossless, picture_bytes=None if lossless else picture_bytes, quantization_matrix=qm)


def noise_pictures(cf, n=1, seed=0):
    vp = cf["video_parameters"]
    W, H = vp["frame_width"], vp["frame_height"]
    if cf["picture_coding_mode"] == 1:
        H //= 2
    cw = W // (2 if vp["color_diff_format_index"] in (1, 2) else 1)
    ch = H // (2 if vp["color_diff_format_index"] == 2 else 1)
    top = vp["luma_excursion"] + 1
    rand = np.random.RandomState(seed)
    return [{"Y": rand.randint(0, top, (H, W)).tolist(), "C1": rand.randint(0, top, (ch, cw)).tolist(),
             "C2": rand.randint(0, top, (ch, cw)).tolist(), "pic_num": 100 + i} for i in range(n)]


def serialise(ref, sequence):
    f = BytesIO()
    ref.bitstream.autofill_and_serialise_stream(f, ref.bitstream.Stream(sequences=[sequence]))
    return f.getvalue()


def decode(ref, data, file_type=BytesIO):
    out = []
    state = ref.State(_output_picture_callback=lambda pic, vp, pcm: out.append((copy.deepcopy(pic), dict(vp), pcm)))
    ref.decoder.init_io(state, file_type(data))
    ref.decoder.parse_stream(state)
    return out


class Unseekable(object):
    """A file-like object with read() and tell() only: all decoder/io.py needs."""

    def __init__(self, data):
        self._f = BytesIO(data)

    def read(self, n=-1):
        return self._f.read(n)

    def tell(self):
        return self._f.tell()


CONFIGS = [
    # test_lossless (tests/encoder/test_pictures.py:1355): symmetric, then asymmetric
    dict(lossless=True),
    dict(lossless=True, wi=3, wiho=1, d=2, dho=1),
    # test_lossy (:1412) style, both profiles, uneven 5 x 7 slices
    dict(picture_bytes=64 * 32 * 2 // 4),
    dict(profile="ld", picture_bytes=64 * 32 * 2 // 4),
    # every filter family the kernels special-case, custom depths, 4:4:4 / 4:2:0, fields, more bits
    dict(wi=0, d=3, sx=4, sy=2, depth=10, picture_bytes=1500),
    dict(wi=2, d=1, sx=8, sy=4, cdf=0, depth=12, picture_bytes=3000),
    dict(wi=4, d=2, sx=4, sy=4, profile="ld", cdf=2, fields=True, picture_bytes=700),
    dict(wi=6, d=2, sx=4, sy=4, depth=10, picture_bytes=1200),
    dict(wi=5, d=1, sx=2, sy=2, depth=10, picture_bytes=1500),
    dict(wi=0, d=1, dho=2, sx=4, sy=4, depth=10, picture_bytes=1100),
    # fragments
    dict(fragment_slice_count=3, picture_bytes=800),
    dict(profile="ld", fragment_slice_count=4, picture_bytes=600, sx=4, sy=4, wi=4),
    dict(lossless=True, fragment_slice_count=6),
]


@pytest.mark.parametrize("kw", CONFIGS, ids=lambda kw: ",".join("%s=%s" % kv for kv in sorted(kw.items())))
def test_encoder_and_decoder_drop_in(ref, gpu, kw):
    """make_sequence -> serialise -> parse_stream, the reference alone against the reference with the
    CUDA path installed: same bytes out of the encoder, same pictures out of the decoder (and, for
    lossless coding, the pictures that went in)."""
    cf = codec_features(ref, **kw)
    pictures = noise_pictures(cf, n=2)
    plain_bytes = serialise(ref, ref.encoder.make_sequence(cf, copy.deepcopy(pictures)))
    plain_out = decode(ref, plain_bytes)
    assert len(plain_out) == 2

    from vc2_conformance_b200 import capi

    before = capi.lib().vc2b_launch_count()
    rebound = gpu.install()
    assert ("vc2_conformance.encoder.pictures", "quantize_to_fit") in rebound
    assert ("vc2_conformance.encoder.pictures", "make_transform_data_hq_lossy") in rebound
    gpu_bytes = serialise(ref, ref.encoder.make_sequence(cf, copy.deepcopy(pictures)))
    assert gpu_bytes == plain_bytes, "the encoder with install() serialises different bytes"
    enc_launches = capi.lib().vc2b_launch_count() - before
    assert enc_launches > 0, "the encoder did not reach the CUDA library"
    gpu_out = decode(ref, plain_bytes)
    assert capi.lib().vc2b_launch_count() - before > enc_launches, "the decoder did not reach the CUDA library"
    assert gpu_out == plain_out
    assert decode(ref, plain_bytes, Unseekable) == plain_out, "reader without seek()"
    if kw.get("lossless"):
        assert [o[0] for o in gpu_out] == pictures


def _function_level_cases():
    rng = np.random.RandomState(5)
    for n, scale in [(1, 3), (7, 40), (64, 300), (300, 5000), (33, 1 << 20)]:
        vals = (rng.randint(-scale, scale + 1, n) * (rng.rand(n) < 0.7)).tolist()
        qms = rng.randint(0, 9, n).tolist()
        yield vals, qms


def test_quantiser_functions_drop_in(ref, gpu):
    """calculate_coeffs_bits / quantize_coeffs / quantize_to_fit / forward_quant on the reference's lists
    (tests/encoder/test_pictures.py:226-360)."""
    from vc2_conformance.encoder import pictures as EP

    plain = []
    cases = list(_function_level_cases())
    for vals, qms in cases:
        row = [EP.calculate_coeffs_bits(vals), EP.quantize_coeffs(9, vals, qms)]
        for target, align, minq in [(0, 1, 0), (40, 8, 0), (200, 16, 3), (10 ** 6, 1, 0), (17, 1, 11)]:
            sets = [EP.ComponentCoeffs(vals, qms), EP.ComponentCoeffs(vals[::-1], qms[::-1]), EP.ComponentCoeffs([], [])]
            row.append(EP.quantize_to_fit(target, sets, align, minq))
        plain.append(row)
    assert EP.calculate_coeffs_bits([]) == 0
    gpu.install()
    assert EP.calculate_coeffs_bits([]) == 0
    assert EP.calculate_coeffs_bits([0, 0, 0]) == 0
    for (vals, qms), want in zip(cases, plain):
        got = [EP.calculate_coeffs_bits(vals), EP.quantize_coeffs(9, vals, qms)]
        for target, align, minq in [(0, 1, 0), (40, 8, 0), (200, 16, 3), (10 ** 6, 1, 0), (17, 1, 11)]:
            sets = [EP.ComponentCoeffs(vals, qms), EP.ComponentCoeffs(vals[::-1], qms[::-1]), EP.ComponentCoeffs([], [])]
            q, lists = EP.quantize_to_fit(target, sets, align, minq)
            got.append((q, lists))
        assert got[0] == want[0]
        assert got[1] == want[1]
        for g, w in zip(got[2:], want[2:]):
            assert g[0] == w[0] and [list(x) for x in g[1]] == [list(x) for x in w[1]]


def test_slice_coeffs_view_matches_reference(ref, gpu):
    """transform_and_slice_picture (encoder/pictures.py:350): the lazy [[SliceCoeffs]] view of the device
    coefficients equals the reference's lists, LD (DC prediction applied) and HQ."""
    from vc2_conformance.encoder import pictures as EP

    for kw in [dict(), dict(profile="ld", wi=4, sx=4, sy=4), dict(wi=0, d=1, dho=2, sx=3, sy=5, depth=10)]:
        cf = codec_features(ref, **kw)
        pic = noise_pictures(cf, seed=3)[0]
        want = EP.transform_and_slice_picture(cf, copy.deepcopy(pic))
        gpu.install()
        got = EP.transform_and_slice_picture(cf, copy.deepcopy(pic))
        gpu.uninstall()
        assert len(got) == len(want)
        for row_g, row_w in zip(got, want):
            assert len(row_g) == len(row_w)
            for sg, sw in zip(row_g, row_w):
                for c in range(3):
                    assert list(sg[c].coeff_values) == list(sw[c].coeff_values)
                    assert list(sg[c].quant_matrix_values) == list(sw[c].quant_matrix_values)


def _raises(fn):
    try:
        fn()
    except Exception as e:  # noqa: BLE001 -- the point is to compare whatever the reference raises
        return e
    return None


def _same_exception(a, b):
    assert type(a) is type(b), (a, b)
    assert a is None or vars(a) == vars(b), (vars(a), vars(b))
    if a is not None and hasattr(a, "explain"):
        assert a.explain() == b.explain()


@pytest.mark.parametrize("kw", [dict(picture_bytes=64 * 32 * 2 // 4), dict(profile="ld", picture_bytes=64 * 32 * 2 // 4),
                                dict(fragment_slice_count=3, picture_bytes=800),
                                dict(profile="ld", fragment_slice_count=4, picture_bytes=600, sx=4, sy=4, wi=4)],
                         ids=["hq", "ld", "hq_fragments", "ld_fragments"])
def test_truncated_streams_raise_the_reference_exceptions(ref, gpu, kw):
    """A stream cut inside the transform data: the same exception type (UnexpectedEndOfStream,
    decoder/io.py:197) with and without install(), and the same pictures before it."""
    cf = codec_features(ref, **kw)
    data = serialise(ref, ref.encoder.make_sequence(cf, noise_pictures(cf, n=2)))
    cuts = sorted(set([len(data) // 3, len(data) // 2, len(data) - 40, len(data) - 14, len(data) - 13]))
    plain = [_raises(lambda: decode(ref, data[:cut])) for cut in cuts]
    assert any(e is not None for e in plain)
    gpu.install()
    for cut, want in zip(cuts, plain):
        _same_exception(_raises(lambda: decode(ref, data[:cut])), want)
        _same_exception(_raises(lambda: decode(ref, data[:cut], Unseekable)), want)


@pytest.mark.parametrize("fragments", [0, 4], ids=["pictures", "fragments"])
def test_invalid_slice_y_length_carries_the_reference_arguments(ref, gpu, fragments):
    """InvalidSliceYLength(slice_y_length, slice_bytes, sx, sy) (decoder/transform_data_syntax.py:166-172,
    decoder/exceptions.py:1645): corrupt the length field of one LD slice."""
    cf = codec_features(ref, profile="ld", picture_bytes=600, sx=4, sy=4, wi=4, fragment_slice_count=fragments)
    seq = ref.encoder.make_sequence(cf, noise_pictures(cf, n=1))
    def corrupt(ld_slice):  # all-zero coefficients are 1-bits, which may run past a bounded block
        ld_slice["y_transform"] = [0] * len(ld_slice["y_transform"])
        ld_slice["c_transform"] = [0] * len(ld_slice["c_transform"])
        ld_slice["slice_y_length"] = 500

    for du in seq["data_units"]:
        if "picture_parse" in du:
            corrupt(du["picture_parse"]["wavelet_transform"]["transform_data"]["ld_slices"][6])
        if "fragment_parse" in du and du["fragment_parse"]["fragment_header"].get("fragment_slice_count"):
            if du["fragment_parse"]["fragment_header"]["fragment_y_offset"] == 1:
                corrupt(du["fragment_parse"]["fragment_data"]["ld_slices"][2])
    data = serialise(ref, seq)
    want = _raises(lambda: decode(ref, data))
    assert type(want).__name__ == "InvalidSliceYLength"
    gpu.install()
    got = _raises(lambda: decode(ref, data))
    _same_exception(got, want)
    assert (got.slice_y_length, got.sx, got.sy) == (500, 2, 1)
