"""
Stand-in for the un-vendored ``vc2_data_tables`` PyPI package (pinned
``>=0.1.1,<2.0`` by /root/reference/setup.py:48).  TEST INFRASTRUCTURE ONLY.

The reference's hot-path arithmetic needs exactly three things from that
package: the ``LiftingFilterTypes`` enum, the ``LIFTING_FILTERS`` table
(SMPTE ST 2042-1 Table 15.x lifting stages; call sites
/root/reference/vc2_conformance/pseudocode/picture_decoding.py:10,190,200,264
and picture_encoding.py:10,207,214) and the ``WaveletFilters`` indices.  They
are restated here from the published standard.  Everything else in this module
exists only so that the reference's modules *import*; header-level tables are
deliberately minimal and are never consulted by the picture path.

This package is only ever put on ``sys.path`` by ``oracle/ref_import.py``
(golden-vector generation inside the build container).  Nothing in the product
package or in the GPU tests imports it.
"""

from collections import namedtuple
from enum import IntEnum

__all__ = []  # populated at the bottom


class ParseCodes(IntEnum):
    sequence_header = 0x00
    end_of_sequence = 0x10
    auxiliary_data = 0x20
    padding_data = 0x30
    low_delay_picture = 0xC8
    high_quality_picture = 0xE8
    low_delay_picture_fragment = 0xCC
    high_quality_picture_fragment = 0xEC


class PictureCodingModes(IntEnum):
    pictures_are_frames = 0
    pictures_are_fields = 1


class ColorDifferenceSamplingFormats(IntEnum):
    color_4_4_4 = 0
    color_4_2_2 = 1
    color_4_2_0 = 2


class SourceSamplingModes(IntEnum):
    progressive = 0
    interlaced = 1


class Profiles(IntEnum):
    low_delay = 0
    high_quality = 3


class Levels(IntEnum):
    unconstrained = 0
    sub_sd = 1
    sd = 2
    hd = 3
    digital_cinema_2k = 4
    digital_cinema_4k = 5
    uhdtv_4k = 6
    uhdtv_8k = 7
    hd_over_sd_sdi = 64
    hd_over_hd_sdi = 65
    uhd_over_hd_sdi = 66


class WaveletFilters(IntEnum):
    deslauriers_dubuc_9_7 = 0
    le_gall_5_3 = 1
    deslauriers_dubuc_13_7 = 2
    haar_no_shift = 3
    haar_with_shift = 4
    fidelity = 5
    daubechies_9_7 = 6


class LiftingFilterTypes(IntEnum):
    even_add_odd = 1
    even_subtract_odd = 2
    odd_add_even = 3
    odd_subtract_even = 4


class BaseVideoFormats(IntEnum):
    custom_format = 0
    qsif525 = 1
    qcif = 2
    sif525 = 3
    cif = 4
    _4sif525 = 5
    _4cif = 6
    sdi480i_60 = 7
    sdi576i_50 = 8
    hd720p_60 = 9
    hd720p_50 = 10
    hd1080i_60 = 11
    hd1080i_50 = 12
    hd1080p_60 = 13
    hd1080p_50 = 14
    dc2k = 15
    dc4k = 16
    uhdtv_4k_60 = 17
    uhdtv_4k_50 = 18
    uhdtv_8k_60 = 19
    uhdtv_8k_50 = 20
    hd1080p_24 = 21
    sd_pro486 = 22


class PresetFrameRates(IntEnum):
    fps_24_over_1_001 = 1
    fps_24 = 2
    fps_25 = 3
    fps_30_over_1_001 = 4
    fps_30 = 5
    fps_50 = 6
    fps_60_over_1_001 = 7
    fps_60 = 8
    fps_15_over_1_001 = 9
    fps_25_over_2 = 10
    fps_48 = 11
    fps_48_over_1_001 = 12
    fps_96 = 13
    fps_100 = 14
    fps_120_over_1_001 = 15
    fps_120 = 16


class PresetPixelAspectRatios(IntEnum):
    ratio_1_1 = 1
    ratio_4_3_525_line = 2
    ratio_4_3_625_line = 3
    ratio_16_9_525_line = 4
    ratio_16_9_625_line = 5
    reduced_horizontal_resolution = 6


class PresetSignalRanges(IntEnum):
    video_8bit_full_range = 1
    video_8bit = 2
    video_10bit = 3
    video_12bit = 4
    video_10bit_full_range = 5
    video_12bit_full_range = 6
    video_16bit = 7
    video_16bit_full_range = 8


class PresetColorSpecs(IntEnum):
    custom = 0
    sdtv_525 = 1
    sdtv_625 = 2
    hdtv = 3
    d_cinema = 4
    uhdtv = 5
    hdr_tv_pq = 6
    hdr_tv_hlg = 7


class PresetColorPrimaries(IntEnum):
    hdtv = 0
    sdtv_525 = 1
    sdtv_625 = 2
    d_cinema = 3
    uhdtv = 4


class PresetColorMatrices(IntEnum):
    hdtv = 0
    sdtv = 1
    reversible = 2
    rgb = 3
    uhdtv = 4


class PresetTransferFunctions(IntEnum):
    tv_gamma = 0
    extended_gamut = 1
    linear = 2
    d_cinema = 3
    perceptual_quantizer = 4
    hybrid_log_gamma = 5


PARSE_INFO_PREFIX = 0x42424344
PARSE_INFO_HEADER_BYTES = 13

LiftingStage = namedtuple("LiftingStage", "lift_type,S,L,D,taps")
LiftingFilterParameters = namedtuple(
    "LiftingFilterParameters", "stages,filter_bit_shift"
)

_T = LiftingFilterTypes

# SMPTE ST 2042-1 (15.4.4.3) synthesis lifting stages, in order of
# application.  Restated from the published standard.
LIFTING_FILTERS = {
    WaveletFilters.deslauriers_dubuc_9_7: LiftingFilterParameters(
        stages=[
            LiftingStage(_T.even_subtract_odd, 2, 2, 0, [1, 1]),
            LiftingStage(_T.odd_add_even, 4, 4, -1, [-1, 9, 9, -1]),
        ],
        filter_bit_shift=1,
    ),
    WaveletFilters.le_gall_5_3: LiftingFilterParameters(
        stages=[
            LiftingStage(_T.even_subtract_odd, 2, 2, 0, [1, 1]),
            LiftingStage(_T.odd_add_even, 1, 2, 0, [1, 1]),
        ],
        filter_bit_shift=1,
    ),
    WaveletFilters.deslauriers_dubuc_13_7: LiftingFilterParameters(
        stages=[
            LiftingStage(_T.even_subtract_odd, 5, 4, -1, [-1, 9, 9, -1]),
            LiftingStage(_T.odd_add_even, 4, 4, -1, [-1, 9, 9, -1]),
        ],
        filter_bit_shift=1,
    ),
    WaveletFilters.haar_no_shift: LiftingFilterParameters(
        stages=[
            LiftingStage(_T.even_subtract_odd, 1, 1, 1, [1]),
            LiftingStage(_T.odd_add_even, 0, 1, 0, [1]),
        ],
        filter_bit_shift=0,
    ),
    WaveletFilters.haar_with_shift: LiftingFilterParameters(
        stages=[
            LiftingStage(_T.even_subtract_odd, 1, 1, 1, [1]),
            LiftingStage(_T.odd_add_even, 0, 1, 0, [1]),
        ],
        filter_bit_shift=1,
    ),
    WaveletFilters.fidelity: LiftingFilterParameters(
        stages=[
            LiftingStage(_T.odd_add_even, 8, 8, -3, [-2, 10, -25, 81, 81, -25, 10, -2]),
            LiftingStage(
                _T.even_subtract_odd, 8, 8, -3, [-8, 21, -46, 161, 161, -46, 21, -8]
            ),
        ],
        filter_bit_shift=0,
    ),
    WaveletFilters.daubechies_9_7: LiftingFilterParameters(
        stages=[
            LiftingStage(_T.even_subtract_odd, 12, 2, 0, [1817, 1817]),
            LiftingStage(_T.odd_subtract_even, 12, 2, 0, [3616, 3616]),
            LiftingStage(_T.even_add_odd, 12, 2, 0, [217, 217]),
            LiftingStage(_T.odd_add_even, 12, 2, 0, [6497, 6497]),
        ],
        filter_bit_shift=1,
    ),
}

# ---------------------------------------------------------------------------
# Default quantisation matrices (SMPTE ST 2042-1 Annex D).
#   key = (wavelet_index, wavelet_index_ho, dwt_depth, dwt_depth_ho) -> {level: {orient: value}}
# Pinning: the Deslauriers-Dubuc (9,7) tables for dwt_depth_ho = 0..4 are quoted verbatim in the
# reference tree (/root/reference/util/parse_default_quantisation_matrix_table.py:10-55); the
# dwt_depth_ho = 0 tables of the other six filters are restated from the published standard and
# are NOT pinned by anything under /root/reference.  Asymmetric tables of the other
# filters are not restated: such streams need custom_quant_matrix (the reference raises
# NoQuantisationMatrixAvailable, decoder/picture_syntax.py:331).
# ---------------------------------------------------------------------------


def _sym(ll, rows, depth):
    """{level: {orient: v}} of a 2-D only transform of `depth` levels."""
    m = {0: {"LL": ll if depth else 0}}
    for level in range(1, depth + 1):
        hl, lh, hh = rows[level - 1]
        m[level] = {"HL": hl, "LH": lh, "HH": hh}
    return m


def _asym(l, hs, rows, depth, depth_ho):
    m = {0: {"L": l}}
    for level in range(1, depth_ho + 1):
        m[level] = {"H": hs[level - 1]}
    for level in range(depth_ho + 1, depth_ho + depth + 1):
        hl, lh, hh = rows[level - depth_ho - 1]
        m[level] = {"HL": hl, "LH": lh, "HH": hh}
    return m


_SYM_TABLES = {
    # wavelet: (LL for depth >= 1, per-level (HL, LH, HH)); LL may depend on the depth
    WaveletFilters.deslauriers_dubuc_9_7: (lambda d: 5, [(3, 3, 0), (4, 4, 1), (5, 5, 2), (6, 6, 3)]),
    WaveletFilters.le_gall_5_3: (lambda d: 4, [(2, 2, 0), (4, 4, 2), (5, 5, 3), (7, 7, 5)]),
    WaveletFilters.deslauriers_dubuc_13_7: (lambda d: 5, [(3, 3, 0), (4, 4, 1), (5, 5, 2), (6, 6, 3)]),
    WaveletFilters.haar_with_shift: (lambda d: 8, [(4, 4, 0), (4, 4, 0), (4, 4, 0), (4, 4, 0)]),
    WaveletFilters.fidelity: (lambda d: 0, [(4, 4, 8), (8, 8, 12), (13, 13, 17), (17, 17, 21)]),
    WaveletFilters.daubechies_9_7: (lambda d: 3, [(1, 1, 0), (4, 4, 2), (6, 6, 5), (9, 9, 7)]),
}

QUANTISATION_MATRICES = {}
for _w, (_ll, _rows) in _SYM_TABLES.items():
    for _d in range(0, 5):
        QUANTISATION_MATRICES[(_w, _w, _d, 0)] = _sym(_ll(_d), _rows, _d)
# Haar without shift: every level adds 4 from the top down (Table D.4)
for _d in range(0, 5):
    QUANTISATION_MATRICES[(WaveletFilters.haar_no_shift, WaveletFilters.haar_no_shift, _d, 0)] = _sym(
        4 * (_d + 1), [(4 * (_d - _l), 4 * (_d - _l), 4 * (_d - _l - 1)) for _l in range(_d)], _d)
# Deslauriers-Dubuc (9,7) with horizontal-only levels (quoted in the reference's util script)
_DD = WaveletFilters.deslauriers_dubuc_9_7
_DD_HO = {
    1: ([0], [(3, 3, 0), (4, 4, 1), (5, 5, 2), (6, 6, 3)], 4),
    2: ([0, 3], [(5, 5, 3), (6, 6, 4), (7, 7, 5)], 3),
    3: ([0, 3, 5], [(8, 8, 5), (9, 9, 6)], 2),
    4: ([0, 3, 5, 8], [(10, 10, 8)], 1),
}
for _dho, (_hs, _rows, _maxd) in _DD_HO.items():
    for _d in range(0, _maxd + 1):
        QUANTISATION_MATRICES[(_DD, _DD, _d, _dho)] = _asym(3, _hs, _rows, _d, _dho)

# ---------------------------------------------------------------------------
# Header-level tables (SMPTE ST 2042-1 clause 11 / Annex C), restated from the published standard.
# They are consulted by the reference's sequence-header code (set_source_defaults,
# pseudocode/video_parameters.py:133; encoder/sequence_header.py) -- never by the picture path --
# and are NOT pinned by anything under /root/reference.
# ---------------------------------------------------------------------------
from fractions import Fraction  # noqa: E402

ProfileParameters = namedtuple("ProfileParameters", "allowed_parse_codes")
PROFILES = {
    Profiles.low_delay: ProfileParameters(
        [
            ParseCodes.sequence_header,
            ParseCodes.end_of_sequence,
            ParseCodes.auxiliary_data,
            ParseCodes.padding_data,
            ParseCodes.low_delay_picture,
            ParseCodes.low_delay_picture_fragment,
        ]
    ),
    Profiles.high_quality: ProfileParameters(
        [
            ParseCodes.sequence_header,
            ParseCodes.end_of_sequence,
            ParseCodes.auxiliary_data,
            ParseCodes.padding_data,
            ParseCodes.high_quality_picture,
            ParseCodes.high_quality_picture_fragment,
        ]
    ),
}
LevelParameters = namedtuple("LevelParameters", "standard")
LEVELS = {level: LevelParameters("n/a") for level in Levels}

BaseVideoFormatParameters = namedtuple(
    "BaseVideoFormatParameters",
    "frame_width,frame_height,color_diff_format_index,source_sampling,"
    "top_field_first,frame_rate_index,pixel_aspect_ratio_index,clean_width,"
    "clean_height,left_offset,top_offset,signal_range_index,color_spec_index",
)
_C = ColorDifferenceSamplingFormats
_P, _I = SourceSamplingModes.progressive, SourceSamplingModes.interlaced
# index: (width, height, chroma format, sampling, top field first, frame rate, pixel aspect ratio,
#         clean width, clean height, left, top, signal range, colour spec)   -- Annex C, Table C.1
_BVF = {
    0: (640, 480, _C.color_4_2_0, _P, False, 1, 1, 640, 480, 0, 0, 1, 0),
    1: (176, 120, _C.color_4_2_0, _P, False, 9, 2, 176, 120, 0, 0, 1, 1),
    2: (176, 144, _C.color_4_2_0, _P, True, 10, 3, 176, 144, 0, 0, 1, 2),
    3: (352, 240, _C.color_4_2_0, _P, False, 9, 2, 352, 240, 0, 0, 1, 1),
    4: (352, 288, _C.color_4_2_0, _P, True, 10, 3, 352, 288, 0, 0, 1, 2),
    5: (704, 480, _C.color_4_2_0, _P, False, 9, 2, 704, 480, 0, 0, 1, 1),
    6: (704, 576, _C.color_4_2_0, _P, True, 10, 3, 704, 576, 0, 0, 1, 2),
    7: (720, 480, _C.color_4_2_2, _I, False, 4, 2, 704, 480, 8, 0, 3, 1),
    8: (720, 576, _C.color_4_2_2, _I, True, 3, 3, 704, 576, 8, 0, 3, 2),
    9: (1280, 720, _C.color_4_2_2, _P, True, 7, 1, 1280, 720, 0, 0, 3, 3),
    10: (1280, 720, _C.color_4_2_2, _P, True, 6, 1, 1280, 720, 0, 0, 3, 3),
    11: (1920, 1080, _C.color_4_2_2, _I, True, 4, 1, 1920, 1080, 0, 0, 3, 3),
    12: (1920, 1080, _C.color_4_2_2, _I, True, 3, 1, 1920, 1080, 0, 0, 3, 3),
    13: (1920, 1080, _C.color_4_2_2, _P, True, 7, 1, 1920, 1080, 0, 0, 3, 3),
    14: (1920, 1080, _C.color_4_2_2, _P, True, 6, 1, 1920, 1080, 0, 0, 3, 3),
    15: (2048, 1080, _C.color_4_4_4, _P, True, 2, 1, 2048, 1080, 0, 0, 4, 4),
    16: (4096, 2160, _C.color_4_4_4, _P, True, 2, 1, 4096, 2160, 0, 0, 4, 4),
    17: (3840, 2160, _C.color_4_2_2, _P, True, 7, 1, 3840, 2160, 0, 0, 3, 5),
    18: (3840, 2160, _C.color_4_2_2, _P, True, 6, 1, 3840, 2160, 0, 0, 3, 5),
    19: (7680, 4320, _C.color_4_2_2, _P, True, 7, 1, 7680, 4320, 0, 0, 3, 5),
    20: (7680, 4320, _C.color_4_2_2, _P, True, 6, 1, 7680, 4320, 0, 0, 3, 5),
    21: (1920, 1080, _C.color_4_2_2, _P, True, 1, 1, 1920, 1080, 0, 0, 3, 3),
    22: (720, 486, _C.color_4_2_2, _I, False, 4, 2, 720, 486, 0, 0, 3, 3),
}
BASE_VIDEO_FORMAT_PARAMETERS = {
    BaseVideoFormats(i): BaseVideoFormatParameters(*row) for i, row in _BVF.items()
}
PRESET_FRAME_RATES = {
    PresetFrameRates(i): f
    for i, f in {
        1: Fraction(24000, 1001), 2: Fraction(24, 1), 3: Fraction(25, 1), 4: Fraction(30000, 1001),
        5: Fraction(30, 1), 6: Fraction(50, 1), 7: Fraction(60000, 1001), 8: Fraction(60, 1),
        9: Fraction(15000, 1001), 10: Fraction(25, 2), 11: Fraction(48, 1), 12: Fraction(48000, 1001),
        13: Fraction(96, 1), 14: Fraction(100, 1), 15: Fraction(120000, 1001), 16: Fraction(120, 1),
    }.items()
}
PRESET_PIXEL_ASPECT_RATIOS = {
    PresetPixelAspectRatios(i): f
    for i, f in {1: Fraction(1, 1), 2: Fraction(10, 11), 3: Fraction(12, 11), 4: Fraction(40, 33),
                 5: Fraction(16, 11), 6: Fraction(4, 3)}.items()
}
SignalRangeParameters = namedtuple(
    "SignalRangeParameters",
    "luma_offset,luma_excursion,color_diff_offset,color_diff_excursion",
)
PRESET_SIGNAL_RANGES = {
    PresetSignalRanges(i): SignalRangeParameters(*row)
    for i, row in {
        1: (0, 255, 128, 255), 2: (16, 219, 128, 224), 3: (64, 876, 512, 896), 4: (256, 3504, 2048, 3584),
        5: (0, 1023, 512, 1023), 6: (0, 4095, 2048, 4095), 7: (4096, 56064, 32768, 57344),
        8: (0, 65535, 32768, 65535),
    }.items()
}
ColorSpecificiation = namedtuple(
    "ColorSpecificiation", "color_primaries_index,color_matrix_index,transfer_function_index"
)
_CP, _CM, _TF = PresetColorPrimaries, PresetColorMatrices, PresetTransferFunctions
PRESET_COLOR_SPECS = {
    PresetColorSpecs(i): ColorSpecificiation(*row)
    for i, row in {
        0: (_CP.hdtv, _CM.hdtv, _TF.tv_gamma),
        1: (_CP.sdtv_525, _CM.sdtv, _TF.tv_gamma),
        2: (_CP.sdtv_625, _CM.sdtv, _TF.tv_gamma),
        3: (_CP.hdtv, _CM.hdtv, _TF.tv_gamma),
        4: (_CP.d_cinema, _CM.reversible, _TF.d_cinema),
        5: (_CP.uhdtv, _CM.uhdtv, _TF.tv_gamma),
        6: (_CP.uhdtv, _CM.uhdtv, _TF.perceptual_quantizer),
        7: (_CP.uhdtv, _CM.uhdtv, _TF.hybrid_log_gamma),
    }.items()
}
# colour science tables (color_conversion.py) are out of scope: import-only placeholders
PRESET_COLOR_PRIMARIES = {}
PRESET_COLOR_MATRICES = {}
PRESET_TRANSFER_FUNCTIONS = {}

__all__ = [name for name in dir() if not name.startswith("_")]
