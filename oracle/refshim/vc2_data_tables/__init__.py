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

# Default quantisation matrices (Annex D) are NOT restated: every stream this
# repo generates carries a custom quantisation matrix.
QUANTISATION_MATRICES = {}

# ---------------------------------------------------------------------------
# Header-level tables: import-only placeholders (never used on the picture
# path).
# ---------------------------------------------------------------------------
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
BASE_VIDEO_FORMAT_PARAMETERS = {}
PRESET_FRAME_RATES = {}
PRESET_PIXEL_ASPECT_RATIOS = {}
SignalRangeParameters = namedtuple(
    "SignalRangeParameters",
    "luma_offset,luma_excursion,color_diff_offset,color_diff_excursion",
)
PRESET_SIGNAL_RANGES = {}
PRESET_COLOR_SPECS = {}
PRESET_COLOR_PRIMARIES = {}
PRESET_COLOR_MATRICES = {}
PRESET_TRANSFER_FUNCTIONS = {}

__all__ = [name for name in dir() if not name.startswith("_")]
