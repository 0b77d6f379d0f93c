"""Golden vectors for the stream indexer vc2b_index_stream: whole VC-2 streams assembled
from the reference's own containers, serialised by autofill_and_serialise_stream and read back by the
reference's deserialiser (bitstream.parse_stream), which is the ground truth for every header field.

    python tools/make_golden_index.py     ->  tests/golden/index.npz

Needs /root/reference (this container only); the vectors travel with the repository.

The sequence headers override every video parameter the picture path needs (frame size, colour
difference format, signal range), so the contents of the base video format table -- part of the
un-vendored vc2_data_tables package -- never matter; the placeholder rows installed below only have to
exist for the reference's set_source_defaults() to run.
"""
import os
import sys
from collections import namedtuple
from fractions import Fraction
from io import BytesIO

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tools")]
import make_golden as MG  # noqa: E402  (installs the reference import shim)

import vc2_data_tables as T  # noqa: E402
from vc2_conformance import bitstream as B  # noqa: E402
from vc2_conformance.encoder import pictures as encp  # noqa: E402

# placeholder rows (never relied upon, see above)
T.BASE_VIDEO_FORMAT_PARAMETERS[T.BaseVideoFormats(0)] = T.BaseVideoFormatParameters(
    640, 480, T.ColorDifferenceSamplingFormats.color_4_2_0, T.SourceSamplingModes.progressive, False, 1, 1,
    640, 480, 0, 0, 1, 0)
T.PRESET_FRAME_RATES[1] = Fraction(24000, 1001)
T.PRESET_PIXEL_ASPECT_RATIOS[1] = Fraction(1, 1)
T.PRESET_SIGNAL_RANGES[1] = T.SignalRangeParameters(0, 255, 128, 255)
T.PRESET_COLOR_SPECS[0] = namedtuple(
    "ColorSpecificiation", "color_primaries_index,color_matrix_index,transfer_function_index")(0, 0, 0)

O = MG.O
PC = T.ParseCodes

# name, W, H, chroma (wdiv, hdiv), depth, wi, wiho, d, dho, sx, sy, profile, bytes, fragment_slice_count, fields, zero_next
CASES = [
    ("hq_v2_picture", 64, 32, (2, 1), 10, 0, 0, 2, 0, 4, 4, "hq", 1536, 0, False, False),
    ("hq_v3_asym_picture_no_next_offset", 64, 32, (1, 1), 8, 1, 3, 1, 2, 4, 2, "hq", 2048, 0, False, True),
    ("ld_fields_picture_no_next_offset", 64, 64, (2, 2), 10, 4, 4, 2, 0, 4, 4, "ld", 1400, 0, True, True),
    ("hq_fragments", 96, 32, (2, 1), 12, 1, 1, 2, 0, 6, 2, "hq", 2400, 5, False, False),
    ("ld_fragments", 64, 32, (2, 1), 8, 4, 4, 1, 0, 4, 4, "ld", 1100, 3, False, False),
]


def sequence_header_unit(W, H, cdiv, depth, fields, major):
    exc = (1 << depth) - 1
    sh = B.SequenceHeader(
        parse_parameters=B.ParseParameters(major_version=B.AUTO, minor_version=0, profile=0, level=0),
        base_video_format=0,
        video_parameters=B.SourceParameters(
            frame_size=B.FrameSize(custom_dimensions_flag=True, frame_width=W, frame_height=H),
            color_diff_sampling_format=B.ColorDiffSamplingFormat(
                custom_color_diff_format_flag=True, color_diff_format_index={(1, 1): 0, (2, 1): 1, (2, 2): 2}[cdiv]),
            scan_format=B.ScanFormat(custom_scan_format_flag=True, source_sampling=1 if fields else 0),
            frame_rate=B.FrameRate(custom_frame_rate_flag=True, index=0, frame_rate_numer=30000, frame_rate_denom=1001),
            pixel_aspect_ratio=B.PixelAspectRatio(custom_pixel_aspect_ratio_flag=False),
            clean_area=B.CleanArea(custom_clean_area_flag=True, clean_width=W, clean_height=H, left_offset=0, top_offset=0),
            signal_range=B.SignalRange(custom_signal_range_flag=True, index=0, luma_offset=0, luma_excursion=exc,
                                       color_diff_offset=1 << (depth - 1), color_diff_excursion=exc),
            color_spec=B.ColorSpec(custom_color_spec_flag=False),
        ),
        picture_coding_mode=1 if fields else 0,
    )
    return B.DataUnit(parse_info=B.ParseInfo(parse_code=PC.sequence_header), sequence_header=sh)


def main():
    rng = np.random.RandomState(23)
    rec, names = {}, []
    for (name, W, H, cdiv, depth, wi, wiho, d, dho, sx, sy, profile, nbytes, frag, fields, zero_next) in CASES:
        qm = MG.make_quant_matrix(d, dho, rng)
        ph = H // 2 if fields else H
        pics = MG.make_picture(rng, W, ph, cdiv, depth, "ramp")
        cf = {
            "wavelet_index": wi, "wavelet_index_ho": wiho, "dwt_depth": d, "dwt_depth_ho": dho,
            "slices_x": sx, "slices_y": sy, "picture_coding_mode": 1 if fields else 0,
            "profile": T.Profiles.low_delay if profile == "ld" else T.Profiles.high_quality,
            "level": T.Levels(0), "lossless": False, "picture_bytes": nbytes, "fragment_slice_count": frag,
            "quantization_matrix": qm,
            "video_parameters": {
                "frame_width": W, "frame_height": H,
                "color_diff_format_index": {(1, 1): 0, (2, 1): 1, (2, 2): 2}[cdiv],
                "luma_excursion": (1 << depth) - 1, "color_diff_excursion": (1 << depth) - 1,
            },
        }
        asym = (wiho != wi) or dho != 0
        major = 3 if (asym or frag) else 2
        units = [sequence_header_unit(W, H, cdiv, depth, fields, major)]
        for k in range(2):
            picture = {"Y": pics[0].tolist(), "C1": pics[1].tolist(), "C2": pics[2].tolist(), "pic_num": 100 + k}
            units += list(encp.make_picture_data_units(cf, picture))
            if k == 0:
                units.append(B.DataUnit(parse_info=B.ParseInfo(parse_code=PC.padding_data),
                                        padding=B.Padding(bytes=b"\x42\x42\x43\x44padding")))
        units.append(B.DataUnit(parse_info=B.ParseInfo(parse_code=PC.end_of_sequence)))
        if zero_next:
            for u in units:
                if u["parse_info"]["parse_code"] in (PC.high_quality_picture, PC.low_delay_picture):
                    u["parse_info"]["next_parse_offset"] = 0
        f = BytesIO()
        B.autofill_and_serialise_stream(f, B.Stream(sequences=[B.Sequence(data_units=units)]))
        data = f.getvalue()

        # ground truth: the reference deserialiser
        reader = B.BitstreamReader(BytesIO(data))
        with B.Deserialiser(reader) as des:
            B.parse_stream(des, MG.State())
        got_units = des.context["sequences"][0]["data_units"]
        rows = []
        for u in got_units:
            pi = u["parse_info"]
            row = dict(offset=pi["_offset"], parse_code=int(pi["parse_code"]), next=int(pi["next_parse_offset"]),
                       picture_number=-1, frag=(0, 0, 0, 0), tp=None)
            if "picture_parse" in u:
                pp = u["picture_parse"]
                row["picture_number"] = int(pp["picture_header"]["picture_number"])
                row["tp"] = pp["wavelet_transform"]["transform_parameters"]
            elif "fragment_parse" in u:
                fp = u["fragment_parse"]
                fh = fp["fragment_header"]
                row["picture_number"] = int(fh["picture_number"])
                row["frag"] = (int(fh["fragment_data_length"]), int(fh["fragment_slice_count"]),
                               int(fh.get("fragment_x_offset", 0)), int(fh.get("fragment_y_offset", 0)))
                row["tp"] = fp.get("transform_parameters")
            rows.append(row)

        # one int64 row per data unit:
        # [offset, parse_code, next_parse_offset, picture_number, frag_len, frag_count, frag_x, frag_y, has_tp,
        #  wavelet_index, dwt_depth, wavelet_index_ho, dwt_depth_ho, slices_x, slices_y, sb_num, sb_den, prefix, scaler]
        table = np.zeros((len(rows), 19), np.int64)
        for i, r in enumerate(rows):
            table[i, :4] = [r["offset"], r["parse_code"], r["next"], r["picture_number"]]
            table[i, 4:8] = r["frag"]
            tp = r["tp"]
            if tp is not None:
                etp = tp.get("extended_transform_parameters", {})
                sp = tp["slice_parameters"]
                table[i, 8] = 1
                table[i, 9:19] = [
                    tp["wavelet_index"], tp["dwt_depth"],
                    etp["wavelet_index_ho"] if etp.get("asym_transform_index_flag") else tp["wavelet_index"],
                    etp["dwt_depth_ho"] if etp.get("asym_transform_flag") else 0,
                    sp["slices_x"], sp["slices_y"], sp.get("slice_bytes_numerator", 0),
                    sp.get("slice_bytes_denominator", 1), sp.get("slice_prefix_bytes", 0),
                    sp.get("slice_size_scaler", 1)]
                assert tp["quant_matrix"]["custom_quant_matrix"]
        # byte offset of the transform data of whole pictures: where the reference's serialisation of the
        # deserialised TransformData sits in the stream
        data_offsets = np.full(len(rows), -1, np.int64)
        for i, (u, r) in enumerate(zip(got_units, rows)):
            if "picture_parse" not in u:
                continue
            t = table[i]
            pp = O.make_params(W, ph, W // cdiv[0], ph // cdiv[1], depth, depth, int(t[9]), int(t[11]), int(t[10]),
                               int(t[12]), int(t[13]), int(t[14]), profile == "ld", int(t[17]), int(t[18]),
                               int(t[15]), int(t[16]) if profile == "ld" else 1, qm)
            td_bytes = MG.serialise_transform_data(u["picture_parse"]["wavelet_transform"]["transform_data"],
                                                   MG.state_from_params(pp, qm))
            at = data.find(td_bytes, int(t[0]) + 13)
            assert at > 0 and data.find(td_bytes, at + 1) in (-1, data.find(td_bytes, at + 1)), name
            data_offsets[i] = at
        major = int(got_units[0]["sequence_header"]["parse_parameters"]["major_version"])
        rec[name + ".data_offsets"] = data_offsets
        rec[name + ".stream"] = np.frombuffer(data, np.uint8)
        rec[name + ".units"] = table
        rec[name + ".quant_matrix"] = MG.qm_to_array(qm)
        rec[name + ".dims"] = np.array([W, ph, W // cdiv[0], ph // cdiv[1], depth, depth, major, 1 if fields else 0],
                                       np.int64)
        names.append(name)
        print("  %-36s %6d bytes, %d data units" % (name, len(data), len(rows)))
    rec["names"] = np.array(names)
    path = os.path.join(MG.GOLDEN, "index.npz")
    np.savez_compressed(path, **rec)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
