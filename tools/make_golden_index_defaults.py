"""Golden vectors for the tables of the stream indexer (vc2b_index_stream): streams whose sequence headers
rely on a BASE VIDEO FORMAT and whose pictures rely on the DEFAULT QUANTISATION MATRIX, written by the
unmodified reference encoder (encoder.make_sequence picks the closest base video format and only
overrides what differs; quantization_matrix=None selects the default matrix) and parsed by the unmodified
reference decoder, whose per-picture state is the ground truth.

    python tools/make_golden_index_defaults.py     ->  tests/golden/index_defaults.npz

Needs a reference tree (oracle/ref_import.py).  NOTE on pinning: the tables themselves live in the
un-vendored vc2_data_tables package; the reference runs here on the stand-in oracle/refshim/vc2_data_tables
(restated from SMPTE ST 2042-1), so these vectors pin csrc/index.cu against that restatement and against
the reference's own header logic -- not against a third copy of the standard.  The Deslauriers-Dubuc (9,7)
matrices are additionally quoted inside the reference tree (util/parse_default_quantisation_matrix_table.py).
"""
import copy
import os
import sys
from io import BytesIO

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import ref_import  # noqa: E402

ref_import.install()

import vc2_data_tables as T  # noqa: E402
from vc2_conformance import bitstream, decoder, encoder  # noqa: E402
from vc2_conformance.codec_features import CodecFeatures  # noqa: E402
from vc2_conformance.decoder import fragment_syntax, picture_syntax  # noqa: E402
from vc2_conformance.pseudocode.state import State  # noqa: E402
from vc2_conformance.pseudocode.video_parameters import set_source_defaults  # noqa: E402

# name, base video format, overrides of its video parameters, profile, wavelet, depth, depth_ho, slices, bytes, fragments
CASES = [
    # native sizes (everything from the base video format)
    ("cif_haar_d2_ld", T.BaseVideoFormats.cif, {}, "ld", 4, 2, 0, (4, 3), 1500, 0),
    ("qsif525_legall_d3", T.BaseVideoFormats.qsif525, {}, "hq", 1, 3, 0, (2, 3), 1200, 0),
    ("qcif_haar_noshift_d4", T.BaseVideoFormats.qcif, {}, "hq", 3, 4, 0, (2, 2), 2000, 0),
    # small custom frame sizes (the Python reference takes minutes per full-size picture); colour difference
    # format and signal range still come from the base video format
    ("hd1080p50_dd97_d2_ho1", T.BaseVideoFormats.hd1080p_50, dict(frame_width=128, frame_height=64, clean_width=128,
                                                                  clean_height=64), "hq", 0, 2, 1, (4, 4), 1500, 0),
    ("dc2k_dd137_d4", T.BaseVideoFormats.dc2k, dict(frame_width=96, frame_height=64, clean_width=96, clean_height=64),
     "hq", 2, 4, 0, (3, 2), 1500, 0),
    ("uhd4k60_fidelity_d1_10bit_full", T.BaseVideoFormats.uhdtv_4k_60,
     dict(frame_width=64, frame_height=64, clean_width=64, clean_height=64, left_offset=0, top_offset=0, luma_offset=0, luma_excursion=1023,
          color_diff_offset=512, color_diff_excursion=1023), "hq", 5, 1, 0, (4, 4), 1500, 0),
    ("sd576i_daub97_d3_fragments_fields", T.BaseVideoFormats.sdi576i_50,
     dict(frame_width=64, frame_height=64, clean_width=64, clean_height=64, left_offset=0, top_offset=0), "hq", 6, 3, 0, (2, 2), 900, 3),
]


def main():
    rec, names = {}, []
    for name, base, overrides, profile, wi, d, dho, (sx, sy), nbytes, frag in CASES:
        vp = set_source_defaults(base)
        vp.update(overrides)
        fields = vp["source_sampling"] == T.SourceSamplingModes.interlaced
        cf = CodecFeatures(
            name=name, level=T.Levels.unconstrained,
            profile=T.Profiles.high_quality if profile == "hq" else T.Profiles.low_delay,
            picture_coding_mode=T.PictureCodingModes(1 if fields else 0), video_parameters=vp,
            wavelet_index=T.WaveletFilters(wi), wavelet_index_ho=T.WaveletFilters(wi), dwt_depth=d, dwt_depth_ho=dho,
            slices_x=sx, slices_y=sy, fragment_slice_count=frag, lossless=False, picture_bytes=nbytes,
            quantization_matrix=None)
        W, H = vp["frame_width"], vp["frame_height"] // (2 if fields else 1)
        cdf = int(vp["color_diff_format_index"])
        cw, ch = W // (2 if cdf else 1), H // (2 if cdf == 2 else 1)
        pic = {"Y": np.full((H, W), vp["luma_offset"] + 7).tolist(),
               "C1": np.full((ch, cw), vp["color_diff_offset"]).tolist(),
               "C2": np.full((ch, cw), vp["color_diff_offset"] - 3).tolist(), "pic_num": 4}
        pics = [pic] if not fields else [pic, dict(pic, pic_num=5)]
        seq = encoder.make_sequence(cf, pics)
        f = BytesIO()
        bitstream.autofill_and_serialise_stream(f, bitstream.Stream(sequences=[seq]))
        data = f.getvalue()

        # ground truth: the state of the reference decoder when it reaches the picture's transform data
        snaps = []

        def snap(state):
            qm = np.zeros((16, 4), np.int64)
            for level, orients in state["quant_matrix"].items():
                for orient, v in orients.items():
                    qm[level][{"LL": 0, "L": 0, "H": 1, "HL": 1, "LH": 2, "HH": 3}[orient]] = v
            snaps.append((np.array([state[k] for k in ("luma_width", "luma_height", "color_diff_width",
                                                      "color_diff_height", "luma_depth", "color_diff_depth",
                                                      "wavelet_index", "wavelet_index_ho", "dwt_depth", "dwt_depth_ho",
                                                      "slices_x", "slices_y")], np.int64), qm))

        orig_td, orig_ifs = picture_syntax.transform_data, fragment_syntax.initialize_fragment_state

        def td(state):
            snap(state)
            return orig_td(state)

        def ifs(state):
            snap(state)
            return orig_ifs(state)

        picture_syntax.transform_data, fragment_syntax.initialize_fragment_state = td, ifs
        try:
            out = []
            state = State(_output_picture_callback=lambda p_, v_, m_: out.append(copy.deepcopy(p_)))
            decoder.init_io(state, BytesIO(data))
            decoder.parse_stream(state)
        finally:
            picture_syntax.transform_data, fragment_syntax.initialize_fragment_state = orig_td, orig_ifs
        assert len(snaps) == len(pics) and len(out) == len(pics)
        # the stream really does rely on the tables
        sh = seq["data_units"][0]["sequence_header"]
        base = sh["base_video_format"]  # the encoder picks the closest format: possibly a sibling of the one asked for
        srcp = sh["video_parameters"]
        assert not srcp["color_diff_sampling_format"]["custom_color_diff_format_flag"], name
        assert "signal_range" not in overrides and not srcp["signal_range"]["custom_signal_range_flag"] or \
            "luma_excursion" in overrides, name
        rec[name + ".stream"] = np.frombuffer(data, np.uint8)
        rec[name + ".state"] = snaps[0][0]
        rec[name + ".quant_matrix"] = snaps[0][1]
        rec[name + ".picture_Y"] = np.array(out[0]["Y"], np.int64)
        rec[name + ".picture_C1"] = np.array(out[0]["C1"], np.int64)
        rec[name + ".picture_C2"] = np.array(out[0]["C2"], np.int64)
        names.append(name)
        print("  %-40s %6d bytes  base format %2d  %s" % (name, len(data), int(base), snaps[0][0].tolist()))
    rec["names"] = np.array(names)
    path = os.path.join(ROOT, "tests", "golden", "index_defaults.npz")
    np.savez_compressed(path, **rec)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
