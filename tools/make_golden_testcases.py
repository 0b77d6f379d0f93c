"""Parity streams restating the reference's decoder test-case generators `slice_padding_data`,
`dangling_bounded_block_data` (test_cases/decoder/pictures.py:645-880) and `lossless_quantization`
(test_cases/decoder/lossless_quantization.py:117).

The generators themselves need whole sequences and the codec-feature tables; what makes their streams
special is done by four helpers that act on one slice -- fill_hq_slice_padding (:89),
fill_ld_slice_padding (:168), cut_off_value_at_end_of_hq_slice (:475), cut_off_value_at_end_of_ld_slice
(:566).  Here the UNMODIFIED helpers are applied to every slice of pictures encoded by the reference
encoder, the result is serialised by the reference serialiser and decoded by the reference decoder:

    python tools/make_golden_testcases.py     ->  tests/golden/testcases.npz

Needs /root/reference (this container only); the vectors travel with the repository.
"""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tools")]
import make_golden as MG  # noqa: E402  (installs the reference import shim)

from vc2_conformance.encoder import pictures as encp  # noqa: E402
from vc2_conformance.pseudocode.picture_decoding import filter_bit_shift  # noqa: E402
from vc2_conformance.pseudocode.state import State  # noqa: E402

# `vc2_conformance.test_cases/__init__.py` imports every generator module, and the encoder ones need the
# absent `vc2_bit_widths` package.  Only test_cases/decoder/pictures.py is wanted here, so the two package
# modules are replaced by empty shells that point at the real directories; pictures.py and the helper
# functions taken from it are the reference's own, unmodified files.
import importlib  # noqa: E402
import types  # noqa: E402
from collections import namedtuple  # noqa: E402

_tc_dir = os.path.join(MG.ref_import.REFERENCE_ROOT, "vc2_conformance", "test_cases")
_pkg = types.ModuleType("vc2_conformance.test_cases")
_pkg.__path__ = [_tc_dir]
_pkg.TestCase = namedtuple("TestCase", "value,subcase_name,case_name,metadata")
_pkg.decoder_test_case_generator = lambda f: f
sys.modules["vc2_conformance.test_cases"] = _pkg
_dpkg = types.ModuleType("vc2_conformance.test_cases.decoder")
_dpkg.__path__ = [os.path.join(_tc_dir, "decoder")]
sys.modules["vc2_conformance.test_cases.decoder"] = _dpkg
tcp = importlib.import_module("vc2_conformance.test_cases.decoder.pictures")
tcl = importlib.import_module("vc2_conformance.test_cases.decoder.lossless_quantization")
from vc2_data_tables import Profiles  # noqa: E402

O = MG.O

# name, W, H, chroma (wdiv, hdiv), depth, wi, wiho, d, dho, sx, sy, profile, picture bytes / slice bytes numerator
CASES = [
    ("hq_dd97_d2", 64, 32, (2, 1), 10, 0, 0, 2, 0, 4, 4, "hq", 1536),
    ("hq_legall_asym", 64, 32, (1, 1), 8, 1, 1, 1, 1, 4, 2, "hq", 2048),
    ("ld_haar_shift_d2", 64, 32, (2, 1), 10, 4, 4, 2, 0, 4, 4, "ld", 1400),
    ("ld_legall_uneven", 80, 36, (2, 1), 8, 1, 1, 2, 0, 5, 3, "ld", 1500),
]

FILLERS = [(b"\x00", False, "all_zeros"), (b"\xFF", False, "all_ones"), (b"\xAA", False, "alternating_1s_and_0s"),
           (b"\x55", False, "alternating_0s_and_1s"), (tcp.make_dummy_end_of_sequence(), True, "dummy_end_of_sequence")]


def main():
    rng = np.random.RandomState(11)
    rec, names = {}, []
    for (name, W, H, cdiv, depth, wi, wiho, d, dho, sx, sy, profile, arg) in CASES:
        qm = MG.make_quant_matrix(d, dho, rng)
        pics = MG.make_picture(rng, W, H, cdiv, depth, "ramp")
        picture = {"Y": pics[0].tolist(), "C1": pics[1].tolist(), "C2": pics[2].tolist(), "pic_num": 7}
        cf = {
            "wavelet_index": wi, "wavelet_index_ho": wiho, "dwt_depth": d, "dwt_depth_ho": dho,
            "slices_x": sx, "slices_y": sy, "picture_coding_mode": 0,
            "profile": Profiles.low_delay if profile == "ld" else Profiles.high_quality,
            "quantization_matrix": qm,
            "video_parameters": {
                "frame_width": W, "frame_height": H,
                "color_diff_format_index": {(1, 1): 0, (2, 1): 1, (2, 2): 2}[cdiv],
                "luma_excursion": (1 << depth) - 1, "color_diff_excursion": (1 << depth) - 1,
            },
        }
        slice_coeffs = encp.transform_and_slice_picture(cf, picture)
        if profile == "hq":
            scaler, td = encp.make_transform_data_hq_lossy(arg, slice_coeffs, 0, 1)
            p = O.make_params(W, H, W // cdiv[0], H // cdiv[1], depth, depth, wi, wiho, d, dho, sx, sy,
                              False, 0, scaler, 0, 1, qm)
            key, comps = "hq_slices", ["Y", "C1", "C2"]
        else:
            td = encp.make_transform_data_ld_lossy(arg, slice_coeffs)
            p = O.make_params(W, H, W // cdiv[0], H // cdiv[1], depth, depth, wi, wiho, d, dho, sx, sy,
                              True, 0, 1, arg, sx * sy, qm)
            key, comps = "ld_slices", ["Y", "C"]
        state = MG.state_from_params(p, qm)
        rec[name + ".params"] = MG.params_to_array(p)
        rec[name + ".quant_matrix"] = MG.qm_to_array(qm)
        # magnitude of the dangling value as in dangling_bounded_block_data (:800-810)
        shift = filter_bit_shift(State(wavelet_index=wi, wavelet_index_ho=wiho))
        magnitude = (d + dho) * shift + 1

        variants = []
        for filler, align, expl in FILLERS:
            for comp in comps:
                variants.append(("slice_padding_data[slice_%s_%s]" % (comp, expl), "pad", (comp, filler, align)))
        for dangle in tcp.DanglingTransformValueType:
            for comp in comps:
                variants.append(("dangling_bounded_block_data[%s_%s]" % (dangle.name, comp), "dangle", (comp, dangle)))

        if profile == "hq":
            variants.append(("lossless_quantization", "lossless_q", ()))

        for vname, kind, args in variants:
            t = copy.deepcopy(td)
            pv, statev = p, state
            if kind == "lossless_q":
                # lossless_quantization (test_cases/decoder/lossless_quantization.py:117-190): a lossless picture
                # whose every coefficient is 1 at a qindex high enough for all quantisation factors to differ
                # (largest matrix entry + MINIMUM_DISTINCT_QINDEX), length fields from the reference's
                # calculate_hq_length_field, slice_size_scaler raised until they fit eight bits.
                _scaler0, t = encp.make_transform_data_hq_lossless(slice_coeffs)
                qi = max(v for sub in qm.values() for v in sub.values()) + tcl.MINIMUM_DISTINCT_QINDEX
                longest = 0
                for sl in t[key]:
                    sl["qindex"] = qi
                    for c in ("y", "c1", "c2"):
                        sl[c + "_transform"] = [1] * len(sl[c + "_transform"])
                        sl["slice_%s_length" % c] = encp.calculate_hq_length_field(sl[c + "_transform"], 1)
                        longest = max(longest, sl["slice_%s_length" % c])
                new_scaler = max(1, (longest + 254) // 255)
                for sl in t[key]:
                    for c in ("y", "c1", "c2"):
                        sl["slice_%s_length" % c] = (sl["slice_%s_length" % c] + new_scaler - 1) // new_scaler
                pv = O.make_params(W, H, W // cdiv[0], H // cdiv[1], depth, depth, wi, wiho, d, dho, sx, sy,
                                   False, 0, new_scaler, 0, 1, qm)
                statev = MG.state_from_params(pv, qm)
                rec["%s.%s.params" % (name, vname)] = MG.params_to_array(pv)
            try:
                for n, sl in enumerate(t[key]):
                    sy_, sx_ = divmod(n, sx)
                    if kind == "lossless_q":
                        break
                    if kind == "pad" and profile == "hq":
                        tcp.fill_hq_slice_padding(state, sx_, sy_, sl, args[0], args[1], args[2])
                    elif kind == "pad":
                        tcp.fill_ld_slice_padding(state, sx_, sy_, sl, args[0], args[1], args[2])
                    elif profile == "hq":
                        tcp.cut_off_value_at_end_of_hq_slice(state, sx_, sy_, sl, args[0], args[1], magnitude, 0)
                    else:
                        tcp.cut_off_value_at_end_of_ld_slice(state, sx_, sy_, sl, args[0], args[1], magnitude)
            except tcp.UnsatisfiableBlockSizeError:
                print("   (slices too small for %s %s)" % (name, vname))
                continue
            data = MG.serialise_transform_data(t, statev)
            res = MG.ref_decode(pv, qm, data + b"\x00")
            assert res["status"] == 0 and res["consumed"] == len(data), (name, vname, res.get("consumed"), len(data))
            MG.check_oracle_decode(pv, data + b"\x00", res, name + "." + vname)
            if kind == "lossless_q":  # the generator drops pictures that clip (:176-190)
                assert all(0 < int(a.min()) and int(a.max()) < (1 << depth) - 1 for a in res["picture"]), name
            full = "%s.%s" % (name, vname)
            rec[full + ".data"] = np.frombuffer(data, np.uint8)
            for c in range(3):
                rec[full + ".dec_coeffs%d" % c] = res["coeffs"][c].astype(np.int32)
                rec[full + ".decoded%d" % c] = res["picture"][c].astype(np.uint16)
            names.append(full)
        print("  %-20s %d streams so far" % (name, len(names)))
    rec["names"] = np.array(names)
    path = os.path.join(MG.GOLDEN, "testcases.npz")
    np.savez_compressed(path, **rec)
    print(path, os.path.getsize(path), "bytes;", len(names), "streams")


if __name__ == "__main__":
    main()
