"""Helpers shared by the oracle and GPU tests: load tests/golden/*.npz (vectors produced by
tools/make_golden.py from the unmodified Python reference)."""

import os

import numpy as np

import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

PARAM_FIELDS = [
    "luma_width", "luma_height", "color_diff_width", "color_diff_height", "luma_depth",
    "color_diff_depth", "wavelet_index", "wavelet_index_ho", "dwt_depth", "dwt_depth_ho",
    "slices_x", "slices_y", "is_ld", "slice_prefix_bytes", "slice_size_scaler",
    "slice_bytes_numerator", "slice_bytes_denominator",
]


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def params_from_arrays(arr, qm):
    p = O.Params()
    for f, v in zip(PARAM_FIELDS, arr):
        setattr(p, f, int(v))
    for lv in range(O.MAX_LEVELS):
        for o in range(4):
            p.quant_matrix[lv][o] = int(qm[lv][o])
    return p


def picture_case_names():
    z = load("pictures")
    return [str(n) for n in z["names"]]


def corrupt_variants(z, name):
    out = []
    for v in range(4):
        key = "%s.corrupt%d" % (name, v)
        if key + ".data" in z.files:
            out.append(key)
    return out
