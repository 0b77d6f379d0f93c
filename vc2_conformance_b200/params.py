"""
Picture-path parameters: the subset of the reference's ``State``
(/root/reference/vc2_conformance/pseudocode/state.py:22-390) the CUDA path reads.
"""

from . import capi

ORIENT_SLOT = {"LL": 0, "L": 0, "H": 1, "HL": 1, "LH": 2, "HH": 3}

FIELDS = [
    "luma_width", "luma_height", "color_diff_width", "color_diff_height", "luma_depth",
    "color_diff_depth", "wavelet_index", "wavelet_index_ho", "dwt_depth", "dwt_depth_ho",
    "slices_x", "slices_y", "is_ld", "slice_prefix_bytes", "slice_size_scaler",
    "slice_bytes_numerator", "slice_bytes_denominator",
]


def make_params(
    luma_width,
    luma_height,
    color_diff_width=None,
    color_diff_height=None,
    luma_depth=10,
    color_diff_depth=None,
    wavelet_index=1,
    wavelet_index_ho=None,
    dwt_depth=3,
    dwt_depth_ho=0,
    slices_x=1,
    slices_y=1,
    is_ld=False,
    slice_prefix_bytes=0,
    slice_size_scaler=1,
    slice_bytes_numerator=0,
    slice_bytes_denominator=1,
    quant_matrix=None,
):
    """Build a ``capi.Params``.  ``quant_matrix`` is ``{level: {orient: value}}`` exactly as
    ``state["quant_matrix"]`` in the reference (pseudocode/state.py:120), or None (all zero)."""
    p = capi.Params()
    p.luma_width = int(luma_width)
    p.luma_height = int(luma_height)
    p.color_diff_width = int(luma_width if color_diff_width is None else color_diff_width)
    p.color_diff_height = int(luma_height if color_diff_height is None else color_diff_height)
    p.luma_depth = int(luma_depth)
    p.color_diff_depth = int(luma_depth if color_diff_depth is None else color_diff_depth)
    p.wavelet_index = int(wavelet_index)
    p.wavelet_index_ho = int(wavelet_index if wavelet_index_ho is None else wavelet_index_ho)
    p.dwt_depth = int(dwt_depth)
    p.dwt_depth_ho = int(dwt_depth_ho)
    p.slices_x = int(slices_x)
    p.slices_y = int(slices_y)
    p.is_ld = 1 if is_ld else 0
    p.slice_prefix_bytes = int(slice_prefix_bytes)
    p.slice_size_scaler = int(slice_size_scaler)
    p.slice_bytes_numerator = int(slice_bytes_numerator)
    p.slice_bytes_denominator = int(slice_bytes_denominator)
    if quant_matrix:
        for level, orients in quant_matrix.items():
            for orient, value in orients.items():
                p.quant_matrix[int(level)][ORIENT_SLOT[orient]] = int(value)
    return p


def params_from_state(state):
    """``capi.Params`` from a reference ``State`` (or plain dict) as left by
    ``transform_parameters`` (decoder/picture_syntax.py:94) + ``set_coding_parameters``
    (pseudocode/video_parameters.py:170)."""
    is_ld = (state.get("parse_code", 0xE8) & 0xF8) == 0xC8  # parse_code_functions.py:47
    return make_params(
        state["luma_width"],
        state["luma_height"],
        state["color_diff_width"],
        state["color_diff_height"],
        state.get("luma_depth", 8),
        state.get("color_diff_depth", state.get("luma_depth", 8)),
        state["wavelet_index"],
        state.get("wavelet_index_ho", state["wavelet_index"]),
        state["dwt_depth"],
        state.get("dwt_depth_ho", 0),
        state.get("slices_x", 1),
        state.get("slices_y", 1),
        is_ld,
        state.get("slice_prefix_bytes", 0),
        state.get("slice_size_scaler", 1),
        state.get("slice_bytes_numerator", 0),
        state.get("slice_bytes_denominator", 1),
        state.get("quant_matrix"),
    )


def params_from_arrays(values, quant_matrix):
    """From the flat layout used by tests/golden (FIELDS order, [16][4] matrix)."""
    p = capi.Params()
    for f, v in zip(FIELDS, values):
        setattr(p, f, int(v))
    for lv in range(capi.MAX_LEVELS):
        for o in range(4):
            p.quant_matrix[lv][o] = int(quant_matrix[lv][o])
    return p


def copy_params(p, **changes):
    q = capi.Params()
    for f in FIELDS:
        setattr(q, f, getattr(p, f))
    for lv in range(capi.MAX_LEVELS):
        for o in range(4):
            q.quant_matrix[lv][o] = p.quant_matrix[lv][o]
    for k, v in changes.items():
        setattr(q, k, v)
    return q


class Geometry(object):
    """Subband layout of a parameter set, queried from the library (never re-derived here)."""

    def __init__(self, p):
        import ctypes as C

        L = capi.lib()
        capi.check("vc2b_check_params", L.vc2b_check_params(C.byref(p)))
        self.params = p
        self.num_subbands = L.vc2b_num_subbands(C.byref(p))
        self.bands = []  # [comp][band] -> dict
        for comp in range(3):
            row = []
            for b in range(self.num_subbands):
                level, slot, w, h = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
                off = C.c_int64()
                capi.check(
                    "vc2b_subband_info",
                    L.vc2b_subband_info(C.byref(p), comp, b, C.byref(level), C.byref(slot), C.byref(w),
                                        C.byref(h), C.byref(off)),
                )
                row.append(dict(level=level.value, slot=slot.value, w=w.value, h=h.value, off=off.value))
            self.bands.append(row)
        self.comp_coeffs = [int(L.vc2b_component_coeff_count(C.byref(p), c)) for c in range(3)]
        self.picture_coeffs = int(L.vc2b_picture_coeff_count(C.byref(p)))
        self.picture_samples = int(L.vc2b_picture_sample_count(C.byref(p)))
        self.scratch_bytes = int(L.vc2b_transform_scratch_bytes(C.byref(p)))
        self.padded = []
        for comp in range(3):
            w, h = C.c_int32(), C.c_int32()
            L.vc2b_padded_dims(C.byref(p), comp, C.byref(w), C.byref(h))
            self.padded.append((w.value, h.value))
        self.dims = [
            (p.luma_width, p.luma_height),
            (p.color_diff_width, p.color_diff_height),
            (p.color_diff_width, p.color_diff_height),
        ]
        self.comp_off = [0, self.comp_coeffs[0], self.comp_coeffs[0] + self.comp_coeffs[1]]
        s0 = self.dims[0][0] * self.dims[0][1]
        s1 = self.dims[1][0] * self.dims[1][1]
        self.plane_off = [0, s0, s0 + s1]
        self.num_slices = p.slices_x * p.slices_y

    def orient_name(self, band):
        p = self.params
        b = self.bands[0][band]
        if b["level"] == 0:
            return "LL" if p.dwt_depth_ho == 0 else "L"
        if b["level"] <= p.dwt_depth_ho:
            return "H"
        return ["HL", "LH", "HH"][b["slot"] - 1]
