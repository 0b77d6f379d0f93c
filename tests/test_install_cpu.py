"""Host-side logic that needs no GPU: the C-ABI library loads and exports every symbol the header
declares, geometry agrees with the oracle, and install()/uninstall() rebind the reference's call
sites (when the reference tree is present, i.e. in the build container)."""

import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from vc2_conformance_b200 import capi

    L = capi.lib()
    header = open(os.path.join(ROOT, "include", "vc2_b200.h")).read()
    declared = set(re.findall(r"\b(vc2b_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(L, name), name
        assert name in capi.SIGNATURES, "capi.py does not bind %s" % name
    assert L.vc2b_version().startswith(b"vc2b200")


def test_params_struct_layout_matches_oracle():
    import oracle as O
    from vc2_conformance_b200 import capi

    assert C.sizeof(capi.Params) == C.sizeof(O.Params)
    for (n1, t1), (n2, t2) in zip(capi.Params._fields_, O.Params._fields_):
        assert n1 == n2 and C.sizeof(t1) == C.sizeof(t2)


@pytest.mark.parametrize("kw", [
    dict(luma_width=1920, luma_height=1080, color_diff_width=960, color_diff_height=1080, dwt_depth=3),
    dict(luma_width=50, luma_height=30, color_diff_width=25, color_diff_height=15, dwt_depth=2, dwt_depth_ho=1),
    dict(luma_width=3840, luma_height=2160, color_diff_width=1920, color_diff_height=2160, dwt_depth=4),
    dict(luma_width=7, luma_height=5, dwt_depth=0, dwt_depth_ho=3),
])
def test_geometry_matches_oracle(kw):
    import oracle as O
    import synth
    from vc2_conformance_b200.params import Geometry, make_params

    p = make_params(**kw)
    g = Geometry(p)
    po = synth.oracle_params(p)
    for comp in range(3):
        level, orient, w, h, off, total = O.layout(po, comp)
        assert g.comp_coeffs[comp] == total
        assert g.padded[comp] == O.padded_dims(po, comp)
        for b in range(g.num_subbands):
            B = g.bands[comp][b]
            assert (B["level"], B["slot"], B["w"], B["h"], B["off"]) == (level[b], orient[b], w[b], h[b], off[b])


def test_bad_params_are_rejected():
    from vc2_conformance_b200 import capi
    from vc2_conformance_b200.params import make_params

    L = capi.lib()
    assert L.vc2b_check_params(C.byref(make_params(0, 10))) == capi.E_BAD_PARAMS
    assert L.vc2b_check_params(C.byref(make_params(16, 16, wavelet_index=9))) == capi.E_BAD_PARAMS
    assert L.vc2b_check_params(C.byref(make_params(16, 16, luma_depth=20))) == capi.E_UNSUPPORTED
    assert L.vc2b_check_params(C.byref(make_params(16, 16, slices_x=0))) == capi.E_BAD_PARAMS
    assert L.vc2b_check_params(C.byref(make_params(16, 16))) == 0


def test_envelope_bound_is_sane():
    from vc2_conformance_b200 import capi
    from vc2_conformance_b200.params import make_params

    L = capi.lib()
    # 10-bit pictures through 4 levels of any filter stay far inside the int32 envelope
    for wi in range(7):
        p = make_params(3840, 2160, wavelet_index=wi, dwt_depth=4)
        assert L.vc2b_idwt_max_safe_coeff(C.byref(p)) > (1 << 17)


def test_host_quant_tables_match_oracle():
    import oracle as O
    from vc2_conformance_b200 import pseudocode as P

    for i in range(120):
        assert P.quant_factor(i) == O.lib().vc2o_quant_factor(i)
        assert P.quant_offset(i) == O.lib().vc2o_quant_offset(i)


def test_install_rebinds_reference_call_sites():
    import ref_import

    if not ref_import.available():
        pytest.skip("reference tree not present (only in the build container)")
    ref_import.install()
    import vc2_conformance.decoder.stream as stream
    import vc2_conformance.decoder.picture_syntax as picture_syntax
    import vc2_conformance.encoder.pictures as enc_pictures
    from vc2_conformance_b200 import install, pseudocode

    orig = (stream.picture_decode, picture_syntax.transform_data, enc_pictures.picture_encode)
    done = install.install()
    try:
        assert ("vc2_conformance.decoder.stream", "picture_decode") in done
        assert stream.picture_decode is pseudocode.picture_decode
        assert picture_syntax.transform_data is pseudocode.transform_data
        assert enc_pictures.picture_encode is pseudocode.picture_encode
    finally:
        install.uninstall()
    assert (stream.picture_decode, picture_syntax.transform_data, enc_pictures.picture_encode) == orig


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "vc2_conformance_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "vc2_oracle" not in src and "libvc2oracle" not in src, f
