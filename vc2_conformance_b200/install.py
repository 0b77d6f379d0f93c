"""
Rebinds the reference's per-picture call sites to the CUDA implementations.

The reference has no plugin API: its modules bind the pseudocode functions by name at import time
(``from vc2_conformance.pseudocode... import picture_decode``) and look them up in their own module
namespace at call time.  ``install()`` replaces exactly those attributes; ``uninstall()`` restores
them.  Nothing else of the reference is touched: header parsing, level constraints, the
sequence-level loop and the output callback stay the reference's own code.

    import vc2_conformance_b200.install as gpu
    gpu.install()
    # ... vc2_conformance.decoder.parse_stream(state) now unpacks and transforms on the B200
"""

import importlib

from . import pseudocode as P

# (module, attribute) -> replacement; reference file:line of the call site in comments
BINDINGS = [
    # decoder/stream.py:50,129,133: picture_decode(state) after picture_parse / last fragment
    ("vc2_conformance.decoder.stream", "picture_decode", P.picture_decode),
    # decoder/picture_syntax.py:52,90: wavelet_transform -> transform_data(state)
    ("vc2_conformance.decoder.picture_syntax", "transform_data", P.transform_data),
    # decoder/fragment_syntax.py:28-32,140,168,175-181: fragmented pictures buffer slice bytes
    ("vc2_conformance.decoder.fragment_syntax", "initialize_fragment_state", P.initialize_fragment_state),
    ("vc2_conformance.decoder.fragment_syntax", "slice", P.fragment_slice),
    ("vc2_conformance.decoder.fragment_syntax", "dc_prediction", P.dc_prediction),
    # encoder/pictures.py:115,386: transform_and_slice_picture -> picture_encode(state, picture)
    ("vc2_conformance.encoder.pictures", "picture_encode", P.picture_encode),
    # encoder/pictures.py:131,217,251,270,295: the per-slice quantiser functions (lists in, lists out)
    ("vc2_conformance.encoder.pictures", "forward_quant", P.forward_quant),
    ("vc2_conformance.encoder.pictures", "calculate_coeffs_bits", P.calculate_coeffs_bits),
    ("vc2_conformance.encoder.pictures", "quantize_coeffs", P.quantize_coeffs),
    ("vc2_conformance.encoder.pictures", "quantize_to_fit", P.quantize_to_fit),
    # encoder/pictures.py:350,528,617,724 as called by make_picture_parse (:903-990): whole pictures, the
    # coefficients stay on the device between the transform and the quantiser
    ("vc2_conformance.encoder.pictures", "transform_and_slice_picture", P.transform_and_slice_picture),
    ("vc2_conformance.encoder.pictures", "make_transform_data_hq_lossy", P.make_transform_data_hq_lossy),
    ("vc2_conformance.encoder.pictures", "make_transform_data_hq_lossless", P.make_transform_data_hq_lossless),
    ("vc2_conformance.encoder.pictures", "make_transform_data_ld_lossy", P.make_transform_data_ld_lossy),
    # the star namespace third parties use (pseudocode/__init__.py:96-107)
    ("vc2_conformance.pseudocode", "picture_decode", P.picture_decode),
    ("vc2_conformance.pseudocode", "picture_encode", P.picture_encode),
    ("vc2_conformance.pseudocode", "idwt", P.idwt),
    ("vc2_conformance.pseudocode", "dwt", P.dwt),
    ("vc2_conformance.pseudocode", "h_synthesis", P.h_synthesis),
    ("vc2_conformance.pseudocode", "vh_synthesis", P.vh_synthesis),
    ("vc2_conformance.pseudocode", "h_analysis", P.h_analysis),
    ("vc2_conformance.pseudocode", "vh_analysis", P.vh_analysis),
    ("vc2_conformance.pseudocode", "offset_picture", P.offset_picture),
    ("vc2_conformance.pseudocode", "clip_picture", P.clip_picture),
    ("vc2_conformance.pseudocode", "remove_offset_picture", P.remove_offset_picture),
    ("vc2_conformance.pseudocode.picture_decoding", "picture_decode", P.picture_decode),
    ("vc2_conformance.pseudocode.picture_decoding", "idwt", P.idwt),
    ("vc2_conformance.pseudocode.picture_encoding", "picture_encode", P.picture_encode),
    ("vc2_conformance.pseudocode.picture_encoding", "dwt", P.dwt),
]

_saved = []


def install(strict=True):
    """Rebind the call sites.  Fails loudly (no CPU fallback) if the CUDA library is missing.
    Returns the list of (module, attribute) actually rebound."""
    from . import capi

    capi.lib()  # raises LibraryMissing if libvc2b200.so has not been built
    if _saved:
        return [(m, a) for m, a, _ in _saved]
    done = []
    for modname, attr, repl in BINDINGS:
        try:
            mod = importlib.import_module(modname)
        except ImportError:
            if strict:
                uninstall()
                raise
            continue
        if not hasattr(mod, attr):
            continue
        _saved.append((mod, attr, getattr(mod, attr)))
        setattr(mod, attr, repl)
        done.append((modname, attr))
    return done


def uninstall():
    while _saved:
        mod, attr, orig = _saved.pop()
        setattr(mod, attr, orig)


def installed():
    return bool(_saved)


def original(modname, attr):
    """The reference's own ``modname.attr``: the saved original while installed, else the live attribute."""
    for mod, a, orig in _saved:
        if mod.__name__ == modname and a == attr:
            return orig
    return getattr(importlib.import_module(modname), attr)
