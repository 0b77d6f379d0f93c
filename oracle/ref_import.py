"""
ref_import -- put the UNMODIFIED Python reference (/root/reference) on
``sys.path`` together with the stand-in packages under ``oracle/refshim/``.

TEST INFRASTRUCTURE ONLY.  /root/reference exists only inside the build
container, never on the GPU box: the only callers are ``tools/make_golden.py``
(golden-vector generation, vectors committed under ``tests/golden/``),
``tools/time_reference.py`` and the ``-m "not gpu"`` tests that cross-check
the C oracle against the live reference *when it is present* (they skip
otherwise).  Nothing under ``vc2_conformance_b200/`` may import this module.

Stand-ins (oracle/refshim/): ``vc2_data_tables`` (un-vendored PyPI dependency,
pinned ``>=0.1.1,<2.0`` by /root/reference/setup.py:48; lifting tables restated
from SMPTE ST 2042-1), ``bitarray`` and ``sentinels`` (import-only helpers).
"""

import os
import sys

REFERENCE_ROOT = os.environ.get("VC2_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "vc2_conformance"))


def install():
    """Make ``import vc2_conformance`` resolve to the unmodified reference."""
    if not available():
        raise ImportError("reference tree not present at %s" % REFERENCE_ROOT)
    for p in (_SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import vc2_conformance  # noqa: F401

    return vc2_conformance
