"""
ref_import -- put the UNMODIFIED Python reference (/root/reference) on
``sys.path`` together with the stand-in packages under ``oracle/refshim/``.

TEST INFRASTRUCTURE ONLY.  /root/reference exists only inside the build
container; the GPU box gets the git-ignored ``baseline/_ref`` copy made by
``pip install --no-deps --target baseline/_ref /root/reference`` (unmodified
sources).  Callers: ``tools/make_golden*.py`` (golden-vector
generation, vectors committed under ``tests/golden/``), ``tools/time_reference.py``
(the Python reference timed as a CPU baseline), and the tests that drive the
unmodified reference through ``vc2_conformance_b200.install`` (they skip when no
reference tree is present).  Nothing under ``vc2_conformance_b200/`` may import
this module.

Stand-ins (oracle/refshim/): ``vc2_data_tables`` (un-vendored PyPI dependency,
pinned ``>=0.1.1,<2.0`` by /root/reference/setup.py:48; lifting tables restated
from SMPTE ST 2042-1), ``bitarray`` and ``sentinels`` (import-only helpers).
"""

import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "refshim")
# where the unmodified reference may live: the read-only tree of the build container, or the
# git-ignored `pip install --target baseline/_ref /root/reference` copy that travels to the GPU box
_CANDIDATES = [os.environ.get("VC2_REFERENCE_ROOT"), "/root/reference",
               os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]


def _find():
    for root in _CANDIDATES:
        if root and os.path.isdir(os.path.join(root, "vc2_conformance")):
            return root
    return None


REFERENCE_ROOT = _find()


def available():
    return REFERENCE_ROOT is not None


def install():
    """Make ``import vc2_conformance`` resolve to the unmodified reference."""
    if not available():
        raise ImportError("reference tree not present (looked in %s)" % [c for c in _CANDIDATES if c])
    for p in (_SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import vc2_conformance  # noqa: F401

    return vc2_conformance
