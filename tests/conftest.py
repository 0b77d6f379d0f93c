import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    # The CUDA library and the C oracle are built artefacts (git-ignored): build them if a fresh
    # checkout is being tested.  nvcc cross-compiles without a GPU.
    from vc2_conformance_b200 import _build

    if not os.path.exists(_build.LIB):
        _build.build()
    import oracle

    oracle.build()


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when no device is present and -m gpu was not forced."""
    try:
        import torch

        have = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(params=["auto", "warp", "coeff16"])
def unpack_kernel(request):
    """Runs a test once with the library's own choice of unpack kernel (lane-per-slice where the
    slices are small), once with the warp-per-slice kernel forced (vc2b_set_unpack_kernel) and once with
    BatchCodec decoding through the int16 coefficient plane (vc2b_unpack16 / vc2b_idwt16, with its
    fall-back to the int32 plane for pictures that do not fit)."""
    from vc2_conformance_b200 import capi, device

    L = capi.lib()
    old16 = device.DEFAULT_COEFF16
    L.vc2b_set_unpack_kernel(1 if request.param == "warp" else 0)
    device.DEFAULT_COEFF16 = request.param == "coeff16"
    yield request.param
    L.vc2b_set_unpack_kernel(0)
    device.DEFAULT_COEFF16 = old16
