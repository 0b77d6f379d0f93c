"""
Raw planar picture files and picture comparison on the device.

Mirrors ``vc2_conformance.file_format.write_picture`` / ``read_picture`` (file_format.py:158-194,
:263-314) and ``scripts/vc2_picture_compare.py`` ``psnr`` / ``measure_differences`` (:370-420) for
pictures that live in a :class:`~vc2_conformance_b200.device.BatchCodec` layout (planes Y, C1, C2 as
uint16): the per-sample work runs in CUDA kernels behind the C ABI, the host only moves whole buffers
and evaluates the two logarithms of the PSNR.
"""

import ctypes as C
import math

import numpy as np

from . import capi
from .device import _DeviceStream, _ptr, _torch

COMPONENTS = ("Y", "C1", "C2")


def raw_picture_bytes(params):
    n = int(capi.lib().vc2b_raw_picture_bytes(C.byref(params)))
    if n < 0:
        raise capi.Vc2bError("vc2b_raw_picture_bytes", n)
    return n


def pictures_to_raw(params, pictures, n, stream=None):
    """``pictures``: device uint16 tensor holding ``n`` pictures.  Returns a pinned host uint8 tensor
    ``[n, raw_picture_bytes]``: exactly the bytes ``write_picture`` writes for each picture."""
    torch = _torch()
    nbytes = raw_picture_bytes(params)
    # allocation, kernel and read-back all on ONE stream of the pictures' device
    with _DeviceStream(torch, pictures.device, stream):
        raw = torch.empty((n, nbytes), dtype=torch.uint8, device=pictures.device)
        sp = C.c_void_p(torch.cuda.current_stream(pictures.device).cuda_stream)
        capi.check("vc2b_pictures_to_raw",
                   capi.lib().vc2b_pictures_to_raw(C.byref(params), _ptr(pictures), n, _ptr(raw), nbytes, sp))
        host = torch.empty((n, nbytes), dtype=torch.uint8).pin_memory()
        host.copy_(raw, non_blocking=False)
    return host


def write_pictures(params, pictures, n, file):
    """Appends ``n`` device pictures to ``file`` (open for binary writing) in the raw format."""
    file.write(pictures_to_raw(params, pictures, n).numpy().tobytes())


def read_pictures(params, file, n, device="cuda:0"):
    """Reads ``n`` raw pictures from ``file`` into a device uint16 tensor ``[n * picture_samples]``;
    bits above the sample depth are masked off as ``read_picture`` does."""
    torch = _torch()
    nbytes = raw_picture_bytes(params)
    data = file.read(n * nbytes)
    if len(data) != n * nbytes:
        raise EOFError("raw picture file ends early: wanted %d bytes, got %d" % (n * nbytes, len(data)))
    samples = int(capi.lib().vc2b_picture_sample_count(C.byref(params)))
    with _DeviceStream(torch, torch.device(device), None):
        raw = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(device)
        pictures = torch.empty(n * samples, dtype=torch.uint16, device=device)
        sp = C.c_void_p(torch.cuda.current_stream(pictures.device).cuda_stream)
        capi.check("vc2b_raw_to_pictures",
                   capi.lib().vc2b_raw_to_pictures(C.byref(params), _ptr(raw), nbytes, n, _ptr(pictures), sp))
    return pictures


def psnr_from_sums(sum_sq, count, max_value):
    """``psnr`` of vc2_picture_compare.py:370 from the sum of squared errors of ``count`` samples;
    None when the pictures are identical."""
    if sum_sq == 0:
        return None
    mse = float(sum_sq) / float(count)
    return (20 * (math.log(max_value) / math.log(10))) - (10 * (math.log(mse) / math.log(10)))


def compare_pictures(params, a, b, n):
    """Per picture and component: (samples that differ, PSNR in dB or None).  ``a``, ``b``: device
    uint16 tensors with ``n`` pictures each."""
    torch = _torch()
    with _DeviceStream(torch, a.device, None):
        out = torch.zeros(n * 6, dtype=torch.int64, device=a.device)
        sp = C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)
        capi.check("vc2b_picture_compare",
                   capi.lib().vc2b_picture_compare(C.byref(params), _ptr(a), _ptr(b), n, _ptr(out), sp))
        res = out.cpu().numpy().reshape(n, 3, 2)
    dims = [(params.luma_width * params.luma_height, params.luma_depth)] + \
        [(params.color_diff_width * params.color_diff_height, params.color_diff_depth)] * 2
    return [[(int(res[i, c, 0]), psnr_from_sums(int(res[i, c, 1]), dims[c][0], (1 << dims[c][1]) - 1))
             for c in range(3)] for i in range(n)]


def measure_differences(params, a, b, index=0, n=1):
    """The (identical, text) pair of vc2_picture_compare.py:384 ``measure_differences`` for picture
    ``index`` of the two device buffers."""
    rows = compare_pictures(params, a, b, max(n, index + 1))[index]
    sizes = [params.luma_width * params.luma_height] + [params.color_diff_width * params.color_diff_height] * 2
    identical = all(p is None for _c, p in rows)
    lines = []
    for name, (count, p), size in zip(COMPONENTS, rows, sizes):
        if p is None:
            lines.append("{}: Identical".format(name))
        else:
            lines.append("{}: Different: PSNR = {:.1f} dB, {} pixel{} ({:.1f}%) differ{}".format(
                name, p, count, "s" if count != 1 else "", (count * 100.0) / size, "s" if count == 1 else ""))
    return identical, "\n".join(lines)
