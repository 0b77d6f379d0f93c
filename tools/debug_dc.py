"""Times vc2b_dc_prediction (decoder and encoder direction) on the cfg3 DC bands (developer tool)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import bench  # noqa: E402
from vc2_conformance_b200 import capi  # noqa: E402

w = bench.make_workload("cfg3", 64)
p = w["p"]
L = capi.lib()
n = 64
ncoef = L.vc2b_picture_coeff_count(ctypes.byref(p))
coeffs = torch.randint(-200, 200, (n, ncoef), dtype=torch.int32, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for mode in [0, 1]:
    for _ in range(2):
        capi.check("vc2b_dc_prediction", L.vc2b_dc_prediction(ctypes.byref(p), coeffs.data_ptr(), n, mode, s))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        capi.check("vc2b_dc_prediction", L.vc2b_dc_prediction(ctypes.byref(p), coeffs.data_ptr(), n, mode, s))
    e1.record()
    torch.cuda.synchronize()
    print("inverse=%d: %.1f us per %d fields" % (mode,
          e0.elapsed_time(e1) * 1000 / 5, n))
