"""Stage micro-benchmarks (device timed): python tools/bench_stage.py idwt|dwt|unpack|walk|quant|pack|dc|dcinv [--workload cfg2] [--pictures 16]"""
import argparse, os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, torch
import bench
from vc2_conformance_b200 import capi
from vc2_conformance_b200.device import BatchCodec, _ptr

ap = argparse.ArgumentParser()
ap.add_argument("stage")
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--pictures", type=int, default=16)
ap.add_argument("--group", type=int, default=128)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--int32-plane", action="store_true")
a = ap.parse_args()
w = bench.make_workload(a.workload, a.pictures)
p, NP, S = w["p"], w["pictures"], w["samples"]
codec = BatchCodec(p, NP, max_stream_bytes=NP * (w["picture_bytes"] + 64), group=a.group, coeff16=not a.int32_plane)
pics = bench.synth_pictures(w, 0, 0, min(NP, 4))
streams = codec.encode_pictures(pics, w["picture_bytes"])
streams = [streams[i % len(streams)] for i in range(NP)]
host, offs = codec.pack_host_streams(streams)
codec.decode_host(host, offs, NP)
torch.cuda.synchronize()
_st = codec.status_host(NP)
codec.redo_coeff16(_st)   # pictures whose coefficients do not fit the int16 plane (12-bit 8K) go through int32
assert (_st[:, 0] == 0).all()
L = codec.lib
sp = codec._stream_ptr(None)
padded = sum(codec.g.comp_coeffs)

def run_idwt():
    codec.idwt_device(NP)
def run_dwt():
    codec.dwt_device(NP)
def run_unpack():
    codec.unpack_device(codec.data, codec.pic_offset, NP)
def run_walk():
    L.vc2b_hq_slice_offsets(C.byref(p), _ptr(codec.data), _ptr(codec.pic_offset), NP, _ptr(codec.slice_offset), _ptr(codec.status), sp)
def run_dc():      # apply_dc_prediction then dc_prediction: the band returns to its values, so iterations are alike
    L.vc2b_dc_prediction(C.byref(p), _ptr(codec.coeffs), NP, 1, sp)
    L.vc2b_dc_prediction(C.byref(p), _ptr(codec.coeffs), NP, 0, sp)
def run_dcinv():
    L.vc2b_dc_prediction(C.byref(p), _ptr(codec.coeffs), NP, 1, sp)
def run_quant():
    codec.quantize_device(NP, w["picture_bytes"])
def run_pack():
    codec.pack_device(NP, w["picture_bytes"])
fn = {"idwt": run_idwt, "dwt": run_dwt, "unpack": run_unpack, "walk": run_walk, "quant": run_quant, "pack": run_pack,
      "dc": run_dc, "dcinv": run_dcinv}[a.stage]
nbytes = {"idwt": (4 * padded + 2 * S) * NP, "dwt": (4 * padded + 2 * S) * NP,
          "unpack": int(offs[NP]) + 4 * padded * NP, "walk": int(offs[NP]), "quant": 4 * padded * NP,
          "pack": int(offs[NP]) + 4 * padded * NP, "dc": 0, "dcinv": 0}[a.stage]
for _ in range(3):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
print("[plane %s] %s %s: %.3f ms per %d pictures = %.1f us/picture, %.1f GB/s algorithmic (%.1f%% of 6542)" % (
    "int16" if codec.coeff16 else "int32", a.stage, a.workload, ms, NP, 1000 * ms / NP, nbytes / ms / 1e6, 100 * nbytes / ms / 1e6 / 6542.4))
