"""Experiment: overlap the (issue-bound) unpack of sub-batch k+1 with the (HBM-bound) IDWT of sub-batch k."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, torch
import bench
from vc2_conformance_b200 import capi
from vc2_conformance_b200.device import BatchCodec, _ptr

w = bench.make_workload("cfg2", 120)
p, NP, S = w["p"], w["pictures"], w["samples"]
codec = BatchCodec(p, NP, max_stream_bytes=NP * (w["picture_bytes"] + 64))
pics = bench.synth_pictures(w, 0, 0, 4)
streams = codec.encode_pictures(pics, w["picture_bytes"])
streams = [streams[i % 4] for i in range(NP)]
host, offs = codec.pack_host_streams(streams)
codec.decode_host(host, offs, NP)
torch.cuda.synchronize()
L = codec.lib
g = codec.g

def seq():
    codec.decode_device(codec.data, codec.pic_offset, NP)

main = torch.cuda.current_stream()
def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

print("sequential: %.3f ms per %d pictures" % (timeit(seq), NP))

for K in (2, 3, 4, 6):
    for prio in (False, True):
        sA = torch.cuda.Stream(priority=-1 if prio else 0)   # IDWT stream (high priority)
        sB = torch.cuda.Stream()                             # unpack stream
        sub = NP // K
        scratch2 = [torch.empty(max(g.scratch_bytes * sub, 16), dtype=torch.uint8, device="cuda") for _ in range(2)]
        offs_dev = [codec.pic_offset[k * sub:(k + 1) * sub + 1] for k in range(K)]
        def ov():
            ev0 = torch.cuda.Event(); ev0.record(main)
            sA.wait_event(ev0); sB.wait_event(ev0)
            done = []
            for k in range(K):
                first = k * sub
                spB = C.c_void_p(sB.cuda_stream)
                co = codec.coeffs[first * g.picture_coeffs:]
                stt = codec.status[first * 4:]
                qi = codec.qindex[first * g.num_slices:]
                so = codec.slice_offset[first * g.num_slices:]
                L.vc2b_hq_slice_offsets(C.byref(p), _ptr(codec.data), _ptr(offs_dev[k]), sub, _ptr(so), _ptr(stt), spB)
                L.vc2b_unpack(C.byref(p), _ptr(codec.data), _ptr(offs_dev[k]), sub, _ptr(so), _ptr(co), _ptr(qi), _ptr(stt), spB)
                ev = torch.cuda.Event(); ev.record(sB)
                sA.wait_event(ev)
                spA = C.c_void_p(sA.cuda_stream)
                L.vc2b_idwt(C.byref(p), _ptr(co), sub, _ptr(codec.pictures[first * g.picture_samples:]),
                            _ptr(scratch2[k % 2]), scratch2[k % 2].numel(), spA)
            evA = torch.cuda.Event(); evA.record(sA); main.wait_event(evA)
            evB = torch.cuda.Event(); evB.record(sB); main.wait_event(evB)
        print("overlap K=%d prio=%s: %.3f ms" % (K, prio, timeit(ov)))
