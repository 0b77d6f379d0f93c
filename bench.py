#!/usr/bin/env python
"""
bench.py -- decoded pictures/s of the VC-2 picture path on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference            # the CPU arm (oracle port on all host cores)

One "step" decodes one synthetic sequence (default: BASELINE config 2 -- 120 pictures of
3840x2160 10-bit 4:2:2 HQ, Deslauriers-Dubuc (9,7) depth 4, 120x135 slices): slice-offset
index, slice unpack + inverse quantisation, IDWT + pad removal + clip + offset.

  value  pictures/s with the compressed pictures already resident in HBM (device timed)
  e2e    pictures/s through SequenceDecoder with HOST buffers: pinned host -> device copies of
         the streams and device -> pinned host copies of the decoded pictures inside the
         timed region
  roofline     the dominant stage of the step (by device time inside the timed region)
  cpu_baseline the oracle (CPU port of the reference algorithm) on the box's host cores on a
               bounded sample of the same workload

Multi-GPU: pictures are independent, so ranks decode their own sequences (weak scaling); there
is no collective on the data path -- only a MAX all-reduce of the device-timed elapsed time.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "cfg2": dict(
        name="3840x2160 10-bit 4:2:2 HQ, Deslauriers-Dubuc(9,7) depth 4, 120x135 slices, 120-picture sequence",
        params=dict(luma_width=3840, luma_height=2160, color_diff_width=1920, color_diff_height=2160,
                    luma_depth=10, wavelet_index=0, dwt_depth=4, slices_x=120, slices_y=135),
        pictures=120, bits_per_sample=2.5),
    # BASELINE.json configs[0]
    "cfg1": dict(
        name="1920x1080 10-bit 4:2:2 HQ, LeGall(5,3) depth 3, 60x135 slices, 120-picture sequence",
        params=dict(luma_width=1920, luma_height=1080, color_diff_width=960, color_diff_height=1080,
                    luma_depth=10, wavelet_index=1, dwt_depth=3, slices_x=60, slices_y=135),
        pictures=120, bits_per_sample=2.5),
    # BASELINE.json configs[0] read literally: 8 x 8 slices (uneven: 135/8 rows), 31980 + 2*15990 coefficients each
    "cfg1_8x8": dict(
        name="1920x1080 10-bit 4:2:2 HQ, LeGall(5,3) depth 3, 8x8 slices (uneven), 120-picture sequence",
        params=dict(luma_width=1920, luma_height=1080, color_diff_width=960, color_diff_height=1080,
                    luma_depth=10, wavelet_index=1, dwt_depth=3, slices_x=8, slices_y=8),
        pictures=120, bits_per_sample=2.5),
    "cfg3": dict(
        name="1920x1080 interlaced LD (1920x540 fields) 10-bit 4:2:2, Haar-with-shift depth 2, 60x135 slices, 4:1",
        params=dict(luma_width=1920, luma_height=540, color_diff_width=960, color_diff_height=540,
                    luma_depth=10, wavelet_index=4, dwt_depth=2, slices_x=60, slices_y=135, is_ld=True),
        pictures=240, bits_per_sample=2.5),
    "cfg4": dict(
        name="1920x1080 12-bit 4:4:4 HQ v3 asymmetric: Fidelity (V) + Daubechies(9,7) (H), depth 2 + ho 3, 60x270 slices",
        params=dict(luma_width=1920, luma_height=1080, luma_depth=12, wavelet_index=5, wavelet_index_ho=6,
                    dwt_depth=2, dwt_depth_ho=3, slices_x=60, slices_y=270),
        pictures=120, bits_per_sample=3.0),
    # BASELINE.json configs[4] (encoder path; used by tools/bench_stage.py dwt / quant)
    "cfg5": dict(
        name="7680x4320 12-bit 4:4:4 HQ encoder path, Deslauriers-Dubuc(9,7) depth 4, 240x270 slices",
        params=dict(luma_width=7680, luma_height=4320, luma_depth=12, wavelet_index=0, dwt_depth=4,
                    slices_x=240, slices_y=270),
        pictures=8, bits_per_sample=3.0),
}


# DRAM traffic per picture (dram__bytes_read.sum + dram__bytes_write.sum) from `ncu --set full`
# captures of these kernels on this workload; the summaries are committed under profiles/.
NCU_TRAFFIC_PER_PICTURE = {
    "cfg2": {
        "unpack": (71.6e6, "profiles/r1_v9_unpack_lanes_ncu_full_raw.csv (unpack_lanes_kernel, 32 pictures: 0.215 GB read + 2.077 GB written)"),
        "idwt": (144.0e6, "profiles/r1_v4_idwt_levels_ncu.csv (4 idwt_stream8_kernel launches, 16 pictures)"),
    },
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--pictures", type=int, default=0, help="override pictures per sequence")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--chunk", type=int, default=8)
    ap.add_argument("--group", type=int, default=128, help="pictures per IDWT launch group")
    return ap.parse_args()


def make_workload(name, npictures=0):
    import synth
    from vc2_conformance_b200.params import make_params

    w = dict(WORKLOADS[name])
    kw = dict(w["params"])
    d, dho = kw.get("dwt_depth", 0), kw.get("dwt_depth_ho", 0)
    kw["quant_matrix"] = synth.default_quant_matrix(d, dho)
    S = kw["luma_width"] * kw["luma_height"] + 2 * kw.get("color_diff_width", kw["luma_width"]) * kw.get(
        "color_diff_height", kw["luma_height"])
    nslices = kw["slices_x"] * kw["slices_y"]
    picture_bytes = int(S * w["bits_per_sample"] / 8) // 4 * 4
    if kw.get("is_ld"):
        kw["slice_bytes_numerator"] = picture_bytes
        kw["slice_bytes_denominator"] = nslices
    else:
        # smallest scaler whose 8-bit length fields can hold any slice (encoder/pictures.py:586)
        max_slice = (picture_bytes + nslices - 1) // nslices
        kw["slice_size_scaler"] = max(1, (max_slice - 4 + 254) // 255)
    w["p"] = make_params(**kw)
    w["picture_bytes"] = picture_bytes
    w["samples"] = S
    if npictures:
        w["pictures"] = npictures
    return w


def synth_pictures(w, first, count):
    """Half white noise (the reference's recipe), half integer ramps, alternating."""
    import synth

    out = np.empty((count, w["samples"]), np.uint16)
    for i in range(count):
        seed = first + i
        planes = synth.noise_picture(w["p"], seed) if seed % 2 == 0 else synth.ramp_picture(w["p"], seed)
        out[i] = synth.flat_picture(planes)
    return out


class ClockSampler(object):
    """SM clock and throttle reasons sampled DURING the timed regions (B200_PROFILING.md clocks line):
    NVML polled every 2 ms from a thread (nvidia_ml_py); nvidia-smi -lms as the fallback."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, uuid=None):
        self.index = index
        self.uuid = uuid
        self.samples = []   # (time, sm_mhz, sm_max_mhz, [reasons])
        self.proc = None
        self.thread = None
        self.running = False
        self.source = None

    def _nvml_handle(self):
        import pynvml as N

        N.nvmlInit()
        if self.uuid:
            try:
                return N, N.nvmlDeviceGetHandleByUUID(("GPU-" + str(self.uuid)).encode())
            except Exception:
                pass
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = self.index
        if vis:
            try:
                idx = int(vis.split(",")[self.index])
            except Exception:
                pass
        return N, N.nvmlDeviceGetHandleByIndex(idx)

    def _poll_nvml(self, N, h):
        bits = [(N.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"),
                (N.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (N.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                (N.nvmlClocksEventReasonSwPowerCap, "sw_power_cap")]
        try:
            smmax = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
        except Exception:
            smmax = 0.0
        while self.running:
            try:
                sm = float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM))
                try:
                    mask = N.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((time.time(), sm, smmax, [n for b, n in bits if mask & b]))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            N, h = self._nvml_handle()
            N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
            self.running = True
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, args=(N, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.running = False
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            try:
                sm, smmax = float(f[0]), float(f[1])
            except Exception:
                continue
            self.samples.append((time.time(), sm, smmax,
                                 [n for n, v in zip(self.NAMES, f[3:7]) if v.lower().startswith("active")]))

    def stop(self):
        self.running = False
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, windows):
        """windows: [(t0, t1), ...] wall-clock bounds of the timed regions"""
        sm, smmax, reasons = [], 0, set()
        for t, f, fmax, rs in list(self.samples):
            if not any(t0 <= t <= t1 for t0, t1 in windows):
                continue
            sm.append(f)
            smmax = max(smmax, fmax)
            reasons.update(rs)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": smmax,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# -------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm on host threads
# -------------------------------------------------------------------------------------------

def cpu_decode_rate(w, streams, budget_s, max_threads=None):
    """Decodes the given streams with the C oracle on all host cores (one picture per thread at
    a time; ctypes releases the GIL).  Returns (pictures/s, cores, pictures decoded, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    import synth

    po = synth.oracle_params(w["p"])
    O.lib()
    cores = max_threads or os.cpu_count() or 1
    done = [0]
    lock = threading.Lock()
    t_end = time.time() + budget_s
    order = list(range(len(streams)))

    def work(tid):
        k = tid
        while True:
            if time.time() > t_end and done[0] >= cores:
                return
            st, pics, consumed = O.decode_picture(po, streams[order[k % len(order)]])
            assert st == 0
            with lock:
                done[0] += 1
            k += cores
            if time.time() > t_end:
                return

    t0 = time.time()
    threads = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    dt = time.time() - t0
    return done[0] / dt, cores, done[0], dt


def oracle_streams(w, count):
    """A few streams of the workload encoded by the oracle itself (CPU arm input)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    import synth

    po = synth.oracle_params(w["p"])
    out = []
    for i in range(count):
        planes = synth.noise_picture(w["p"], i) if i % 2 == 0 else synth.ramp_picture(w["p"], i)
        coeffs = O.encode_picture(po, planes)
        if w["p"].is_ld:
            data, _q = O.encode_ld_lossy(po, coeffs)
        else:
            data, _q = O.encode_hq_lossy(po, coeffs, w["picture_bytes"])
        out.append(data)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w = make_workload(args.workload, args.pictures)
    # bounded sample: two distinct pictures of the workload, decoded repeatedly on all cores
    streams = oracle_streams(w, 2)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    import synth

    O.lib()
    po = synth.oracle_params(w["p"])
    cores = os.cpu_count() or 1
    per_step = cores  # pictures per step: one per core
    t_steps = []

    def work(i):
        st, pics, consumed = O.decode_picture(po, streams[i % len(streams)])
        assert st == 0

    for step in range(args.warmup + args.steps):
        t0 = time.time()
        threads = [threading.Thread(target=work, args=(i,)) for i in range(per_step)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if step >= args.warmup:
            t_steps.append(time.time() - t0)
    total = sum(t_steps)
    value = per_step * args.steps / total
    line = {
        "impl": "reference",
        "metric": "decoded pictures/s",
        "value": value,
        "unit": "pictures/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1000.0 * total / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "int64",
        "data": "synthetic",
        "config": {"workload": w["name"], "bits_per_sample": w["bits_per_sample"],
                   "sample": "%d pictures per step (one per host core), 2 distinct pictures" % per_step},
        "cpu_baseline": {"value": value, "unit": "pictures/s", "cores": cores, "kind": "port",
                         "sample": "%d full-size pictures per step x %d steps on %d threads"
                                   % (per_step, args.steps, cores)},
        "e2e": {"value": value, "unit": "pictures/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# -------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist

    from vc2_conformance_b200.device import PAD, BatchCodec, SequenceDecoder

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = "cuda:%d" % local_rank
    numa_cpus = 0
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
        if not os.environ.get("VC2B_NO_NUMA_BIND"):
            # one process per GPU: keep the pinned host buffers of this rank in the GPU's NUMA node
            from vc2_conformance_b200.sharding import bind_process_to_gpu

            try:
                uuid0 = torch.cuda.get_device_properties(local_rank).uuid
            except Exception:
                uuid0 = None
            numa_cpus = bind_process_to_gpu(local_rank, uuid0)

    w = make_workload(args.workload, args.pictures)
    p = w["p"]
    NP = w["pictures"]
    S = w["samples"]

    # ---- synthetic sequence, encoded on the device by the encoder mirror ---------------------
    codec = BatchCodec(p, NP, max_stream_bytes=NP * (w["picture_bytes"] + 64), device=dev, group=args.group)
    streams = []
    gen = 8
    for i0 in range(0, NP, gen):
        n = min(gen, NP - i0)
        pics = synth_pictures(w, rank * NP + i0, n)
        streams += codec.encode_pictures(pics, w["picture_bytes"])
    host, offs = codec.pack_host_streams(streams)
    stream_bytes = int(offs[NP])
    cur = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # resident copy of the compressed sequence
    codec._reserve_data(host.numel())
    codec.data[: host.numel()].copy_(host)
    codec.pic_offset[: NP + 1].copy_(offs)
    torch.cuda.synchronize()

    # correctness gate of the timed path: decode once, compare with the source of a lossless
    # re-encode?  (parity itself is the job of tests/; here: status words only)
    codec.decode_device(codec.data, codec.pic_offset, NP)
    st = codec.status_host(NP)
    assert (st[:, 0] == 0).all(), "decode status not OK"
    assert (st[:, 2] == np.diff(offs.numpy())).all(), "consumed bytes mismatch"

    try:
        uuid = torch.cuda.get_device_properties(local_rank).uuid
    except Exception:
        uuid = None
    sampler = ClockSampler(local_rank, uuid)
    sampler.start()

    # ---- device-resident timed region --------------------------------------------------------
    for _ in range(args.warmup):
        codec.decode_device(codec.data, codec.pic_offset, NP)
    barrier()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = codec.launches
    t_wall0 = time.time()
    e_start.record(cur)
    for k in range(args.steps):
        ev[k][0].record(cur)
        # stage 1: slice index + unpack (+ dc prediction for LD)
        import ctypes as C

        from vc2_conformance_b200 import capi
        from vc2_conformance_b200.device import _ptr

        L = codec.lib
        sp = codec._stream_ptr(None)
        if not p.is_ld:
            capi.check("vc2b_hq_slice_offsets",
                       L.vc2b_hq_slice_offsets(C.byref(p), _ptr(codec.data), _ptr(codec.pic_offset), NP,
                                               _ptr(codec.slice_offset), _ptr(codec.status), sp))
        capi.check("vc2b_unpack",
                   L.vc2b_unpack(C.byref(p), _ptr(codec.data), _ptr(codec.pic_offset), NP,
                                 _ptr(codec.slice_offset), _ptr(codec.coeffs), _ptr(codec.qindex),
                                 _ptr(codec.status), sp))
        ev[k][1].record(cur)
        codec.idwt_device(NP)
        ev[k][2].record(cur)
    e_end.record(cur)
    barrier()
    t_wall1 = time.time()
    launches = codec.launches - launches0
    ms_total = e_start.elapsed_time(e_end)
    ms_unpack = sum(ev[k][0].elapsed_time(ev[k][1]) for k in range(args.steps)) / args.steps
    ms_idwt = sum(ev[k][1].elapsed_time(ev[k][2]) for k in range(args.steps)) / args.steps
    t_ms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * NP * args.steps / (ms_max / 1000.0)
    clock_windows = [(t_wall0, t_wall1)]

    # ---- end-to-end timed region (host buffers) ----------------------------------------------
    e2e = None
    if not args.no_e2e:
        chunk = args.chunk
        max_chunk = max(int(offs[min(NP, i + chunk)] - offs[i]) for i in range(0, NP, chunk)) + PAD + 64
        seq = SequenceDecoder(p, chunk=chunk, nstreams=3, max_chunk_bytes=max_chunk, device=dev)
        chunks = seq.plan(host, offs.numpy())
        out_host = torch.empty((NP, S), dtype=torch.uint16).pin_memory()
        for _ in range(max(1, args.warmup)):
            seq.fork_from(cur)
            seq.decode(chunks, out_host)
            seq.join_to(cur)
        barrier()
        # the e2e path must produce exactly what the resident path produced
        ref_last = codec.pictures[(NP - 1) * S: NP * S].cpu()
        assert torch.equal(out_host[NP - 1].view(torch.int16), ref_last.view(torch.int16)), "e2e output mismatch"
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = seq.launches
        t_e2e0 = time.time()
        s0.record(cur)
        for _ in range(args.steps):
            seq.fork_from(cur)
            seq.decode(chunks, out_host)
            seq.join_to(cur)
        s1.record(cur)
        barrier()
        clock_windows.append((t_e2e0, time.time()))
        e2e_ms = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
        e2e = {
            "value": world * NP * args.steps / (float(e2e_ms.item()) / 1000.0),
            "unit": "pictures/s",
            "h2d_bytes_per_step": int(stream_bytes + 8 * (NP + len(chunks))),
            "d2h_bytes_per_step": int(2 * S * NP),
            "gpu_launches": seq.launches - l0,
            "pipeline": "%d-picture chunks over 3 CUDA streams" % chunk,
            "numa_bound_cpus": numa_cpus,
        }
    sampler.stop()
    clocks = sampler.summary(clock_windows)

    # ---- roofline of the stages (algorithmic bytes) -----------------------------
    peak, peak_src = measured_peak()
    padded = sum(codec.g.comp_coeffs)
    idwt_bytes = (4 * padded + 2 * S) * NP            # int32 coefficients in, uint16 pixels out
    unpack_bytes = stream_bytes + 4 * padded * NP     # compressed bytes in, int32 coefficients out
    idwt_gbs = idwt_bytes / (ms_idwt / 1000.0) / 1e9
    unpack_gbs = unpack_bytes / (ms_unpack / 1000.0) / 1e9
    stages = {
        "unpack": {"ms_per_step": ms_unpack, "achieved": unpack_gbs, "frac": unpack_gbs / peak,
                   "algorithmic_bytes_per_picture": unpack_bytes // NP,
                   "kernels": "hq_walk_kernel + unpack_lanes_kernel<HQ>" if not p.is_ld
                   else "unpack_lanes_kernel<LD> + dc_prediction_kernel"},
        "idwt": {"ms_per_step": ms_idwt, "achieved": idwt_gbs, "frac": idwt_gbs / peak,
                 "algorithmic_bytes_per_picture": idwt_bytes // NP,
                 "kernels": "idwt_stream8_kernel x %d levels (last fused with crop+clip+offset)"
                            % (p.dwt_depth + p.dwt_depth_ho)},
    }
    dom = "unpack" if ms_unpack >= ms_idwt else "idwt"
    traffic, traffic_src = None, None
    if args.workload in NCU_TRAFFIC_PER_PICTURE and dom in NCU_TRAFFIC_PER_PICTURE[args.workload]:
        per_pic, traffic_src = NCU_TRAFFIC_PER_PICTURE[args.workload][dom]
        traffic = per_pic * NP  # per launch of the stage, like `achieved`
    roofline = {"bound": "hbm", "stage": dom, "achieved": stages[dom]["achieved"], "peak": peak, "unit": "GB/s",
                "frac": stages[dom]["frac"], "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes": stages[dom]["algorithmic_bytes_per_picture"] * NP, "peak_source": peak_src}

    # ---- CPU baseline (rank 0, N=1 only) -------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = streams[:4]
        rate, cores, npics, secs = cpu_decode_rate(w, sample, budget_s=12.0)
        cpu = {"value": rate, "unit": "pictures/s", "cores": cores, "kind": "port",
               "sample": "%d full-size pictures (4 distinct) decoded by the C oracle on %d threads in %.1f s"
                         % (npics, cores, secs)}

    if rank == 0:
        line = {
            "metric": "decoded pictures/s",
            "value": value,
            "unit": "pictures/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "int32",
            "data": "synthetic",
            "config": {"workload": w["name"], "pictures_per_step_per_gpu": NP,
                       "bits_per_sample": w["bits_per_sample"], "stream_bytes_per_picture": stream_bytes // NP,
                       "slice_size_scaler": int(p.slice_size_scaler),
                       "pictures": "even seeds white noise (picture_generators.py:600 recipe), odd seeds integer ramps",
                       "l2": "inputs larger than L2 (%.0f MB of streams, %.0f MB of coefficients per step)"
                             % (stream_bytes / 1e6, 4 * padded * NP / 1e6),
                       "idwt_group": args.group},
            "e2e": e2e,
            "roofline": roofline,
            "stages": stages,
            "cpu_baseline": cpu,
            "gpu_launches": launches,
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
