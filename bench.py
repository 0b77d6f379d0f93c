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


# DRAM traffic per picture (dram__bytes_read.sum + dram__bytes_write.sum) from `ncu --set full` captures of
# these kernels on this workload: profiles/traffic.json names the committed summary each figure comes from.
def ncu_traffic(workload, plane, stage):
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            e = json.load(f)[workload][plane][stage]
        return float(e["bytes_per_picture"]), e["source"]
    except (OSError, KeyError, ValueError):
        return None, None


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="decode", choices=["decode", "encode"],
                    help="decode: unpack + IDWT (the headline metric); encode: DWT + quantize_to_fit + slice packing "
                         "(BASELINE config 5: --workload cfg5 --mode encode)")
    ap.add_argument("--no-py-reference", action="store_true",
                    help="skip the unmodified Python reference leg of the CPU baseline (baseline/_ref)")
    ap.add_argument("--pictures", type=int, default=0, help="override pictures per sequence")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-config-sweep", action="store_true",
                    help="skip the stage timings of the other BASELINE configurations (stages_by_config)")
    ap.add_argument("--int32-plane", action="store_true",
                    help="stage coefficients as int32 between unpack and IDWT (default: int16 high-pass bands)")
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


def picture_kind(w, i):
    """Picture i of a sequence: the first half white noise (the reference's recipe), the second half integer
    ramps -- each kind a contiguous run, so that the bench can also time them separately."""
    return "noise" if i < (w["pictures"] + 1) // 2 else "ramp"


def synth_picture(w, rank, i):
    import synth

    seed = rank * w["pictures"] + i
    return synth.noise_picture(w["p"], seed) if picture_kind(w, i) == "noise" else synth.ramp_picture(w["p"], seed)


def synth_pictures(w, rank, first, count):
    import synth

    out = np.empty((count, w["samples"]), np.uint16)
    for i in range(count):
        out[i] = synth.flat_picture(synth_picture(w, rank, first + i))
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
# CPU arm: the oracle port of the reference algorithm on host threads (and, where the reference tree is
# present, the unmodified Python reference itself: tools/time_reference.py)
# -------------------------------------------------------------------------------------------

def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O

    O.lib()
    return O


def oracle_encode(O, po, w, planes):
    coeffs = O.encode_picture(po, planes)
    if w["p"].is_ld:
        return O.encode_ld_lossy(po, coeffs)[0]
    return O.encode_hq_lossy(po, coeffs, w["picture_bytes"])[0]


def cpu_rate(w, mode, items, budget_s, check=None, max_threads=None):
    """Runs the C oracle over ``items`` on all host cores (one picture per thread at a time; ctypes releases
    the GIL) until ``budget_s`` has passed and every thread has finished a picture.
    decode: items are streams, encode: items are [Y, C1, C2] planes.  ``check(i, result)`` sees the first
    result of every item (the oracle doubling as the checker of the GPU path's output).
    Returns (pictures/s, cores, pictures, seconds)."""
    import synth

    O = _oracle()
    po = synth.oracle_params(w["p"])
    cores = max_threads or os.cpu_count() or 1
    done = [0]
    lock = threading.Lock()
    seen = set()
    t_end = time.time() + budget_s
    errors = []

    def work(tid):
        k = tid
        while True:
            if time.time() > t_end and done[0] >= cores:
                return
            i = k % len(items)
            try:
                if mode == "decode":
                    st, res, _consumed = O.decode_picture(po, items[i])
                    assert st == 0
                else:
                    res = oracle_encode(O, po, w, items[i])
                with lock:
                    done[0] += 1
                    first = i not in seen
                    seen.add(i)
                if first and check is not None:
                    check(i, res)
            except Exception as e:  # noqa: BLE001 -- reported by the caller
                errors.append(e)
                return
            k += cores
            if time.time() > t_end:
                return

    t0 = time.time()
    threads = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    dt = time.time() - t0
    return done[0] / dt, cores, done[0], dt


def python_reference_baseline(args, mode):
    """The unmodified Python reference on a cropped picture, one process per core (tools/time_reference.py)."""
    if args.no_py_reference:
        return None
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import time_reference

        return time_reference.measure(args.workload, (512, 256), None, mode)
    except Exception as e:  # noqa: BLE001 -- a reported baseline must not take the GPU line down
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


def metric_name(mode):
    return "decoded pictures/s" if mode == "decode" else "encoded pictures/s"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import synth

    w = make_workload(args.workload, args.pictures)
    O = _oracle()
    po = synth.oracle_params(w["p"])
    # bounded sample: two distinct pictures of the workload (one noise, one ramp), processed repeatedly
    planes = [synth.noise_picture(w["p"], 0), synth.ramp_picture(w["p"], 1)]
    items = planes if args.mode == "encode" else [oracle_encode(O, po, w, pl) for pl in planes]
    cores = os.cpu_count() or 1
    per_step = cores  # pictures per step: one per core
    t_steps = []

    def work(i):
        if args.mode == "decode":
            st, _pics, _consumed = O.decode_picture(po, items[i % len(items)])
            assert st == 0
        else:
            oracle_encode(O, po, w, items[i % len(items)])

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
        "metric": metric_name(args.mode),
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
        "config": {"workload": w["name"], "mode": args.mode, "bits_per_sample": w["bits_per_sample"],
                   "sample": "%d pictures per step (one per host core), 2 distinct pictures" % per_step},
        "cpu_baseline": {"value": value, "unit": "pictures/s", "cores": cores, "kind": "port",
                         "sample": "%d full-size pictures per step x %d steps on %d threads (C restatement of the "
                                   "reference algorithm, oracle/vc2_oracle.c; the reference itself is Python)"
                                   % (per_step, args.steps, cores)},
        # the unmodified Python reference on a cropped picture, for scale
        "cpu_baseline_reference": python_reference_baseline(args, args.mode),
        "e2e": {"value": value, "unit": "pictures/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# -------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------

class Rig(object):
    """Process-level plumbing shared by the decode and encode benches: device, torch.distributed, clocks."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = "cuda:%d" % self.local_rank
        self.numa_cpus = 0
        try:
            self.uuid = torch.cuda.get_device_properties(self.local_rank).uuid
        except Exception:
            self.uuid = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device(self.dev))
            if not os.environ.get("VC2B_NO_NUMA_BIND"):
                # one process per GPU: keep the pinned host buffers of this rank in the GPU's NUMA node
                from vc2_conformance_b200.sharding import bind_process_to_gpu

                self.numa_cpus = bind_process_to_gpu(self.local_rank, self.uuid)
        self.cur = torch.cuda.current_stream()
        self.sampler = ClockSampler(self.local_rank, self.uuid)
        self.windows = []

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_ms(self, ms):
        t = self.torch.tensor([ms], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def event(self):
        return self.torch.cuda.Event(enable_timing=True)

    def finish(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def timed(rig, steps, warmup, step_fn, nmarks):
    """W warm-up steps, then K steps bracketed by barrier + synchronize; ``step_fn(mark)`` calls ``mark()``
    after each of its ``nmarks`` stages.  Returns (total ms, [mean ms per stage])."""
    for _ in range(warmup):
        step_fn(lambda: None)
    rig.barrier()
    ev = [[rig.event() for _ in range(nmarks + 1)] for _ in range(steps)]
    e0, e1 = rig.event(), rig.event()
    t0 = time.time()
    e0.record(rig.cur)
    for k in range(steps):
        it = iter(ev[k])
        next(it).record(rig.cur)
        step_fn(lambda: next(it).record(rig.cur))
    e1.record(rig.cur)
    rig.barrier()
    rig.windows.append((t0, time.time()))
    stage_ms = [sum(ev[k][i].elapsed_time(ev[k][i + 1]) for k in range(steps)) / steps for i in range(nmarks)]
    return e0.elapsed_time(e1), stage_ms


def decode_stages_of(rig, name, npictures, steps=4, warmup=2, coeff16=True):
    """Device-timed unpack / IDWT stages of another BASELINE configuration (one resident batch), so that
    every configuration has a driver-run figure; the pictures of the first two (noise, ramp) are checked
    against the C oracle's decode when ``check`` streams are returned."""
    from vc2_conformance_b200.device import BatchCodec

    torch = rig.torch
    w = make_workload(name, npictures)
    p, NP, S = w["p"], w["pictures"], w["samples"]
    codec = BatchCodec(p, NP, max_stream_bytes=NP * (w["picture_bytes"] + 64), device=rig.dev, group=128,
                       coeff16=coeff16)
    distinct = min(NP, 4)
    pics = np.stack([__import__("synth").flat_picture(synth_picture(dict(w, pictures=distinct), rig.rank, i))
                     for i in range(distinct)])
    streams = codec.encode_pictures(pics, w["picture_bytes"])
    streams = [streams[i % distinct] for i in range(NP)]
    host, offs = codec.pack_host_streams(streams)
    codec._reserve_data(host.numel())
    codec.data[: host.numel()].copy_(host)
    codec.pic_offset[: NP + 1].copy_(offs)
    torch.cuda.synchronize()
    codec.decode_device(codec.data, codec.pic_offset, NP)
    assert (codec.status_host(NP)[:, 0] == 0).all(), "decode status not OK (%s)" % name

    def step(mark):
        codec.unpack_device(codec.data, codec.pic_offset, NP)
        mark()
        codec.idwt_device(NP)
        mark()

    _ms, (ms_u, ms_i) = timed(rig, steps, warmup, step, 2)
    peak, _src = measured_peak()
    padded = sum(codec.g.comp_coeffs)
    ub = (int(offs[NP]) + 4 * padded * NP) / (ms_u / 1000.0) / 1e9
    ib = ((4 * padded + 2 * S) * NP) / (ms_i / 1000.0) / 1e9
    out = {"workload": w["name"], "pictures": NP, "coefficient_plane": "int16" if codec.coeff16 else "int32",
           "unpack_us_per_picture": 1000.0 * ms_u / NP, "unpack_frac": ub / peak,
           "idwt_us_per_picture": 1000.0 * ms_i / NP, "idwt_frac": ib / peak,
           "pictures_per_s": NP / ((ms_u + ms_i) / 1000.0)}
    # parity of what was just timed: the first picture of each kind against the oracle
    O = _oracle()
    import synth
    po = synth.oracle_params(p)
    for i in range(min(2, distinct)):
        st, dec, _consumed = O.decode_picture(po, streams[i])
        got = codec.picture_planes(i)
        assert st == 0 and all((got[c] == dec[c]).all() for c in range(3)), "GPU picture differs from the oracle (%s)" % name
    out["checked"] = "pictures 0..%d bit-exact against the oracle" % (min(2, distinct) - 1)
    del codec
    torch.cuda.empty_cache()
    return out


def run_decode(args):
    import ctypes as C

    from vc2_conformance_b200 import capi
    from vc2_conformance_b200.device import PAD, BatchCodec, SequenceDecoder, _ptr

    rig = Rig(args)
    torch, dev, world, rank, cur = rig.torch, rig.dev, rig.world, rig.rank, rig.cur
    w = make_workload(args.workload, args.pictures)
    p = w["p"]
    NP = w["pictures"]
    S = w["samples"]

    # ---- synthetic sequence, encoded on the device by the encoder mirror ---------------------
    codec = BatchCodec(p, NP, max_stream_bytes=NP * (w["picture_bytes"] + 64), device=dev, group=args.group,
                       coeff16=not args.int32_plane)
    streams = []
    gen = 8
    for i0 in range(0, NP, gen):
        n = min(gen, NP - i0)
        streams += codec.encode_pictures(synth_pictures(w, rank, i0, n), w["picture_bytes"])
    host, offs = codec.pack_host_streams(streams)
    stream_bytes = int(offs[NP])

    # resident copy of the compressed sequence
    codec._reserve_data(host.numel())
    codec.data[: host.numel()].copy_(host)
    codec.pic_offset[: NP + 1].copy_(offs)
    torch.cuda.synchronize()

    # gate of the timed path: status words and consumed bytes of every picture (values are compared with the
    # oracle's decode of sample pictures in the cpu_baseline leg below)
    codec.decode_device(codec.data, codec.pic_offset, NP)
    st = codec.status_host(NP)
    # (a picture that did not fit the int16 coefficient plane would have to be decoded again through the int32
    # plane inside the timed region: the synthetic sequence has none, and the line says which plane ran)
    assert (st[:, 0] == 0).all(), "decode status not OK"
    assert (st[:, 2] == np.diff(offs.numpy())).all(), "consumed bytes mismatch"

    rig.sampler.start()
    L = codec.lib
    sp = codec._stream_ptr(None)

    def stage_unpack(first, n):
        codec.unpack_device(codec.data, codec.pic_offset[first:], n, None, first)

    def make_step(first, n):
        def step(mark):
            stage_unpack(first, n)          # stage 1: slice index + unpack (+ dc prediction for LD)
            mark()
            codec.idwt_device(n, None, first)  # stage 2: IDWT + crop + clip + offset
            mark()
        return step

    # ---- device-resident timed region --------------------------------------------------------
    launches0 = codec.launches
    ms_total, (ms_unpack, ms_idwt) = timed(rig, args.steps, args.warmup, make_step(0, NP), 2)
    launches = codec.launches - launches0
    ms_max = rig.max_ms(ms_total)
    value = world * NP * args.steps / (ms_max / 1000.0)

    # the two kinds of picture on their own (each a contiguous run of the sequence): what the content costs
    by_content = {}
    n_noise = (NP + 1) // 2
    for kind, first, n in (("noise", 0, n_noise), ("ramp", n_noise, NP - n_noise)):
        if n < 1:
            continue
        k_steps = max(3, args.steps // 4)
        ms_k, (ms_u, ms_i) = timed(rig, k_steps, 1, make_step(first, n), 2)
        by_content[kind] = {"pictures": n, "pictures_per_s": n * k_steps / (ms_k / 1000.0),
                            "unpack_us_per_picture": 1000.0 * ms_u / n, "idwt_us_per_picture": 1000.0 * ms_i / n}
    codec.decode_device(codec.data, codec.pic_offset, NP)  # leave every picture decoded for the checks below

    # ---- end-to-end timed region (host buffers) ----------------------------------------------
    e2e = None
    if not args.no_e2e:
        chunk = args.chunk
        max_chunk = max(int(offs[min(NP, i + chunk)] - offs[i]) for i in range(0, NP, chunk)) + PAD + 64
        seq = SequenceDecoder(p, chunk=chunk, nstreams=3, max_chunk_bytes=max_chunk, device=dev,
                              coeff16=not args.int32_plane)
        chunks = seq.plan(host, offs.numpy())
        out_host = torch.empty((NP, S), dtype=torch.uint16).pin_memory()

        def e2e_step(mark):
            seq.fork_from(cur)
            seq.decode(chunks, out_host)
            seq.join_to(cur)
            mark()

        l0 = seq.launches
        e2e_ms, _ = timed(rig, args.steps, max(1, args.warmup), e2e_step, 1)
        e2e_launches = (seq.launches - l0) * args.steps // (args.steps + max(1, args.warmup))
        # the e2e path must produce exactly what the resident path produced, for every picture
        resident = codec.pictures[: NP * S].view(NP, S).view(torch.int16)
        assert (seq.finish(chunks, out_host)[:, 0] == 0).all(), "e2e decode status not OK"
        same = torch.equal(out_host.view(torch.int16), resident.cpu())
        assert same, "e2e output differs from the resident path"
        e2e = {
            "value": world * NP * args.steps / (rig.max_ms(e2e_ms) / 1000.0),
            "unit": "pictures/s",
            "h2d_bytes_per_step": int(stream_bytes + 8 * (NP + len(chunks))),
            "d2h_bytes_per_step": int(2 * S * NP),
            "gpu_launches": e2e_launches,
            "pipeline": "%d-picture chunks over 3 CUDA streams" % chunk,
            "numa_bound_cpus": rig.numa_cpus,
            "checked": "all %d pictures equal to the resident path" % NP,
        }
    # ---- the other decoder configurations of BASELINE.json (device-timed stages, one resident batch each) ----
    by_config = None
    if rank == 0 and world == 1 and not args.no_config_sweep and args.workload == "cfg2":
        by_config = {}
        for name, npic in (("cfg1", 192), ("cfg3", 256), ("cfg4", 128)):
            by_config[name] = decode_stages_of(rig, name, npic, coeff16=not args.int32_plane)
    rig.sampler.stop()
    clocks = rig.sampler.summary(rig.windows)

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
                   "kernels": "hq_walk_kernel + unpack kernel" if not p.is_ld else "LD unpack kernel + dc_prediction_kernel"},
        "idwt": {"ms_per_step": ms_idwt, "achieved": idwt_gbs, "frac": idwt_gbs / peak,
                 "algorithmic_bytes_per_picture": idwt_bytes // NP,
                 "kernels": "%d transform levels (last fused with crop+clip+offset)" % (p.dwt_depth + p.dwt_depth_ho)},
    }
    dom = "unpack" if ms_unpack >= ms_idwt else "idwt"
    per_pic, traffic_src = ncu_traffic(args.workload, "int16" if codec.coeff16 else "int32", dom)
    traffic = per_pic * NP if per_pic else None  # per launch of the stage, like `achieved`
    for st_name in stages:
        pp, _src = ncu_traffic(args.workload, "int16" if codec.coeff16 else "int32", st_name)
        stages[st_name]["dram_traffic_per_picture"] = pp
    roofline = {"bound": "hbm", "stage": dom, "achieved": stages[dom]["achieved"], "peak": peak, "unit": "GB/s",
                "frac": stages[dom]["frac"], "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes": stages[dom]["algorithmic_bytes_per_picture"] * NP, "peak_source": peak_src}

    # ---- CPU baseline (rank 0, N=1 only): the oracle also checks the GPU's pictures ---------------
    cpu, cpu_ref = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        idx = sorted(set([0, 1, NP - 2, NP - 1]) & set(range(NP)))  # noise and ramp pictures
        gpu_pictures = {i: codec.picture_planes(i) for i in idx}
        checked = []

        def check(k, planes):
            for c in range(3):
                assert (planes[c] == gpu_pictures[idx[k]][c]).all(), "GPU picture %d differs from the oracle" % idx[k]
            checked.append(idx[k])

        rate, cores, npics, secs = cpu_rate(w, "decode", [streams[i] for i in idx], 12.0, check)
        cpu = {"value": rate, "unit": "pictures/s", "cores": cores, "kind": "port",
               "sample": "%d full-size pictures (%d distinct) decoded by the C oracle on %d threads in %.1f s"
                         % (npics, len(idx), cores, secs),
               "checked": "GPU pictures %s bit-exact against the oracle" % sorted(checked)}
        cpu_ref = python_reference_baseline(args, "decode")

    if rank == 0:
        line = {
            "metric": metric_name("decode"),
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
            "config": {"workload": w["name"], "mode": "decode", "pictures_per_step_per_gpu": NP,
                       "bits_per_sample": w["bits_per_sample"], "stream_bytes_per_picture": stream_bytes // NP,
                       "slice_size_scaler": int(p.slice_size_scaler),
                       "pictures": "first half white noise (picture_generators.py:600 recipe), second half integer ramps",
                       "l2": "inputs larger than L2 (%.0f MB of streams, %.0f MB of coefficients per step)"
                             % (stream_bytes / 1e6, 4 * padded * NP / 1e6),
                       "idwt_group": args.group,
                       "coefficient_plane": ("int16 high-pass bands + int32 level-0 band between unpack and IDWT "
                                             "(vc2b_unpack16 / vc2b_idwt16; every picture of the sequence fits)"
                                             if codec.coeff16 else "int32")},
            "e2e": e2e,
            "roofline": roofline,
            "stages": stages,
            "stages_by_content": by_content,
            "stages_by_config": by_config,
            "cpu_baseline": cpu,
            "cpu_baseline_reference": cpu_ref,
            "gpu_launches": launches,
            "clocks": clocks,
        }
        print(json.dumps(line))
    rig.finish()
    return 0


def run_encode(args):
    """BASELINE config 5 style: picture_encode (offset removal, padding, DWT) + quantize_to_fit + slice packing."""
    import ctypes as C

    from vc2_conformance_b200.device import BatchCodec, SequenceEncoder

    rig = Rig(args)
    torch, dev, world, rank, cur = rig.torch, rig.dev, rig.world, rig.rank, rig.cur
    w = make_workload(args.workload, args.pictures)
    p = w["p"]
    NP = w["pictures"]
    S = w["samples"]
    pb = w["picture_bytes"]

    codec = BatchCodec(p, NP, max_stream_bytes=1 << 16, device=dev, group=min(args.group, NP))
    pics = synth_pictures(w, rank, 0, NP)
    host_pics = torch.from_numpy(pics.view(np.int16)).view(torch.uint16).pin_memory()
    codec.upload_pictures(pics)
    torch.cuda.synchronize()
    nbytes = int(codec.lib.vc2b_packed_bytes(C.byref(p), pb))

    # gate: the streams decode without error and consume exactly their bytes
    first_streams = codec.encode_pictures(pics, pb)
    dec = BatchCodec(p, 1, max_stream_bytes=nbytes + 64, device=dev)
    for i in sorted(set([0, NP - 1])):
        dec.decode_streams([first_streams[i]])
        assert int(dec.status_host(1)[0][2]) == len(first_streams[i]), "decoder did not consume the stream"
    codec.upload_pictures(pics)
    out_holder = [None]

    def step(mark):
        codec.dwt_device(NP)
        mark()
        codec.quantize_device(NP, pb)
        mark()
        out_holder[0] = codec.pack_device(NP, pb)[0]
        mark()

    rig.sampler.start()
    launches0 = codec.launches
    ms_total, (ms_dwt, ms_fit, ms_pack) = timed(rig, args.steps, args.warmup, step, 3)
    launches = codec.launches - launches0
    ms_max = rig.max_ms(ms_total)
    value = world * NP * args.steps / (ms_max / 1000.0)
    resident_streams = out_holder[0][:, :nbytes].cpu().numpy()

    e2e = None
    if not args.no_e2e:
        chunk = max(1, min(args.chunk, NP // 3 if NP >= 3 else 1))
        seq = SequenceEncoder(p, pb, chunk=chunk, nstreams=3, device=dev)
        out_host = torch.empty((NP, seq.stride), dtype=torch.uint8).pin_memory()

        def e2e_step(mark):
            seq.fork_from(cur)
            seq.encode(host_pics, out_host)
            seq.join_to(cur)
            mark()

        l0 = seq.launches
        e2e_ms, _ = timed(rig, args.steps, max(1, args.warmup), e2e_step, 1)
        e2e_launches = (seq.launches - l0) * args.steps // (args.steps + max(1, args.warmup))
        assert (out_host.numpy()[:, :nbytes] == resident_streams).all(), "e2e streams differ from the resident path"
        e2e = {
            "value": world * NP * args.steps / (rig.max_ms(e2e_ms) / 1000.0),
            "unit": "pictures/s",
            "h2d_bytes_per_step": int(2 * S * NP),
            "d2h_bytes_per_step": int(seq.stride * NP),
            "gpu_launches": e2e_launches,
            "pipeline": "%d-picture chunks over 3 CUDA streams" % chunk,
            "numa_bound_cpus": rig.numa_cpus,
            "checked": "all %d streams equal to the resident path" % NP,
        }
    rig.sampler.stop()
    clocks = rig.sampler.summary(rig.windows)

    peak, peak_src = measured_peak()
    padded = sum(codec.g.comp_coeffs)
    dwt_bytes = (2 * S + 4 * padded) * NP              # uint16 pixels in, int32 coefficients out (6 B/sample)
    fit_bytes = 4 * padded * NP                        # every coefficient read once (the probes run from shared memory)
    pack_bytes = (4 * padded + nbytes) * NP            # coefficients in, packed bytes out
    stages = {}
    for name, ms, nb_, kern in (("dwt", ms_dwt, dwt_bytes, "DWT levels (first fused with offset removal + padding)"),
                                ("quantize_to_fit", ms_fit, fit_bytes, "quantize_fit_kernel"),
                                ("pack", ms_pack, pack_bytes, "pack_slices_kernel")):
        gbs = nb_ / (ms / 1000.0) / 1e9
        stages[name] = {"ms_per_step": ms, "achieved": gbs, "frac": gbs / peak,
                        "algorithmic_bytes_per_picture": nb_ // NP, "kernels": kern}
    dom = max(stages, key=lambda k: stages[k]["ms_per_step"])
    roofline = {"bound": "hbm", "stage": dom, "achieved": stages[dom]["achieved"], "peak": peak, "unit": "GB/s",
                "frac": stages[dom]["frac"], "traffic": None, "traffic_source": None,
                "algorithmic_bytes": stages[dom]["algorithmic_bytes_per_picture"] * NP, "peak_source": peak_src}

    cpu, cpu_ref = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        idx = sorted(set([0, NP - 1]))
        checked = []

        def check(k, data):
            assert data == resident_streams[idx[k]].tobytes(), "GPU stream %d differs from the oracle's" % idx[k]
            checked.append(idx[k])

        items = [synth_picture(w, rank, i) for i in idx]
        rate, cores, npics, secs = cpu_rate(w, "encode", items, 12.0, check)
        cpu = {"value": rate, "unit": "pictures/s", "cores": cores, "kind": "port",
               "sample": "%d full-size pictures (%d distinct) encoded by the C oracle on %d threads in %.1f s"
                         % (npics, len(idx), cores, secs),
               "checked": "GPU streams %s byte-exact against the oracle" % sorted(checked)}
        cpu_ref = python_reference_baseline(args, "encode")

    if rank == 0:
        line = {
            "metric": metric_name("encode"),
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
            "config": {"workload": w["name"], "mode": "encode", "pictures_per_step_per_gpu": NP,
                       "bits_per_sample": w["bits_per_sample"], "stream_bytes_per_picture": nbytes,
                       "slice_size_scaler": int(p.slice_size_scaler),
                       "pictures": "first half white noise (picture_generators.py:600 recipe), second half integer ramps",
                       "l2": "inputs larger than L2 (%.0f MB of pictures, %.0f MB of coefficients per step)"
                             % (2 * S * NP / 1e6, 4 * padded * NP / 1e6),
                       "sharding": "sequence sharded by picture: every rank encodes its own pictures, no collective"},
            "e2e": e2e,
            "roofline": roofline,
            "stages": stages,
            "cpu_baseline": cpu,
            "cpu_baseline_reference": cpu_ref,
            "gpu_launches": launches,
            "clocks": clocks,
        }
        print(json.dumps(line))
    rig.finish()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_encode(args) if args.mode == "encode" else run_decode(args)


if __name__ == "__main__":
    sys.exit(main())
