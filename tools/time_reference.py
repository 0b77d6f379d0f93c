#!/usr/bin/env python
"""
time_reference.py -- the UNMODIFIED Python reference timed on the host cores (BASELINE.md section 3,
SURVEY.md 8(d) "CPU baseline beside it").

The per-picture entry points the reference's decoder calls (decoder/stream.py:128-133) --
``decoder.transform_data_syntax.transform_data`` and ``pseudocode.picture_decoding.picture_decode`` -- or,
with ``--mode encode``, ``encoder.pictures.transform_and_slice_picture`` + ``make_transform_data_hq_lossy``
(encoder/pictures.py:350,617), run on a CROPPED picture of the named workload (same coding parameters and
slice size, fewer slices), one picture per process on every host core at once, and scaled linearly in the
number of samples to the full picture size.  Prints one JSON object.

    python tools/time_reference.py --workload cfg2 [--crop 512x256] [--procs N] [--mode decode|encode]

The reference tree is found by oracle/ref_import.py (/root/reference in the build container, the git-ignored
baseline/_ref copy on the GPU box); the cropped stream is produced by the C oracle's encoder.  This script is
test / measurement infrastructure: nothing under vc2_conformance_b200/ uses it.
"""

import argparse
import json
import multiprocessing as mp
import os
import sys
import time
from io import BytesIO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def cropped_workload(name, crop_w, crop_h):
    """The workload's parameters on a crop_w x crop_h luma picture with the same slice dimensions."""
    import bench
    from vc2_conformance_b200.params import copy_params

    w = bench.make_workload(name)
    p = w["p"]
    sw, sh = p.luma_width // p.slices_x, p.luma_height // p.slices_y  # slice size in luma samples
    crop_w = max(sw, crop_w // sw * sw)
    crop_h = max(sh, crop_h // sh * sh)
    cdw, cdh = p.luma_width // p.color_diff_width, p.luma_height // p.color_diff_height
    sx, sy = crop_w // sw, crop_h // sh
    S = crop_w * crop_h + 2 * (crop_w // cdw) * (crop_h // cdh)
    picture_bytes = int(S * w["bits_per_sample"] / 8) // 4 * 4
    kw = dict(luma_width=crop_w, luma_height=crop_h, color_diff_width=crop_w // cdw, color_diff_height=crop_h // cdh,
              slices_x=sx, slices_y=sy)
    if p.is_ld:
        kw.update(slice_bytes_numerator=picture_bytes, slice_bytes_denominator=sx * sy)
    q = copy_params(p, **kw)
    return dict(p=q, samples=S, full_samples=w["samples"], picture_bytes=picture_bytes, name=w["name"],
                crop="%dx%d" % (crop_w, crop_h))


def _reference_state(p):
    from vc2_conformance.pseudocode.state import State

    qm = {}
    top = p.dwt_depth + p.dwt_depth_ho
    for level in range(top + 1):
        if level == 0:
            qm[0] = {("LL" if p.dwt_depth_ho == 0 else "L"): p.quant_matrix[0][0]}
        elif level <= p.dwt_depth_ho:
            qm[level] = {"H": p.quant_matrix[level][1]}
        else:
            qm[level] = {"HL": p.quant_matrix[level][1], "LH": p.quant_matrix[level][2], "HH": p.quant_matrix[level][3]}
    st = State(parse_code=0xC8 if p.is_ld else 0xE8, luma_width=p.luma_width, luma_height=p.luma_height,
               color_diff_width=p.color_diff_width, color_diff_height=p.color_diff_height, luma_depth=p.luma_depth,
               color_diff_depth=p.color_diff_depth, wavelet_index=p.wavelet_index, wavelet_index_ho=p.wavelet_index_ho,
               dwt_depth=p.dwt_depth, dwt_depth_ho=p.dwt_depth_ho, slices_x=p.slices_x, slices_y=p.slices_y,
               quant_matrix=qm, picture_number=0)
    if p.is_ld:
        st["slice_bytes_numerator"] = p.slice_bytes_numerator
        st["slice_bytes_denominator"] = p.slice_bytes_denominator
    else:
        st["slice_prefix_bytes"] = p.slice_prefix_bytes
        st["slice_size_scaler"] = p.slice_size_scaler
    return st, qm


def _worker(args):
    """One process: the reference's own functions on one cropped picture.  Returns (seconds, checksum)."""
    fields, qm_rows, data, planes, mode, picture_bytes = args
    import ref_import

    ref_import.install()
    from vc2_conformance_b200 import capi  # only the plain ctypes struct; no device code runs here

    p = capi.Params()
    for k, v in fields.items():
        setattr(p, k, v)
    for lv, row in enumerate(qm_rows):
        for o, v in enumerate(row):
            p.quant_matrix[lv][o] = v
    st, qm = _reference_state(p)
    if mode == "decode":
        from vc2_conformance.decoder import io as dio
        from vc2_conformance.decoder import transform_data_syntax as tds
        from vc2_conformance.pseudocode import picture_decoding

        out = []
        st["_output_picture_callback"] = lambda pic, vp, pcm: out.append(pic)
        st["video_parameters"] = {}
        st["picture_coding_mode"] = 0
        t0 = time.perf_counter()
        dio.init_io(st, BytesIO(data))
        tds.transform_data(st)
        picture_decoding.picture_decode(st)
        dt = time.perf_counter() - t0
        return dt, [[int(v) for row in out[0][c] for v in row] for c in ("Y", "C1", "C2")]
    # encode: the reference's picture_encode + quantize_to_fit for every slice
    from vc2_conformance.encoder import pictures as encp
    from vc2_conformance.pseudocode.video_parameters import VideoParameters
    from vc2_data_tables import Profiles

    cdf = {(1, 1): 0, (2, 1): 1, (2, 2): 2}[(p.luma_width // p.color_diff_width, p.luma_height // p.color_diff_height)]
    cf = {"wavelet_index": p.wavelet_index, "wavelet_index_ho": p.wavelet_index_ho, "dwt_depth": p.dwt_depth,
          "dwt_depth_ho": p.dwt_depth_ho, "slices_x": p.slices_x, "slices_y": p.slices_y, "picture_coding_mode": 0,
          "profile": Profiles.low_delay if p.is_ld else Profiles.high_quality, "quantization_matrix": qm,
          "video_parameters": VideoParameters(frame_width=p.luma_width, frame_height=p.luma_height,
                                              color_diff_format_index=cdf, luma_excursion=(1 << p.luma_depth) - 1,
                                              color_diff_excursion=(1 << p.color_diff_depth) - 1)}
    picture = {"Y": planes[0], "C1": planes[1], "C2": planes[2], "pic_num": 0}
    t0 = time.perf_counter()
    coeffs = encp.transform_and_slice_picture(cf, picture)
    if p.is_ld:
        td = encp.make_transform_data_ld_lossy(picture_bytes, coeffs)
        q = [s["qindex"] for s in td["ld_slices"]]
    else:
        _scaler, td = encp.make_transform_data_hq_lossy(picture_bytes, coeffs, 0, p.slice_size_scaler)
        q = [s["qindex"] for s in td["hq_slices"]]
    dt = time.perf_counter() - t0
    return dt, q


def measure(workload="cfg2", crop=(512, 256), procs=None, mode="decode"):
    """Returns the baseline dict (pictures/s of the full-size workload on `procs` cores)."""
    import ref_import

    if not ref_import.available():
        return {"unavailable": "no reference tree (neither /root/reference nor baseline/_ref)"}
    import numpy as np

    import oracle as O
    import synth
    from vc2_conformance_b200.params import FIELDS

    w = cropped_workload(workload, *crop)
    p = w["p"]
    po = synth.oracle_params(p)
    procs = procs or os.cpu_count() or 1
    jobs, expect = [], []
    fields = {k: int(getattr(p, k)) for k in FIELDS}
    qm_rows = [[int(p.quant_matrix[lv][o]) for o in range(4)] for lv in range(16)]
    for i in range(procs):
        planes = synth.noise_picture(p, i) if i % 2 == 0 else synth.ramp_picture(p, i)
        coeffs = O.encode_picture(po, planes)
        if p.is_ld:
            data, q = O.encode_ld_lossy(po, coeffs)
        else:
            data, q = O.encode_hq_lossy(po, coeffs, w["picture_bytes"])
        if mode == "decode":
            st, dec, _consumed = O.decode_picture(po, data)
            assert st == 0
            expect.append([a.astype(np.int64).reshape(-1).tolist() for a in dec])
        else:
            expect.append([int(v) for v in q])
        jobs.append((fields, qm_rows, data, [a.astype(np.int64).tolist() for a in planes], mode, w["picture_bytes"]))
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(procs) as pool:
        results = pool.map(_worker, jobs)
    wall = time.perf_counter() - t0
    # the oracle and the reference agree on what was computed (the baseline times the right thing)
    for (dt, got), want in zip(results, expect):
        assert got == want, "reference and oracle disagree on the cropped picture"
    per_picture = sum(dt for dt, _ in results) / len(results)   # seconds per cropped picture per core
    scale = w["full_samples"] / float(w["samples"])
    rate = procs / (per_picture * scale)
    return {
        "value": rate, "unit": "pictures/s", "cores": procs, "kind": "reference",
        "per_core": 1.0 / (per_picture * scale),
        "sample": "unmodified Python reference (%s): %d processes x 1 picture cropped to %s (%d samples, %.2f s each), "
                  "scaled x%.1f in samples to the full picture; wall %.1f s incl. interpreter start"
                  % ("transform_data + picture_decode" if mode == "decode" else
                     "transform_and_slice_picture + make_transform_data_*_lossy", procs, w["crop"], w["samples"],
                     per_picture, scale, wall),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--crop", default="512x256")
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--mode", default="decode", choices=["decode", "encode"])
    a = ap.parse_args()
    cw, ch = [int(v) for v in a.crop.lower().split("x")]
    print(json.dumps(measure(a.workload, (cw, ch), a.procs or None, a.mode)))


if __name__ == "__main__":
    main()
