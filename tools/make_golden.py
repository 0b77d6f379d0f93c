#!/usr/bin/env python
"""
make_golden.py -- generate tests/golden/*.npz from the UNMODIFIED Python
reference (/root/reference, imported through oracle/ref_import.py) and check
the C oracle (oracle/vc2_oracle.c) against every vector as it is produced.

Run inside the build container only (the reference does not travel to the GPU
box).  The vectors are small and committed; tests/test_oracle_golden.py re-checks
the oracle against them on every run, and tests/test_gpu_*.py check the CUDA
path against the same files.

    python tools/make_golden.py            # regenerate everything
"""

import os
import sys
from io import BytesIO

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_import  # noqa: E402

ref_import.install()

import oracle as O  # noqa: E402

from vc2_conformance import bitstream, decoder  # noqa: E402
from vc2_conformance.bitstream import exp_golomb  # noqa: E402
from vc2_conformance.pseudocode import (  # noqa: E402
    picture_decoding,
    picture_encoding,
    quantization,
    slice_sizes,
    vc2_math,
)
from vc2_conformance.pseudocode.state import State  # noqa: E402
from vc2_conformance.decoder import transform_data_syntax as tds  # noqa: E402
from vc2_conformance.decoder import io as dio  # noqa: E402
from vc2_conformance.decoder.exceptions import (  # noqa: E402
    InvalidSliceYLength,
    UnexpectedEndOfStream,
)
from vc2_conformance.encoder import pictures as encp  # noqa: E402
from vc2_data_tables import Profiles, WaveletFilters  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

ORIENTS_2D = ["HL", "LH", "HH"]


# ---------------------------------------------------------------------------
# conversions between the reference's nested containers and flat arrays
# ---------------------------------------------------------------------------


def make_quant_matrix(dwt_depth, dwt_depth_ho, rng):
    qm = {}
    if dwt_depth_ho == 0:
        qm[0] = {"LL": int(rng.randint(0, 4))}
    else:
        qm[0] = {"L": int(rng.randint(0, 4))}
        for level in range(1, dwt_depth_ho + 1):
            qm[level] = {"H": int(rng.randint(0, 6))}
    for level in range(dwt_depth_ho + 1, dwt_depth_ho + dwt_depth + 1):
        qm[level] = {o: int(rng.randint(0, 8)) for o in ORIENTS_2D}
    return qm


def qm_to_array(qm):
    a = np.zeros((O.MAX_LEVELS, 4), np.int32)
    for level, orients in qm.items():
        for orient, value in orients.items():
            a[level][O.ORIENT_SLOT[orient]] = value
    return a


def qm_from_array(a, dwt_depth, dwt_depth_ho):
    qm = {}
    if dwt_depth_ho == 0:
        qm[0] = {"LL": int(a[0][0])}
    else:
        qm[0] = {"L": int(a[0][0])}
        for level in range(1, dwt_depth_ho + 1):
            qm[level] = {"H": int(a[level][1])}
    for level in range(dwt_depth_ho + 1, dwt_depth_ho + dwt_depth + 1):
        qm[level] = {o: int(a[level][O.ORIENT_SLOT[o]]) for o in ORIENTS_2D}
    return qm


def coeff_dict_to_flat(p, comp, d):
    level, orient, w, h, off, total = O.layout(p, comp)
    names = {0: ["LL", "L"], 1: ["HL", "H"], 2: ["LH"], 3: ["HH"]}
    out = np.zeros(total, np.int64)
    for b in range(len(level)):
        band = None
        for nm in names[int(orient[b])]:
            if nm in d[int(level[b])]:
                band = d[int(level[b])][nm]
        arr = np.array(band, dtype=object).astype(np.int64)
        assert arr.shape == (h[b], w[b]), (arr.shape, h[b], w[b])
        out[off[b] : off[b] + w[b] * h[b]] = arr.reshape(-1)
    return out


def flat_to_coeff_dict(p, comp, flat):
    level, orient, w, h, off, total = O.layout(p, comp)
    d = {}
    for b in range(len(level)):
        lv = int(level[b])
        if lv == 0:
            nm = "LL" if p.dwt_depth_ho == 0 else "L"
        elif lv <= p.dwt_depth_ho:
            nm = "H"
        else:
            nm = ORIENTS_2D[int(orient[b]) - 1]
        band = flat[off[b] : off[b] + w[b] * h[b]].reshape(h[b], w[b])
        d.setdefault(lv, {})[nm] = [[int(v) for v in row] for row in band]
    return d


PARAM_FIELDS = [
    "luma_width", "luma_height", "color_diff_width", "color_diff_height", "luma_depth",
    "color_diff_depth", "wavelet_index", "wavelet_index_ho", "dwt_depth", "dwt_depth_ho",
    "slices_x", "slices_y", "is_ld", "slice_prefix_bytes", "slice_size_scaler",
    "slice_bytes_numerator", "slice_bytes_denominator",
]


def params_to_array(p):
    return np.array([getattr(p, f) for f in PARAM_FIELDS], np.int64)


def state_from_params(p, qm):
    st = State(
        parse_code=0xC8 if p.is_ld else 0xE8,
        luma_width=p.luma_width,
        luma_height=p.luma_height,
        color_diff_width=p.color_diff_width,
        color_diff_height=p.color_diff_height,
        luma_depth=p.luma_depth,
        color_diff_depth=p.color_diff_depth,
        wavelet_index=p.wavelet_index,
        wavelet_index_ho=p.wavelet_index_ho,
        dwt_depth=p.dwt_depth,
        dwt_depth_ho=p.dwt_depth_ho,
        slices_x=p.slices_x,
        slices_y=p.slices_y,
        quant_matrix=qm,
        picture_number=7,
    )
    if p.is_ld:
        st["slice_bytes_numerator"] = p.slice_bytes_numerator
        st["slice_bytes_denominator"] = p.slice_bytes_denominator
    else:
        st["slice_prefix_bytes"] = p.slice_prefix_bytes
        st["slice_size_scaler"] = p.slice_size_scaler
    return st


# ---------------------------------------------------------------------------
# 1. maths / quantiser / geometry
# ---------------------------------------------------------------------------


def gen_quant():
    rng = np.random.RandomState(1)
    idx = np.arange(0, 120)
    qf = np.array([quantization.quant_factor(int(i)) for i in idx], np.int64)
    qo = np.array([quantization.quant_offset(int(i)) for i in idx], np.int64)
    vals = np.concatenate([np.arange(-40, 41), rng.randint(-(1 << 20), 1 << 20, 400)]).astype(np.int64)
    qis = rng.randint(0, 64, vals.size).astype(np.int32)
    qis[:81] = np.arange(81) % 16
    iq = np.array([quantization.inverse_quant(int(v), int(q)) for v, q in zip(vals, qis)], np.int64)
    fq = np.array([quantization.forward_quant(int(v), int(q)) for v, q in zip(vals, qis)], np.int64)
    ns = np.arange(1, 300).astype(np.int64)
    il2 = np.array([vc2_math.intlog2(int(n)) for n in ns], np.int32)
    sv = np.concatenate([np.arange(-70, 71), rng.randint(-(1 << 30), 1 << 30, 100)]).astype(np.int64)
    slen = np.array([exp_golomb.signed_exp_golomb_length(int(v)) for v in sv], np.int32)
    # oracle check
    L = O.lib()
    assert all(L.vc2o_quant_factor(int(i)) == qf[k] for k, i in enumerate(idx))
    assert all(L.vc2o_quant_offset(int(i)) == qo[k] for k, i in enumerate(idx))
    assert all(L.vc2o_inverse_quant(int(v), int(q)) == r for v, q, r in zip(vals, qis, iq))
    assert all(L.vc2o_forward_quant(int(v), int(q)) == r for v, q, r in zip(vals, qis, fq))
    assert all(L.vc2o_intlog2(int(n)) == r for n, r in zip(ns, il2))
    assert all(L.vc2o_signed_exp_golomb_length(int(v)) == r for v, r in zip(sv, slen))
    np.savez_compressed(
        os.path.join(GOLDEN, "quant.npz"),
        idx=idx, quant_factor=qf, quant_offset=qo, vals=vals, qis=qis, inverse_quant=iq,
        forward_quant=fq, intlog2_n=ns, intlog2=il2, sint_vals=sv, sint_len=slen,
    )
    print("quant.npz ok")


def gen_geometry():
    rng = np.random.RandomState(2)
    cases = []
    for _ in range(40):
        d = int(rng.randint(0, 5))
        dho = int(rng.randint(0, 4))
        lw = int(rng.randint(1, 400))
        lh = int(rng.randint(1, 300))
        cw = lw // int(rng.choice([1, 2]))
        ch = lh // int(rng.choice([1, 2]))
        cw, ch = max(cw, 1), max(ch, 1)
        sx = int(rng.randint(1, 12))
        sy = int(rng.randint(1, 12))
        num = int(rng.randint(1, 5000))
        den = int(rng.randint(1, 200))
        p = O.make_params(lw, lh, cw, ch, dwt_depth=d, dwt_depth_ho=dho, slices_x=sx, slices_y=sy,
                          slice_bytes_numerator=num, slice_bytes_denominator=den)
        st = State(luma_width=lw, luma_height=lh, color_diff_width=cw, color_diff_height=ch,
                   dwt_depth=d, dwt_depth_ho=dho, slices_x=sx, slices_y=sy,
                   slice_bytes_numerator=num, slice_bytes_denominator=den)
        row = list(params_to_array(p))
        for comp, cn in ((0, "Y"), (1, "C1")):
            for level in range(0, d + dho + 1):
                w = slice_sizes.subband_width(st, level, cn)
                h = slice_sizes.subband_height(st, level, cn)
                assert O.lib().vc2o_subband_width(p, level, comp) == w
                assert O.lib().vc2o_subband_height(p, level, comp) == h
                row += [w, h]
            row += [-1] * (2 * (8 - (d + dho + 1)))
        sb = [slice_sizes.slice_bytes(st, x, y) for y in range(sy) for x in range(sx)]
        assert sb == [O.lib().vc2o_slice_bytes(p, x, y) for y in range(sy) for x in range(sx)]
        row += [sum(sb), sb[0], sb[-1]]
        cases.append(row)
    np.savez_compressed(os.path.join(GOLDEN, "geometry.npz"), cases=np.array(cases, np.int64))
    print("geometry.npz ok")


# ---------------------------------------------------------------------------
# 2. exp-Golomb bounded blocks
# ---------------------------------------------------------------------------


def ref_read_block(data, bit_offset, bits_left, n):
    """Read n values with the reference's decoder/io.py read_sintb from a bounded block."""
    st = State()
    dio.init_io(st, BytesIO(bytes(data)))
    for _ in range(bit_offset):
        dio.read_bit(st)
    st["bits_left"] = bits_left
    out = []
    try:
        for _ in range(n):
            out.append(dio.read_sintb(st))
        status = 0
    except UnexpectedEndOfStream:
        status = 1
    return status, out


def gen_expgolomb():
    rng = np.random.RandomState(3)
    blobs, meta, outs = [], [], []
    for case in range(60):
        n = int(rng.randint(1, 80))
        scale = int(rng.choice([1, 2, 4, 30, 1000, 100000]))
        vals = (rng.standard_normal(n) * scale).astype(np.int64)
        if case % 7 == 0:
            vals[:] = 0
        f = BytesIO()
        w = bitstream.BitstreamWriter(f)
        pre = int(rng.randint(0, 8))
        w.write_nbits(pre, 0)
        for v in vals:
            w.write_sint(int(v))
        total_bits = w.tell()[0] * 8 + (7 - w.tell()[1]) - pre
        w.flush()
        data = f.getvalue() + bytes(rng.randint(0, 256, 3).astype(np.uint8).tolist())
        # choose a bounded block length: exact, truncated (dangling), or longer
        mode = case % 3
        if mode == 0:
            bits_left = total_bits
        elif mode == 1:
            bits_left = int(rng.randint(0, total_bits + 1))
        else:
            bits_left = min(total_bits + int(rng.randint(0, 16)), len(data) * 8 - pre)
        nread = n + int(rng.randint(0, 5))
        status, out = ref_read_block(data, pre, bits_left, nread)
        assert status == 0
        arr = np.frombuffer(data, np.uint8).copy()
        o = np.zeros(nread, np.int64)
        st = O.lib().vc2o_read_sintb_block(arr, arr.size, pre, bits_left, o, nread, None)
        assert st == 0 and list(o) == out, (case, list(o), out)
        blobs.append(arr)
        meta.append([pre, bits_left, nread, arr.size])
        outs.append(np.array(out, np.int64))
    np.savez_compressed(
        os.path.join(GOLDEN, "expgolomb.npz"),
        data=np.concatenate(blobs), meta=np.array(meta, np.int64), out=np.concatenate(outs),
    )
    print("expgolomb.npz ok")


# ---------------------------------------------------------------------------
# 3. lifting / idwt / dwt
# ---------------------------------------------------------------------------


def gen_lifting():
    rng = np.random.RandomState(4)
    rec = {}
    k = 0
    for filt in range(7):
        for n in (2, 4, 6, 10, 16, 34, 64):
            a = rng.randint(-2000, 2000, n).astype(np.int64)
            s = [int(v) for v in a]
            picture_decoding.oned_synthesis(s, WaveletFilters(filt))
            an = [int(v) for v in a]
            picture_encoding.oned_analysis(an, WaveletFilters(filt))
            os_ = a.copy()
            O.lib().vc2o_oned_synthesis(os_, 1, n, filt)
            oa = a.copy()
            O.lib().vc2o_oned_analysis(oa, 1, n, filt)
            assert list(os_) == s and list(oa) == an, (filt, n)
            rec["in_%d" % k] = a
            rec["synth_%d" % k] = np.array(s, np.int64)
            rec["anal_%d" % k] = np.array(an, np.int64)
            rec["filt_%d" % k] = np.array([filt], np.int32)
            k += 1
    rec["count"] = np.array([k], np.int32)
    np.savez_compressed(os.path.join(GOLDEN, "lifting.npz"), **rec)
    print("lifting.npz ok (%d cases)" % k)


TRANSFORM_CASES = [
    # (wavelet_index, wavelet_index_ho, dwt_depth, dwt_depth_ho, dc_w, dc_h)
    (1, 1, 3, 0, 6, 5),
    (0, 0, 2, 0, 7, 6),
    (2, 2, 2, 0, 8, 5),
    (3, 3, 2, 0, 5, 4),
    (4, 4, 2, 0, 5, 4),
    (5, 5, 1, 0, 12, 10),
    (6, 6, 2, 0, 7, 6),
    (5, 6, 2, 3, 3, 5),   # BASELINE config 4: Fidelity vertical, Daubechies horizontal
    (3, 1, 2, 1, 4, 4),   # reference test_lossless asymmetric case
    (1, 0, 0, 2, 9, 6),   # horizontal only
    (0, 4, 1, 1, 8, 8),
    (6, 2, 1, 2, 5, 7),
    (1, 1, 0, 0, 5, 4),   # no transform at all
]


def gen_transform():
    rng = np.random.RandomState(5)
    rec = {}
    for k, (wi, wiho, d, dho, dcw, dch) in enumerate(TRANSFORM_CASES):
        W = dcw << (d + dho)
        H = dch << d
        p = O.make_params(W, H, wavelet_index=wi, wavelet_index_ho=wiho, dwt_depth=d, dwt_depth_ho=dho)
        total = O.layout(p, 0)[5]
        assert total == W * H
        flat = rng.randint(-600, 600, total).astype(np.int64)
        st = State(wavelet_index=wi, wavelet_index_ho=wiho, dwt_depth=d, dwt_depth_ho=dho)
        ref_pic = picture_decoding.idwt(st, flat_to_coeff_dict(p, 0, flat))
        ref_pic = np.array(ref_pic, dtype=object).astype(np.int64)
        assert (O.idwt(p, 0, flat) == ref_pic).all(), ("idwt", k)
        pic = rng.randint(-512, 512, (H, W)).astype(np.int64)
        ref_co = picture_encoding.dwt(st, [[int(v) for v in row] for row in pic])
        ref_flat = coeff_dict_to_flat(p, 0, ref_co)
        assert (O.dwt(p, 0, pic) == ref_flat).all(), ("dwt", k)
        rec["case_%d" % k] = np.array([wi, wiho, d, dho, W, H], np.int32)
        rec["coeffs_%d" % k] = flat.astype(np.int32)
        rec["idwt_%d" % k] = ref_pic.astype(np.int32)
        rec["pic_%d" % k] = pic.astype(np.int32)
        rec["dwt_%d" % k] = ref_flat.astype(np.int32)
    rec["count"] = np.array([len(TRANSFORM_CASES)], np.int32)
    np.savez_compressed(os.path.join(GOLDEN, "transform.npz"), **rec)
    print("transform.npz ok (%d cases)" % len(TRANSFORM_CASES))


def gen_dc_prediction():
    rng = np.random.RandomState(6)
    rec = {}
    shapes = [(1, 1), (1, 7), (6, 1), (5, 9), (16, 12), (33, 40)]
    for k, (h, w) in enumerate(shapes):
        band = rng.randint(-300, 300, (h, w)).astype(np.int64)
        b = [[int(v) for v in row] for row in band]
        tds.dc_prediction(b)
        a = [[int(v) for v in row] for row in band]
        encp.apply_dc_prediction(a)
        assert (O.dc_prediction(band) == np.array(b)).all()
        assert (O.apply_dc_prediction(band) == np.array(a)).all()
        rec["in_%d" % k] = band.astype(np.int32)
        rec["pred_%d" % k] = np.array(b, np.int32)
        rec["apply_%d" % k] = np.array(a, np.int32)
    rec["count"] = np.array([len(shapes)], np.int32)
    np.savez_compressed(os.path.join(GOLDEN, "dc_prediction.npz"), **rec)
    print("dc_prediction.npz ok")


# ---------------------------------------------------------------------------
# 4. whole pictures through the reference encoder and decoder
# ---------------------------------------------------------------------------

PICTURE_CASES = [
    # name, W, H, chroma (wdiv, hdiv), depth, wi, wiho, d, dho, sx, sy, profile, mode, arg
    ("hq_lossless_legall_uneven", 64, 32, (2, 1), 8, 1, 1, 2, 0, 5, 7, "hq", "lossless", None),
    ("hq_lossless_asym_haar_legall", 64, 32, (2, 1), 8, 3, 1, 2, 1, 5, 7, "hq", "lossless", None),
    ("hq_lossy_legall_d3", 96, 48, (2, 1), 10, 1, 1, 3, 0, 6, 3, "hq", "lossy", 1800),
    ("hq_lossy_dd97_d4", 128, 64, (2, 1), 10, 0, 0, 4, 0, 4, 4, "hq", "lossy", 3000),
    ("hq_lossy_dd97_d2_prefix_scaler", 64, 48, (2, 2), 10, 0, 0, 2, 0, 4, 3, "hq", "lossy_prefix", 2500),
    ("hq_lossy_fidelity_daub_asym", 128, 32, (1, 1), 12, 5, 6, 2, 3, 2, 4, "hq", "lossy", 6000),
    ("hq_lossy_dd137_444", 48, 40, (1, 1), 12, 2, 2, 2, 0, 3, 5, "hq", "lossy", 4000),
    ("hq_lossy_daub_16bit", 40, 24, (2, 1), 16, 6, 6, 2, 0, 5, 3, "hq", "lossy", 3000),
    ("hq_padded_dims", 50, 30, (2, 1), 10, 1, 1, 3, 0, 3, 2, "hq", "lossy", 1500),
    ("ld_lossy_haar_shift_d2", 64, 32, (2, 1), 10, 4, 4, 2, 0, 4, 4, "ld", "lossy", 1400),
    ("ld_lossy_legall_uneven", 80, 36, (2, 1), 8, 1, 1, 2, 0, 5, 7, "ld", "lossy", 1100),
    ("ld_lossy_haar_noshift_asym", 64, 32, (1, 1), 10, 3, 3, 1, 2, 4, 4, "ld", "lossy", 2100),
    ("ld_tiny_slices", 32, 16, (2, 1), 8, 4, 4, 2, 0, 8, 4, "ld", "lossy", 160),
]


def make_picture(rng, W, H, cdiv, depth, kind):
    cw, ch = W // cdiv[0], H // cdiv[1]
    if kind == "noise":
        return [rng.randint(0, 1 << depth, (H, W)).astype(np.int64),
                rng.randint(0, 1 << depth, (ch, cw)).astype(np.int64),
                rng.randint(0, 1 << depth, (ch, cw)).astype(np.int64)]
    out = []
    for (h, w) in ((H, W), (ch, cw), (ch, cw)):
        x = (np.arange(w) * ((1 << depth) - 1)) // max(w - 1, 1)
        y = (np.arange(h) * 31) // max(h - 1, 1)
        a = np.clip(x[None, :] + y[:, None] + rng.randint(-3, 4, (h, w)), 0, (1 << depth) - 1)
        out.append(a.astype(np.int64))
    return out


def serialise_transform_data(td, state):
    f = BytesIO()
    w = bitstream.BitstreamWriter(f)
    with bitstream.Serialiser(w, td, bitstream.vc2_default_values) as ser:
        bitstream.transform_data(ser, state)
    w.flush()
    return f.getvalue()


def ref_decode(p, qm, data):
    """Run the reference's transform_data + picture_decode; returns dict of results."""
    st = state_from_params(p, qm)
    dio.init_io(st, BytesIO(bytes(data)))
    res = {}
    try:
        tds.transform_data(st)
    except UnexpectedEndOfStream:
        res["status"] = O.UNEXPECTED_END_OF_STREAM
        return res
    except InvalidSliceYLength:
        res["status"] = O.INVALID_SLICE_Y_LENGTH
        return res
    res["status"] = 0
    res["consumed"] = dio.tell(st)[0] if dio.tell(st)[1] == 7 else dio.tell(st)[0] + 1
    res["coeffs"] = [coeff_dict_to_flat(p, c, st[k]) for c, k in
                     enumerate(["y_transform", "c1_transform", "c2_transform"])]
    out = []
    st["_output_picture_callback"] = lambda pic, vp, pcm: out.append(pic)
    st["video_parameters"] = {}
    st["picture_coding_mode"] = 0
    picture_decoding.picture_decode(st)
    pic = out[0]
    assert pic["pic_num"] == 7
    res["picture"] = [np.array(pic[c], dtype=object).astype(np.int64) for c in ("Y", "C1", "C2")]
    return res


def check_oracle_decode(p, data, res, name):
    st, coeffs, qindex, slen, consumed, bad = O.transform_data(p, data)
    assert st == res["status"], (name, st, res["status"])
    if st != 0:
        return
    assert consumed == res["consumed"], (name, consumed, res["consumed"])
    for c in range(3):
        assert (coeffs[c] == res["coeffs"][c]).all(), (name, "coeffs", c)
        pic = O.decode_component(p, c, coeffs[c])
        assert (pic.astype(np.int64) == res["picture"][c]).all(), (name, "picture", c)


def gen_pictures():
    rng = np.random.RandomState(7)
    rec = {}
    names = []
    for case in PICTURE_CASES:
        (name, W, H, cdiv, depth, wi, wiho, d, dho, sx, sy, profile, mode, arg) = case
        qm = make_quant_matrix(d, dho, rng)
        kind = "noise" if "lossless" in mode else ("ramp" if rng.randint(0, 2) else "noise")
        pics = make_picture(rng, W, H, cdiv, depth, kind)
        picture = {"Y": pics[0].tolist(), "C1": pics[1].tolist(), "C2": pics[2].tolist(), "pic_num": 7}
        cf = {
            "wavelet_index": wi, "wavelet_index_ho": wiho, "dwt_depth": d, "dwt_depth_ho": dho,
            "slices_x": sx, "slices_y": sy, "picture_coding_mode": 0,
            "profile": Profiles.low_delay if profile == "ld" else Profiles.high_quality,
            "quantization_matrix": qm,
            "video_parameters": {
                "frame_width": W, "frame_height": H,
                "color_diff_format_index": {(1, 1): 0, (2, 1): 1, (2, 2): 2}[cdiv],
                "luma_excursion": (1 << depth) - 1, "color_diff_excursion": (1 << depth) - 1,
            },
        }
        # --- reference encoder -------------------------------------------------
        slice_coeffs = encp.transform_and_slice_picture(cf, picture)
        prefix = 0
        if profile == "hq":
            if mode == "lossless":
                scaler, td = encp.make_transform_data_hq_lossless(slice_coeffs)
                qidx = [0] * (sx * sy)
            else:
                min_scaler = 3 if mode == "lossy_prefix" else 1
                prefix = 2 if mode == "lossy_prefix" else 0
                scaler, td = encp.make_transform_data_hq_lossy(arg, slice_coeffs, 0, min_scaler)
                qidx = [s["qindex"] for s in td["hq_slices"]]
            if prefix:
                for s in td["hq_slices"]:
                    s["prefix_bytes"] = bytes(rng.randint(0, 256, prefix).astype(np.uint8).tolist())
            p = O.make_params(W, H, W // cdiv[0], H // cdiv[1], depth, depth, wi, wiho, d, dho, sx, sy,
                              False, prefix, scaler, 0, 1, qm)
        else:
            td = encp.make_transform_data_ld_lossy(arg, slice_coeffs)
            qidx = [s["qindex"] for s in td["ld_slices"]]
            p = O.make_params(W, H, W // cdiv[0], H // cdiv[1], depth, depth, wi, wiho, d, dho, sx, sy,
                              True, 0, 1, arg, sx * sy, qm)
        state = state_from_params(p, qm)
        data = serialise_transform_data(td, state)

        # --- reference picture_encode (coefficients before DC prediction) ------
        est = State(wavelet_index=wi, wavelet_index_ho=wiho, dwt_depth=d, dwt_depth_ho=dho,
                    luma_width=W, luma_height=H, color_diff_width=W // cdiv[0],
                    color_diff_height=H // cdiv[1], luma_depth=depth, color_diff_depth=depth)
        import copy
        picture_encoding.picture_encode(est, copy.deepcopy(picture))
        enc_coeffs = [coeff_dict_to_flat(p, c, est[k]) for c, k in
                      enumerate(["y_transform", "c1_transform", "c2_transform"])]

        # --- oracle encoder must reproduce the reference's bytes ---------------
        o_coeffs = O.encode_picture(p, [a.astype(np.uint16) for a in pics])
        for c in range(3):
            assert (o_coeffs[c] == enc_coeffs[c]).all(), (name, "picture_encode", c)
        if profile == "hq" and mode == "lossless":
            o_data = O.encode_hq_fixed(p, o_coeffs, 0)
            o_q = np.zeros(sx * sy, np.int32)
        elif profile == "hq":
            if prefix == 0:
                o_data, o_q = O.encode_hq_lossy(p, o_coeffs, arg)
            else:  # the oracle writes zero prefix bytes; compare with prefix zeroed
                o_data, o_q = None, None
        else:
            o_data, o_q = O.encode_ld_lossy(p, o_coeffs)
        if o_data is not None:
            assert o_data == data, (name, "bytes", len(o_data), len(data))
            assert list(o_q) == qidx, (name, "qindex")

        # --- reference decoder --------------------------------------------------
        res = ref_decode(p, qm, data + b"\x00")
        assert res["status"] == 0 and res["consumed"] == len(data), (name, res.get("consumed"), len(data))
        check_oracle_decode(p, data + b"\x00", res, name)
        if mode == "lossless":
            for c in range(3):
                assert (res["picture"][c] == pics[c]).all()

        rec[name + ".params"] = params_to_array(p)
        rec[name + ".quant_matrix"] = qm_to_array(qm)
        rec[name + ".data"] = np.frombuffer(data, np.uint8)
        rec[name + ".qindex"] = np.array(qidx, np.int32)
        rec[name + ".arg"] = np.array([arg or 0], np.int64)
        for c in range(3):
            rec[name + ".pic%d" % c] = pics[c].astype(np.uint16)
            rec[name + ".enc_coeffs%d" % c] = enc_coeffs[c].astype(np.int32)
            rec[name + ".dec_coeffs%d" % c] = res["coeffs"][c].astype(np.int32)
            rec[name + ".decoded%d" % c] = res["picture"][c].astype(np.uint16)
        names.append(name)
        print("  picture case %-34s %6d bytes ok" % (name, len(data)))

        # --- corrupted variants (dangling / truncated bounded blocks, EOF) ------
        arr = np.frombuffer(data, np.uint8).copy()
        for v in range(4):
            bad = arr.copy()
            vname = "%s.corrupt%d" % (name, v)
            if v == 0:      # random byte noise all over: lengths change, garbage codes
                idx = rng.randint(0, bad.size, max(2, bad.size // 40))
                bad[idx] = rng.randint(0, 256, idx.size)
            elif v == 1:    # zero runs (long codes) and 0xFF runs
                a = int(rng.randint(0, bad.size - 8))
                bad[a:a + 6] = 0
                b = int(rng.randint(0, bad.size - 8))
                bad[b:b + 5] = 0xFF
            elif v == 2:    # truncated stream
                bad = bad[: int(bad.size * 0.6)]
            else:           # noise in the first slices' headers
                bad[: min(12, bad.size)] = rng.randint(0, 256, min(12, bad.size))
            bdata = bad.tobytes()
            try:
                r = ref_decode(p, qm, bdata)
            except OverflowError:
                continue
            big = r["status"] == 0 and any(np.abs(r["coeffs"][c]).max() >= (1 << 31) for c in range(3))
            # outside the fixed-width envelope (the reference is arbitrary precision): not a vector
            if big or O.transform_data(p, bdata)[0] == O.INT64_ENVELOPE:
                continue
            check_oracle_decode(p, bdata, r, vname)
            rec[vname + ".data"] = np.frombuffer(bdata, np.uint8)
            rec[vname + ".status"] = np.array([r["status"]], np.int32)
            if r["status"] == 0:
                rec[vname + ".consumed"] = np.array([r["consumed"]], np.int64)
                for c in range(3):
                    rec[vname + ".dec_coeffs%d" % c] = r["coeffs"][c].astype(np.int64)
                    rec[vname + ".decoded%d" % c] = r["picture"][c].astype(np.uint16)
    rec["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLDEN, "pictures.npz"), **rec)
    print("pictures.npz ok (%d cases)" % len(names))


def gen_quantize_to_fit():
    rng = np.random.RandomState(8)
    rec = {}
    k = 0
    for _ in range(25):
        nsets = int(rng.choice([1, 2, 3]))
        sets = []
        for s in range(nsets):
            n = int(rng.randint(1, 60))
            scale = int(rng.choice([3, 50, 3000]))
            c = (rng.standard_normal(n) * scale).astype(np.int64)
            if rng.randint(0, 4) == 0:
                c[n // 2:] = 0
            qm = rng.randint(0, 6, n).astype(np.int32)
            sets.append((c, qm))
        align = int(rng.choice([1, 8, 16]))
        full = sum(encp.calculate_coeffs_bits([int(v) for v in c]) for c, _ in sets)
        target = int(rng.randint(0, full + 20))
        minq = int(rng.choice([0, 0, 3]))
        qi, qsets = encp.quantize_to_fit(
            target,
            [encp.ComponentCoeffs([int(v) for v in c], [int(v) for v in qm]) for c, qm in sets],
            align, minq)
        import ctypes as C
        carr = (C.c_void_p * nsets)(*[c.ctypes.data for c, _ in sets])
        qarr = (C.c_void_p * nsets)(*[q.ctypes.data for _, q in sets])
        narr = (C.c_size_t * nsets)(*[c.size for c, _ in sets])
        outs = [np.zeros(c.size, np.int64) for c, _ in sets]
        oarr = (C.c_void_p * nsets)(*[o.ctypes.data for o in outs])
        f = O.lib().vc2o_quantize_to_fit
        f.restype = C.c_int
        f.argtypes = [C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        oq = f(target, nsets, carr, qarr, narr, align, minq, oarr)
        assert oq == qi, (oq, qi)
        for o, q in zip(outs, qsets):
            assert list(o) == q
        rec["meta_%d" % k] = np.array([nsets, align, target, minq, qi], np.int64)
        for s in range(nsets):
            rec["c_%d_%d" % (k, s)] = sets[s][0]
            rec["qm_%d_%d" % (k, s)] = sets[s][1]
            rec["q_%d_%d" % (k, s)] = np.array(qsets[s], np.int64)
            rec["bits_%d_%d" % (k, s)] = np.array([encp.calculate_coeffs_bits(qsets[s])], np.int64)
        k += 1
    rec["count"] = np.array([k], np.int32)
    np.savez_compressed(os.path.join(GOLDEN, "quantize_to_fit.npz"), **rec)
    print("quantize_to_fit.npz ok")


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    O.build(force=True)
    gen_quant()
    gen_geometry()
    gen_expgolomb()
    gen_lifting()
    gen_transform()
    gen_dc_prediction()
    gen_quantize_to_fit()
    gen_pictures()
    print("all golden vectors regenerated and the C oracle agrees with the reference on each")


if __name__ == "__main__":
    main()
