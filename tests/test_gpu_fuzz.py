"""Seeded random parameter fuzzing of the whole CUDA path against the C oracle: odd and tiny
dimensions, every filter pair, depth / horizontal-only depth mixes, slice counts that do not divide
the subbands (empty and ragged rectangles), HQ with prefix bytes and scalers, LD with odd slice
sizes.  Decode must reproduce the oracle's pictures and the device encoder the oracle's bytes."""

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def _random_case(rng):
    d = int(rng.randint(0, 5))
    dho = int(rng.randint(0, 4)) if rng.randint(0, 3) == 0 else 0
    while d + dho > 5:
        d -= 1
    W = int(rng.choice([1, 2, 3, 5, 8, 17, 31, 32, 33, 64, 100, 129, 200, 260]))
    H = int(rng.choice([1, 2, 3, 4, 7, 16, 30, 33, 64, 90, 131]))
    fmt = int(rng.randint(0, 3))
    cw = max(1, W // (2 if fmt >= 1 else 1))
    ch = max(1, H // (2 if fmt == 2 else 1))
    depth = int(rng.choice([8, 10, 12, 16]))
    wi = int(rng.randint(0, 7))
    wiho = wi if rng.randint(0, 3) else int(rng.randint(0, 7))
    sx = int(rng.randint(1, 9))
    sy = int(rng.randint(1, 9))
    is_ld = bool(rng.randint(0, 3) == 0)
    return dict(W=W, H=H, cw=cw, ch=ch, depth=depth, wi=wi, wiho=wiho, d=d, dho=dho, sx=sx, sy=sy, is_ld=is_ld,
                prefix=int(rng.choice([0, 0, 1, 3])), kind=int(rng.randint(0, 3)))


CASES = [_random_case(np.random.RandomState(1000 + i)) for i in range(48)]


@pytest.mark.parametrize("k", range(len(CASES)))
def test_fuzz_decode_encode_vs_oracle(k, unpack_kernel):
    import ctypes as C

    import oracle as O
    from vc2_conformance_b200 import capi
    from vc2_conformance_b200.device import BatchCodec
    from vc2_conformance_b200.params import make_params

    c = CASES[k]
    rng = np.random.RandomState(2000 + k)
    S = c["W"] * c["H"] + 2 * c["cw"] * c["ch"]
    ns = c["sx"] * c["sy"]
    qm = synth.default_quant_matrix(c["d"], c["dho"])
    kw = dict(luma_width=c["W"], luma_height=c["H"], color_diff_width=c["cw"], color_diff_height=c["ch"],
              luma_depth=c["depth"], wavelet_index=c["wi"], wavelet_index_ho=c["wiho"], dwt_depth=c["d"],
              dwt_depth_ho=c["dho"], slices_x=c["sx"], slices_y=c["sy"], quant_matrix=qm)
    if c["is_ld"]:
        pbytes = max(ns * 2, int(S * c["depth"] / 8 / (2 + c["kind"])))
        kw.update(is_ld=True, slice_bytes_numerator=pbytes, slice_bytes_denominator=ns)
        picture_bytes = 0
    else:
        picture_bytes = max(ns * 6, int(S * c["depth"] / 8 / (1.5 + c["kind"]))) // 4 * 4 + 4 * ns
        kw.update(slice_prefix_bytes=c["prefix"],
                  slice_size_scaler=O.safe_lossy_hq_slice_size_scaler(picture_bytes, ns) + int(rng.randint(0, 2)))
    p = make_params(**kw)
    po = synth.oracle_params(p)
    pics = [synth.noise_picture(p, 3 * k), synth.ramp_picture(p, 3 * k + 1)]
    streams, decoded = [], []
    for planes in pics:
        coeffs = O.encode_picture(po, planes)
        if c["is_ld"]:
            try:
                data, _q = O.encode_ld_lossy(po, coeffs)
            except ValueError:
                pytest.skip("slice too small for its header (InsufficientLDPictureBytesError)")
        else:
            data, _q = O.encode_hq_lossy(po, coeffs, picture_bytes)
        st, dec, consumed = O.decode_picture(po, data)
        assert st == 0 and consumed == len(data)
        streams.append(data)
        decoded.append(dec)
    codec = BatchCodec(p, 2)
    codec.decode_streams(streams, check=False)
    st = codec.status_host(2)
    if (st[:, 0] == capi.ST_INT32_ENVELOPE).any():
        # only legitimate when the coefficients really leave the documented envelope
        safe = capi.lib().vc2b_idwt_max_safe_coeff(C.byref(p))
        assert int(st[:, 3].max()) > safe, (c, st)
        pytest.skip("outside the int32 envelope (flagged, not wrapped)")
    assert (st[:, 0] == 0).all(), (c, st)
    for i in range(2):
        planes = codec.picture_planes(i)
        for comp in range(3):
            assert (planes[comp] == decoded[i][comp]).all(), (c, i, comp)
    flat = np.stack([synth.flat_picture(pl) for pl in pics])
    try:
        out = codec.encode_pictures(flat, picture_bytes)
    except capi.Vc2bError as e:
        if e.code == capi.E_UNSUPPORTED:
            pytest.skip("analysis outside the int32 envelope for this depth/filter/levels")
        raise
    assert out[0] == streams[0] and out[1] == streams[1], c
