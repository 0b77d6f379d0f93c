"""Synthetic inputs shared by tests and bench.py: the reference's white-noise recipe
(vc2_conformance/picture_generators.py:600-639: RandomState(seed).randint(0, 1<<depth, (h,w)) for
Y, C1, C2 in that order) and a pure-integer ramp, plus stream generation through the oracle
encoder (test infrastructure)."""

import numpy as np


def noise_picture(p, seed):
    rand = np.random.RandomState(seed)
    dims = [(p.luma_height, p.luma_width, p.luma_depth),
            (p.color_diff_height, p.color_diff_width, p.color_diff_depth),
            (p.color_diff_height, p.color_diff_width, p.color_diff_depth)]
    return [rand.randint(0, 1 << d, (h, w)).astype(np.uint16) for (h, w, d) in dims]


def ramp_picture(p, seed):
    rand = np.random.RandomState(seed)
    dims = [(p.luma_height, p.luma_width, p.luma_depth),
            (p.color_diff_height, p.color_diff_width, p.color_diff_depth),
            (p.color_diff_height, p.color_diff_width, p.color_diff_depth)]
    out = []
    for (h, w, d) in dims:
        x = (np.arange(w, dtype=np.int64) * ((1 << d) - 1)) // max(w - 1, 1)
        y = (np.arange(h, dtype=np.int64) * ((1 << d) // 8)) // max(h - 1, 1)
        a = x[None, :] // 2 + y[:, None] + rand.randint(0, 4, (h, w))
        out.append(np.clip(a, 0, (1 << d) - 1).astype(np.uint16))
    return out


def flat_picture(planes):
    return np.concatenate([a.reshape(-1) for a in planes]).astype(np.uint16)


def oracle_params(p):
    """vc2_conformance_b200 Params -> oracle Params (same field names)."""
    import oracle as O

    q = O.Params()
    for f, _t in O.Params._fields_:
        if f == "quant_matrix":
            for lv in range(O.MAX_LEVELS):
                for o in range(4):
                    q.quant_matrix[lv][o] = p.quant_matrix[lv][o]
        else:
            setattr(q, f, getattr(p, f))
    return q


def default_quant_matrix(dwt_depth, dwt_depth_ho):
    """A custom quantisation matrix (streams never rely on Annex D defaults)."""
    qm = {}
    if dwt_depth_ho == 0:
        qm[0] = {"LL": 0}
    else:
        qm[0] = {"L": 0}
        for level in range(1, dwt_depth_ho + 1):
            qm[level] = {"H": min(level, 3)}
    for level in range(dwt_depth_ho + 1, dwt_depth_ho + dwt_depth + 1):
        k = level - dwt_depth_ho
        qm[level] = {"HL": min(k, 4), "LH": min(k, 4), "HH": min(k + 1, 5)}
    return qm
