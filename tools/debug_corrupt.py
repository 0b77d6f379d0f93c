import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
from golden_utils import load
from vc2_conformance_b200.params import params_from_arrays, Geometry
from vc2_conformance_b200.device import BatchCodec

name, v = sys.argv[1], sys.argv[2]
z = load("pictures")
p = params_from_arrays(z[name + ".params"], z[name + ".quant_matrix"])
g = Geometry(p)
key = "%s.corrupt%s" % (name, v)
data = z[key + ".data"].tobytes()
print("first bytes", [hex(b) for b in data[:16]], "len", len(data))
codec = BatchCodec(p, 1)
codec.decode_streams([data], check=False)
print("status", codec.status_host(1))
print("qindex", codec.qindex[:8].cpu().numpy())
co = codec.coeff_components(0)
for c in range(3):
    exp = z[key + ".dec_coeffs%d" % c]
    bad = np.nonzero(co[c] != exp)[0]
    print("comp", c, "mismatches", bad.size)
    for i in bad[:12]:
        for b, B in enumerate(g.bands[c]):
            if B["off"] <= i < B["off"] + B["w"] * B["h"]:
                l = i - B["off"]
                print("   idx", i, "band", b, "y", l // B["w"], "x", l % B["w"], "got", co[c][i], "exp", exp[i])
