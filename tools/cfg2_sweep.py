"""One cfg2 sweep (tt_compress! of a random Float64 TT, d=40, rank 512 -> 64), repeated: target of the ncu launch lists."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ttn_b200 as t

d, rmax, mb = 40, 512, 64
rks = [min(2 ** k, 2 ** (d - k), rmax) for k in range(d + 1)]
rng = np.random.default_rng(1)
x = t.DeviceTT.upload(t.TTvector(d, [np.asfortranarray(rng.standard_normal((2, rks[k], rks[k + 1])) / math.sqrt(2 * rks[k + 1])) for k in range(d)],
                                 (2,) * d, rks))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    y = t.tt_compress_(x.copy(), mb)
t.synchronize()
print("ok", max(y.ttv_rks), t.get_option("gram_fallbacks"))
