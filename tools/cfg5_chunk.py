"""One cfg5 chunk (296 vectors) through the fused apply + tt_compress!, twice: target of the ncu launch lists."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ttn_b200 as t

nvec, d, r, W = int(os.environ.get("NVEC", "296")), 30, 64, 4
rks = [min(2 ** k, 2 ** (d - k), r) for k in range(d + 1)]
Rk = [min(4 ** k, 4 ** (d - k), W) for k in range(d + 1)]
rng = np.random.default_rng(7)
A = t.TToperator(d, [np.asfortranarray((rng.standard_normal((2, 2, Rk[k], Rk[k + 1])) + 1j * rng.standard_normal((2, 2, Rk[k], Rk[k + 1])))
                                       / math.sqrt(2.0 * Rk[k + 1])) for k in range(d)], (2,) * d, Rk)
Ad = t.DeviceTTO.upload(A)
g = np.random.default_rng(100)
cores = [np.asfortranarray((g.standard_normal((2, rks[k], rks[k + 1], nvec)) + 1j * g.standard_normal((2, rks[k], rks[k + 1], nvec)))
                           / math.sqrt(4.0 * rks[k + 1])) for k in range(d)]
xd = t.DeviceTT.upload_batched(cores, (2,) * d, rks)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    y = t.apply_compress(Ad, xd, r)
t.synchronize()
print("ok", y.ttv_rks[:4], t.get_option("gram_fallbacks"))
