"""TTO x TTV apply at the cfg5 shapes (ComplexF64, d=30, r=64, W=4, `nvec` vectors): device time of the apply kernels
(per-family CUDA events) and achieved HBM GB/s = bytes written + read / time.  usage: apply_probe.py [nvec] [reps]"""
import json
import math
import sys

import numpy as np

sys.path.insert(0, '.')
import ttn_b200 as t

nvec = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
d, r, W = 30, 64, 4
rks = [min(2 ** k, 2 ** (d - k), r) for k in range(d + 1)]
Rk = [min(4 ** k, 4 ** (d - k), W) for k in range(d + 1)]
rng = np.random.default_rng(7)
A = t.TToperator(d, [np.asfortranarray((rng.standard_normal((2, 2, Rk[k], Rk[k + 1])) + 1j * rng.standard_normal((2, 2, Rk[k], Rk[k + 1])))
                                       / math.sqrt(2.0 * Rk[k + 1])) for k in range(d)], (2,) * d, Rk)
cores = [np.asfortranarray((rng.standard_normal((2, rks[k], rks[k + 1], nvec)) + 1j * rng.standard_normal((2, rks[k], rks[k + 1], nvec)))
                           / math.sqrt(4.0 * rks[k + 1])) for k in range(d)]
xs = [t.TTvector(d, [c[..., b] for c in cores], (2,) * d, rks) for b in range(nvec)]
Ad, xd = t.DeviceTTO.upload(A), t.DeviceTT.upload(xs)
y = t.apply(Ad, xd); del y
t.synchronize()
t.profile(True)
for _ in range(reps):
    y = t.apply(Ad, xd)
    del y
t.synchronize()
fam = t.profile_read(); t.profile(False)
ms = fam["apply"][0] / reps
out_bytes = sum(2 * Rk[k] * rks[k] * Rk[k + 1] * rks[k + 1] for k in range(d)) * 16 * nvec
in_bytes = sum(2 * rks[k] * rks[k + 1] for k in range(d)) * 16 * nvec
print(json.dumps({"nvec": nvec, "apply_ms": ms, "bytes_written": out_bytes, "bytes_read": in_bytes,
                  "GBps": (out_bytes + in_bytes) / ms / 1e6, "launches_per_apply": fam["apply"][1] // reps}))
