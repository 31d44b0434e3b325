"""Diagnostics of the Gram path of tt_compress! at the cfg2 / cfg5 shapes: was it accepted, how long does a sweep take.
python tools/gram_probe.py [cfg2|cfg5|both]"""
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ttn_b200 as t


def cfg2(d=40, rmax=512, mb=64):
    rks = [min(2 ** k, 2 ** (d - k), rmax) for k in range(d + 1)]
    rng = np.random.default_rng(1)
    cores = [np.asfortranarray(rng.standard_normal((2, rks[k], rks[k + 1])) / math.sqrt(2 * rks[k + 1])) for k in range(d)]
    x = t.DeviceTT.upload(t.TTvector(d, cores, (2,) * d, rks))
    out = {}
    for rep in range(3):
        y = x.copy()
        t.synchronize(); t0 = time.perf_counter()
        t.reset_launch_count()
        t.tt_compress_(y, mb)
        t.synchronize(); out[f"ms_{rep}"] = (time.perf_counter() - t0) * 1e3
    out["launches"] = t.launch_count()
    for k in ("gram_calls", "gram_fallbacks", "gram_last_flags"):
        out[k] = t.get_option(k)
    y = x.copy()
    t.profile(True); t.tt_compress_(y, mb); out["families"] = {k: round(v[0], 3) for k, v in t.profile_read().items()}; t.profile(False)
    return out


def cfg5(nvec=int(os.environ.get("NVEC", "256")), d=30, r=64, W=4):
    rks = [min(2 ** k, 2 ** (d - k), r) for k in range(d + 1)]
    Rk = [min(4 ** k, 4 ** (d - k), W) for k in range(d + 1)]
    rng = np.random.default_rng(7)
    A = t.TToperator(d, [np.asfortranarray((rng.standard_normal((2, 2, Rk[k], Rk[k + 1])) + 1j * rng.standard_normal((2, 2, Rk[k], Rk[k + 1])))
                                           / math.sqrt(2.0 * Rk[k + 1])) for k in range(d)], (2,) * d, Rk)
    Ad = t.DeviceTTO.upload(A)
    g = np.random.default_rng(100)
    cores = [np.asfortranarray((g.standard_normal((2, rks[k], rks[k + 1], nvec)) + 1j * g.standard_normal((2, rks[k], rks[k + 1], nvec)))
                               / math.sqrt(4.0 * rks[k + 1])) for k in range(d)]
    xs = [t.TTvector(d, [c[..., b] for c in cores], (2,) * d, rks) for b in range(nvec)]
    xd = t.DeviceTT.upload(xs)
    out = {}
    for rep in range(3):
        t.synchronize(); t0 = time.perf_counter()
        t.reset_launch_count()
        t.tt_compress_(t.apply(Ad, xd), r)
        t.synchronize(); out[f"ms_{rep}"] = (time.perf_counter() - t0) * 1e3
    out["vectors_per_s"] = nvec / (out["ms_2"] * 1e-3)
    for rep in range(3):
        t.synchronize(); t0 = time.perf_counter()
        t.apply_compress(Ad, xd, r)
        t.synchronize(); out[f"fused_ms_{rep}"] = (time.perf_counter() - t0) * 1e3
    out["fused_vectors_per_s"] = nvec / (out["fused_ms_2"] * 1e-3)
    t.profile(True); t.apply_compress(Ad, xd, r); out["fused_families"] = {k: round(v[0], 3) for k, v in t.profile_read().items()}; t.profile(False)
    out["launches"] = t.launch_count()
    for k in ("gram_calls", "gram_fallbacks", "gram_last_flags"):
        out[k] = t.get_option(k)
    t.profile(True); t.tt_compress_(t.apply(Ad, xd), r); out["families"] = {k: round(v[0], 3) for k, v in t.profile_read().items()}; t.profile(False)
    return out


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "both"
    res = {}
    if which in ("cfg2", "both"):
        res["cfg2"] = cfg2()
    if which in ("cfg5", "both"):
        res["cfg5"] = cfg5()
    print(json.dumps(res))
