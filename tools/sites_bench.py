"""Timing of the site-surgery compositions on the GPU (own input builder; the oracle is not used).
usage: python tools/sites_bench.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import ttn_b200 as t
    rng = np.random.default_rng(0)
    out = {}

    def rand_tt(dims, rmax):
        d = len(dims)
        rks = [1] * (d + 1)
        for k in range(1, d):
            rks[k] = int(min(rmax, np.prod([float(n) for n in dims[:k]]), np.prod([float(n) for n in dims[k:]])))
        cores = [np.asfortranarray(rng.standard_normal((dims[k], rks[k], rks[k + 1]))) for k in range(d)]
        return t.TTvector(d, cores, dims, rks)

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        return min(ts)

    # reorder: 2 dims x 12 bits, rank 32, serial -> interleaved (66 swaps) with a 1e-10 relative threshold
    bits = 12
    x = rand_tt((2,) * (2 * bits), 32)
    xd = t.DeviceTT.upload(x)
    out["reorder_2x12_r32_s"] = timed(lambda: t.reorder(xd, 2, bits, "serial", "interleaved", threshold=1e-10))
    y = t.reorder(xd, 2, bits, "serial", "interleaved", threshold=1e-10)
    out["reorder_max_rank"] = max(y.ttv_rks)

    # hadamard_ttm: d = 16, ranks 24 x 24, tol 1e-10, rmax 64
    a = rand_tt((2,) * 16, 24); b = rand_tt((2,) * 16, 24)
    out["hadamard_ttm_d16_r24_s"] = timed(lambda: t.hadamard_ttm(a, b, tol=1e-10, rmax=64), reps=2)

    # exact hadamard + rounding (the route the apply + rounding kernels serve)
    ad, bd = a, t.DeviceTT.upload(b)
    out["hadamard_then_compress_s"] = timed(lambda: t.tt_compress_(t.hadamard(ad, bd), 64), reps=2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
