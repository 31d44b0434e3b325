"""Timing of the site-surgery compositions on the GPU; CPU oracle timed beside on the same input.
usage: python tools/sites_bench.py [--cpu]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import torch
    import ttn_b200 as t
    import ttn_oracle as o
    cpu = "--cpu" in sys.argv
    rng = np.random.default_rng(0)
    out = {}

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        return min(ts)

    # reorder: 2 dims x 12 bits, rank 32, serial -> interleaved (66 swaps) with a 1e-10 relative threshold
    bits = 12
    x = o.rand_tt((2,) * (2 * bits), 32, rng=rng)
    xd = t.DeviceTT.upload(x)
    out["reorder_2x12_r32_s"] = timed(lambda: t.reorder(xd, 2, bits, "serial", "interleaved", threshold=1e-10))
    y = t.reorder(xd, 2, bits, "serial", "interleaved", threshold=1e-10)
    out["reorder_max_rank"] = max(y.ttv_rks)
    if cpu:
        t0 = time.perf_counter(); o.reorder(x, 2, bits, "serial", "interleaved", threshold=1e-10); out["reorder_cpu_s"] = time.perf_counter() - t0

    # hadamard_ttm: d = 16, ranks 24 x 24, tol 1e-10, rmax 64
    a = o.rand_tt((2,) * 16, 24, rng=rng); b = o.rand_tt((2,) * 16, 24, rng=rng)
    out["hadamard_ttm_d16_r24_s"] = timed(lambda: t.hadamard_ttm(a, b, tol=1e-10, rmax=64), reps=2)
    if cpu:
        t0 = time.perf_counter(); o.hadamard_ttm(a, b, tol=1e-10, rmax=64); out["hadamard_ttm_cpu_s"] = time.perf_counter() - t0

    # exact hadamard + rounding (the route the apply + rounding kernels serve)
    ad, bd = a, t.DeviceTT.upload(b)
    out["hadamard_then_compress_s"] = timed(lambda: t.tt_compress_(t.hadamard(ad, bd), 64), reps=2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
