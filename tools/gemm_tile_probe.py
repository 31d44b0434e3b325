"""ComplexF64 DMMA GEMM, tile variants (ttn_set_option("gemm_compact", v)): 0 = 64x128 / 64x64 tiles with 4 pipeline stages (one CTA
per SM), 1 = 64x64 tile with 2 stages (68 KB, two to three CTAs per SM).  Prints TFLOP/s (8 M N K) per shape and variant."""
import ctypes as C
import json
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import ttn_b200 as t                                   # noqa: E402
from ttn_b200 import _lib                              # noqa: E402


def real_tiles(lib):
    """Float64: option gemm_real_tile 0 (128x128, one CTA per SM; TMA-staged when aligned) vs 1 / 2 (128x64, 3 / 2 stages) vs 3 (64x64)"""
    shapes = [(5120, 4096, 1024, 1), (2048, 2048, 2048, 1), (1024, 1024, 5120, 1), (1024, 512, 5120, 1), (4096, 4096, 1024, 1)]
    out = []
    for (M, N, K, nb) in shapes:
        ptrs = []
        for n in (M * K * nb, K * N * nb, M * N * nb):
            p = C.c_void_p()
            _lib.check(lib.ttn_dev_alloc(n * 8, C.byref(p)))
            h = np.random.default_rng(1).standard_normal(n)
            _lib.check(lib.ttn_h2d(p, h.ctypes.data, h.nbytes))
            ptrs.append(p)
        row = {"M": M, "N": N, "K": K}
        for v in (0, 1, 4, 5):
            t.set_option("gemm_real_tile", v)

            def run():
                _lib.check(lib.ttn_gemm(0, M, N, K, ptrs[0], 1, M, 0, ptrs[1], 1, K, 0, ptrs[2], 1, M, 1.0, 0.0, nb, M * K, K * N, M * N))
            run()
            t.synchronize()
            reps = 10
            t0 = time.perf_counter()
            for _ in range(reps):
                run()
            t.synchronize()
            el = (time.perf_counter() - t0) / reps
            row[f"tflops_v{v}"] = round(2.0 * M * N * K * nb / el / 1e12, 2)
        out.append(row)
        for p in ptrs:
            _lib.check(lib.ttn_dev_free(p))
    t.set_option("gemm_real_tile", 0)
    print(json.dumps(out))


def main():
    lib = _lib.lib()
    if len(sys.argv) > 1 and sys.argv[1] == "real":
        return real_tiles(lib)
    shapes = [(128, 128, 512, 296), (128, 512, 128, 296), (256, 128, 128, 296), (2048, 2048, 2048, 1), (4096, 1024, 1024, 1),
              (1024, 1024, 5120, 1), (512, 512, 512, 8)]
    out = []
    for (M, N, K, nb) in shapes:
        es = 16
        ptrs = []
        for n in (M * K * nb, K * N * nb, M * N * nb):
            p = C.c_void_p()
            _lib.check(lib.ttn_dev_alloc(n * es, C.byref(p)))
            h = np.random.default_rng(1).standard_normal(2 * n)
            _lib.check(lib.ttn_h2d(p, h.ctypes.data, h.nbytes))
            ptrs.append(p)
        row = {"M": M, "N": N, "K": K, "batch": nb}
        for v in (0, 1):
            t.set_option("gemm_compact", v)

            def run():
                _lib.check(lib.ttn_gemm(2, M, N, K, ptrs[0], 1, M, 0, ptrs[1], 1, K, 0, ptrs[2], 1, M, 1.0, 0.0, nb, M * K, K * N, M * N))
            run()
            t.synchronize()
            reps = 10
            t0 = time.perf_counter()
            for _ in range(reps):
                run()
            t.synchronize()
            el = (time.perf_counter() - t0) / reps
            row[f"tflops_v{v}"] = round(8.0 * M * N * K * nb / el / 1e12, 2)
        out.append(row)
        for p in ptrs:
            _lib.check(lib.ttn_dev_free(p))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
