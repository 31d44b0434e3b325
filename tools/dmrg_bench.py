"""DMRG two-site sweep at cfg4-like shapes (Heisenberg XYZ chain, MPO rank 5): one full sweep
(`sweep_schedule=[2]`) from a random orthogonalised TT with bond cap chi and a fixed Lanczos budget
(krylovdim x 1 restart), as described in SURVEY.md §8(d)-4.  Prints one JSON line.

usage: python tools/dmrg_bench.py L chi [krylovdim]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ttn_b200 as t  # noqa: E402


def heisenberg(d, jx=1.1, jy=0.8, jz=1.2):
    """src/tt_operators.jl:162-218 with λ = 0 (real encoding of σyσy)"""
    X = np.array([[0.0, 1.0], [1.0, 0.0]]); Z = np.array([[1.0, 0.0], [0.0, -1.0]])
    yr = np.array([[0.0, -1.0], [1.0, 0.0]]); Y1, Y2, I2 = -yr, yr, np.eye(2)
    cores = []
    c = np.zeros((2, 2, 1, 5)); c[:, :, 0, 1] = jx * X; c[:, :, 0, 2] = jy * Y1; c[:, :, 0, 3] = jz * Z; c[:, :, 0, 4] = I2
    cores.append(c)
    for _ in range(d - 2):
        c = np.zeros((2, 2, 5, 5))
        c[:, :, 0, 0] = I2; c[:, :, 1, 0] = X; c[:, :, 2, 0] = Y2; c[:, :, 3, 0] = Z
        c[:, :, 4, 1] = jx * X; c[:, :, 4, 2] = jy * Y1; c[:, :, 4, 3] = jz * Z; c[:, :, 4, 4] = I2
        cores.append(c)
    c = np.zeros((2, 2, 5, 1)); c[:, :, 0, 0] = I2; c[:, :, 1, 0] = X; c[:, :, 2, 0] = Y2; c[:, :, 3, 0] = Z
    cores.append(c)
    return t.TToperator(d, cores, (2,) * d, [1] + [5] * (d - 1) + [1])


def rand_tt(d, chi, seed=3):
    rks = [min(2 ** k, 2 ** (d - k), chi) for k in range(d + 1)]
    rng = np.random.default_rng(seed)
    cores = [np.asfortranarray(rng.standard_normal((2, rks[k], rks[k + 1])) / np.sqrt(2 * rks[k + 1])) for k in range(d)]
    return t.TTvector(d, cores, (2,) * d, rks)


def main():
    L, chi = int(sys.argv[1]), int(sys.argv[2])
    kd = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    H = t.DeviceTTO.upload(heisenberg(L))
    x0 = t.DeviceTT.upload(rand_tt(L, chi))
    t.synchronize()
    t.reset_launch_count()
    t.profile(True)
    t0 = time.perf_counter()
    E, x, rh = t.dmrg_eigsolve(H, x0, N=2, tol=1e-12, sweep_schedule=[2], rmax_schedule=[chi], linsolv_maxiter=1,
                               linsolv_tol=1e-10, krylovdim=kd)
    t.synchronize()
    el = time.perf_counter() - t0
    fam = t.profile_read()
    t.profile(False)
    print(json.dumps({"metric": "DMRG sweep s", "L": L, "chi": chi, "krylovdim": kd, "sweep_s": el, "bond_steps": len(E),
                      "E_first": float(E[0]), "E_last": float(E[-1]), "max_rank": max(rh), "launches": t.launch_count(),
                      "family_ms": {k: round(v[0], 2) for k, v in fam.items()}}))


if __name__ == "__main__":
    main()
