"""cfg3 (BASELINE.json configs[2]): 2-D Laplace in interleaved QTT format, 2 x `bits` bits, `mals_linsolve` with rank cap
`rmax`.  One call = one forward and one backward two-site sweep (mals.jl:240-309), local systems
solved matrix-free by GMRES on the DMMA three-GEMM matvec.  Prints one JSON line: wall seconds, relative residual
||Ax-b||/||b|| (TT arithmetic on the device), ranks, launches.

usage: python tools/cfg3_bench.py [--bits 20] [--rmax 128] [--tol 1e-12] [--maxiter 200] [--krylovdim 30]
The operator / right-hand side builders below restate the reference generators (tt_operators.jl:4-19,282-284,654-656,
qtt_tools.jl:138-154) for this tool; the oracle package is test infrastructure and is not imported here.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ttn_b200 as t  # noqa: E402


def toeplitz_cores(alpha, beta, gamma, d):
    I2, J = np.eye(2), np.array([[0.0, 1.0], [0.0, 0.0]])
    cores = []
    for k in range(d):
        rl, rr = (1 if k == 0 else 3), (1 if k == d - 1 else 3)
        c = np.zeros((2, 2, rl, rr))
        for i in range(2):
            for j in range(2):
                if d == 1:
                    c[i, j, 0, 0] = alpha * I2[i, j] + beta * J[i, j] + gamma * J[j, i]
                elif k == 0:
                    c[i, j, 0, :] = [I2[i, j], J[j, i], J[i, j]]
                elif k == d - 1:
                    c[i, j, :, 0] = [alpha * I2[i, j] + beta * J[i, j] + gamma * J[j, i], gamma * J[i, j], beta * J[j, i]]
                else:
                    c[i, j] = np.array([[I2[i, j], J[j, i], J[i, j]], [0.0, J[i, j], 0.0], [0.0, 0.0, J[j, i]]])
        cores.append(c)
    return cores


def interleave(cores, own_first):
    out = []
    for c in cores:
        Rl, Rr = c.shape[2], c.shape[3]
        R = Rr if own_first else Rl
        p = np.zeros((2, 2, R, R))
        for s in range(2):
            p[s, s] = np.eye(R)
        out += [c.copy(), p] if own_first else [p, c.copy()]
    return out


def add_ops(xc, yc):
    d, out = len(xc), []
    for k, (a, b) in enumerate(zip(xc, yc)):
        if k == 0:
            c = np.concatenate([a, b], axis=3)
        elif k == d - 1:
            c = np.concatenate([a, b], axis=2)
        else:
            c = np.zeros((2, 2, a.shape[2] + b.shape[2], a.shape[3] + b.shape[3]))
            c[:, :, :a.shape[2], :a.shape[3]] = a
            c[:, :, a.shape[2]:, a.shape[3]:] = b
        out.append(c)
    return out


def laplace2d(bits):
    lap = toeplitz_cores(2.0, -1.0, -1.0, bits)
    cores = add_ops(interleave(lap, True), interleave(lap, False))
    h = 1.0 / (2 ** bits - 1)
    cores[0] = cores[0] / h ** 2
    d = 2 * bits
    return t.TToperator(d, [np.asfortranarray(c) for c in cores], (2,) * d, [1] + [c.shape[3] for c in cores])


def qtt_sin_cores(d, lam=1.0, a=0.0, b=1.0):
    """sin(lam*pi*x) on the 2^d points of [a,b] as a rank-2 QTT, site 1 = most significant bit (qtt_tools.jl:138-154)"""
    h = (b - a) / (2 ** d - 1)
    w = lam * np.pi
    cores = [np.zeros((2, 1, 2))] + [np.zeros((2, 2, 2)) for _ in range(d - 2)] + [np.zeros((2, 2, 1))]
    cores[0][0, 0, :] = [np.sin(w * a), np.cos(w * a)]
    t1 = w * (a + h * 2 ** (d - 1))
    cores[0][1, 0, :] = [np.sin(t1), np.cos(t1)]
    for k in range(2, d):
        tk = w * h * 2 ** (d - k)
        cores[k - 1][0] = np.eye(2)
        cores[k - 1][1] = [[np.cos(tk), -np.sin(tk)], [np.sin(tk), np.cos(tk)]]
    cores[d - 1][0, 0, 0] = 1.0
    cores[d - 1][1, :, 0] = [np.cos(w * h), np.sin(w * h)]
    return cores


def sin2d(bits):
    sx = qtt_sin_cores(bits); sy = qtt_sin_cores(bits)
    vec = []
    for k in range(bits):
        cx, cy = sx[k], sy[k]
        ryl = cy.shape[1]
        gx = np.einsum("sab,cd->sacbd", cx, np.eye(ryl)).reshape(2, cx.shape[1] * ryl, cx.shape[2] * ryl, order="F")
        rxr = cx.shape[2]
        gy = np.einsum("ab,scd->sacbd", np.eye(rxr), cy).reshape(2, rxr * cy.shape[1], rxr * cy.shape[2], order="F")
        vec += [gx, gy]
    d = 2 * bits
    return t.TTvector(d, [np.asfortranarray(c) for c in vec], (2,) * d, [1] + [c.shape[2] for c in vec])


def run(bits=20, rmax=128, tol=1e-12, maxiter=200, krylovdim=30, x0rank=8, profile=False):
    d = 2 * bits
    A, b = laplace2d(bits), sin2d(bits)
    rks = [min(2 ** k, 2 ** (d - k), x0rank) for k in range(d + 1)]
    rng = np.random.default_rng(2)
    x0 = t.TTvector(d, [np.asfortranarray(rng.standard_normal((2, rks[k], rks[k + 1])) / np.sqrt(2 * rks[k + 1])) for k in range(d)],
                    (2,) * d, rks)
    Ad, bd, xd = t.DeviceTTO.upload(A), t.DeviceTT.upload(b), t.DeviceTT.upload(x0)
    t.mals_linsolve(Ad, bd, xd, tol=tol, rmax=rmax, linsolv_maxiter=maxiter, krylovdim=krylovdim)     # warm-up (allocator, attributes)
    t.synchronize()
    t.reset_launch_count()
    if profile:
        t.profile(True)
    t0 = time.perf_counter()
    x, info = t.mals_linsolve(Ad, bd, xd, tol=tol, rmax=rmax, return_info=True, linsolv_maxiter=maxiter, krylovdim=krylovdim)
    t.synchronize()
    el = time.perf_counter() - t0
    fam = None
    if profile:
        fam = {k: round(v[0], 1) for k, v in t.profile_read().items()}
        t.profile(False)
    launches = int(t.launch_count())
    r = t.sub(t.apply(Ad, x), bd)
    res = t.norm(r) / t.norm(bd)
    return {"metric": "cfg3 mals_linsolve s", "value": el, "unit": "s", "bits": bits, "d": d, "rmax": rmax, "tol": tol,
            "relative_residual": float(res), "solver_residual": float(info["residual"]), "max_rank": int(max(x.ttv_rks)),
            "mpo_rank": int(max(A.tto_rks)), "gpu_launches": launches, "family_ms": fam,
            "note": "one mals_linsolve call = one forward + one backward two-site sweep from a random rank-8 start "
                    "(mals.jl:240-309); the residual is that of the reference algorithm after one call, not of a converged solve"}


def heat_operator(bits, dt_scale=100.0):
    """I + dt * L with L the interleaved 2-D Laplace operator above and dt = dt_scale / ||L||: the operator of one implicit Euler
    step of the heat equation on the 2^bits x 2^bits grid (what `implicit_euler_method`, src/solvers/euler.jl:99-135, hands to
    `mals_linsolve`); condition number ~ dt_scale, so the matrix-free local solves converge and the residual is meaningful"""
    lap = toeplitz_cores(2.0, -1.0, -1.0, bits)
    lcores = add_ops(interleave(lap, True), interleave(lap, False))
    h = 1.0 / (2 ** bits - 1)
    dt = dt_scale / (8.0 / h ** 2)
    lcores[0] = lcores[0] * (dt / h ** 2)
    d = 2 * bits
    ident = [np.eye(2).reshape(2, 2, 1, 1).copy() for _ in range(d)]
    cores = add_ops(ident, lcores)
    return t.TToperator(d, [np.asfortranarray(c) for c in cores], (2,) * d, [1] + [c.shape[3] for c in cores]), dt


def heat_problem(bits, rmax, start_rank, dt_scale=100.0):
    """operator, a known solution of rank `rmax`, and a random start of rank `start_rank` (host objects of the package's mirror types)"""
    d = 2 * bits
    A, dt = heat_operator(bits, dt_scale)
    rng = np.random.default_rng(5)

    def rand_tt(rank):
        rks = [min(2 ** k, 2 ** (d - k), rank) for k in range(d + 1)]
        return t.TTvector(d, [np.asfortranarray(rng.standard_normal((2, rks[k], rks[k + 1])) / np.sqrt(2 * rks[k + 1])) for k in range(d)],
                          (2,) * d, rks)

    return A, dt, rand_tt(rmax), rand_tt(start_rank)


def run_heat(bits=20, rmax=128, tol=1e-10, maxiter=20, krylovdim=40, start_rank=64, dt_scale=100.0, calls=3):
    """cfg3 at its stated size with ranks that actually reach the cap: `mals_linsolve(I + dt L, b, x0; tol, rmax)` on the interleaved
    2 x `bits`-bit grid with b = (I + dt L) x_true, x_true a random train of rank `rmax`, x0 a random train of rank `start_rank`;
    the call is repeated on its own result (each call = one forward + one backward two-site sweep, mals.jl:240-309) and the error
    against x_true is reported per call.  Windows of n^2 r^2 = 65 536 unknowns are solved matrix-free (GMRES on the three-GEMM
    matvec); the dense local matrix of mals.jl:148-169 would be 34 GB.  bench.py times the NumPy port of the
    reference algorithm (dense local K, so only at small `rmax`) on the inputs `heat_problem` returns."""
    d = 2 * bits
    A, dt, xt, x0 = heat_problem(bits, rmax, start_rank, dt_scale)
    Ad, xtd, xd = t.DeviceTTO.upload(A), t.DeviceTT.upload(xt), t.DeviceTT.upload(x0)
    bd = t.apply(Ad, xtd)
    nb, nxt = t.norm(bd), t.norm(xtd)
    t.synchronize()
    per_call = []
    for c in range(calls):
        t.reset_launch_count()
        t.set_option("reset_flops", 1)
        t.synchronize()
        t0 = time.perf_counter()
        xd = t.mals_linsolve(Ad, bd, xd, tol=tol, rmax=rmax, linsolv_maxiter=maxiter, krylovdim=krylovdim)
        t.synchronize()
        el = time.perf_counter() - t0
        gflop = t.get_option("gemm_flops") / 1e9
        res = t.norm(t.sub(t.apply(Ad, xd), bd)) / nb
        err = t.norm(t.sub(xd, xtd)) / nxt
        per_call.append({"s": el, "relative_residual": float(res), "relative_error": float(err), "max_rank": int(max(xd.ttv_rks)),
                         "ranks_at_cap": int(sum(1 for v in xd.ttv_rks if v == rmax)), "gpu_launches": int(t.launch_count()),
                         "gemm_gflop": gflop, "gemm_tflops": gflop / 1e3 / el})
        if res < 10 * tol:
            break
    out = {"metric": "cfg3 (implicit-Euler heat step) mals_linsolve s per call", "value": per_call[0]["s"], "unit": "s", "bits": bits, "d": d,
           "rmax": rmax, "tol": tol, "dt": dt, "condition_number_about": dt_scale, "mpo_rank": int(max(A.tto_rks)),
           "rhs_rank": int(max(bd.ttv_rks)), "start_rank": start_rank, "calls": per_call,
           "note": "b = (I + dt L) x_true with a random rank-rmax x_true; every call is one forward + one backward two-site sweep; local "
                   "systems of up to n^2 rmax^2 unknowns solved matrix-free by GMRES"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bits", type=int, default=20)
    ap.add_argument("--rmax", type=int, default=128)
    ap.add_argument("--tol", type=float, default=1e-12)
    ap.add_argument("--maxiter", type=int, default=200)
    ap.add_argument("--krylovdim", type=int, default=30)
    ap.add_argument("--x0rank", type=int, default=8)
    ap.add_argument("--heat", action="store_true", help="the well-conditioned implicit-Euler variant whose ranks reach the cap")
    ap.add_argument("--start-rank", type=int, default=64)
    args = ap.parse_args()
    if args.heat:
        print(json.dumps(run_heat(args.bits, args.rmax, start_rank=args.start_rank)), flush=True)
        return
    print(json.dumps(run(args.bits, args.rmax, args.tol, args.maxiter, args.krylovdim, args.x0rank, profile=True)), flush=True)


if __name__ == "__main__":
    main()
