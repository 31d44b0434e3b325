"""Oracle restatement of `krylov_linsolve` (src/solvers/euler.jl:34-74) with the TT vector operations of the reference's
VectorInterface extension (ext/TensorTrainNumericsVectorInterfaceExt/...jl:11-110).  Test infrastructure only.

KrylovKit (the third-party solver the reference calls) is not part of the reference tree; GMRES / CG / BiCGStab are restated
in textbook form, and parity with the CUDA path is on the converged solution.
"""
from __future__ import annotations

import math

import numpy as np

from .core import copy_tt
from .ops import add, scale, dot, norm, apply, orthogonalize, tt_compress


class _Ops:
    def __init__(self, A, max_bond):
        self.A, self.max_bond = A, int(max_bond)

    def rnd(self, x):
        return tt_compress(copy_tt(x), self.max_bond) if self.max_bond > 0 else orthogonalize(x)

    def op(self, x):
        y = apply(self.A, x)
        return tt_compress(y, self.max_bond) if self.max_bond > 0 else y

    def axpby(self, alpha, x, beta, y):
        return self.rnd(add(scale(beta, y), scale(alpha, x)))

    def scale(self, x, alpha):
        return orthogonalize(scale(alpha, x))

    dot = staticmethod(dot)
    norm = staticmethod(norm)


def _gmres(ops, b, x, krylovdim, maxiter, tol):
    for _ in range(max(1, maxiter)):
        r = ops.axpby(-1.0, ops.op(x), 1.0, b)
        beta = ops.norm(r)
        if beta <= tol:
            return x
        V = [ops.scale(r, 1.0 / beta)]
        m = max(1, krylovdim)
        H = np.zeros((m + 1, m), dtype=np.result_type(b.dtype, np.float64))
        y, k = None, 0
        for j in range(m):
            w = ops.op(V[j])
            for i in range(j + 1):            # modified Gram-Schmidt in TT arithmetic
                H[i, j] = ops.dot(V[i], w)
                w = ops.axpby(-H[i, j], V[i], 1.0, w)
            hn = ops.norm(w)
            H[j + 1, j] = hn
            k = j + 1
            e1 = np.zeros(k + 1, dtype=H.dtype); e1[0] = beta
            y, *_ = np.linalg.lstsq(H[:k + 1, :k], e1, rcond=None)
            res = np.linalg.norm(H[:k + 1, :k] @ y - e1)
            if res <= tol or hn <= 1e-14 * beta:
                break
            V.append(ops.scale(w, 1.0 / hn))
        for i in range(k):
            x = ops.axpby(y[i], V[i], 1.0, x)
    return x


def _cg(ops, b, x, maxiter, tol):
    r = ops.axpby(-1.0, ops.op(x), 1.0, b)
    p = r
    rs = ops.dot(r, r).real if np.iscomplexobj(ops.dot(r, r)) else float(ops.dot(r, r))
    for _ in range(max(1, maxiter)):
        if math.sqrt(abs(rs)) <= tol:
            break
        Ap = ops.op(p)
        alpha = rs / ops.dot(p, Ap)
        x = ops.axpby(alpha, p, 1.0, x)
        r = ops.axpby(-alpha, Ap, 1.0, r)
        rs_new = ops.dot(r, r)
        rs_new = rs_new.real if np.iscomplexobj(rs_new) else float(rs_new)
        p = ops.axpby(rs_new / rs, p, 1.0, r)
        rs = rs_new
    return x


def _bicgstab(ops, b, x, maxiter, tol):
    r = ops.axpby(-1.0, ops.op(x), 1.0, b)
    r0 = r
    rho = alpha = omega = 1.0
    v = p = None
    for it in range(max(1, maxiter)):
        if ops.norm(r) <= tol:
            break
        rho_new = ops.dot(r0, r)
        if it == 0:
            p = r
        else:
            beta = (rho_new / rho) * (alpha / omega)
            p = ops.axpby(beta, ops.axpby(-omega, v, 1.0, p), 1.0, r)
        v = ops.op(p)
        alpha = rho_new / ops.dot(r0, v)
        s = ops.axpby(-alpha, v, 1.0, r)
        if ops.norm(s) <= tol:
            x = ops.axpby(alpha, p, 1.0, x)
            break
        t = ops.op(s)
        omega = ops.dot(t, s) / ops.dot(t, t)
        x = ops.axpby(omega, s, 1.0, ops.axpby(alpha, p, 1.0, x))
        r = ops.axpby(-omega, t, 1.0, s)
        rho = rho_new
    return x


def krylov_linsolve(A, b, guess, max_bond=0, krylov_solver="auto", krylovdim=8, maxiter=20, rtol=1e-8, atol=1e-12, tol=None,
                    issymmetric=False, ishermitian=None, isposdef=False, verbosity=0):
    if ishermitian is None:
        ishermitian = issymmetric
    solver = str(krylov_solver).lstrip(":")
    if solver == "auto" and isposdef and (issymmetric or ishermitian):
        solver = "cg"
    if solver == "auto":
        solver = "bicgstab" if max_bond > 0 else "gmres"
    if solver not in ("gmres", "bicgstab", "cg"):
        raise ValueError(f"Unknown Krylov solver: {krylov_solver}. Use :auto, :bicgstab, :cg, or :gmres.")
    ops = _Ops(A, max_bond)
    tol_value = max(atol, rtol * ops.norm(b)) if tol is None else float(tol)
    if solver == "gmres":
        return _gmres(ops, b, guess, krylovdim, maxiter, tol_value)
    if solver == "cg":
        return _cg(ops, b, guess, krylovdim * maxiter, tol_value)
    return _bicgstab(ops, b, guess, maxiter, tol_value)
