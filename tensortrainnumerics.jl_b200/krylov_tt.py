"""`krylov_linsolve` of src/solvers/euler.jl:34-74 on device-resident trains.

The reference hands TT vectors to KrylovKit through its VectorInterface extension
(ext/TensorTrainNumericsVectorInterfaceExt/...jl:11-110): every Krylov vector operation is TT arithmetic — `add` is
`beta*y + alpha*x` followed by `tt_compress!(., max_bond)` (or `orthogonalize` when `max_bond == 0`), `scale` is
`orthogonalize(alpha*x)`, `inner` is `dot` — and the operator is `x -> tt_compress!(A*x, max_bond)` (euler.jl:56).
KrylovKit itself is a third-party dependency that is not part of the reference tree, so the three
solvers it is asked for (GMRES, BiCGStab, CG; `_krylov_algorithm`, euler.jl:9-32) are restated here from their textbook
form; parity is on the converged solution.  Every train stays in HBM; each vector operation is one or two C-ABI calls.
"""
from __future__ import annotations

import math

import numpy as np

from . import api as _a


class _Ops:
    """the VectorInterface operations of the reference's extension, on DeviceTT"""

    def __init__(self, A, max_bond):
        self.A = A if isinstance(A, _a.DeviceTTO) else _a.DeviceTTO.upload(A)
        self.max_bond = int(max_bond)

    def rnd(self, x):                         # `_round`, VectorInterfaceExt:11-14
        return _a.tt_compress_(x, self.max_bond) if self.max_bond > 0 else _a.orthogonalize(x)

    def op(self, x):                          # euler.jl:56
        y = _a.apply(self.A, x)
        return _a.tt_compress_(y, self.max_bond) if self.max_bond > 0 else y

    def axpby(self, alpha, x, beta, y):       # `add(y, x, alpha, beta)` = round(beta*y + alpha*x), VectorInterfaceExt:31-33
        return self.rnd(_a.add(_a.scale(beta, y), _a.scale(alpha, x)))

    def scale(self, x, alpha):                # VectorInterfaceExt:69-71
        return _a.orthogonalize(_a.scale(alpha, x))

    dot = staticmethod(_a.dot)
    norm = staticmethod(_a.norm)


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
    """euler.jl:34-74.  `krylov_solver` in {"auto", "gmres", "bicgstab", "cg"} (Julia symbols :auto, ...)."""
    if ishermitian is None:
        ishermitian = issymmetric
    solver = str(krylov_solver).lstrip(":")
    if solver == "auto" and isposdef and (issymmetric or ishermitian):
        solver = "cg"                                      # euler.jl:57
    if solver == "auto":
        solver = "bicgstab" if max_bond > 0 else "gmres"   # euler.jl:17
    if solver not in ("gmres", "bicgstab", "cg"):
        raise ValueError(f"Unknown Krylov solver: {krylov_solver}. Use :auto, :bicgstab, :cg, or :gmres.")
    ops = _Ops(A, max_bond)
    bd, host = _a._dev(b)
    xd, _ = _a._dev(guess)
    bd, xd = _a._match(bd, xd)
    tol_value = max(atol, rtol * ops.norm(bd)) if tol is None else float(tol)
    if solver == "gmres":
        x = _gmres(ops, bd, xd, krylovdim, maxiter, tol_value)
    elif solver == "cg":
        x = _cg(ops, bd, xd, krylovdim * maxiter, tol_value)      # euler.jl:29
    else:
        x = _bicgstab(ops, bd, xd, maxiter, tol_value)
    return _a._ret(x, host)
