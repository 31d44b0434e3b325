"""Oracle restatement of the time steppers of src/solvers/euler.jl (test infrastructure only).

euler_method :76-97, implicit_euler_method :99-143, crank_nicholson_method :145-192, rk4_method :194-222, each a loop of
the oracle's own `apply`, `add`, `scale`, `orthogonalize`, `tt_compress`, `dot`/`norm` and linear TT solvers.
"""
from __future__ import annotations

import numpy as np

from .core import copy_tt
from .ops import add, scale, sub, dot, norm, apply, orthogonalize, tt_compress
from .generators import id_tto, tto_add, tto_scale
from .als import als_linsolve
from .mals import mals_linsolve
from .dmrg import dmrg_linsolve


def _shifted(A, alpha):
    return tto_add(id_tto(A.N, dtype=A.dtype), tto_scale(alpha, A))


def _lin(tt_solver, M, rhs, guess, kw):
    from .krylov_tt import krylov_linsolve
    fn = {"mals": mals_linsolve, "als": als_linsolve, "dmrg": dmrg_linsolve, "krylov": krylov_linsolve}.get(tt_solver)
    if fn is None:
        raise ValueError(f"Unknown TT solver: {tt_solver}")
    return fn(M, rhs, guess, **kw)


def euler_method(A, u0, steps, normalize=True, return_error=False):
    sol = u0
    for h in steps:
        upd = apply(A, sol)
        sol = orthogonalize(add(sol, scale(h, upd)))
        if normalize:
            sol = scale(1.0 / np.sqrt(abs(dot(sol, sol))), sol)
    if return_error:
        h = steps[-1]
        res = sub(sol, apply(_shifted(A, h), sol))
        return sol, norm(res) / norm(sol)
    return sol


def _implicit(A, u0, guess, steps, normalize, return_error, tt_solver, max_bond, kw, theta):
    sol, prev = u0, u0
    for h in steps:
        M = _shifted(A, -theta * h)
        rhs = sol if theta == 1.0 else apply(_shifted(A, (1.0 - theta) * h), sol)
        nxt = _lin(tt_solver, M, rhs, guess, dict(kw, max_bond=max_bond) if tt_solver == "krylov" else kw)
        if normalize:
            nxt = scale(1.0 / norm(nxt), nxt)
        prev = sol
        sol = tt_compress(copy_tt(nxt), max_bond) if max_bond > 0 else orthogonalize(nxt)
        guess = sol
    if return_error:
        h = steps[-1]
        M = _shifted(A, -theta * h)
        rhs = prev if theta == 1.0 else apply(_shifted(A, (1.0 - theta) * h), prev)
        res = sub(apply(M, sol), rhs)
        return sol, norm(res) / norm(sol)
    return sol


def implicit_euler_method(A, u0, guess, steps, normalize=True, return_error=False, tt_solver="mals", max_bond=0, **kw):
    return _implicit(A, u0, guess, steps, normalize, return_error, tt_solver, max_bond, kw, 1.0)


def crank_nicholson_method(A, u0, guess, steps, normalize=True, return_error=False, tt_solver="mals", max_bond=0, **kw):
    return _implicit(A, u0, guess, steps, normalize, return_error, tt_solver, max_bond, kw, 0.5)


def rk4_method(A, u0, steps, max_bond, normalize=True, return_error=False):
    def rnd(x):
        return tt_compress(copy_tt(x), max_bond)

    def incr_of(u, h):
        k1 = apply(A, u)
        k2 = apply(A, rnd(add(u, scale(h / 2, k1))))
        k3 = apply(A, rnd(add(u, scale(h / 2, k2))))
        k4 = apply(A, rnd(add(u, scale(h, k3))))
        return scale(h / 6, rnd(add(add(add(k1, scale(2.0, k2)), scale(2.0, k3)), k4)))

    u = u0
    for h in steps:
        un = rnd(add(u, incr_of(u, h)))
        if normalize:
            un = scale(1.0 / np.sqrt(abs(dot(un, un))), un)
        u = un
    if return_error:
        incr = incr_of(u, steps[-1])
        res = rnd(sub(sub(u, sub(u, incr)), incr))
        return u, norm(res) / max(norm(u), np.finfo(float).eps)
    return u
