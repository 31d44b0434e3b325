"""Time steppers of src/solvers/euler.jl on device-resident trains.

`euler_method` (euler.jl:76-97), `implicit_euler_method` (:99-143), `crank_nicholson_method` (:145-192) and `rk4_method`
(:194-222) are loops of hot-path calls — `A*x`, `+`, scalar `*`, `orthogonalize`, `tt_compress!`, `dot`/`norm` and one
linear TT solve per step — so they are compositions of the C-ABI entry points; every train stays in HBM between the calls
(arguments are uploaded once, the result is downloaded once).  The operator algebra `I ± h·A` acts on the MPO cores on
the host (a few KB) exactly as `+`/scalar `*` of src/tt_operations.jl:71-96,268-278 do.
"""
from __future__ import annotations

import math

import numpy as np

from . import api as _a


# ---- TToperator algebra on the host (containers only; src/tt_operations.jl:71-96, 268-278, tt_operators.jl:524-532) ----
def id_tto(d, dtype=np.float64, n=2):
    c = np.zeros((n, n, 1, 1), dtype=dtype)
    c[:, :, 0, 0] = np.eye(n)
    return _a.TToperator(d, [c.copy() for _ in range(d)], (n,) * d, [1] * (d + 1))


def tto_scale(a, A):
    cores = [np.array(c, copy=True) for c in A.tto_vec]
    dt = np.result_type(cores[0].dtype, np.asarray(a).dtype)
    cores = [c.astype(dt) for c in cores]
    cores[0] = cores[0] * a                       # the scalar multiplies the first core (tt_operations.jl:268-278)
    return _a.TToperator(A.N, cores, A.tto_dims, A.tto_rks)


def tto_add(x, y):
    assert tuple(x.tto_dims) == tuple(y.tto_dims), "Incompatible dimensions"
    d = x.N
    dt = np.result_type(x.tto_vec[0].dtype, y.tto_vec[0].dtype)
    cores, rks = [], [1]
    for k in range(d):
        a, b = x.tto_vec[k], y.tto_vec[k]
        n = a.shape[0]
        rl = 1 if k == 0 else a.shape[2] + b.shape[2]
        rr = 1 if k == d - 1 else a.shape[3] + b.shape[3]
        c = np.zeros((n, n, rl, rr), dtype=dt)
        if d == 1:
            c[:] = a + b
        elif k == 0:
            c[:, :, 0, :a.shape[3]] = a[:, :, 0, :]
            c[:, :, 0, a.shape[3]:] = b[:, :, 0, :]
        elif k == d - 1:
            c[:, :, :a.shape[2], 0] = a[:, :, :, 0]
            c[:, :, a.shape[2]:, 0] = b[:, :, :, 0]
        else:
            c[:, :, :a.shape[2], :a.shape[3]] = a
            c[:, :, a.shape[2]:, a.shape[3]:] = b
        cores.append(np.asfortranarray(c))
        rks.append(rr)
    return _a.TToperator(d, cores, x.tto_dims, rks)


def _shifted(A, alpha):
    """I + alpha*A as a TToperator (host cores)"""
    return tto_add(id_tto(A.N, A.tto_vec[0].dtype, A.tto_dims[0]), tto_scale(alpha, A))


def _lin(tt_solver, M, rhs, guess, kw):
    if tt_solver == "mals":
        return _a.mals_linsolve(M, rhs, guess, **kw)
    if tt_solver == "als":
        return _a.als_linsolve(M, rhs, guess, **kw)
    if tt_solver == "dmrg":
        return _a.dmrg_linsolve(M, rhs, guess, **kw)
    if tt_solver == "krylov":
        from .krylov_tt import krylov_linsolve
        return krylov_linsolve(M, rhs, guess, **kw)
    raise ValueError(f"Unknown TT solver: {tt_solver}")


def _host_op(A):
    if isinstance(A, _a.DeviceTTO):
        raise TypeError("the time steppers build I ± h·A from the host cores: pass a TToperator")
    return A


def euler_method(A, u0, steps, normalize=True, return_error=False):
    """euler.jl:76-97."""
    A = _host_op(A)
    Ad = _a.DeviceTTO.upload(A)
    sol, host = _a._dev(u0)
    for h in steps:
        upd = _a.apply(Ad, sol)
        sol = _a.orthogonalize(_a.add(sol, _a.scale(h, upd)))
        if normalize:
            n2 = _a.dot(sol, sol)
            sol = _a.scale(1.0 / math.sqrt(abs(n2)), sol)
    if return_error:
        h = steps[-1]
        res = _a.sub(sol, _a.apply(_a.DeviceTTO.upload(_shifted(A, h)), sol))
        return _a._ret(sol, host), _a.norm(res) / _a.norm(sol)
    return _a._ret(sol, host)


def _implicit(A, u0, guess, steps, normalize, return_error, tt_solver, max_bond, kw, theta):
    """theta = 1: implicit Euler (euler.jl:99-143); theta = 1/2: Crank-Nicolson (euler.jl:145-192)."""
    A = _host_op(A)
    sol, host = _a._dev(u0)
    guess, _ = _a._dev(guess)
    prev = sol
    for h in steps:
        M = _a.DeviceTTO.upload(_shifted(A, -theta * h))
        rhs = sol if theta == 1.0 else _a.apply(_a.DeviceTTO.upload(_shifted(A, (1.0 - theta) * h)), sol)
        nxt = _lin(tt_solver, M, rhs, guess, dict(kw, max_bond=max_bond) if tt_solver == "krylov" else kw)
        if normalize:
            nxt = _a.scale(1.0 / _a.norm(nxt), nxt)
        prev = sol
        sol = _a.tt_compress_(nxt, max_bond) if max_bond > 0 else _a.orthogonalize(nxt)
        guess = sol
    if return_error:
        h = steps[-1]
        M = _a.DeviceTTO.upload(_shifted(A, -theta * h))
        rhs = prev if theta == 1.0 else _a.apply(_a.DeviceTTO.upload(_shifted(A, (1.0 - theta) * h)), prev)
        res = _a.sub(_a.apply(M, sol), rhs)
        return _a._ret(sol, host), _a.norm(res) / _a.norm(sol)
    return _a._ret(sol, host)


def implicit_euler_method(A, u0, guess, steps, normalize=True, return_error=False, tt_solver="mals", max_bond=0, **kw):
    """euler.jl:99-143: (I - h A) u_{n+1} = u_n by one TT linear solve per step."""
    return _implicit(A, u0, guess, steps, normalize, return_error, tt_solver, max_bond, kw, 1.0)


def crank_nicholson_method(A, u0, guess, steps, normalize=True, return_error=False, tt_solver="mals", max_bond=0, **kw):
    """euler.jl:145-192: (I - h/2 A) u_{n+1} = (I + h/2 A) u_n."""
    return _implicit(A, u0, guess, steps, normalize, return_error, tt_solver, max_bond, kw, 0.5)


def rk4_method(A, u0, steps, max_bond, normalize=True, return_error=False):
    """euler.jl:194-222 (every intermediate stage is rounded to `max_bond`)."""
    Ad = A if isinstance(A, _a.DeviceTTO) else _a.DeviceTTO.upload(A)
    u, host = _a._dev(u0)

    def rnd(x):
        return _a.tt_compress_(x, max_bond)

    def incr_of(u, h):
        k1 = _a.apply(Ad, u)
        k2 = _a.apply(Ad, rnd(_a.add(u, _a.scale(h / 2, k1))))
        k3 = _a.apply(Ad, rnd(_a.add(u, _a.scale(h / 2, k2))))
        k4 = _a.apply(Ad, rnd(_a.add(u, _a.scale(h, k3))))
        s = _a.add(_a.add(_a.add(k1, _a.scale(2.0, k2)), _a.scale(2.0, k3)), k4)
        return _a.scale(h / 6, rnd(s))

    for h in steps:
        un = rnd(_a.add(u, incr_of(u, h)))
        if normalize:
            un = _a.scale(1.0 / math.sqrt(abs(_a.dot(un, un))), un)
        u = un
    if return_error:
        incr = incr_of(u, steps[-1])
        res = rnd(_a.sub(_a.sub(u, _a.sub(u, incr)), incr))
        return _a._ret(u, host), _a.norm(res) / max(_a.norm(u), np.finfo(float).eps)
    return _a._ret(u, host)

