"""TDVP (oracle; test infrastructure only).  Follows src/solvers/tdvp.jl.

tdvp.jl:29-43 (_applyH1_lsr, _applyH0, env updates), :45-152 (tdvp1sweep!), :154-203 (tdvp),
:205-208 (_applyH2_lsr), :210-301 (tdvp2sweep!), :303-357 (tdvp2).

KrylovKit.exponentiate (tdvp.jl:75,95,…) is third-party and absent from the reference tree; the oracle
evaluates exp(t·H_loc)·v exactly (dense `expm` of the assembled local operator), i.e. the converged
limit of the Krylov exponential.  Layouts: cores (l,s,r), MPO (a, s_out, b, s_in), FL[bra,mpo,ket],
FR[ket,mpo,bra] (tdvp.jl:54-55,29-43).

With `imaginary_time=True` the reference promotes to ComplexF64 internally but every number stays
real (t = +h is real, tdvp.jl:74); the oracle keeps real arithmetic in that case.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from .core import TTvector, TToperator, complex_tt, complex_tto, copy_tt
from .ops import orthogonalize, norm, scale, svdtrunc, apply, sub


def apply_H1_lsr(AC, FL, FR, M):
    """tdvp.jl:29-31."""
    return np.einsum("xay,ytz,asbt,zbw->xsw", FL, AC, M, FR, optimize=True)


def apply_H0(C, FL, FR):
    """tdvp.jl:33-35."""
    return np.einsum("xay,yz,zaw->xw", FL, C, FR, optimize=True)


def apply_H2_lsr(AAC, FL, FR, M1, M2):
    """tdvp.jl:205-208."""
    return np.einsum("xay,ytuz,asbt,bvcu,zcw->xsvw", FL, AAC, M1, M2, FR, optimize=True)


def update_left_env(A, M, FL):
    """tdvp.jl:37-39: FLnext[α,a,β] = FL[α′,a′,β′] A[β′,s′,β] M[a′,s,a,s′] conj(A[α′,s,α])."""
    return np.einsum("xpy,ytb,psat,xsc->cab", FL, A, M, np.conj(A), optimize=True)


def update_right_env(A, M, FR):
    """tdvp.jl:41-43: FRprev[α,a,β] = A[α,s′,α′] FR[α′,a′,β′] M[a,s,a′,s′] conj(A[β,s,β′])."""
    return np.einsum("ctx,xpy,aspt,bsy->cab", A, FR, M, np.conj(A), optimize=True)


def _expm_apply(fun, t, v):
    """exp(t·H)·v for the linear map `fun` (dense assembly; local problems in the oracle tests are small)."""
    shape = v.shape
    n = v.size
    K = np.empty((n, n), dtype=v.dtype)
    e = np.zeros(n, dtype=v.dtype)
    for j in range(n):
        e[:] = 0
        e[j] = 1
        K[:, j] = np.reshape(fun(np.reshape(e, shape, order="F")), -1, order="F")
    out = sla.expm(t * K) @ np.reshape(v, -1, order="F")
    return np.reshape(out, shape, order="F")


def _roc(z):
    """tdvp.jl:22 `_real_or_complex_t`."""
    return float(np.real(z)) if np.imag(z) == 0 else complex(z)


def _build_envs(A_lsr, M_asbs, Tc):
    N = len(A_lsr)
    F = [None] * (N + 2)
    F[0] = np.ones((1, 1, 1), dtype=Tc)
    F[N + 1] = np.ones((1, 1, 1), dtype=Tc)
    for k in range(N, 0, -1):
        F[k] = update_right_env(A_lsr[k - 1], M_asbs[k - 1], F[k + 1])
    return F


def tdvp1sweep(dt, psi: TTvector, H: TToperator, F=None):
    """tdvp.jl:45-152.  `dt` is the reference's dt_eff (complex: +i·h for imaginary time, h for real)."""
    N = psi.N
    Tc = np.result_type(psi.dtype, H.dtype, type(_roc(-1j * dt)))
    A = [np.transpose(c, (1, 0, 2)).astype(Tc) for c in psi.ttv_vec]
    M = [np.transpose(c, (2, 0, 3, 1)) for c in H.tto_vec]
    if F is None:
        F = _build_envs(A, M, Tc)
    AC = A[0]
    t1, t0 = _roc(-1j * dt), _roc(+1j * dt)
    for k in range(1, N):
        AC = _expm_apply(lambda x: apply_H1_lsr(x, F[k - 1], F[k + 1], M[k - 1]), t1, AC)
        Dl, d, Dr = AC.shape
        Q, R = sla.qr(np.reshape(AC, (Dl * d, Dr), order="F"), mode="economic")
        r = min(Dl * d, Dr)
        A[k - 1] = np.reshape(Q[:, :r], (Dl, d, r), order="F")
        F[k] = update_left_env(A[k - 1], M[k - 1], F[k - 1])
        C = R[:r, :]
        C = _expm_apply(lambda X: apply_H0(X, F[k], F[k + 1]), t0, C)
        AC = np.einsum("ag,gsb->asb", C, A[k])
    k = N
    AC = _expm_apply(lambda x: apply_H1_lsr(x, F[k - 1], F[k + 1], M[k - 1]), t1, AC)
    for k in range(N - 1, 0, -1):
        Dl, d, Dr = AC.shape
        Am = np.reshape(AC, (Dl, d * Dr), order="F")
        Q, R = sla.qr(Am.conj().T, mode="economic")
        r = min(Dl, d * Dr)
        L = R[:r, :].conj().T
        A[k] = np.reshape(Q[:, :r].conj().T, (r, d, Dr), order="F")
        F[k + 1] = update_right_env(A[k], M[k], F[k + 2])
        C = _expm_apply(lambda X: apply_H0(X, F[k], F[k + 1]), t0, L)
        AC = np.einsum("asg,gb->asb", A[k - 1], C)
        AC = _expm_apply(lambda x: apply_H1_lsr(x, F[k - 1], F[k + 1], M[k - 1]), t1, AC)
    A[0] = AC
    for k in range(N):
        psi.ttv_vec[k] = np.ascontiguousarray(np.transpose(A[k], (1, 0, 2)))
    psi.ttv_rks = [a.shape[0] for a in A] + [A[-1].shape[2]]
    psi.ttv_ot = [0] * N
    return psi, F


def tdvp2sweep(dt, psi: TTvector, H: TToperator, F=None, max_bond=None, truncerr=0.0):
    """tdvp.jl:210-301."""
    N = psi.N
    Tc = np.result_type(psi.dtype, H.dtype, type(_roc(-1j * dt)))
    dth = dt / 2
    A = [np.transpose(c, (1, 0, 2)).astype(Tc) for c in psi.ttv_vec]
    M = [np.transpose(c, (2, 0, 3, 1)) for c in H.tto_vec]
    if F is None:
        F = _build_envs(A, M, Tc)
    AC = A[0]
    t2, t1 = _roc(-1j * dth), _roc(+1j * dth)
    for k in range(1, N):
        AAC = np.einsum("asg,gtb->astb", AC, A[k])
        AAC = _expm_apply(lambda X: apply_H2_lsr(X, F[k - 1], F[k + 2], M[k - 1], M[k]), t2, AAC)
        Dl, d1, d2, Dr = AAC.shape
        U, s, Vt = svdtrunc(np.reshape(AAC, (Dl * d1, d2 * Dr), order="F"), max_bond=max_bond, truncerr=truncerr)
        A[k - 1] = np.reshape(U, (Dl, d1, U.shape[1]), order="F")
        F[k] = update_left_env(A[k - 1], M[k - 1], F[k - 1])
        AC = np.reshape(s[:, None] * Vt, (len(s), d2, Dr), order="F")
        if k < N - 1:
            AC = _expm_apply(lambda x: apply_H1_lsr(x, F[k], F[k + 2], M[k]), t1, AC)
    for k in range(N - 1, 0, -1):
        AAC = np.einsum("asg,gtb->astb", A[k - 1], AC)
        AAC = _expm_apply(lambda X: apply_H2_lsr(X, F[k - 1], F[k + 2], M[k - 1], M[k]), t2, AAC)
        Dl, d1, d2, Dr = AAC.shape
        U, s, Vt = svdtrunc(np.reshape(AAC, (Dl * d1, d2 * Dr), order="F"), max_bond=max_bond, truncerr=truncerr)
        A[k] = np.reshape(Vt, (Vt.shape[0], d2, Dr), order="F")
        F[k + 1] = update_right_env(A[k], M[k], F[k + 2])
        AC = np.reshape(U * s[None, :], (Dl, d1, len(s)), order="F")
        if k > 1:
            AC = _expm_apply(lambda x: apply_H1_lsr(x, F[k - 1], F[k + 1], M[k - 1]), t1, AC)
    A[0] = AC
    for k in range(N):
        psi.ttv_vec[k] = np.ascontiguousarray(np.transpose(A[k], (1, 0, 2)))
    psi.ttv_rks = [a.shape[0] for a in A] + [A[-1].shape[2]]
    psi.ttv_ot = [0] * N
    return psi, F


def _drive(sweep, H, u0, steps, normalize, sweeps, imaginary_time, return_error, **kw):
    psi = orthogonalize(u0)
    if not imaginary_time:
        psi = complex_tt(psi)
        Hc = complex_tto(H)
    else:
        Hc = H
    psi_prev = psi
    for h in steps:
        psi_prev_step = copy_tt(psi)
        dt_eff = (1j * h) if imaginary_time else complex(h)
        F = None
        for _ in range(sweeps):
            psi, F = sweep(dt_eff, psi, Hc, F, **kw)
        if normalize:
            psi = scale(1.0 / norm(psi), psi)
        psi = orthogonalize(psi)
        psi_prev = psi_prev_step
    if return_error:
        h = steps[-1]
        if imaginary_time:
            res = sub(scale(1.0 / h, sub(psi, psi_prev)), apply(Hc, psi))
        else:
            res = sub(scale(1.0 / h, sub(psi, psi_prev)), scale(-1j, apply(Hc, psi)))
        return psi, norm(res) / norm(psi)
    return psi


def tdvp(H, u0, steps, normalize=True, sweeps=1, imaginary_time=False, return_error=False):
    """tdvp.jl:154-203."""
    return _drive(tdvp1sweep, H, u0, steps, normalize, sweeps, imaginary_time, return_error)


def tdvp2(H, u0, steps, normalize=True, sweeps=1, max_bond=None, truncerr=0.0, imaginary_time=False,
          return_error=False):
    """tdvp.jl:303-357."""
    return _drive(tdvp2sweep, H, u0, steps, normalize, sweeps, imaginary_time, return_error,
                  max_bond=max_bond, truncerr=truncerr)
