"""MALS (oracle; test infrastructure only).  Follows src/solvers/mals.jl line by line.

mals.jl:10-40 (5-index right environments), :42-56 (sv_trunc), :60-92 (RHS environments),
:94-146 (SVD core moves), :148-169 (dense K + `\\`), :171-218 (K_eigmin_mals),
:240-309 (mals_linsolve: exactly one forward and one backward sweep), :335-425 (mals_eigsolve).
Rank-padded buffers + views in the reference are replaced by exactly-sized arrays (same values).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from .core import TTvector, TToperator
from .ops import orthogonalize, apply, sub, norm
from .als import update_G, update_Gb


def updateH_mals(x_vec, A_vec, Hi):
    """mals.jl:10-13: Him[a,i,α,l,β] = conj(x)[j,α,x] (Hi[z,j,x,k,y] x[k,β,y]) A[i,l,a,z]."""
    return np.einsum("jax,zjxky,kby,iluz->uialb", np.conj(x_vec), Hi, x_vec, A_vec, optimize=True)


def init_H_mals(x: TTvector, A: TToperator):
    """mals.jl:15-40.  H[i] (1-based i = 1..d-1) = right environment of the window (i, i+1) already
    contracted with A_{i+1}: shape (R_i, n_{i+1}, r_{i+1}, n_{i+1}, r_{i+1})."""
    d = x.N
    H = [None] * (d - 1)
    H[d - 2] = np.reshape(np.transpose(A.tto_vec[d - 1], (2, 0, 1, 3)),
                          (-1, x.ttv_dims[d - 1], 1, x.ttv_dims[d - 1], 1), order="F").astype(x.dtype)
    for i in range(d - 1, 1, -1):  # 1-based i = d-1 .. 2
        H[i - 2] = updateH_mals(x.ttv_vec[i], A.tto_vec[i - 1], H[i - 1])
    return H


def sv_trunc(s, tol):
    """mals.jl:42-56 (squared-weight rule; the element that crosses the threshold is kept)."""
    s = np.asarray(s)
    if tol == 0.0:
        return s
    d = len(s)
    i = 0
    weight = 0.0
    norm2 = float(np.sum(np.abs(s) ** 2))
    while i < d and weight < tol * norm2:
        weight += s[d - i - 1] ** 2
        i += 1
    return s[: d - i + 1]


def updateHb_mals(x_vec, b_vec, Hbi):
    """mals.jl:60-66: Hbim[β,i,χ] = conj(x)[j,χ,a] Hbi[γ,j,a] b[i,β,γ]."""
    return np.einsum("jca,gja,ibg->bic", np.conj(x_vec), Hbi, b_vec, optimize=True)


def init_Hb_mals(x: TTvector, b: TTvector):
    """mals.jl:68-92."""
    d = x.N
    Hb = [None] * (d - 1)
    Hb[d - 2] = np.reshape(np.transpose(b.ttv_vec[d - 1], (1, 0, 2)),
                           (b.ttv_rks[d - 1], b.ttv_dims[d - 1], 1), order="F").astype(x.dtype)
    for i in range(d - 1, 1, -1):
        Hb[i - 2] = updateHb_mals(x.ttv_vec[i], b.ttv_vec[i - 1], Hb[i - 1])
    return Hb


def left_core_move_mals(x: TTvector, i, V, tol, rmax):
    """mals.jl:94-119 (i 1-based): core i+1 <- Vt (right-orthogonal), core i <- U·S."""
    u, s, vh = sla.svd(np.reshape(V, (V.shape[0] * V.shape[1], -1), order="F"),
                       full_matrices=False, lapack_driver="gesdd")
    st = sv_trunc(s, tol)
    r = min(len(st), rmax)
    x.ttv_rks[i] = r
    x.ttv_vec[i] = np.ascontiguousarray(
        np.transpose(np.reshape(vh[:r, :], (r, V.shape[2], V.shape[3]), order="F"), (1, 0, 2)))
    x.ttv_vec[i - 1] = np.ascontiguousarray(
        np.reshape(u[:, :r] * st[None, :r], (V.shape[0], V.shape[1], -1), order="F"))
    x.ttv_ot[i] = 1
    x.ttv_ot[i - 1] = 0
    return x


def right_core_move_mals(x: TTvector, i, V, tol, rmax):
    """mals.jl:121-146 (i 1-based): core i <- U (left-orthogonal), core i+1 <- S·Vt."""
    u, s, vh = sla.svd(np.reshape(V, (V.shape[0] * V.shape[1], -1), order="F"),
                       full_matrices=False, lapack_driver="gesdd")
    st = sv_trunc(s, tol)
    r = min(len(st), rmax)
    x.ttv_rks[i] = r
    x.ttv_vec[i - 1] = np.ascontiguousarray(np.reshape(u[:, :r], (V.shape[0], V.shape[1], r), order="F"))
    x.ttv_ot[i - 1] = -1
    x.ttv_vec[i] = np.ascontiguousarray(
        np.transpose(np.reshape(st[:r, None] * vh[:r, :], (r, V.shape[2], V.shape[3]), order="F"), (1, 0, 2)))
    x.ttv_ot[i] = 0
    return x


def K_full_mals(Gi, Hi):
    """mals.jl:148-157: K[(a,b,c,d),(e,f,g,h)] = Σ_z G[a,b,e,f,z] H[z,c,d,g,h]."""
    dims = (Gi.shape[0], Gi.shape[1], Hi.shape[1], Hi.shape[2])
    K8 = np.einsum("abefz,zcdgh->abcdefgh", Gi, Hi)
    n = int(np.prod(dims))
    return np.reshape(K8, (n, n), order="F"), dims


def Ksolve_mals(Gi, Hi, G_bi, H_bi):
    """mals.jl:159-169."""
    K, dims = K_full_mals(Gi, Hi)
    Pb = np.einsum("abz,zcd->abcd", G_bi, H_bi)
    V = np.linalg.solve(K, np.reshape(Pb, -1, order="F"))
    return np.reshape(V, dims, order="F")


def K_eigmin_mals(Gi, Hi):
    """mals.jl:171-218, dense branch (`eigen(Hermitian(K), 1:1)`)."""
    K, dims = K_full_mals(Gi, Hi)
    w, v = sla.eigh(K, lower=False, subset_by_index=[0, 0])
    return float(np.real(w[0])), np.reshape(v[:, 0], dims, order="F")


def _residual(A, x, b):
    return norm(sub(apply(A, x), b)) / max(norm(b), np.finfo(float).eps)


def mals_linsolve(A: TToperator, b: TTvector, tt_start: TTvector, tol=1e-12, rmax=None, return_info=False):
    """mals.jl:240-309."""
    T = tt_start.dtype
    d = b.N
    dims = tt_start.ttv_dims
    if rmax is None:
        rmax = int(round(np.sqrt(float(np.prod([float(n) for n in dims])))))
    x = orthogonalize(tt_start)
    G = [None] * d
    Gb = [None] * d
    G[0] = np.reshape(A.tto_vec[0][:, :, 0, :], (dims[0], 1, dims[0], 1, -1), order="F").astype(T)
    Gb[0] = np.reshape(b.ttv_vec[0], (dims[0], 1, -1), order="F").astype(T)
    H = init_H_mals(x, A)
    Hb = init_Hb_mals(x, b)
    for i in range(1, d):
        V = Ksolve_mals(G[i - 1], H[i - 1], Gb[i - 1], Hb[i - 1])
        x = right_core_move_mals(x, i, V, tol, rmax)
        G[i] = update_G(x.ttv_vec[i - 1], A.tto_vec[i], G[i - 1])
        Gb[i] = update_Gb(x.ttv_vec[i - 1], b.ttv_vec[i], Gb[i - 1])
    for i in range(d - 1, 0, -1):
        V = Ksolve_mals(G[i - 1], H[i - 1], Gb[i - 1], Hb[i - 1])
        x = left_core_move_mals(x, i, V, tol, rmax)
        if i > 1:
            H[i - 2] = updateH_mals(x.ttv_vec[i], A.tto_vec[i - 1], H[i - 1])
            Hb[i - 2] = updateHb_mals(x.ttv_vec[i], b.ttv_vec[i - 1], Hb[i - 1])
    if return_info:
        return x, {"residual": _residual(A, x, b)}
    return x


def mals_eigsolve(A: TToperator, tt_start: TTvector, tol=1e-12, sweep_schedule=(2,), rmax_schedule=None):
    """mals.jl:335-425 with the dense local eigensolver."""
    d = A.N
    dims = tt_start.ttv_dims
    T = tt_start.dtype
    if rmax_schedule is None:
        rmax_schedule = [int(round(np.sqrt(float(np.prod([float(n) for n in dims])))))]
    assert len(rmax_schedule) == len(sweep_schedule), "Sweep schedule error"
    x = orthogonalize(tt_start)
    E, r_hist = [], []
    G = [None] * d
    G[0] = np.reshape(A.tto_vec[0][:, :, 0, :], (dims[0], 1, dims[0], 1, -1), order="F").astype(T)
    H = init_H_mals(x, A)
    nsweeps = 0
    i_sched = 1
    while i_sched <= len(sweep_schedule):
        nsweeps += 1
        if nsweeps == sweep_schedule[i_sched - 1]:
            i_sched += 1
            if i_sched > len(sweep_schedule):
                return np.array(E), x, r_hist
        for i in range(1, d):
            lam, V = K_eigmin_mals(G[i - 1], H[i - 1])
            E.append(lam)
            x = right_core_move_mals(x, i, V, tol, rmax_schedule[i_sched - 1])
            r_hist.append(max(x.ttv_rks))
            G[i] = update_G(x.ttv_vec[i - 1], A.tto_vec[i], G[i - 1])
        for i in range(d - 1, 0, -1):
            lam, V = K_eigmin_mals(G[i - 1], H[i - 1])
            E.append(lam)
            x = left_core_move_mals(x, i, V, tol, rmax_schedule[i_sched - 1])
            r_hist.append(max(x.ttv_rks))
            if i > 1:
                H[i - 2] = updateH_mals(x.ttv_vec[i], A.tto_vec[i - 1], H[i - 1])
    return np.array(E), x, r_hist
