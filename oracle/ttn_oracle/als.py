"""ALS (oracle; test infrastructure only).  Follows src/solvers/als.jl line by line.

als.jl:9-55 (environments), :58-70 (K_full / Ksolve: dense assembly + `\\`), :72-88 (K_eigmin),
:104-136 (QR core moves), :161-225 (als_linsolve), :251-321 (als_eigsolve).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from .core import TTvector, TToperator, increase_ranks
from .ops import orthogonalize, apply, sub, norm


def update_H(x_vec, A_vec, Hi):
    """als.jl:23-26: Him[a,α,β] = conj(x)[j,α,φ] Hi[z,φ,χ] x[k,β,χ] A[j,k,a,z]."""
    return np.einsum("jaf,zfc,kbc,jkuz->uab", np.conj(x_vec), Hi, x_vec, A_vec, optimize=True)


def init_H(x: TTvector, A: TToperator):
    """als.jl:9-21."""
    d = x.N
    H = [None] * d
    H[d - 1] = np.ones((1, 1, 1), dtype=x.dtype)
    for i in range(d - 1, 0, -1):
        H[i - 1] = update_H(x.ttv_vec[i], A.tto_vec[i], H[i])
    return H


def update_Hb(x_vec, b_vec, H_bi):
    """als.jl:42-45: H_bim[α,β] = H_bi[φ,χ] b[i,β,χ] conj(x)[i,α,φ]."""
    return np.einsum("fc,ibc,iaf->ab", H_bi, b_vec, np.conj(x_vec), optimize=True)


def init_Hb(x: TTvector, b: TTvector):
    """als.jl:28-40."""
    d = x.N
    Hb = [None] * d
    Hb[d - 1] = np.ones((1, 1), dtype=x.dtype)
    for i in range(d - 1, 0, -1):
        Hb[i - 1] = update_Hb(x.ttv_vec[i], b.ttv_vec[i], Hb[i])
    return Hb


def update_G(x_vec, A_vec, Gi):
    """als.jl:47-50: Gip[j,α,k,β,J] = conj(x)[l,φ,α] (Gi[l,φ,m,χ,L] x[m,χ,β]) A[j,k,L,J]."""
    return np.einsum("lfa,lfmcL,mcb,jkLJ->jakbJ", np.conj(x_vec), Gi, x_vec, A_vec, optimize=True)


def update_Gb(x_vec, b_vec, G_bi):
    """als.jl:52-55: G_bip[i,α,β] = b[i,φ,β] G_bi[j,χ,φ] conj(x)[j,χ,α]."""
    return np.einsum("ifb,jcf,jca->iab", b_vec, G_bi, np.conj(x_vec), optimize=True)


def K_full(Gi, Hi):
    """als.jl:58-63: K[(a,b,c),(d,e,f)] = Σ_z G[a,b,d,e,z] H[z,c,f]."""
    dims = (Gi.shape[0], Gi.shape[1], Hi.shape[1])
    K6 = np.einsum("abdez,zcf->abcdef", Gi, Hi)
    n = int(np.prod(dims))
    return np.reshape(K6, (n, n), order="F"), dims


def Ksolve(Gi, G_bi, Hi, H_bi):
    """als.jl:65-70."""
    K, dims = K_full(Gi, Hi)
    Pb = np.einsum("iab,cb->iac", G_bi, H_bi)
    V = np.linalg.solve(K, np.reshape(Pb, -1, order="F"))
    return np.reshape(V, dims, order="F")


def K_eigmin(Gi, Hi):
    """als.jl:72-88, dense branch (`eigen(Hermitian(K), 1:1)`; Hermitian() reads the upper triangle).
    The lobpcg branch converges to the same eigenpair; the oracle always uses the dense solve."""
    K, dims = K_full(Gi, Hi)
    w, v = sla.eigh(K, lower=False, subset_by_index=[0, 0])
    return float(np.real(w[0])), np.reshape(v[:, 0], dims, order="F")


def right_core_move(x: TTvector, V, i, rks):
    """als.jl:122-136 (i is 1-based)."""
    rim, ri = rks[i - 1], rks[i]
    ni = x.ttv_dims[i - 1]
    Q, R = sla.qr(np.reshape(V, (ni * rim, -1), order="F"), mode="full")
    x.ttv_vec[i - 1] = np.ascontiguousarray(np.reshape(Q[:, :ri], (ni, rim, -1), order="F"))
    x.ttv_ot[i - 1] = -1
    x.ttv_vec[i] = np.einsum("bz,azc->abc", R[:ri, :], x.ttv_vec[i])
    x.ttv_ot[i] = 0
    return x


def left_core_move(x: TTvector, V, i, rks):
    """als.jl:104-120 (i is 1-based)."""
    rim, ri = rks[i - 1], rks[i]
    ni = x.ttv_dims[i - 1]
    Q, R = sla.qr(np.reshape(np.transpose(V, (0, 2, 1)), (ni * ri, -1), order="F"), mode="full")
    x.ttv_vec[i - 1] = np.ascontiguousarray(
        np.transpose(np.reshape(Q[:, :rim], (ni, ri, -1), order="F"), (0, 2, 1)))
    x.ttv_ot[i - 1] = 1
    x.ttv_vec[i - 2] = np.einsum("abz,cz->abc", x.ttv_vec[i - 2], R[:rim, :])
    x.ttv_ot[i - 2] = 0
    return x


def _residual(A, x, b):
    return norm(sub(apply(A, x), b)) / max(norm(b), np.finfo(float).eps)


def als_linsolve(A: TToperator, b: TTvector, tt_start: TTvector, sweep_count=2, return_info=False):
    """als.jl:161-225.  NB: `sweep_count` counts half-sweeps (:198-222)."""
    T = tt_start.dtype
    d = A.N
    x = orthogonalize(tt_start)
    dims = tt_start.ttv_dims
    rks = list(tt_start.ttv_rks)
    G = [None] * d
    Gb = [None] * d
    G[0] = np.reshape(A.tto_vec[0][:, :, 0, :], (dims[0], 1, dims[0], 1, -1), order="F").astype(T)
    Gb[0] = np.reshape(b.ttv_vec[0], (dims[0], 1, -1), order="F").astype(T)
    H = init_H(x, A)
    Hb = init_Hb(x, b)
    nsweeps = 0
    while nsweeps < sweep_count:
        nsweeps += 1
        for i in range(1, d):
            V = Ksolve(G[i - 1], Gb[i - 1], H[i - 1], Hb[i - 1])
            x = right_core_move(x, V, i, rks)
            G[i] = update_G(x.ttv_vec[i - 1], A.tto_vec[i], G[i - 1])
            Gb[i] = update_Gb(x.ttv_vec[i - 1], b.ttv_vec[i], Gb[i - 1])
        if nsweeps == sweep_count:
            pass  # als.jl:210-211 evaluates an expression and falls out of the loop
        else:
            nsweeps += 1
            for i in range(d, 1, -1):
                V = Ksolve(G[i - 1], Gb[i - 1], H[i - 1], Hb[i - 1])
                x = left_core_move(x, V, i, rks)
                H[i - 2] = update_H(x.ttv_vec[i - 1], A.tto_vec[i - 1], H[i - 1])
                Hb[i - 2] = update_Hb(x.ttv_vec[i - 1], b.ttv_vec[i - 1], Hb[i - 1])
    if return_info:
        return x, {"residual": _residual(A, x, b)}
    return x


def als_eigsolve(A: TToperator, tt_start: TTvector, sweep_schedule=(2,), rmax_schedule=None):
    """als.jl:251-321 with noise_schedule == 0 and the dense local eigensolver."""
    d = A.N
    if rmax_schedule is None:
        rmax_schedule = [max(tt_start.ttv_rks)]
    assert len(rmax_schedule) == len(sweep_schedule), "Sweep schedule error"
    x = orthogonalize(tt_start)
    dims = tt_start.ttv_dims
    T = tt_start.dtype
    E = []
    G = [None] * d
    G[0] = np.reshape(A.tto_vec[0][:, :, 0, :], (dims[0], 1, dims[0], 1, -1), order="F").astype(T)
    H = init_H(x, A)
    nsweeps = 0
    i_sched = 1
    while i_sched <= len(sweep_schedule):
        nsweeps += 1
        if nsweeps == sweep_schedule[i_sched - 1]:
            i_sched += 1
            if i_sched > len(sweep_schedule):
                return np.array(E), x
            x = increase_ranks(x, rmax_schedule[i_sched - 1])
            x = orthogonalize(x)
            H = init_H(x, A)
            # the zero-padded G[i+1] of als.jl:294-298 are all overwritten by update_G before use
        for i in range(1, d):
            lam, V = K_eigmin(G[i - 1], H[i - 1])
            E.append(lam)
            x = right_core_move(x, V, i, x.ttv_rks)
            G[i] = update_G(x.ttv_vec[i - 1], A.tto_vec[i], G[i - 1])
        for i in range(d, 1, -1):
            lam, V = K_eigmin(G[i - 1], H[i - 1])
            E.append(lam)
            x = left_core_move(x, V, i, x.ttv_rks)
            H[i - 2] = update_H(x.ttv_vec[i - 1], A.tto_vec[i - 1], H[i - 1])
    return np.array(E), x


def K_eiggenmin(Gi, Hi, Ki, Li):
    """als.jl:89-102, dense branch: `eigen(K, S)` of the general (non-symmetric) LAPACK driver, eigenvalues in Julia's default
    order (by real part, then imaginary part), first pair returned.  Note the index order `Gi[d,e,a,b,z] * Hi[z,f,c]` (:91-92):
    the matrices are the transposes of `K_full`'s."""
    dims = (Gi.shape[0], Gi.shape[1], Hi.shape[1])
    n = int(np.prod(dims))
    K = np.reshape(np.einsum("deabz,zfc->abcdef", Gi, Hi), (n, n), order="F")
    S = np.reshape(np.einsum("deabz,zfc->abcdef", Ki, Li), (n, n), order="F")
    w, v = sla.eig(K, S)
    order = sorted(range(n), key=lambda j: (w[j].real, w[j].imag))
    j = order[0]
    return float(np.real(w[j])), np.reshape(v[:, j], dims, order="F")


def als_gen_eigsolv(A: TToperator, S: TToperator, tt_start: TTvector, sweep_schedule=(2,), rmax_schedule=None):
    """als.jl:344-440 (dense local solves): lowest pair of A x = λ S x by ALS sweeps with two sets of environments."""
    d = A.N
    if rmax_schedule is None:
        rmax_schedule = [max(tt_start.ttv_rks)]
    x = orthogonalize(tt_start)
    dims = tt_start.ttv_dims
    T = np.result_type(tt_start.dtype, A.tto_vec[0].dtype, S.tto_vec[0].dtype)
    E = []
    G = [None] * d
    Kc = [None] * d
    G[0] = np.reshape(A.tto_vec[0][:, :, 0, :], (dims[0], 1, dims[0], 1, -1), order="F").astype(T)
    Kc[0] = np.reshape(S.tto_vec[0][:, :, 0, :], (dims[0], 1, dims[0], 1, -1), order="F").astype(T)
    H = init_H(x, A)
    L = init_H(x, S)
    nsweeps = 0
    i_sched = 1
    while i_sched <= len(sweep_schedule):
        nsweeps += 1
        if nsweeps == sweep_schedule[i_sched - 1]:
            i_sched += 1
            if i_sched > len(sweep_schedule):
                return np.array(E), x
            x = increase_ranks(x, rmax_schedule[i_sched - 1])
            x = orthogonalize(x)
            H = init_H(x, A)
            L = init_H(x, S)
        for i in range(1, d):
            if x.ttv_ot[i - 1] == 0:                                              # :404
                lam, V = K_eiggenmin(G[i - 1], H[i - 1], Kc[i - 1], L[i - 1])
                E.append(lam)
                x = right_core_move(x, V.astype(T), i, x.ttv_rks)
            G[i] = update_G(x.ttv_vec[i - 1], A.tto_vec[i], G[i - 1])
            Kc[i] = update_G(x.ttv_vec[i - 1], S.tto_vec[i], Kc[i - 1])
        for i in range(d, 1, -1):
            lam, V = K_eiggenmin(G[i - 1], H[i - 1], Kc[i - 1], L[i - 1])
            E.append(lam)
            x = left_core_move(x, V.astype(T), i, x.ttv_rks)
            H[i - 2] = update_H(x.ttv_vec[i - 1], A.tto_vec[i - 1], H[i - 1])
            L[i - 2] = update_H(x.ttv_vec[i - 1], S.tto_vec[i - 1], L[i - 1])
    return np.array(E), x
