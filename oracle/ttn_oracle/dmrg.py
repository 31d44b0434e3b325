"""DMRG (oracle; test infrastructure only).  Follows src/solvers/dmrg.jl for N ∈ {1, 2}.

dmrg.jl:10-35 (operator environments), :38-46 (Amid), :49-54 (K_full), :56-90 (RHS environments,
b_mid), :92-177 (Ksolve!), :179-185 (cut_off_index), :187-232 (SVD core moves), :235-259 (K_eigmin),
:261-296 (workspaces), :312-342 (update_right / update_left), :385-473 (dmrg_linsolve),
:501-578 (dmrg_eigsolve).

Third-party Krylov pieces (KrylovKit.eigsolve / linsolve, dmrg.jl:170,245) are not under the reference
tree; the oracle replaces them by *converged* dense solves of the same (symmetrised) local operator, so
parity is claimed on converged energies / solutions, not on 1e-6-tolerance sweep trajectories
(SURVEY.md §7.3 "Krylov equivalence").  Rank-padded buffers + views become exactly-sized arrays.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.sparse.linalg as spla

from .core import TTvector, TToperator, r_and_d_to_rks, increase_ranks
from .ops import orthogonalize, apply, sub, norm


def dmrg_update_H(x_vec, A_vec, Hi):
    """dmrg.jl:27-30: Him[a,α,β] = conj(x)[j,α,φ] Hi[z,φ,χ] x[k,β,χ] A[j,k,a,z]."""
    return np.einsum("jaf,zfc,kbc,jkuz->uab", np.conj(x_vec), Hi, x_vec, A_vec, optimize=True)


def dmrg_update_G(x_vec, A_vec, Gi):
    """dmrg.jl:32-35: Gip[a,α,β] = conj(x)[j,φ,α] Gi[z,φ,χ] x[k,χ,β] A[j,k,z,a]."""
    return np.einsum("jfa,zfc,kcb,jkzu->uab", np.conj(x_vec), Gi, x_vec, A_vec, optimize=True)


def amid(A: TToperator, i: int, j: int):
    """dmrg.jl:38-46 (1-based sites i..j): (R_{i-1}, n_i⋯n_j, n_i⋯n_j, R_j), first site fastest."""
    out = np.transpose(A.tto_vec[i - 1], (2, 0, 1, 3))
    for k in range(i + 1, j + 1):
        C = out  # (α, I, J, ξ)
        t = np.einsum("ijxb,aIJx->aIiJjb", A.tto_vec[k - 1], C)
        a, I, ii, J, jj, b = t.shape
        out = np.reshape(t, (a, I * ii, J * jj, b), order="F")
    return out


def b_mid(b: TTvector, i: int, j: int):
    """dmrg.jl:83-90: (r_{i-1}, n_i⋯n_j, r_j), first site fastest."""
    out = np.transpose(b.ttv_vec[i - 1], (1, 0, 2))
    for k in range(i + 1, j + 1):
        t = np.einsum("aix,jxb->aijb", out, b.ttv_vec[k - 1])
        out = np.reshape(t, (t.shape[0], -1, t.shape[3]), order="F")
    return out


def dmrg_matvec2(G, Am, V, H, symmetrize=True):
    """The K_matfree contraction of dmrg.jl:239-244.
    symmetrize=True : 0.5·(G·Amid·V·H + Gᵀ·Amidᵀ·V·Hᵀ)  (what the reference applies)
    symmetrize=False: the single application G·Amid·V·H."""
    def chain(Gx, Ax, Hx):   # Σ Gx[y,a,d] Ax[y,b,e,z] V[d,e,f] Hx[z,c,f] as three pairwise contractions
        t1 = np.einsum("yad,def->yaef", Gx, V)
        t2 = np.einsum("ybez,yaef->abzf", Ax, t1)
        return np.einsum("abzf,zcf->abc", t2, Hx)
    Y = chain(G, Am, H)
    if symmetrize:
        Y2 = chain(np.transpose(G, (0, 2, 1)), np.transpose(Am, (0, 2, 1, 3)), np.transpose(H, (0, 2, 1)))
        Y = 0.5 * (Y + Y2)
    return Y


def dmrg_matvec2_blas(G, Am, V, H):
    """The single application G·Amid·V·H of dmrg.jl:239-244 lowered to three BLAS GEMMs (tensordot = permute + gemm), which
    is what TensorOperations' `@tensoropt` does for the reference on the CPU.  Same result as dmrg_matvec2(symmetrize=False);
    used as the timed CPU baseline of the matvec (bench.py), where the plain-einsum version would understate the CPU."""
    t1 = np.tensordot(G, V, axes=([2], [0]))              # [y, a, e, f]
    t2 = np.tensordot(Am, t1, axes=([0, 2], [0, 2]))      # [b, z, a, f]
    Y = np.tensordot(t2, H, axes=([1, 3], [0, 2]))        # [b, a, c]
    return np.transpose(Y, (1, 0, 2))


def K_full(G, H, Am):
    """dmrg.jl:49-54 (without the Hermitian wrapper)."""
    dims = (G.shape[1], Am.shape[1], H.shape[1])
    K = np.einsum("yad,zcf,ybez->abcdef", G, H, Am, optimize=True)
    n = int(np.prod(dims))
    return np.reshape(K, (n, n), order="F"), dims


def cut_off_index(s, tol, degen_tol=1e-10):
    """dmrg.jl:179-185."""
    s = np.asarray(s, dtype=float)
    k = int(np.sum(s > np.linalg.norm(s) * tol))
    if k == 0 and len(s) > 0:
        raise IndexError("cut_off_index: no singular value above the threshold (Julia: BoundsError at s[0])")
    # Julia's isapprox(x, y; rtol, atol): |x - y| <= max(atol, rtol * max(|x|, |y|))   (not numpy.isclose's atol + rtol |y|)
    while k < len(s) and abs(s[k - 1] - s[k]) <= max(degen_tol, degen_tol * max(abs(s[k - 1]), abs(s[k]))):
        k += 1
    return k


def _eigmin(G, H, Am, K_dims, it_solver, itslv_thresh):
    """dmrg.jl:235-259.  Dense branch: eigen(Hermitian(K)) (upper triangle).  Iterative branch: lowest
    eigenpair of the symmetrised operator, converged (dense eigh up to order 4096, ARPACK beyond)."""
    n = int(np.prod(K_dims))
    if it_solver or n > itslv_thresh:
        if n <= 4096:
            K, _ = K_full(G, H, Am)
            Ks = 0.5 * (K + K.T)
            w, v = sla.eigh(Ks, subset_by_index=[0, 0])
            return float(w[0]), np.reshape(v[:, 0], K_dims, order="F")
        op = spla.LinearOperator((n, n), dtype=G.dtype, matvec=lambda x: np.reshape(
            dmrg_matvec2(G, Am, np.reshape(x, K_dims, order="F"), H), -1, order="F"))
        w, v = spla.eigsh(op, k=1, which="SA", tol=1e-13)
        return float(w[0]), np.reshape(v[:, 0], K_dims, order="F")
    K, _ = K_full(G, H, Am)
    w, v = sla.eigh(K, lower=False, subset_by_index=[0, 0])
    return float(np.real(w[0])), np.reshape(v[:, 0], K_dims, order="F")


def _ksolve(G, Gb, H, Hb, Am, Bm, it_solver, itslv_thresh):
    """dmrg.jl:92-177 with the linear solve converged."""
    K_dims = (G.shape[1], Am.shape[1], H.shape[1])
    Pb = np.einsum("ab,bic,dc->aid", Gb, Bm, Hb, optimize=True)
    K, _ = K_full(G, H, Am)
    n = K.shape[0]
    if it_solver or n > itslv_thresh:
        Ks = 0.5 * (K + K.T)
    else:
        Ks = np.triu(K) + np.triu(K, 1).T  # Hermitian(K): upper triangle mirrored
    V = np.linalg.solve(Ks, np.reshape(Pb, -1, order="F"))
    return np.reshape(V, K_dims, order="F")


def right_core_move(x: TTvector, V, i, tol, r_max):
    """dmrg.jl:187-209 (i 1-based).  Returns V_move = S·Vt reshaped (r, mid, size(V,3))."""
    u, s, vh = sla.svd(np.reshape(V, (x.ttv_rks[i - 1] * x.ttv_dims[i - 1], -1), order="F"),
                       full_matrices=False, lapack_driver="gesdd")
    r = min(cut_off_index(s, tol), r_max)
    x.ttv_vec[i - 1] = np.ascontiguousarray(
        np.transpose(np.reshape(u[:, :r], (x.ttv_rks[i - 1], x.ttv_dims[i - 1], -1), order="F"), (1, 0, 2)))
    x.ttv_rks[i] = r
    x.ttv_ot[i - 1] = 1
    x.ttv_ot[i] = 0
    Vm = np.reshape(vh[:r, :], (r, -1, V.shape[2]), order="F") * s[:r, None, None]
    return Vm


def left_core_move(x: TTvector, V, j, tol, r_max):
    """dmrg.jl:211-232 (j 1-based).  Returns V_move = U·S reshaped (size(V,1), mid, r)."""
    u, s, vh = sla.svd(np.reshape(V, (-1, x.ttv_dims[j - 1] * x.ttv_rks[j]), order="F"),
                       full_matrices=False, lapack_driver="gesdd")
    r = min(cut_off_index(s, tol), r_max)
    x.ttv_vec[j - 1] = np.ascontiguousarray(
        np.transpose(np.reshape(vh[:r, :], (r, -1, x.ttv_rks[j]), order="F"), (1, 0, 2)))
    x.ttv_rks[j - 1] = r
    x.ttv_ot[j - 1] = -1
    x.ttv_ot[j - 2] = 0
    Vm = np.reshape(u[:, :r], (V.shape[0], -1, r), order="F") * s[None, None, :r]
    return Vm


def _update_right(x, V, i, N, tol, rmax, Ai, Gi):
    """dmrg.jl:312-326: core move, next guess V0 = V_move·core_{i+N}, G[i+1]."""
    Vm = right_core_move(x, V, i, tol, rmax)
    t = np.einsum("aJb,ibc->aJic", Vm, x.ttv_vec[i + N - 1])
    V0 = np.reshape(t, (t.shape[0], -1, t.shape[3]), order="F")
    Gip = dmrg_update_G(x.ttv_vec[i - 1], Ai, Gi)
    return V0, Gip


def _update_left(x, V, i, N, tol, rmax, Aip, Hi):
    """dmrg.jl:328-342 (including the reference's (J, i_{k}) index order of the next guess)."""
    Vm = left_core_move(x, V, i + N - 1, tol, rmax)
    t = np.einsum("bJc,iab->aJic", Vm, x.ttv_vec[i - 2])
    V0 = np.reshape(t, (t.shape[0], -1, t.shape[3]), order="F")
    Him = dmrg_update_H(x.ttv_vec[i + N - 2], Aip, Hi)
    return V0, Him


def _init_H(x: TTvector, A: TToperator, N):
    """dmrg.jl:10-25 with exact-size arrays.  H[i] (1-based i=1..d+1-N) ↔ bond i+N-1."""
    d = x.N
    H = [None] * (d + 1 - N)
    H[d - N] = np.ones((1, 1, 1), dtype=x.dtype)
    for i in range(d + 1 - N, 1, -1):
        H[i - 2] = dmrg_update_H(x.ttv_vec[i + N - 2], A.tto_vec[i + N - 2], H[i - 1])
    return H


def _update_Hb(x_vec, b_vec, Hbi):
    """dmrg.jl:73-76: H_bim[α,β] = H_bi[φ,χ] b[i,β,χ] conj(x)[i,α,φ]."""
    return np.einsum("fc,ibc,iaf->ab", Hbi, b_vec, np.conj(x_vec), optimize=True)


def _update_Gb(x_vec, b_vec, Gbi):
    """dmrg.jl:78-81: G_bip[α,β] = G_bi[φ,χ] b[i,χ,β] conj(x)[i,φ,α]."""
    return np.einsum("fc,icb,ifa->ab", Gbi, b_vec, np.conj(x_vec), optimize=True)


def _init_Hb(x: TTvector, b: TTvector, N):
    """dmrg.jl:56-71."""
    d = x.N
    Hb = [None] * (d + 1 - N)
    Hb[d - N] = np.ones((1, 1), dtype=x.dtype)
    for i in range(d + 1 - N, 1, -1):
        Hb[i - 2] = _update_Hb(x.ttv_vec[i + N - 2], b.ttv_vec[i + N - 2], Hb[i - 1])
    return Hb


def _final_split(x: TTvector, V, N, tol, rmax_last):
    """dmrg.jl:451-462 / :540-551: split the last local solution at site 1 back into cores."""
    if N == 1:
        x.ttv_vec[0] = np.ascontiguousarray(np.transpose(V, (1, 0, 2)))
    else:
        Vm = left_core_move(x, V, N, tol, rmax_last)
        x.ttv_vec[0] = np.ascontiguousarray(np.transpose(np.reshape(Vm, (1, x.ttv_dims[0], -1), order="F"), (1, 0, 2)))
    x.ttv_ot[0] = 0


def dmrg_eigsolve(A: TToperator, tt_start: TTvector, N=2, tol=1e-12, sweep_schedule=(2,), rmax_schedule=None,
                  it_solver=False, itslv_thresh=256):
    """dmrg.jl:501-578.  Returns (E, x, r_hist)."""
    assert N in (1, 2)
    d = tt_start.N
    if rmax_schedule is None:
        rmax_schedule = [int(np.sqrt(float(np.prod([float(n) for n in tt_start.ttv_dims]))))]
    assert len(rmax_schedule) == len(sweep_schedule), "Sweep schedule error"
    x = orthogonalize(tt_start)
    E, r_hist = [], []
    G = [None] * (d + 1 - N)
    Am = [amid(A, i, i + N - 1) for i in range(1, d + 2 - N)]
    G[0] = np.ones((A.tto_rks[0], 1, 1), dtype=x.dtype)
    H = _init_H(x, A, N)
    V0 = b_mid(x, 1, N)
    nsweeps = 0
    i_sched = 1
    while i_sched <= len(sweep_schedule):
        nsweeps += 1
        if nsweeps == sweep_schedule[i_sched - 1]:
            i_sched += 1
            if i_sched > len(sweep_schedule):
                lam, V = _eigmin(G[0], H[0], Am[0], V0.shape, it_solver, itslv_thresh)
                E.append(lam)
                r_hist.append(max(x.ttv_rks))
                _final_split(x, V, N, tol, rmax_schedule[-1])
                return np.array(E), x, r_hist
        rmax = rmax_schedule[i_sched - 1]
        for i in range(1, d - N + 1):
            lam, V = _eigmin(G[i - 1], H[i - 1], Am[i - 1], V0.shape, it_solver, itslv_thresh)
            E.append(lam)
            V0, G[i] = _update_right(x, V, i, N, tol, rmax, A.tto_vec[i - 1], G[i - 1])
            r_hist.append(max(x.ttv_rks))
        for i in range(d - N + 1, 1, -1):
            lam, V = _eigmin(G[i - 1], H[i - 1], Am[i - 1], V0.shape, it_solver, itslv_thresh)
            E.append(lam)
            V0, H[i - 2] = _update_left(x, V, i, N, tol, rmax, A.tto_vec[i + N - 2], H[i - 1])
            r_hist.append(max(x.ttv_rks))
    return np.array(E), x, r_hist


def dmrg_linsolve(A: TToperator, b: TTvector, tt_start: TTvector, N=2, tol=1e-12, sweep_schedule=(2,),
                  rmax_schedule=None, it_solver=True, itslv_thresh=256, return_info=False):
    """dmrg.jl:385-473."""
    assert N in (1, 2)
    d = b.N
    if rmax_schedule is None:
        rmax_schedule = [int(np.sqrt(float(np.prod([float(n) for n in tt_start.ttv_dims]))))]
    rmax_all = max(rmax_schedule)
    if N == 1:
        tt_start = increase_ranks(tt_start, rmax_all)
    x = orthogonalize(tt_start)
    G = [None] * (d + 1 - N)
    Gb = [None] * (d + 1 - N)
    Am = [amid(A, i, i + N - 1) for i in range(1, d + 2 - N)]
    Bm = [b_mid(b, i, i + N - 1) for i in range(1, d + 2 - N)]
    G[0] = np.ones((A.tto_rks[0], 1, 1), dtype=x.dtype)
    Gb[0] = np.ones((1, b.ttv_rks[0]), dtype=x.dtype)
    H = _init_H(x, A, N)
    Hb = _init_Hb(x, b, N)
    nsweeps = 0
    i_sched = 1

    def res():
        return norm(sub(apply(A, x), b)) / max(norm(b), np.finfo(float).eps)

    while i_sched <= len(sweep_schedule):
        nsweeps += 1
        if nsweeps == sweep_schedule[i_sched - 1]:
            i_sched += 1
            if i_sched > len(sweep_schedule):
                V = _ksolve(G[0], Gb[0], H[0], Hb[0], Am[0], Bm[0], it_solver, itslv_thresh)
                _final_split(x, V, N, tol, rmax_schedule[-1])
                return (x, {"residual": res()}) if return_info else x
        rmax = rmax_schedule[i_sched - 1]
        for i in range(1, d - N + 1):
            V = _ksolve(G[i - 1], Gb[i - 1], H[i - 1], Hb[i - 1], Am[i - 1], Bm[i - 1], it_solver, itslv_thresh)
            _, G[i] = _update_right(x, V, i, N, tol, rmax, A.tto_vec[i - 1], G[i - 1])
            Gb[i] = _update_Gb(x.ttv_vec[i - 1], b.ttv_vec[i - 1], Gb[i - 1])
        for i in range(d + 1 - N, 1, -1):
            V = _ksolve(G[i - 1], Gb[i - 1], H[i - 1], Hb[i - 1], Am[i - 1], Bm[i - 1], it_solver, itslv_thresh)
            _, H[i - 2] = _update_left(x, V, i, N, tol, rmax, A.tto_vec[i + N - 2], H[i - 1])
            Hb[i - 2] = _update_Hb(x.ttv_vec[i + N - 2], b.ttv_vec[i + N - 2], Hb[i - 1])
    return (x, {"residual": res()}) if return_info else x
