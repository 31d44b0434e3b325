"""CPU restatement of the other two-site-SVD users.  TEST INFRASTRUCTURE ONLY.

  swap_adjacent_sites / bubble_sort_swaps / reorder     src/qtt_tools.jl:660-694, 704-718, 731-774
  ttm_swap / ttm_contract / hadamard_ttm                src/tt_operations.jl:366-422
  to_qtt                                                src/qtt_tools.jl:254-310
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from .core import TTvector
from .ops import svdtrunc


def swap_adjacent_sites(A: np.ndarray, B: np.ndarray, threshold: float = 0.0):
    """qtt_tools.jl:660-694."""
    d1, rl, rm = A.shape
    d2, _, rr = B.shape
    Cm = np.einsum("alm,bmr->ablr", A, B)                          # (s1, s2, l, r)            :667-672
    M = np.transpose(Cm, (1, 2, 0, 3)).reshape(d2 * rl, d1 * rr, order="F")                    # :674-676
    U, sv, Vt = sla.svd(M, full_matrices=False, lapack_driver="gesdd")
    r_new = max(1, int(np.sum(sv > threshold * sv[0]))) if threshold > 0 else len(sv)          # :680-685
    new_A = U[:, :r_new].reshape(d2, rl, r_new, order="F")
    SV = sv[:r_new, None] * Vt[:r_new, :]
    new_B = np.transpose(SV.reshape(r_new, d1, rr, order="F"), (1, 0, 2))                      # :690-692
    return new_A, new_B


def bubble_sort_swaps(perm):
    """qtt_tools.jl:704-718 (1-based swap positions)."""
    p = list(perm)
    swaps = []
    n = len(p)
    for i in range(1, n + 1):
        for j in range(n - i):
            if p[j] > p[j + 1]:
                p[j], p[j + 1] = p[j + 1], p[j]
                swaps.append(j + 1)
    return swaps


def reorder_perm(n_dims: int, bits_per_dim: int, ordering: str):
    """qtt_tools.jl:741-756: perm[src] = target position (0-based values) when leaving `ordering`."""
    perm = [0] * (n_dims * bits_per_dim)
    for d in range(n_dims):
        for b in range(bits_per_dim):
            if ordering == "serial":
                perm[d * bits_per_dim + b] = b * n_dims + d
            else:
                perm[b * n_dims + d] = d * bits_per_dim + b
    return perm


def reorder(x: TTvector, n_dims: int, bits_per_dim: int, ordering: str, new_ordering: str, threshold: float = 0.0) -> TTvector:
    """qtt_tools.jl:731-774 with the QTTvector metadata passed explicitly."""
    assert new_ordering in ("interleaved", "serial")
    cores = [c.copy() for c in x.ttv_vec]
    if ordering != new_ordering:
        for k in bubble_sort_swaps(reorder_perm(n_dims, bits_per_dim, ordering)):
            cores[k - 1], cores[k] = swap_adjacent_sites(cores[k - 1], cores[k], threshold)
    rks = [1] + [c.shape[2] for c in cores]
    return TTvector(x.N, cores, tuple(c.shape[0] for c in cores), rks, [0] * x.N)


def ttm_swap(cores, rks, j: int, tol: float = 0.0, rmax=None):
    """tt_operations.jl:366-383 (j 1-based)."""
    A, B = cores[j - 1], cores[j]
    dA, rL, _ = A.shape
    dB, _, rR = B.shape
    Cm = np.einsum("xma,yan->xymn", A, B)                          # C[sA, sB, m, n]
    mat = np.transpose(Cm, (2, 1, 0, 3)).reshape(rL * dB, dA * rR, order="F")
    U, S, Vt = svdtrunc(mat, max_bond=rmax, truncerr=tol)
    r = U.shape[1]
    cores[j - 1] = np.transpose(U.reshape(rL, dB, r, order="F"), (1, 0, 2))
    cores[j] = np.transpose((S[:, None] * Vt).reshape(r, dA, rR, order="F"), (1, 0, 2))
    rks[j] = r


def ttm_contract(cores, rks, p: int):
    """tt_operations.jl:385-397 (p 1-based)."""
    A, B = cores[p - 1], cores[p]
    cores[p - 1] = np.einsum("slm,smr->slr", A, B)
    del cores[p]
    del rks[p]


def hadamard_ttm(x: TTvector, y: TTvector, tol: float = 1.0e-14, rmax=None) -> TTvector:
    """tt_operations.jl:399-422."""
    assert tuple(x.ttv_dims) == tuple(y.ttv_dims), "Incompatible TT dimensions"
    d = x.N
    cores = [c.copy() for c in x.ttv_vec] + [np.transpose(y.ttv_vec[d - 1 - k], (0, 2, 1)).copy() for k in range(d)]
    rks = list(x.ttv_rks) + list(reversed(list(y.ttv_rks)))[1:]
    for it in range(1, d + 1):
        for j in range(d, d - it + 1, -1):
            ttm_swap(cores, rks, j, tol=tol, rmax=rmax)
        ttm_contract(cores, rks, d - it + 1)
    return TTvector(d, cores, x.ttv_dims, rks, [0] * d)


def to_qtt(tt: TTvector, split_dims, threshold: float = 0.0) -> TTvector:
    """qtt_tools.jl:254-310."""
    assert len(split_dims) == tt.N, "split_dims must have one entry per TT core"
    for i in range(tt.N):
        assert int(np.prod(split_dims[i])) == tt.ttv_dims[i]
    qtt_cores, new_rks, new_dims = [], [1], []
    for i in range(tt.N):
        core = np.transpose(tt.ttv_vec[i], (1, 0, 2))                      # (r_l, n, r_r)
        rank_prev, rank_next, remaining = new_rks[-1], tt.ttv_rks[i + 1], tt.ttv_dims[i]
        for split_size in split_dims[i][:-1]:
            remaining //= split_size
            core = core.reshape(rank_prev, remaining, split_size, rank_next, order="F")
            core = np.transpose(core, (0, 2, 1, 3))
            M = core.reshape(rank_prev * split_size, remaining * rank_next, order="F")
            U, S, Vt = sla.svd(M, full_matrices=False, lapack_driver="gesdd")
            if threshold > 0.0:
                keep = np.flatnonzero(S / S[0] > threshold)
                U, S, Vt = U[:, keep], S[keep], Vt[keep, :]
            new_rank = len(S)
            qtt_cores.append(np.transpose(U.reshape(rank_prev, split_size, new_rank, order="F"), (1, 0, 2)))
            new_rks.append(new_rank)
            new_dims.append(split_size)
            core = (S[:, None] * Vt).reshape(new_rank, remaining, rank_next, order="F")
            rank_prev = new_rank
        qtt_cores.append(np.transpose(core, (1, 0, 2)))
        new_rks.append(rank_next)
        new_dims.append(remaining)
    n = len(qtt_cores)
    return TTvector(n, qtt_cores, tuple(new_dims), new_rks, [0] * n)
