"""TT algebra, canonicalisation and rounding (oracle; test infrastructure only).

Follows src/tt_operations.jl:10-35 (+), :101-111 (TTO*TTV), :239-250 (dot), :256-266 (scalar *),
:280-282 (-), :452-470 (euclidean_distance, norm); src/tt_tools.jl:511-543 (orthogonalize),
:737-789 (_tt_bond_truncate!, tt_compress!); src/tt_cross_interpolation.jl:149-166 (the `_svdtrunc`
method that actually dispatches for Matrix inputs: relative tail-norm rule).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from .core import TTvector, TToperator, zeros_tt, r_and_d_to_rks, copy_tt


def apply(A: TToperator, v: TTvector) -> TTvector:
    """src/tt_operations.jl:101-111.  y_k[i,(a,ν),(b,μ)] = Σ_j A_k[i,j,a,b] x_k[j,ν,μ];
    the MPO bond index is the fastest one inside each fused bond (:106)."""
    assert A.tto_dims == v.ttv_dims, "Incompatible dimensions"
    dt = np.result_type(A.dtype, v.dtype)
    rks = [a * b for a, b in zip(A.tto_rks, v.ttv_rks)]
    vec = []
    for k in range(v.N):
        y = np.einsum("ijab,jnm->ianbm", A.tto_vec[k], v.ttv_vec[k])
        n = A.tto_dims[k]
        vec.append(np.ascontiguousarray(np.reshape(y, (n, rks[k], rks[k + 1]), order="F").astype(dt)))
    return TTvector(v.N, vec, v.ttv_dims, rks, [0] * v.N)


def add(x: TTvector, y: TTvector) -> TTvector:
    """src/tt_operations.jl:10-35."""
    assert x.ttv_dims == y.ttv_dims, "Incompatible dimensions"
    d = x.N
    dt = np.result_type(x.dtype, y.dtype)
    rks = [a + b for a, b in zip(x.ttv_rks, y.ttv_rks)]
    rks[0] = 1
    rks[d] = 1
    vec = [np.zeros((x.ttv_dims[k], rks[k], rks[k + 1]), dtype=dt) for k in range(d)]
    if d == 1:
        vec[0][:] = x.ttv_vec[0] + y.ttv_vec[0]
        return TTvector(d, vec, x.ttv_dims, rks, [0] * d)
    vec[0][:, :, :x.ttv_rks[1]] = x.ttv_vec[0]
    vec[0][:, :, x.ttv_rks[1]:] = y.ttv_vec[0]
    for k in range(1, d - 1):
        vec[k][:, :x.ttv_rks[k], :x.ttv_rks[k + 1]] = x.ttv_vec[k]
        vec[k][:, x.ttv_rks[k]:, x.ttv_rks[k + 1]:] = y.ttv_vec[k]
    vec[d - 1][:, :x.ttv_rks[d - 1], :] = x.ttv_vec[d - 1]
    vec[d - 1][:, x.ttv_rks[d - 1]:, :] = y.ttv_vec[d - 1]
    return TTvector(d, vec, x.ttv_dims, rks, [0] * d)


def scale(a, A: TTvector) -> TTvector:
    """src/tt_operations.jl:256-266: scales the first core whose ot flag is 0 (core 1 if none)."""
    dt = np.result_type(type(a), A.dtype) if not isinstance(a, (int, float)) else A.dtype
    if a == 0:
        return zeros_tt(dt, A.ttv_dims, A.ttv_rks)
    i = A.ttv_ot.index(0) if 0 in A.ttv_ot else 0
    vec = [c.astype(dt) for c in A.ttv_vec]
    vec[i] = a * vec[i]
    return TTvector(A.N, vec, A.ttv_dims, list(A.ttv_rks), list(A.ttv_ot))


def sub(A: TTvector, B: TTvector) -> TTvector:
    """src/tt_operations.jl:280-282: A - B = (-1.0 * B) + A."""
    return add(scale(-1.0, B), A)


def dot(A: TTvector, B: TTvector):
    """src/tt_operations.jl:239-250: conj on the first argument."""
    assert A.ttv_dims == B.ttv_dims, "TT dimensions are not compatible"
    M = np.ones((1, 1), dtype=np.result_type(A.dtype, B.dtype))
    for k in range(A.N):
        t = np.einsum("zBb,aB->zab", B.ttv_vec[k], M)  # (z, α, b)
        M = np.einsum("zaA,zab->Ab", np.conj(A.ttv_vec[k]), t)
    return M[0, 0]


def norm(a: TTvector) -> float:
    """src/tt_operations.jl:465-470."""
    v = float(np.real(dot(a, a)))
    return float(np.sqrt(max(v, 0.0)))


def euclidean_distance(a: TTvector, b: TTvector) -> float:
    """src/tt_operations.jl:452-455."""
    v = np.real(dot(a, a) - 2 * np.real(dot(b, a)) + dot(b, b))
    return float(np.sqrt(max(v, 0.0)))


def norm_stable(z: TTvector) -> float:
    """‖z‖ via a right-to-left orthogonalisation sweep: backward stable (error ~ eps·Σ‖terms‖), unlike
    sqrt(dot(z, z)) of a difference TT, which cancels and bottoms out at sqrt(eps)."""
    y = orthogonalize(z, i=1)
    return float(np.linalg.norm(y.ttv_vec[0]))


def rel_distance(a: TTvector, b: TTvector) -> float:
    """‖a-b‖/‖b‖ on the (un-rounded) difference TT, evaluated with `norm_stable` so that it resolves down to
    machine precision (the expanded form of euclidean_distance is limited to sqrt(eps) by cancellation)."""
    return norm_stable(sub(a, b)) / max(norm(b), np.finfo(float).tiny)


def orthogonalize(x: TTvector, i: int = 1) -> TTvector:
    """src/tt_tools.jl:511-543, centre `i` is 1-based.  Pure: returns a new TT."""
    d = x.N
    assert 1 <= i <= d, "Impossible orthogonalization"
    T = x.dtype
    y_rks = r_and_d_to_rks(x.ttv_rks, x.ttv_dims)
    y = zeros_tt(T, x.ttv_dims, y_rks)
    FR = np.ones((1, 1), dtype=T)
    for j in range(1, i):  # 1-based j = 1..i-1
        y.ttv_ot[j - 1] = 1
        n = x.ttv_dims[j - 1]
        # temp[α_{j-1}, i_j, α_j] = FR[α_{j-1}, β_{j-1}] x_j[i_j, β_{j-1}, α_j]      (:520)
        temp = np.einsum("ab,ibc->aic", FR, x.ttv_vec[j - 1])
        M = np.reshape(temp, (n * y.ttv_rks[j - 1], -1), order="F")
        Q, R = sla.qr(M, mode="economic")
        y.ttv_rks[j] = Q.shape[1]
        y.ttv_vec[j - 1] = np.ascontiguousarray(
            np.transpose(np.reshape(Q, (y.ttv_rks[j - 1], n, y.ttv_rks[j]), order="F"), (1, 0, 2)))
        FR = R[:y.ttv_rks[j], :]
    FL = np.ones((1, 1), dtype=T)
    for j in range(d, i, -1):  # 1-based j = d..i+1
        y.ttv_ot[j - 1] = -1
        n = x.ttv_dims[j - 1]
        # temp[α_{j-1}, β_j, i_j] = x_j[i_j, α_{j-1}, α_j] FL[α_j, β_j]              (:531)
        temp = np.einsum("iac,cb->abi", x.ttv_vec[j - 1], FL)
        M = np.reshape(temp, (x.ttv_rks[j - 1], -1), order="F")
        # LQ via QR of the conjugate transpose
        Qh, Rh = sla.qr(M.conj().T, mode="economic")
        L, Q = Rh.conj().T, Qh.conj().T
        y.ttv_rks[j - 1] = Q.shape[0]
        y.ttv_vec[j - 1] = np.ascontiguousarray(
            np.transpose(np.reshape(Q, (y.ttv_rks[j - 1], y.ttv_rks[j], n), order="F"), (2, 0, 1)))
        FL = L[:, :y.ttv_rks[j - 1]]
    y.ttv_ot[i - 1] = 0
    c = np.zeros((x.ttv_dims[i - 1], y.ttv_rks[i - 1], y.ttv_rks[i]), dtype=T)
    for k in range(x.ttv_dims[i - 1]):
        c[k] = FR @ x.ttv_vec[i - 1][k] @ FL
    y.ttv_vec[i - 1] = c
    return y


def svdtrunc(A: np.ndarray, max_bond=None, truncerr: float = 0.0):
    """The effective `_svdtrunc` for Matrix arguments: src/tt_cross_interpolation.jl:149-166
    (thin gesdd SVD, relative *tail-norm* rank rule, then the max_bond cap).
    Returns (U[:, :r], s[:r], Vt[:r, :])."""
    U, s, Vt = sla.svd(A, full_matrices=False, lapack_driver="gesdd")
    r = len(s)
    if truncerr > 0:
        nrm = np.linalg.norm(s)
        cum = 0.0
        for i in range(r, 0, -1):
            cum += abs(s[i - 1]) ** 2
            if np.sqrt(cum) > truncerr * nrm:
                r = i
                break
    if max_bond is not None:
        r = min(r, max_bond)
    return U[:, :r], s[:r], Vt[:r, :]


def svdtrunc_abs(A: np.ndarray, max_bond=None, truncerr: float = 0.0):
    """src/tt_tools.jl:737-741 — the absolute-threshold method (shadowed for Matrix inputs; kept so the
    oracle tests can document the dispatch)."""
    U, s, Vt = sla.svd(A, full_matrices=False, lapack_driver="gesdd")
    mb = max(A.shape) if max_bond is None else max_bond
    d = min(mb, int(np.sum(s >= truncerr)))
    return U[:, :d], s[:d], Vt[:d, :]


def tt_bond_truncate(psi: TTvector, k: int, max_bond=None, truncerr: float = 0.0, faithful: bool = True,
                     sigma_out: list | None = None):
    """src/tt_tools.jl:743-770 (`_tt_bond_truncate!`, k 1-based; mutates psi).
    `faithful=True` also performs the reference's trailing `orthogonalize(ψ; i=k)` (:769) and returns
    it; `faithful=False` skips that pure, discarded call and returns None."""
    assert 1 <= k < psi.N, "k must be in 1:(N-1)"
    A = np.transpose(psi.ttv_vec[k - 1], (1, 0, 2))      # (α, s1, γ)
    B = np.transpose(psi.ttv_vec[k], (1, 0, 2))          # (γ, s2, β)
    AAC = np.einsum("asg,gtb->astb", A, B)
    Dl, d1, d2, Dr = AAC.shape
    U, s, Vt = svdtrunc(np.reshape(AAC, (Dl * d1, d2 * Dr), order="F"), max_bond=max_bond, truncerr=truncerr)
    if sigma_out is not None:
        sigma_out.append(s.copy())
    ss = np.sqrt(s)
    U = U * ss[None, :]
    Vt = ss[:, None] * Vt
    new_r = U.shape[1]
    AL = np.reshape(U, (Dl, d1, new_r), order="F")
    AR = np.reshape(Vt, (new_r, d2, Dr), order="F")
    psi.ttv_vec[k - 1] = np.ascontiguousarray(np.transpose(AL, (1, 0, 2)))
    psi.ttv_vec[k] = np.ascontiguousarray(np.transpose(AR, (1, 0, 2)))
    psi.ttv_rks[k] = new_r
    if faithful:
        return orthogonalize(psi, i=k)
    return None


def tt_compress(psi: TTvector, max_bond: int, truncerr: float = 0.0, sweeps: int = 1, faithful: bool = False,
                sigma_out: list | None = None) -> TTvector:
    """src/tt_tools.jl:772-789 (`tt_compress!`): L→R then R→L two-site SVD sweeps on a non-canonical TT,
    √S split to both sides, `ttv_ot` untouched.  Mutates and returns `psi`.
    `faithful=True` additionally executes (and discards, as the reference does) the `orthogonalize`
    of every bond step — used only to time the reference-faithful CPU baseline."""
    assert sweeps >= 1, "sweeps must be >= 1"
    for _ in range(sweeps):
        for k in range(1, psi.N):
            tt_bond_truncate(psi, k, max_bond=max_bond, truncerr=truncerr, faithful=faithful, sigma_out=sigma_out)
        for k in range(psi.N - 1, 0, -1):
            tt_bond_truncate(psi, k, max_bond=max_bond, truncerr=truncerr, faithful=faithful, sigma_out=sigma_out)
    return psi


def hadamard(x: TTvector, y: TTvector) -> TTvector:
    """src/tt_operations.jl:343-360: core_k[s] = kron(x_k[s], y_k[s])."""
    assert x.ttv_dims == y.ttv_dims, "Incompatible TT dimensions"
    vec, rks = [], [x.ttv_rks[k] * y.ttv_rks[k] for k in range(x.N + 1)]
    for k in range(x.N):
        cx, cy = x.ttv_vec[k], y.ttv_vec[k]
        core = np.zeros((cx.shape[0], cx.shape[1] * cy.shape[1], cx.shape[2] * cy.shape[2]), dtype=np.result_type(cx.dtype, cy.dtype))
        for s_ in range(cx.shape[0]):
            core[s_] = np.kron(cx[s_], cy[s_])
        vec.append(core)
    return TTvector(x.N, vec, x.ttv_dims, rks, [0] * x.N)
