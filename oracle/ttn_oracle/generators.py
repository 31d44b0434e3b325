"""Input generators (oracle; test infrastructure only).

Follows src/tt_operators.jl:4-26 (toeplitz_to_qtto, shift), :162-218 (heisenberg_xyz_tto),
:282-284 (Δ), :519-532 (id_tto), src/tt_operations.jl:71-96 (TToperator +), :268-278 (scalar * TTO),
src/qtt_tools.jl:57-71 (qtt_to_vector), :113-154 (qtt_cos, qtt_sin).
The interleaved 2-D Laplacian follows SURVEY.md Appendix F (direct Kronecker-sum construction; the
reference's `reorder` route, src/tt_operators.jl:696-702, explodes the MPO ranks beyond 2x4 bits).
"""
from __future__ import annotations

import numpy as np

from .core import TTvector, TToperator, zeros_tt, zeros_tto, r_and_d_to_rks


def toeplitz_to_qtto(alpha, beta, gamma, d) -> TToperator:
    """src/tt_operators.jl:4-19.  out = zeros_tto(2, d, 3) → ranks r_and_d_to_rks(3.., dims.^2; rmax=3)."""
    dims = (2,) * d
    rks = r_and_d_to_rks([3] * (d + 1), [4] * d, rmax=3)
    out = zeros_tto(np.float64, dims, rks)
    I2 = np.eye(2)
    J = np.zeros((2, 2))
    J[0, 1] = 1.0
    for i in range(2):
        for j in range(2):
            out.tto_vec[0][i, j, 0, :] = [I2[i, j], J[j, i], J[i, j]]
            for k in range(1, d - 1):
                out.tto_vec[k][i, j, :, :] = np.array([[I2[i, j], J[j, i], J[i, j]],
                                                       [0.0, J[i, j], 0.0],
                                                       [0.0, 0.0, J[j, i]]])
            out.tto_vec[d - 1][i, j, :, 0] = [alpha * I2[i, j] + beta * J[i, j] + gamma * J[j, i],
                                              gamma * J[i, j], beta * J[j, i]]
    return out


def shift_op(d):
    """src/tt_operators.jl:24-26."""
    return toeplitz_to_qtto(0, 1, 0, d)


def laplace_dd(d) -> TToperator:
    """Δ(d), src/tt_operators.jl:282-284."""
    return toeplitz_to_qtto(2, -1, -1, d)


def id_tto(d, dtype=np.float64) -> TToperator:
    """src/tt_operators.jl:524-532."""
    vec = []
    for _ in range(d):
        c = np.zeros((2, 2, 1, 1), dtype=dtype)
        c[:, :, 0, 0] = np.eye(2)
        vec.append(c)
    return TToperator(d, vec, (2,) * d, [1] * (d + 1))


def tto_add(x: TToperator, y: TToperator) -> TToperator:
    """src/tt_operations.jl:71-96 (block-diagonal concatenation)."""
    assert x.tto_dims == y.tto_dims, "Incompatible dimensions"
    d = x.N
    dt = np.result_type(x.dtype, y.dtype)
    rks = [a + b for a, b in zip(x.tto_rks, y.tto_rks)]
    rks[0] = 1
    rks[d] = 1
    vec = [np.zeros((x.tto_dims[k], x.tto_dims[k], rks[k], rks[k + 1]), dtype=dt) for k in range(d)]
    vec[0][:, :, :, :x.tto_rks[1]] = x.tto_vec[0]
    vec[0][:, :, :, x.tto_rks[1]:] = y.tto_vec[0]
    for k in range(1, d - 1):
        vec[k][:, :, :x.tto_rks[k], :x.tto_rks[k + 1]] = x.tto_vec[k]
        vec[k][:, :, x.tto_rks[k]:, x.tto_rks[k + 1]:] = y.tto_vec[k]
    vec[d - 1][:, :, :x.tto_rks[d - 1], :] = x.tto_vec[d - 1]
    vec[d - 1][:, :, x.tto_rks[d - 1]:, :] = y.tto_vec[d - 1]
    return TToperator(d, vec, x.tto_dims, rks)


def tto_scale(a, A: TToperator) -> TToperator:
    """src/tt_operations.jl:268-278 (scales the first core with ot == 0)."""
    i = A.tto_ot.index(0) if 0 in A.tto_ot else 0
    vec = [c.copy() for c in A.tto_vec]
    vec[i] = a * vec[i]
    return TToperator(A.N, vec, A.tto_dims, list(A.tto_rks), list(A.tto_ot))


def heisenberg_xyz_tto(d, jx=1.0, jy=1.0, jz=1.0, lam=0.0, field="x") -> TToperator:
    """src/tt_operators.jl:162-218 (real encoding of σʸσʸ, :56-64; λ·P_field on site terms)."""
    assert d >= 2, "Heisenberg XYZ chain needs at least 2 spin sites"
    X = np.array([[0.0, 1.0], [1.0, 0.0]])
    Z = np.array([[1.0, 0.0], [0.0, -1.0]])
    yr = np.array([[0.0, -1.0], [1.0, 0.0]])
    Y1, Y2 = -yr, yr
    cplx = (lam != 0.0 and field == "y")
    T = np.complex128 if cplx else np.float64
    Pf = {"x": X, "z": Z, "y": np.array([[0.0, -1j], [1j, 0.0]])}[field]
    if not cplx:
        Pf = np.real(Pf) if field != "y" else np.zeros((2, 2))
    I2 = np.eye(2)
    cores = []
    c = np.zeros((2, 2, 1, 5), dtype=T)
    c[:, :, 0, 0] = lam * Pf
    c[:, :, 0, 1] = jx * X
    c[:, :, 0, 2] = jy * Y1
    c[:, :, 0, 3] = jz * Z
    c[:, :, 0, 4] = I2
    cores.append(c)
    for _ in range(1, d - 1):
        c = np.zeros((2, 2, 5, 5), dtype=T)
        c[:, :, 0, 0] = I2
        c[:, :, 1, 0] = X
        c[:, :, 2, 0] = Y2
        c[:, :, 3, 0] = Z
        c[:, :, 4, 0] = lam * Pf
        c[:, :, 4, 1] = jx * X
        c[:, :, 4, 2] = jy * Y1
        c[:, :, 4, 3] = jz * Z
        c[:, :, 4, 4] = I2
        cores.append(c)
    c = np.zeros((2, 2, 5, 1), dtype=T)
    c[:, :, 0, 0] = I2
    c[:, :, 1, 0] = X
    c[:, :, 2, 0] = Y2
    c[:, :, 3, 0] = Z
    c[:, :, 4, 0] = lam * Pf
    cores.append(c)
    return TToperator(d, cores, (2,) * d, [1] + [5] * (d - 1) + [1])


def _qtt_trig(d, a, b, lam, first):
    out = zeros_tt(np.float64, (2,) * d, r_and_d_to_rks([2] * (d + 1), (2,) * d))
    h = (b - a) / (2 ** d - 1)
    w = lam * np.pi
    out.ttv_vec[0][0, 0, :] = first(w * a)
    out.ttv_vec[0][1, 0, :] = first(w * (a + h * 2 ** (d - 1)))
    for k in range(2, d):  # 1-based sites 2..d-1
        tk = h * 2 ** (d - k)
        out.ttv_vec[k - 1][0, :, :] = np.eye(2)
        out.ttv_vec[k - 1][1, :, :] = [[np.cos(w * tk), -np.sin(w * tk)], [np.sin(w * tk), np.cos(w * tk)]]
    out.ttv_vec[d - 1][0, 0, 0] = 1.0
    out.ttv_vec[d - 1][1, :, 0] = [np.cos(w * h), np.sin(w * h)]
    return out


def qtt_sin(d, a=0.0, b=1.0, lam=1.0) -> TTvector:
    """src/qtt_tools.jl:138-154: sin(λ·π·x) on 2^d points of [a,b]."""
    return _qtt_trig(d, a, b, lam, lambda t: [np.sin(t), np.cos(t)])


def qtt_cos(d, a=0.0, b=1.0, lam=1.0) -> TTvector:
    """src/qtt_tools.jl:113-130."""
    return _qtt_trig(d, a, b, lam, lambda t: [np.cos(t), -np.sin(t)])


def qtt_to_vector(q: TTvector) -> np.ndarray:
    """src/qtt_tools.jl:57-71 (site 1 = most significant bit)."""
    P = q.ttv_vec[0][:, 0, :]
    for k in range(1, q.N):
        G = q.ttv_vec[k]
        Pn = np.empty((2 * P.shape[0], G.shape[2]), dtype=np.result_type(P.dtype, G.dtype))
        Pn[0::2, :] = P @ G[0]
        Pn[1::2, :] = P @ G[1]
        P = Pn
    return P.reshape(-1)


# ---------------------------------------------------------------------------------------------
# cfg3 inputs: interleaved (x1,y1,x2,y2,…) 2-D Kronecker-sum operator, built directly.
# ---------------------------------------------------------------------------------------------
def _interleave_with_passthrough(A1d: TToperator, own_first: bool) -> TToperator:
    """Place the cores of a 1-D MPO on every other site; the other dimension's sites carry
    δ_ij·δ_ab pass-through cores on the running bond."""
    d = A1d.N
    vec, rks = [], [1]
    for k in range(d):
        core = A1d.tto_vec[k]
        Rl, Rr = core.shape[2], core.shape[3]
        if own_first:
            vec.append(core.copy())
            rks.append(Rr)
            p = np.zeros((2, 2, Rr, Rr))
            for s in range(2):
                p[s, s] = np.eye(Rr)
            vec.append(p)
            rks.append(Rr)
        else:
            p = np.zeros((2, 2, Rl, Rl))
            for s in range(2):
                p[s, s] = np.eye(Rl)
            vec.append(p)
            rks.append(Rl)
            vec.append(core.copy())
            rks.append(Rr)
    return TToperator(2 * d, vec, (2,) * (2 * d), rks)


def laplace2d_interleaved(bits, a=0.0, b=1.0, shift=0.0, scaled=True) -> TToperator:
    """kron(Δ,I)+kron(I,Δ) (DD) on a 2^bits x 2^bits grid in interleaved bit order, scaled by 1/h²
    (src/tt_operators.jl:654-656), plus optional `shift`·I."""
    lap = laplace_dd(bits)
    X = _interleave_with_passthrough(lap, own_first=True)
    Y = _interleave_with_passthrough(lap, own_first=False)
    A = tto_add(X, Y)
    if scaled:
        h = (b - a) / (2 ** bits - 1)
        A = tto_scale(1.0 / h ** 2, A)
    if shift != 0.0:
        A = tto_add(A, tto_scale(shift, id_tto(2 * bits)))
    return A


def qtt_sin2d_interleaved(bits, lam=1.0) -> TTvector:
    """sin(λπx)·sin(λπy) with interleaved bits as a TT (rank ≤ 4): Kronecker product of two
    qtt_sin trains threaded through each other with identity pass-through on the bond."""
    sx = qtt_sin(bits, lam=lam)
    sy = qtt_sin(bits, lam=lam)
    d = bits
    vec, rks = [], [1]
    # bond after x_k carries (rx_k, ry_{k-1}); after y_k carries (rx_k, ry_k); x index fastest
    for k in range(d):
        cx, cy = sx.ttv_vec[k], sy.ttv_vec[k]
        ryl = cy.shape[1]
        gx = np.einsum("sab,cd->sacbd", cx, np.eye(ryl))  # (s, rxl, ryl, rxr, ryl)
        gx = gx.reshape(2, cx.shape[1] * ryl, cx.shape[2] * ryl, order="F")
        vec.append(gx)
        rks.append(gx.shape[2])
        rxr = cx.shape[2]
        gy = np.einsum("ab,scd->sacbd", np.eye(rxr), cy)  # (s, rxr, ryl, rxr, ryr)
        gy = gy.reshape(2, rxr * cy.shape[1], rxr * cy.shape[2], order="F")
        vec.append(gy)
        rks.append(gy.shape[2])
    return TTvector(2 * d, vec, (2,) * (2 * d), rks, [0] * (2 * d))


# ---- quantum-Fourier-transform MPO (src/tt_transformations.jl:1-77) and uniform-grid QTT sampling ---------------------
def _cheb_lobatto(K):
    """tt_transformations.jl:6-11: nodes on [0, 1] and barycentric weights."""
    j = np.arange(K + 1)
    c = 0.5 * (1.0 - np.cos(np.pi * j / K))
    w = np.where((j == 0) | (j == K), 0.5, 1.0) * (-1.0) ** j
    return c, w


def _lagrange_eval(c, w, alpha, x):
    """tt_transformations.jl:13-24."""
    if abs(x - c[alpha]) <= 1.0e-14:
        return 1.0
    with np.errstate(divide="ignore"):           # x on another node: the denominator is Inf and the value 0, as in Julia
        return (w[alpha] / (x - c[alpha])) / np.sum(w / (x - c))


def fourier_qtto(d, sign=-1.0, K=25, normalize=True) -> TToperator:
    """tt_transformations.jl:38-77: rank K+1 interpolative QFT operator (arXiv:2404.03182); output bits come out reversed."""
    assert d >= 1
    c, w = _cheb_lobatto(K)
    r = K + 1
    A = np.empty((2, 2, r, r), dtype=np.complex128)
    for al in range(r):
        for be in range(r):
            for s in range(2):
                lag = _lagrange_eval(c, w, al, 0.5 * (s + c[be]))
                for t in range(2):
                    A[s, t, al, be] = lag * np.exp(1j * np.pi * sign * (s + c[be]) * t)          # :27-34
    AL = A.sum(axis=2, keepdims=True)                                                             # :48-55
    AR = A[:, :, :, :1].copy()                                                                    # :57-60
    cores = [AL.copy()] + [A.copy() for _ in range(d - 2)] + [AR]
    if normalize:
        cores[0] = cores[0] / np.sqrt(2.0 ** d)
    return TToperator(d, cores, (2,) * d, [1] + [r] * (d - 1) + [1])


def function_to_qtt_uniform(f, d) -> TTvector:
    """src/qtt_tools.jl:73-82: samples f(n / 2^d); site 1 carries the LEAST significant bit (`digits` is little-endian)."""
    from .core import ttv_decomp
    N = 2 ** d
    y = np.array([f(n / N) for n in range(N)])
    return ttv_decomp(np.reshape(y, (2,) * d, order="F"))


def matricize(q: TTvector, core: int) -> np.ndarray:
    """src/tt_tools.jl:694-705: entries of the full tensor listed with site 1 as the MOST significant bit."""
    from .core import ttv_to_tensor
    assert core == q.N
    return np.reshape(ttv_to_tensor(q), -1, order="C")


def qtt_exp(d, a=0.0, b=1.0, alpha=1.0, beta=0.0) -> TTvector:
    """src/qtt_tools.jl:160-176: exp(αx + β) on 2^d points of [a, b] (rank 1, coarsest bit first)."""
    h = (b - a) / (2 ** d - 1)
    vec = []
    for k in range(1, d + 1):
        c = np.zeros((2, 1, 1))
        if k == 1:
            c[0, 0, 0] = np.exp(alpha * a + beta)
            c[1, 0, 0] = np.exp(alpha * (a + h * 2 ** (d - 1)) + beta)
        else:
            c[0, 0, 0] = 1.0
            c[1, 0, 0] = np.exp(alpha * h * 2 ** (d - k))
        vec.append(c)
    return TTvector(d, vec, (2,) * d, [1] * (d + 1), [0] * d)
