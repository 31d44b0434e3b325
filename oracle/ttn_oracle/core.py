"""TT containers and constructors (oracle; test infrastructure only).

Follows src/tt_tools.jl:9-65 (types), :100-139 (rand_tt), :186-233 (ttv_decomp), :265-279
(ttv_to_tensor), :407-425 (r_and_d_to_rks), :443-496 (increase_ranks) and
src/tt_operators.jl:534-616 (rand_tto, zeros_tt, zeros_tto).
"""
from __future__ import annotations

import numpy as np


def _prod_i64(xs) -> int:
    """Julia `prod` on Int64 wraps on overflow; r_and_d_to_rks relies on the sign (tt_tools.jl:409-422)."""
    p = 1
    for x in xs:
        p = (p * int(x)) & 0xFFFFFFFFFFFFFFFF
    if p >= 1 << 63:
        p -= 1 << 64
    return p


class TTvector:
    """src/tt_tools.jl:23-29.  `ttv_vec[k]` has shape (n_k, r_{k-1}, r_k)."""

    def __init__(self, N, ttv_vec, ttv_dims, ttv_rks, ttv_ot):
        self.N = int(N)
        self.ttv_vec = list(ttv_vec)
        self.ttv_dims = tuple(int(n) for n in ttv_dims)
        self.ttv_rks = [int(r) for r in ttv_rks]
        self.ttv_ot = [int(o) for o in ttv_ot]

    @property
    def dtype(self):
        return self.ttv_vec[0].dtype


class TToperator:
    """src/tt_tools.jl:48-54.  `tto_vec[k]` has shape (n_k, n_k, R_{k-1}, R_k)."""

    def __init__(self, N, tto_vec, tto_dims, tto_rks, tto_ot=None):
        self.N = int(N)
        self.tto_vec = list(tto_vec)
        self.tto_dims = tuple(int(n) for n in tto_dims)
        self.tto_rks = [int(r) for r in tto_rks]
        self.tto_ot = [0] * self.N if tto_ot is None else [int(o) for o in tto_ot]

    @property
    def dtype(self):
        return self.tto_vec[0].dtype


def r_and_d_to_rks(rks, dims, rmax=1024):
    """src/tt_tools.jl:407-425 (including the overflow-tolerant `prod(...) > 0` tests)."""
    rks = [int(r) for r in rks]
    new = [1] * len(rks)
    for i in range(len(dims)):
        left = _prod_i64(dims[:i])
        right = _prod_i64(dims[i:])
        if right > 0:
            if left > 0:
                new[i] = min(rks[i], left, right, rmax)
            else:
                new[i] = min(rks[i], right, rmax)
        else:
            if left > 0:
                new[i] = min(rks[i], left, rmax)
            else:
                new[i] = min(rks[i], rmax)
    return new


def zeros_tt(dtype, dims, rks, ot=None):
    """src/tt_operators.jl:552-558."""
    assert len(dims) + 1 == len(rks), "Dimensions and ranks are not compatible"
    vec = [np.zeros((dims[i], rks[i], rks[i + 1]), dtype=dtype) for i in range(len(dims))]
    return TTvector(len(dims), vec, dims, rks, [0] * len(dims) if ot is None else ot)


def zeros_tto(dtype, dims, rks):
    """src/tt_operators.jl:604-608."""
    vec = [np.zeros((dims[i], dims[i], rks[i], rks[i + 1]), dtype=dtype) for i in range(len(dims))]
    return TToperator(len(dims), vec, dims, rks)


def _randn(rng, dtype, shape):
    if np.issubdtype(dtype, np.complexfloating):
        # Julia randn(ComplexF64) has unit variance: real/imag each N(0, 1/2)
        return ((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / np.sqrt(2.0)).astype(dtype)
    return rng.standard_normal(shape).astype(dtype)


def rand_tt(dims, rks, rng=None, dtype=np.float64, normalise=False, orthogonal=False):
    """src/tt_tools.jl:119-139.  `rks` may be an int (rmax form, :134-139)."""
    rng = np.random.default_rng(0) if rng is None else rng
    d = len(dims)
    if np.isscalar(rks):
        rmax = int(rks)
        rks = r_and_d_to_rks([rmax] * (d + 1), dims, rmax=rmax)
    y = zeros_tt(dtype, dims, rks)
    for i in range(d):
        c = _randn(rng, np.dtype(dtype), (dims[i], rks[i], rks[i + 1]))
        if normalise:
            c = c * (1.0 / np.sqrt(dims[i] * rks[i + 1]))
            if orthogonal:
                m = np.reshape(np.transpose(c, (0, 2, 1)), (dims[i] * rks[i + 1], rks[i]), order="F")
                q, _ = np.linalg.qr(m)
                c = np.transpose(np.reshape(q, (dims[i], rks[i + 1], rks[i]), order="F"), (0, 2, 1))
        y.ttv_vec[i] = np.ascontiguousarray(c)
    return y


def rand_tto(dims, rmax, rng=None, dtype=np.float64):
    """src/tt_operators.jl:534-545."""
    rng = np.random.default_rng(0) if rng is None else rng
    d = len(dims)
    rks = [1] * (d + 1)
    vec = []
    for i in range(d):
        ri = min(_prod_i64(dims[:i]), _prod_i64(dims[i:]), rmax)
        rip = min(_prod_i64(dims[:i + 1]), _prod_i64(dims[i + 1:]), rmax)
        rks[i + 1] = rip
        vec.append(_randn(rng, np.dtype(dtype), (dims[i], dims[i], ri, rip)))
    return TToperator(d, vec, dims, rks)


def copy_tt(x: TTvector) -> TTvector:
    """src/tt_tools.jl:172-178."""
    return TTvector(x.N, [c.copy() for c in x.ttv_vec], x.ttv_dims, list(x.ttv_rks), list(x.ttv_ot))


def complex_tt(x: TTvector) -> TTvector:
    """src/tt_tools.jl:63-65."""
    return TTvector(x.N, [c.astype(np.complex128) for c in x.ttv_vec], x.ttv_dims, list(x.ttv_rks), list(x.ttv_ot))


def complex_tto(A: TToperator) -> TToperator:
    """src/tt_tools.jl:59-61."""
    return TToperator(A.N, [c.astype(np.complex128) for c in A.tto_vec], A.tto_dims, list(A.tto_rks), list(A.tto_ot))


def ttv_to_tensor(x: TTvector) -> np.ndarray:
    """src/tt_tools.jl:265-279: tensor[s_1,…,s_d]."""
    cur = x.ttv_vec[0][:, 0, :]  # (n1, r1)
    for k in range(1, x.N):
        c = x.ttv_vec[k]  # (n, rl, rr)
        cur = np.einsum("...a,sab->...sb", cur, c)
    return cur[..., 0]


def tto_to_matrix(A: TToperator) -> np.ndarray:
    """Dense matrix of a TToperator with big-endian site order (site 1 most significant), matching
    qtt_to_vector (src/qtt_tools.jl:57-71) and the dense checks of test/test_tt_tools.jl:345-358."""
    cur = A.tto_vec[0][:, :, 0, :]  # (i, j, b)
    rows, cols = cur.shape[0], cur.shape[1]
    for k in range(1, A.N):
        c = A.tto_vec[k]  # (i, j, a, b)
        cur = np.einsum("IJa,ijab->IiJjb", cur, c)
        rows *= c.shape[0]
        cols *= c.shape[1]
        cur = cur.reshape(rows, cols, c.shape[3])
    return cur[:, :, 0]


def ttv_decomp(tensor: np.ndarray, index: int = 1, tol: float = 1e-12) -> TTvector:
    """src/tt_tools.jl:186-233 (HSVD, root at `index`, 1-based; absolute threshold s >= tol)."""
    dims = tensor.shape
    d = len(dims)
    T = tensor.dtype
    vec = [None] * d
    ot = [-1] * d
    ot[index - 1] = 0
    for j in range(index, d):
        ot[j] = 1
    rks = [1] * (d + 1)
    cur = np.asfortranarray(tensor)
    for i in range(index - 1):
        cur = np.reshape(cur, (rks[i] * dims[i], -1), order="F")
        u, s, vh = np.linalg.svd(cur, full_matrices=False)
        rks[i + 1] = int(np.sum(s >= tol))
        core = np.zeros((dims[i], rks[i], rks[i + 1]), dtype=T)
        for x in range(dims[i]):
            core[x] = u[rks[i] * x: rks[i] * (x + 1), :rks[i + 1]]
        vec[i] = core
        cur = s[:rks[i + 1], None] * vh[:rks[i + 1], :]
    for i in range(d - 1, index - 1, -1):
        cur = np.reshape(cur, (-1, dims[i] * rks[i + 1]), order="F")
        u, s, vh = np.linalg.svd(cur, full_matrices=False)
        rks[i] = int(np.sum(s >= tol))
        core = np.zeros((dims[i], rks[i], rks[i + 1]), dtype=T)
        for x in range(dims[i]):
            cols = dims[i] * np.arange(rks[i + 1]) + x
            core[x] = vh[:rks[i], cols]
        vec[i] = core
        cur = u[:, :rks[i]] * s[None, :rks[i]]
    i = index - 1
    cur = np.reshape(cur, (dims[i] * rks[i], -1), order="F")
    core = np.zeros((dims[i], rks[i], rks[i + 1]), dtype=T)
    for x in range(dims[i]):
        core[x] = cur[rks[i] * x: rks[i] * (x + 1), :rks[i + 1]]
    vec[i] = core
    return TTvector(d, vec, dims, rks, ot)


def increase_ranks(x: TTvector, max_bond: int, rks=None, noise: float = 0.0) -> TTvector:
    """src/tt_tools.jl:443-489 with noise == 0 (exact zero padding; the noisy branch draws from the
    Julia RNG and is off the hot path)."""
    assert noise == 0.0, "oracle implements the exact zero-padding branch only"
    d = x.N
    assert max_bond > max(x.ttv_rks), "New bond dimension too low"
    if rks is None:
        rks = [1] + [max_bond] * (d - 1) + [1]
    rks = r_and_d_to_rks(rks, x.ttv_dims, rmax=max_bond)
    out = []
    for i in range(d):
        c = np.zeros((x.ttv_dims[i], rks[i], rks[i + 1]), dtype=x.dtype)
        s = x.ttv_vec[i].shape
        c[:, :s[1], :s[2]] = x.ttv_vec[i]
        out.append(c)
    return TTvector(d, out, x.ttv_dims, rks, [0] * d)
