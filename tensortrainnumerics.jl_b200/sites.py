"""The other users of the two-site truncated split, as compositions of C-ABI calls.

  hadamard                 src/tt_operations.jl:343-360   element-wise product (exact, rank r_x r_y) through the `A*x` kernel
  hadamard_ttm             src/tt_operations.jl:366-422   truncated element-wise product by swap + contract passes
  swap_adjacent_sites_     src/qtt_tools.jl:660-694       exchange two neighbouring physical indices
  reorder                  src/qtt_tools.jl:731-774       serial <-> interleaved QTT ordering by bubble-sorted swaps
  to_qtt                   src/qtt_tools.jl:254-310       split physical indices into QTT sites

The train stays in HBM between the steps: `ttn_swap_sites`, `ttn_merge_sites_diag` and `ttn_split_site` mutate the device
handle (GEMM + QR / Jacobi SVD + re-layout copies), the host only sequences them.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from . import api as _a
from ._lib import check

_NO_CAP = 1 << 62


# ---- Hadamard product ----------------------------------------------------------------------------------
def ttv_to_diag_tto(x):
    """Diagonal TToperator of a host TTvector: D_k[i, j, a, b] = delta_ij x_k[i, a, b] (tt_operations.jl:318-338)."""
    cores = []
    for c in x.ttv_vec:
        n, rl, rr = c.shape
        D = np.zeros((n, n, rl, rr), dtype=c.dtype, order="F")
        for i in range(n):
            D[i, i] = c[i]
        cores.append(D)
    return _a.TToperator(x.N, cores, x.ttv_dims, list(x.ttv_rks))


def hadamard(x, y):
    """Element-wise product of two trains (tt_operations.jl:343-360) as `ttv_to_diag_tto(x) * y` on the device: the same
    tensor with rank r_x * r_y; the fused bond index has x's index fastest (the reference's `kron` puts y's fastest — a
    permutation of the bond basis, i.e. a gauge choice).  `x`'s cores become the MPO, so it is taken from the host."""
    if isinstance(x, _a.DeviceTT):
        x = x.download()
    assert tuple(x.ttv_dims) == tuple(y.ttv_dims), "Incompatible TT dimensions"
    return _a.apply(ttv_to_diag_tto(x), y)


def _host(x):
    return x.download() if isinstance(x, _a.DeviceTT) else x


def hadamard_ttm(x, y, tol: float = 1.0e-14, rmax: int | None = None):
    """`hadamard_ttm(x, y; tol, rmax)`, tt_operations.jl:399-422: the 2d-site chain [x_1 … x_d, y_dᵀ … y_1ᵀ] is folded by
    `_ttm_swap!` passes (truncated with the `_svdtrunc` rule) and `_ttm_contract!` steps; the result has x's dims."""
    host = not (isinstance(x, _a.DeviceTT) or isinstance(y, _a.DeviceTT))
    x, y = _host(x), _host(y)
    assert tuple(x.ttv_dims) == tuple(y.ttv_dims), "Incompatible TT dimensions"
    d = x.N
    dt = np.result_type(x.ttv_vec[0].dtype, y.ttv_vec[0].dtype)
    cores = [np.asfortranarray(c.astype(dt)) for c in x.ttv_vec]
    cores += [np.asfortranarray(np.transpose(y.ttv_vec[d - 1 - k], (0, 2, 1)).astype(dt)) for k in range(d)]
    rks = list(x.ttv_rks) + list(reversed(list(y.ttv_rks)))[1:]
    dims = tuple(x.ttv_dims) + tuple(reversed(tuple(y.ttv_dims)))
    z = _a.DeviceTT.upload(_a.TTvector(2 * d, cores, dims, rks, [0] * (2 * d)))
    lib = _lib.lib()
    cap = _NO_CAP if rmax is None else int(rmax)
    for it in range(1, d + 1):
        for j in range(d, d - it + 1, -1):                      # j = d, d-1, …, d-it+2   (1-based site of the left core)
            check(lib.ttn_swap_sites(z._h, j, 1, cap, float(tol)))
        check(lib.ttn_merge_sites_diag(z._h, d - it + 1))
    return z.download() if host else z


# ---- site swaps / QTT reordering --------------------------------------------------------------------------
def swap_adjacent_sites_(x, k: int, threshold: float = 0.0):
    """`_swap_adjacent_sites(cores[k], cores[k+1]; threshold)`, qtt_tools.jl:660-694, applied in place to sites k, k+1
    (1-based) of a device train (a host TTvector is uploaded, swapped and written back)."""
    xd, host = _a._dev(x)
    check(_lib.lib().ttn_swap_sites(xd._h, int(k), 0, _NO_CAP, float(threshold)))
    if host:
        y = xd.download()
        x.ttv_vec, x.ttv_rks, x.ttv_dims = y.ttv_vec, y.ttv_rks, tuple(y.ttv_dims)
        return x
    return xd


def bubble_sort_swaps(perm):
    """qtt_tools.jl:704-718: adjacent swap positions (1-based) that bubble-sort `perm` into ascending order."""
    p = list(perm)
    swaps = []
    n = len(p)
    for i in range(1, n + 1):
        for j in range(n - i):
            if p[j] > p[j + 1]:
                p[j], p[j + 1] = p[j + 1], p[j]
                swaps.append(j + 1)
    return swaps


class QTTvector(_a.TTvector):
    """src/qtt_tools.jl:370-379: a TTvector with the multi-dimensional QTT metadata (`n_dims`, `bits_per_dim`, `ordering` in
    {"serial", "interleaved"}).  Every TTvector routine accepts it (same field names); `reorder` and `tt_compress_`
    return it re-wrapped like qtt_tools.jl:731-786."""

    def __init__(self, tt, n_dims: int, bits_per_dim: int, ordering: str):
        assert ordering in ("interleaved", "serial"), "ordering must be :interleaved or :serial"
        assert tt.N == n_dims * bits_per_dim, "QTTvector: N must equal n_dims * bits_per_dim"
        super().__init__(tt.N, tt.ttv_vec, tt.ttv_dims, tt.ttv_rks, tt.ttv_ot)
        self.n_dims, self.bits_per_dim, self.ordering = int(n_dims), int(bits_per_dim), ordering


def reorder(x, *args, threshold: float = 0.0):
    """`reorder(q::QTTvector, new_ordering; threshold)`, qtt_tools.jl:731-774.  Either `reorder(q, new_ordering)` with a
    `QTTvector` (returns a QTTvector), or `reorder(x, n_dims, bits_per_dim, ordering, new_ordering)` with the metadata
    passed explicitly for a plain TTvector / DeviceTT (returns a new train of the same kind)."""
    if isinstance(x, QTTvector):
        (new_ordering,) = args
        y = reorder(_a.TTvector(x.N, x.ttv_vec, x.ttv_dims, x.ttv_rks, x.ttv_ot), x.n_dims, x.bits_per_dim, x.ordering,
                    new_ordering, threshold=threshold)
        return QTTvector(y, x.n_dims, x.bits_per_dim, new_ordering)
    n_dims, bits_per_dim, ordering, new_ordering = args
    assert ordering in ("interleaved", "serial") and new_ordering in ("interleaved", "serial"), \
        "ordering must be :interleaved or :serial"
    xd, host = _a._dev(x)
    out = xd.copy()
    if ordering != new_ordering:
        N = xd.N
        assert N == n_dims * bits_per_dim
        perm = [0] * N
        for dd in range(n_dims):
            for b in range(bits_per_dim):
                if ordering == "serial":
                    perm[dd * bits_per_dim + b] = b * n_dims + dd
                else:
                    perm[b * n_dims + dd] = dd * bits_per_dim + b
        lib = _lib.lib()
        for k in bubble_sort_swaps(perm):
            check(lib.ttn_swap_sites(out._h, k, 0, _NO_CAP, float(threshold)))
    return out.download() if host else out


def to_qtt(tt, split_dims, threshold: float = 0.0):
    """`to_qtt(tt, split_dims; threshold)`, qtt_tools.jl:254-310: site i is split into len(split_dims[i]) sites, big-endian
    (the first factor is the coarsest digit)."""
    xd, host = _a._dev(tt)
    dims = xd.ttv_dims
    assert len(split_dims) == xd.N, "split_dims must have one entry per TT core"
    for i, sd in enumerate(split_dims):
        assert int(np.prod(sd)) == dims[i], f"prod(split_dims[{i + 1}]) must equal {dims[i]}"
    out = xd.copy()
    lib = _lib.lib()
    site = 1
    for sd in split_dims:
        for s in sd[:-1]:
            check(lib.ttn_split_site(out._h, site, int(s), 0, _NO_CAP, float(threshold)))
            site += 1
        site += 1
    return out.download() if host else out


def dmrg_cross_superblock_split(superblock, max_bond: int, truncerr: float = 0.0, direction: str = "right"):
    """The truncated superblock split of the DMRG-cross sweeps, tt_cross_interpolation.jl:609-622 (L→R) and :636-647 (R→L):
    `U, S, Vt = _svdtrunc(reshape(superblock, r_l*s1, s2*r_g); max_bond = rmax, truncerr = tol)` on the device
    (`ttn_svdtrunc_host`, the tail-norm rule of :149-166), followed by the two core forms the reference writes at the ends of a
    sweep.  `superblock` has shape (r_l, s1, s2, r_g).  Returns (core_k, core_k1, s, U, Vt): for direction "right"
    core_k = U as (s1, r_l, r), core_k1 = S·Vt as (s2, r, r_g) (lines 618-620); for "left" core_k = U·S, core_k1 = Vt
    (lines 644-646).  The index selection (`maxvol!`) of the interior bonds stays with the caller: TT-cross sampling is host
    work and out of scope (SURVEY.md §8, rows marked out of scope)."""
    sb = np.asarray(superblock)
    r_l, s1, s2, r_g = sb.shape
    A = np.asfortranarray(sb.reshape(r_l * s1, s2 * r_g, order="F"))
    U, s, Vt = _a.svdtrunc(A, max_bond=max_bond, truncerr=truncerr)
    r = len(s)
    if direction == "right":
        core_k = np.transpose(U.reshape(r_l, s1, r, order="F"), (1, 0, 2))
        core_k1 = np.transpose((s[:, None] * Vt).reshape(r, s2, r_g, order="F"), (1, 0, 2))
    else:
        core_k = np.transpose((U * s[None, :]).reshape(r_l, s1, r, order="F"), (1, 0, 2))
        core_k1 = np.transpose(Vt.reshape(r, s2, r_g, order="F"), (1, 0, 2))
    return np.asfortranarray(core_k), np.asfortranarray(core_k1), s, U, Vt
