"""Parity of the site-surgery entry points (`ttn_swap_sites`, `ttn_merge_sites_diag`, `ttn_split_site`; SURVEY.md section
8(f)-4) against the oracle: represented tensors to 1e-10 (the tolerance of the path), ranks and singular-value rules equal."""
import numpy as np
import pytest

import ttn_oracle as o

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("cplx", [False, True])
def test_swap_adjacent_sites_vs_oracle(cplx):
    import ttn_b200 as t
    rng = np.random.default_rng(31)
    x = o.rand_tt((2, 3, 2, 4, 2), 5, rng=rng, dtype=np.complex128 if cplx else np.float64)
    full = o.ttv_to_tensor(x)
    y = o.copy_tt(x)
    t.swap_adjacent_sites_(y, 2)
    assert tuple(y.ttv_dims) == (2, 2, 3, 4, 2)
    A, B = o.swap_adjacent_sites(x.ttv_vec[1], x.ttv_vec[2])
    assert y.ttv_rks[2] == A.shape[2]
    assert _rel(o.ttv_to_tensor(y), np.transpose(full, (0, 2, 1, 3, 4))) < TOL
    # U | S Vt split: the left core is an isometry (qtt_tools.jl:687-692)
    L = y.ttv_vec[1].reshape(-1, y.ttv_rks[2], order="F")
    assert np.allclose(L.conj().T @ L, np.eye(L.shape[1]), atol=1e-12)


def test_swap_threshold_rule_matches_oracle():
    import ttn_b200 as t
    rng = np.random.default_rng(32)
    # a train whose swapped bond has a decaying spectrum: sum of a dominant and a small component
    a = o.rand_tt((2,) * 6, 2, rng=rng); b = o.scale(1e-6, o.rand_tt((2,) * 6, 3, rng=rng))
    x = o.add(a, b)
    for thr in (0.0, 1e-3, 1e-9):
        y = o.copy_tt(x)
        t.swap_adjacent_sites_(y, 3, threshold=thr)
        A, B = o.swap_adjacent_sites(x.ttv_vec[2], x.ttv_vec[3], threshold=thr)
        if thr > 0:
            assert y.ttv_rks[3] == A.shape[2]
        ref = o.TTvector(6, x.ttv_vec[:2] + [A, B] + x.ttv_vec[4:], x.ttv_dims, x.ttv_rks[:3] + [A.shape[2]] + x.ttv_rks[4:], [0] * 6)
        assert _rel(o.ttv_to_tensor(y), o.ttv_to_tensor(ref)) < (TOL if thr == 0 else 1e-8)


@pytest.mark.parametrize("n_dims,bits", [(2, 4), (3, 3)])
def test_reorder_vs_oracle(n_dims, bits):
    import ttn_b200 as t
    rng = np.random.default_rng(33)
    x = o.rand_tt((2,) * (n_dims * bits), 6, rng=rng)
    y = t.reorder(x, n_dims, bits, "serial", "interleaved")
    yo = o.reorder(x, n_dims, bits, "serial", "interleaved")
    assert list(y.ttv_rks) == list(yo.ttv_rks)
    assert _rel(o.ttv_to_tensor(y), o.ttv_to_tensor(yo)) < TOL
    z = t.reorder(y, n_dims, bits, "interleaved", "serial")
    assert _rel(o.ttv_to_tensor(z), o.ttv_to_tensor(x)) < TOL
    same = t.reorder(x, n_dims, bits, "serial", "serial")
    assert all(np.array_equal(a, b) for a, b in zip(same.ttv_vec, x.ttv_vec))


def test_reorder_2d_function_known_answer():
    """f(x, y) = sin(2 pi x) cos(2 pi y) on a 2^5 x 2^5 grid: serial QTT (x bits then y bits) -> interleaved, checked against
    the interleaved generator (test/test_qtt_tools.jl reorder round trip)."""
    import ttn_b200 as t
    bits = 5
    sx, cy = o.qtt_sin(bits, lam=1.0), o.qtt_cos(bits, lam=1.0)
    cores = [c.copy() for c in sx.ttv_vec] + [c.copy() for c in cy.ttv_vec]
    x = o.TTvector(2 * bits, cores, (2,) * (2 * bits), list(sx.ttv_rks) + list(cy.ttv_rks)[1:], [0] * (2 * bits))
    y = t.reorder(x, 2, bits, "serial", "interleaved", threshold=1e-12)
    full = o.ttv_to_tensor(x)
    axes = [0] * (2 * bits)
    for src, tgt in enumerate(o.reorder_perm(2, bits, "serial")):
        axes[tgt] = src
    assert _rel(o.ttv_to_tensor(y), np.transpose(full, axes)) < TOL
    assert max(y.ttv_rks) <= 4                                           # rank-2 x rank-2 separable function


@pytest.mark.parametrize("cplx", [False, True])
def test_hadamard_ttm_vs_oracle(cplx):
    import ttn_b200 as t
    rng = np.random.default_rng(34)
    dt = np.complex128 if cplx else np.float64
    x = o.rand_tt((2, 3, 2, 2, 2, 2), 4, rng=rng, dtype=dt); y = o.rand_tt((2, 3, 2, 2, 2, 2), 3, rng=rng, dtype=dt)
    ref = o.ttv_to_tensor(x) * o.ttv_to_tensor(y)
    z = t.hadamard_ttm(x, y, tol=1e-12)
    zo = o.hadamard_ttm(x, y, tol=1e-12)
    assert tuple(z.ttv_dims) == tuple(x.ttv_dims) and list(z.ttv_rks) == list(zo.ttv_rks)
    assert _rel(o.ttv_to_tensor(z), ref) < TOL
    assert _rel(o.ttv_to_tensor(z), o.ttv_to_tensor(zo)) < TOL
    zt = t.hadamard_ttm(x, y, tol=1e-3, rmax=5)
    zto = o.hadamard_ttm(x, y, tol=1e-3, rmax=5)
    assert list(zt.ttv_rks) == list(zto.ttv_rks)
    assert _rel(o.ttv_to_tensor(zt), o.ttv_to_tensor(zto)) < 1e-8
    with pytest.raises(AssertionError):
        t.hadamard_ttm(x, o.rand_tt((2, 2, 2, 2, 2, 2), 2, rng=rng))


def test_hadamard_ttm_qtt_functions():
    """sin^2 + cos^2 = 1 on a 2^10 grid through the truncated element-wise product."""
    import ttn_b200 as t
    s, c = o.qtt_sin(10, lam=3.0), o.qtt_cos(10, lam=3.0)
    one = t.add(t.hadamard_ttm(s, s, tol=1e-13), t.hadamard_ttm(c, c, tol=1e-13))
    v = o.qtt_to_vector(one)
    assert np.max(np.abs(v - 1.0)) < 1e-10


@pytest.mark.parametrize("cplx", [False, True])
def test_to_qtt_vs_oracle(cplx):
    import ttn_b200 as t
    rng = np.random.default_rng(35)
    x = o.rand_tt((8, 4, 6, 16), 5, rng=rng, dtype=np.complex128 if cplx else np.float64)
    split = [[2, 2, 2], [4], [3, 2], [2, 2, 2, 2]]
    q = t.to_qtt(x, split)
    qo = o.to_qtt(x, split)
    assert tuple(q.ttv_dims) == tuple(qo.ttv_dims) and list(q.ttv_rks) == list(qo.ttv_rks)
    assert _rel(o.ttv_to_tensor(q), o.ttv_to_tensor(x).reshape(q.ttv_dims)) < TOL
    qt = t.to_qtt(x, split, threshold=1e-2)
    qto = o.to_qtt(x, split, threshold=1e-2)
    assert list(qt.ttv_rks) == list(qto.ttv_rks)
    assert _rel(o.ttv_to_tensor(qt), o.ttv_to_tensor(qto)) < 1e-8
    with pytest.raises(AssertionError):
        t.to_qtt(x, [[2, 2, 2], [4], [3, 3], [16]])


def test_site_calls_reject_bad_arguments():
    import ttn_b200 as t
    rng = np.random.default_rng(36)
    xd = t.DeviceTT.upload(o.rand_tt((2, 3, 2), 2, rng=rng))
    with pytest.raises(Exception):
        t.swap_adjacent_sites_(xd, 3)                     # k must be in 1:(N-1)
    with pytest.raises(Exception):
        t._lib.check(t._lib.lib().ttn_merge_sites_diag(xd._h, 1))   # 2 != 3 physical dims
    with pytest.raises(Exception):
        t._lib.check(t._lib.lib().ttn_split_site(xd._h, 2, 2, 0, 1 << 62, 0.0))   # 2 does not divide 3


def test_hadamard_function_reconstruction_on_device():
    """test/test_tt_operations.jl:41-104 (exp / sin / cos cases) through the device `hadamard` and `hadamard_ttm`."""
    import ttn_b200 as t
    d = 8
    xs = np.linspace(0.0, 1.0, 2 ** d)
    A1, A2, A3 = o.qtt_exp(d), o.qtt_sin(d, lam=np.pi), o.qtt_cos(d, lam=np.pi)
    cases = [(A2, A3, np.cos(np.pi ** 2 * xs) * np.sin(np.pi ** 2 * xs)), (A1, A2, np.exp(xs) * np.sin(np.pi ** 2 * xs))]
    for x, y, expected in cases:
        h = t.hadamard(x, y)
        assert np.allclose(o.qtt_to_vector(h), expected, atol=1e-12)
        z = t.hadamard_ttm(x, y)
        assert np.allclose(o.qtt_to_vector(z), expected, atol=1e-10)
        assert t.norm(t.sub(z, h)) / t.norm(h) < 1e-5
    x, y, expected = cases[1]
    assert np.allclose(o.qtt_to_vector(t.hadamard_ttm(x, y, tol=1e-8)), expected, atol=1e-3)


def test_qttvector_wrapper_reorder_and_compress():
    """`QTTvector` (qtt_tools.jl:370-379): `reorder(q, new_ordering)` and `tt_compress!(q, …)` come back re-wrapped with the
    metadata (qtt_tools.jl:731-786); the serial -> interleaved -> serial round trip returns the tensor."""
    import ttn_b200 as t
    rng = np.random.default_rng(37)
    x = o.rand_tt((2,) * 6, 4, rng=rng)
    q = t.QTTvector(x, 2, 3, "serial")
    qi = t.reorder(q, "interleaved")
    assert isinstance(qi, t.QTTvector) and qi.ordering == "interleaved" and (qi.n_dims, qi.bits_per_dim) == (2, 3)
    ref = o.reorder(x, 2, 3, "serial", "interleaved")
    assert _rel(o.ttv_to_tensor(qi), o.ttv_to_tensor(ref)) < TOL
    back = t.reorder(qi, "serial")
    assert _rel(o.ttv_to_tensor(back), o.ttv_to_tensor(x)) < TOL
    same = t.reorder(q, "serial")
    assert isinstance(same, t.QTTvector) and _rel(o.ttv_to_tensor(same), o.ttv_to_tensor(x)) < 1e-15
    vec_list = qi.ttv_vec
    out = t.tt_compress_(qi, 3)
    assert out is qi and qi.ttv_vec is vec_list and max(qi.ttv_rks) <= 3       # same object, same lists mutated
    with pytest.raises(AssertionError):
        t.QTTvector(x, 2, 3, "zigzag")


@pytest.mark.gpu
@pytest.mark.parametrize("cplx", [False, True])
def test_dmrg_cross_superblock_split_vs_oracle(cplx):
    """SURVEY.md §8(f)-4, last item: the truncated superblock split of `tt_cross` with the DMRG strategy
    (tt_cross_interpolation.jl:609-622, 636-647) — `_svdtrunc` of the (r_l s1) x (s2 r_g) superblock with the tail-norm rule
    and the rmax cap, then the two end-of-sweep core forms.  Against the oracle's `svdtrunc` (gesdd): same rank, singular
    values to 1e-10, and the two cores multiply back to the rank-r truncation of the superblock (gauge-free check)."""
    import ttn_b200 as t
    rng = np.random.default_rng(44 + cplx)
    r_l, s1, s2, r_g = 12, 4, 3, 15
    # a superblock sampled from a function of numerical rank ~9 plus a tail decaying below the tolerance
    m, n = r_l * s1, s2 * r_g
    U0, _ = np.linalg.qr(rng.standard_normal((m, m)) + (1j * rng.standard_normal((m, m)) if cplx else 0))
    V0, _ = np.linalg.qr(rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if cplx else 0))
    sv = np.r_[np.logspace(0, -3, 9), np.logspace(-7, -12, min(m, n) - 9)]
    A = (U0[:, :min(m, n)] * sv) @ V0[:, :min(m, n)].conj().T
    sb = A.reshape(r_l, s1, s2, r_g, order="F")
    for max_bond, tol in ((20, 1e-6), (6, 0.0), (64, 1e-10)):
        Ur, Sr, Vtr = o.svdtrunc(A, max_bond=max_bond, truncerr=tol)
        sr = np.diag(Sr) if np.ndim(Sr) == 2 else np.asarray(Sr)
        for direction in ("right", "left"):
            ck, ck1, s, U, Vt = t.dmrg_cross_superblock_split(sb, max_bond, tol, direction=direction)
            r = len(s)
            assert r == len(sr) and np.abs(s - sr).max() < 1e-10 * sr[0]
            assert ck.shape == (s1, r_l, r) and ck1.shape == (s2, r, r_g)
            left = np.transpose(ck, (1, 0, 2)).reshape(m, r, order="F")
            right = np.transpose(ck1, (1, 0, 2)).reshape(r, n, order="F")
            ref = (Ur * sr) @ Vtr
            assert np.linalg.norm(left @ right - ref) < 1e-10 * np.linalg.norm(ref)
            if direction == "right":
                assert np.linalg.norm(left.conj().T @ left - np.eye(r)) < 1e-10       # U is the new left-orthogonal core
