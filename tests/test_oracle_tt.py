"""Pins the NumPy oracle against the reference's own known-answer / dense-oracle tests (CPU only).

Each test cites the reference test (file:line under the reference repo) it ports.
"""
import numpy as np
import pytest

import ttn_oracle as o


def dense_vec(x):
    return o.ttv_to_tensor(x).reshape(-1)  # C-order of tensor[s1..sd] == big-endian site order


def test_readme_quickstart_cfg1():
    # README.md:82-102 — als_linsolve(id_tto(6), qtt_sin(6, λ=π), rand x0; sweep_count=4), rel. err ≈ 4.6e-16
    d = 6
    A = o.id_tto(d)
    b = o.qtt_sin(d, lam=np.pi)
    x0 = o.rand_tt((2,) * d, b.ttv_rks, rng=np.random.default_rng(0))
    x = o.als_linsolve(A, b, x0, sweep_count=4)
    vb = o.qtt_to_vector(b)
    assert np.linalg.norm(o.qtt_to_vector(x) - vb) / np.linalg.norm(vb) < 1e-13


def test_readme_ttv_decomp_roundtrip():
    # README.md:38-53
    rng = np.random.default_rng(1)
    t = rng.standard_normal((2, 2, 2, 2))
    for index in (1, 2, 4):
        x = o.ttv_decomp(t, index=index)
        assert np.linalg.norm(o.ttv_to_tensor(x) - t) / np.linalg.norm(t) < 1e-14


def test_qtt_sin_cos_values():
    # test/test_qtt_tools.jl — analytic QTTs vs direct evaluation; qtt_sin(d; λ) = sin(λπx) (qtt_tools.jl:138-154)
    d = 8
    xs = np.linspace(0.0, 1.0, 2 ** d)
    assert np.abs(o.qtt_to_vector(o.qtt_sin(d, lam=2.0)) - np.sin(2 * np.pi * xs)).max() < 1e-12
    assert np.abs(o.qtt_to_vector(o.qtt_cos(d, lam=3.0)) - np.cos(3 * np.pi * xs)).max() < 1e-12


def test_operator_builders_dense():
    # test/test_tt_operators.jl:320-384 — builders vs explicit matrices
    for d in (3, 5):
        n = 2 ** d
        L = o.tto_to_matrix(o.laplace_dd(d))
        assert np.allclose(L, 2 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1))
        S = o.tto_to_matrix(o.shift_op(d))
        assert np.allclose(S, np.eye(n, k=1))
        assert np.allclose(o.tto_to_matrix(o.id_tto(d)), np.eye(n))


def test_heisenberg_dense():
    # examples/heisenberg_xyz_dmrg.jl:9-19 pattern + tt_operators.jl:162-218
    d = 6
    jx, jy, jz = 1.1, 0.8, 1.2
    H = o.tto_to_matrix(o.heisenberg_xyz_tto(d, jx=jx, jy=jy, jz=jz))
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    Y = np.array([[0, -1j], [1j, 0]])
    Z = np.array([[1, 0], [0, -1]], dtype=complex)

    def site(op, k):
        mats = [np.eye(2, dtype=complex)] * d
        mats[k] = op
        out = mats[0]
        for m in mats[1:]:
            out = np.kron(out, m)
        return out

    ref = sum(jx * site(X, k) @ site(X, k + 1) + jy * site(Y, k) @ site(Y, k + 1) + jz * site(Z, k) @ site(Z, k + 1)
              for k in range(d - 1))
    assert np.abs(ref.imag).max() == 0
    assert np.allclose(H, ref.real, atol=1e-13)


def test_apply_vs_dense():
    # test/test_tt_tools.jl:345-358 (1e-10), test/test_tt_operations.jl:116-136 (1e-12)
    rng = np.random.default_rng(2)
    dims = (2, 3, 2, 2)
    A = o.rand_tto(dims, 3, rng=rng)
    x = o.rand_tt(dims, 4, rng=rng)
    y = o.apply(A, x)
    assert y.ttv_rks == [a * b for a, b in zip(A.tto_rks, x.ttv_rks)]
    ref = o.tto_to_matrix(A) @ dense_vec(x)
    assert np.linalg.norm(dense_vec(y) - ref) / np.linalg.norm(ref) < 1e-12


def test_apply_complex_and_fused_bond_order():
    # tt_operations.jl:106 — MPO bond index fastest inside the fused bond
    rng = np.random.default_rng(3)
    dims = (2, 2, 2)
    A = o.rand_tto(dims, 2, rng=rng, dtype=np.complex128)
    x = o.rand_tt(dims, 3, rng=rng, dtype=np.complex128)
    y = o.apply(A, x)
    k = 1
    yk = y.ttv_vec[k].reshape(2, A.tto_rks[k], x.ttv_rks[k], A.tto_rks[k + 1], x.ttv_rks[k + 1], order="F")
    assert np.allclose(yk, np.einsum("ijab,jnm->ianbm", A.tto_vec[k], x.ttv_vec[k]))
    ref = o.tto_to_matrix(A) @ dense_vec(x)
    assert np.linalg.norm(dense_vec(y) - ref) / np.linalg.norm(ref) < 1e-12


def test_add_scale_dot_norm():
    # test/test_tt_operations.jl — +, scalar *, dot, norm vs dense
    rng = np.random.default_rng(4)
    dims = (2, 2, 3, 2)
    x = o.rand_tt(dims, 3, rng=rng)
    y = o.rand_tt(dims, 2, rng=rng)
    assert np.allclose(dense_vec(o.add(x, y)), dense_vec(x) + dense_vec(y))
    assert np.allclose(dense_vec(o.scale(2.5, x)), 2.5 * dense_vec(x))
    assert np.allclose(dense_vec(o.sub(x, y)), dense_vec(x) - dense_vec(y))
    assert np.isclose(o.dot(x, y), dense_vec(x) @ dense_vec(y))
    assert np.isclose(o.norm(x), np.linalg.norm(dense_vec(x)))
    xc = o.rand_tt(dims, 3, rng=rng, dtype=np.complex128)
    yc = o.rand_tt(dims, 2, rng=rng, dtype=np.complex128)
    assert np.isclose(o.dot(xc, yc), np.vdot(dense_vec(xc), dense_vec(yc)))


@pytest.mark.parametrize("center", [1, 2, 3, 5])
def test_orthogonalize_reconstruction_and_orthonormality(center):
    # test/test_tt_tools.jl:981-1017 — reconstruction + orthonormality ≤ 1e-12; ot flags (tt_tools.jl:519,529,537)
    rng = np.random.default_rng(5)
    dims = (2, 3, 2, 2, 2)
    x = o.rand_tt(dims, 4, rng=rng)
    y = o.orthogonalize(x, i=center)
    assert np.linalg.norm(dense_vec(y) - dense_vec(x)) / np.linalg.norm(dense_vec(x)) < 1e-12
    for j in range(1, center):
        G = y.ttv_vec[j - 1]
        M = np.reshape(np.transpose(G, (1, 0, 2)), (-1, G.shape[2]), order="F")
        assert np.abs(M.T @ M - np.eye(M.shape[1])).max() < 1e-12
        assert y.ttv_ot[j - 1] == 1
    for j in range(center + 1, len(dims) + 1):
        G = y.ttv_vec[j - 1]
        M = np.reshape(np.transpose(G, (1, 2, 0)), (G.shape[1], -1), order="F")
        assert np.abs(M @ M.T - np.eye(M.shape[0])).max() < 1e-12
        assert y.ttv_ot[j - 1] == -1
    assert y.ttv_ot[center - 1] == 0


def test_orthogonalize_bad_center():
    # tt_tools.jl:513
    x = o.rand_tt((2, 2, 2), 2)
    with pytest.raises(AssertionError):
        o.orthogonalize(x, i=0)
    with pytest.raises(AssertionError):
        o.orthogonalize(x, i=4)


def test_r_and_d_to_rks_overflow():
    # test/test_tt_tools.jl:945-946 — overflowing prod(dims) is treated as "no bound"
    d = 70
    rks = o.r_and_d_to_rks([1024] * (d + 1), (2,) * d, rmax=1024)
    assert rks[0] == 1 and rks[-1] == 1
    assert rks[1] == 2 and rks[10] == 1024 and rks[35] == 1024 and rks[d - 1] == 2
    assert o.r_and_d_to_rks([8] * 5, (2, 2, 2, 2)) == [1, 2, 4, 2, 1]


def test_svdtrunc_rules():
    # test/test_tdvp.jl:28-44 — truncerr = 0 rank/σ checks; Appendix B tail-norm rule
    rng = np.random.default_rng(6)
    A = rng.standard_normal((6, 4))
    U, s, Vt = o.svdtrunc(A, max_bond=100, truncerr=0.0)
    assert U.shape == (6, 4) and Vt.shape == (4, 4) and len(s) == 4
    U2, s2, Vt2 = o.svdtrunc(A, max_bond=2)
    assert len(s2) == 2
    assert np.allclose(s2, np.linalg.svd(A, compute_uv=False)[:2], rtol=1e-12)
    # tail-norm: s = [1, 1e-3, 1e-6]; truncerr 1e-4 keeps 2 (tail {1e-6} has norm ≤ 1e-4·‖s‖, tail {1e-3,1e-6} not)
    Q1, _ = np.linalg.qr(rng.standard_normal((5, 3)))
    Q2, _ = np.linalg.qr(rng.standard_normal((4, 3)))
    B = Q1 @ np.diag([1.0, 1e-3, 1e-6]) @ Q2.T
    assert len(o.svdtrunc(B, truncerr=1e-4)[1]) == 2
    assert len(o.svdtrunc(B, truncerr=1e-7)[1]) == 3
    assert len(o.svdtrunc(B, truncerr=0.5)[1]) == 1
    # the shadowed absolute rule (tt_tools.jl:737-741) would answer differently: count(s >= truncerr)
    assert len(o.svdtrunc_abs(B, truncerr=1e-4)[1]) == 2
    assert len(o.svdtrunc_abs(B, truncerr=0.5)[1]) == 1
    assert len(o.svdtrunc_abs(10 * B, truncerr=0.5)[1]) == 1
    assert len(o.svdtrunc(10 * B, truncerr=1e-4)[1]) == 2  # relative rule is scale invariant


def test_bond_truncate_kats():
    # test/test_tt_tools.jl:433-497
    rng = np.random.default_rng(7)
    tt = o.TTvector(3, [rng.standard_normal((2, 1, 4)), rng.standard_normal((2, 4, 4)), rng.standard_normal((2, 4, 1))],
                    (2, 2, 2), [1, 4, 4, 1], [0, 0, 0])
    y = o.tt_bond_truncate(tt, 1, max_bond=2)
    assert tt.ttv_rks[1] <= 2
    assert tt.ttv_vec[0].shape == (2, 1, tt.ttv_rks[1]) and tt.ttv_vec[1].shape == (2, tt.ttv_rks[1], 4)
    assert y.ttv_rks[1] == tt.ttv_rks[1] and y.ttv_vec[0].shape == tt.ttv_vec[0].shape
    # exact rank-1
    u, v, p, q = np.array([1.2, -0.5]), np.array([0.7, 0.3]), np.array([2.0, 3.0]), np.array([4.0, 5.0])
    c1 = np.einsum("s,g->sg", u, p).reshape(2, 1, 2)
    c2 = np.einsum("s,g->sg", v, q).reshape(2, 2, 1)
    t2 = o.TTvector(2, [c1, c2], (2, 2), [1, 2, 1], [0, 0])
    ref = o.ttv_to_tensor(t2).copy()
    o.tt_bond_truncate(t2, 1, max_bond=1)
    assert t2.ttv_rks[1] == 1
    assert np.allclose(o.ttv_to_tensor(t2), ref)
    with pytest.raises(AssertionError):
        o.tt_bond_truncate(tt, 0)
    with pytest.raises(AssertionError):
        o.tt_bond_truncate(tt, tt.N)


def test_tt_compress_behaviour():
    # test/test_tt_tools.jl:500-574 — no-op for large max_bond, shape updates, sweeps ≥ 1
    rng = np.random.default_rng(8)
    tt = o.rand_tt((2, 2, 2), [1, 2, 2, 1], rng=rng)
    ref = o.ttv_to_tensor(tt).copy()
    before = list(tt.ttv_rks)
    y = o.tt_compress(tt, 10, sweeps=1)
    assert y is tt and tt.ttv_rks == before
    assert np.allclose(o.ttv_to_tensor(tt), ref)
    tt4 = o.rand_tt((2, 2, 2, 2), [1, 2, 4, 2, 1], rng=rng)
    o.tt_compress(tt4, 2)
    assert max(tt4.ttv_rks) <= 2
    for k in range(4):
        assert tt4.ttv_vec[k].shape == (2, tt4.ttv_rks[k], tt4.ttv_rks[k + 1])
    with pytest.raises(AssertionError):
        o.tt_compress(tt4, 2, sweeps=0)


def test_tt_compress_faithful_equals_fast():
    # the orthogonalize inside _tt_bond_truncate! (tt_tools.jl:769) is pure and discarded
    rng = np.random.default_rng(9)
    a = o.rand_tt((2,) * 8, 8, rng=rng, normalise=True)
    b = o.copy_tt(a)
    o.tt_compress(a, 3, faithful=True)
    o.tt_compress(b, 3, faithful=False)
    for ca, cb in zip(a.ttv_vec, b.ttv_vec):
        assert np.array_equal(ca, cb)


def test_tt_compress_accuracy_on_function_qtt():
    # test/test_qtt_multidim.jl:577-614 pattern: rank-inflated smooth function compresses back exactly
    d = 8
    s = o.qtt_sin(d, lam=1.0)
    c = o.qtt_cos(d, lam=2.0)
    f = o.add(o.add(s, c), s)  # rank 6, true rank ≤ 4
    ref = o.qtt_to_vector(f).copy()
    o.tt_compress(f, 4)
    assert max(f.ttv_rks) <= 4
    assert np.linalg.norm(o.qtt_to_vector(f) - ref) / np.linalg.norm(ref) < 1e-12
    g = o.add(o.add(s, c), s)
    o.tt_compress(g, 100, truncerr=1e-10)
    assert max(g.ttv_rks) <= 4
    assert np.linalg.norm(o.qtt_to_vector(g) - ref) / np.linalg.norm(ref) < 1e-9


def test_laplace2d_interleaved_dense():
    # test/test_qtt_multidim.jl:244-267 — interleaved 2-D Laplacian equals the bit-permuted kron(Δ,I)+kron(I,Δ)
    bits = 3
    A = o.laplace2d_interleaved(bits, scaled=False)
    M = o.tto_to_matrix(A)
    n = 2 ** bits
    L = 2 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1)
    K = np.kron(L, np.eye(n)) + np.kron(np.eye(n), L)   # index (x, y), x major

    def inter(ix, iy):
        idx = 0
        for k in range(bits):
            idx = (idx << 1) | ((ix >> (bits - 1 - k)) & 1)
            idx = (idx << 1) | ((iy >> (bits - 1 - k)) & 1)
        return idx

    perm = np.array([inter(ix, iy) for ix in range(n) for iy in range(n)])
    ref = np.zeros_like(K)
    ref[np.ix_(perm, perm)] = K
    assert np.allclose(M, ref)
    assert max(A.tto_rks) <= 6
    b = o.qtt_sin2d_interleaved(bits)
    xs = np.linspace(0, 1, n)
    f = np.zeros(n * n)
    for ix in range(n):
        for iy in range(n):
            f[inter(ix, iy)] = np.sin(np.pi * xs[ix]) * np.sin(np.pi * xs[iy])
    assert np.allclose(o.qtt_to_vector(b), f)


def _golden():
    import _golden
    return _golden.load()     # Julia-made outputs (tests/golden/make_golden.jl) take precedence when present


def _tt_from(g, prefix, d, dims=None):
    rks = [int(v) for v in g[prefix + "_rks"]]
    dims = (2,) * d if dims is None else tuple(int(v) for v in dims)
    return o.TTvector(d, [g[f"{prefix}_core{k}"] for k in range(d)], dims, rks, [0] * d)


def test_oracle_reproduces_golden():
    """tests/golden/hotpath_golden.npz (generated by tests/golden/make_golden.py from the oracle and tied to dense ground
    truth there): the oracle must keep reproducing it bit-for-bit-ish (1e-13), so that the GPU parity tests and the
    committed vectors stay anchored to the same mathematics."""
    g = _golden()
    x0 = _tt_from(g, "cfg1_x0", 6)
    x = o.als_linsolve(o.id_tto(6), o.qtt_sin(6, lam=np.pi), x0, sweep_count=4)
    assert np.linalg.norm(o.ttv_to_tensor(x).reshape(-1) - g["cfg1_x"]) < 1e-13 * np.linalg.norm(g["cfg1_x"])
    y = _tt_from(g, "cmp_in", 8)
    sig = []
    z = o.tt_compress(o.copy_tt(y), 5, sigma_out=sig)
    assert list(z.ttv_rks) == [int(v) for v in g["cmp_out_rks"]]
    assert np.linalg.norm(o.ttv_to_tensor(z).reshape(-1) - g["cmp_out"]) < 1e-12 * np.linalg.norm(g["cmp_out"])
    for k, s in enumerate(sig):
        assert np.allclose(s, g["cmp_sigma"][k][:len(s)], rtol=1e-12, atol=1e-14)
    assert np.allclose(o.dmrg_matvec2(g["mv_G"], g["mv_Am"], g["mv_V"], g["mv_H"], symmetrize=False), g["mv_Y"], rtol=1e-13, atol=1e-13)
    U, S, Vt = o.svdtrunc(g["svd_A"])
    sv = np.diag(np.asarray(S)) if np.ndim(S) == 2 else np.asarray(S)
    assert np.abs(sv[:10] - g["svd_s"]).max() < 1e-14 and np.abs(sv[10:]).max() < 1e-14
    # site surgery, generalized ALS and the QFT example (cases 8-12 of make_golden.py)
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    hz = o.hadamard_ttm(_tt_from(g, "had_x", 6, g["had_dims"]), _tt_from(g, "had_y", 6, g["had_dims"]), tol=1e-12)
    assert list(hz.ttv_rks) == [int(v) for v in g["had_out_rks"]] and rel(o.ttv_to_tensor(hz).reshape(-1), g["had_out"]) < 1e-12
    ry = o.reorder(_tt_from(g, "reo_x", 6), 2, 3, "serial", "interleaved")
    assert list(ry.ttv_rks) == [int(v) for v in g["reo_out_rks"]] and rel(o.ttv_to_tensor(ry).reshape(-1), g["reo_out"]) < 1e-12
    qq = o.to_qtt(_tt_from(g, "qtt_x", 3, g["qtt_dims"]), [[2, 2, 2], [4], [3, 2]])
    assert list(qq.ttv_rks) == [int(v) for v in g["qtt_out_rks"]] and rel(o.ttv_to_tensor(qq).reshape(-1), g["qtt_out"]) < 1e-12
    Ag = o.tto_add(o.laplace_dd(5), o.tto_scale(2.0, o.id_tto(5)))
    Sg = o.tto_add(o.id_tto(5), o.tto_scale(-0.15, o.tto_add(o.laplace_dd(5), o.tto_scale(-2.0, o.id_tto(5)))))
    Eg, _ = o.als_gen_eigsolv(Ag, Sg, _tt_from(g, "gen_x0", 5), sweep_schedule=[4], rmax_schedule=[2])
    assert np.abs(Eg - g["gen_E"]).max() < 1e-11
    d, r, coeffs, F, x = _dft_example()
    assert np.array_equal(coeffs, g["dft_coeffs"])
    assert rel(o.matricize(o.tt_compress(o.apply(F, x), 100), d), g["dft_spec"]) < 1e-11


def test_hadamard_vs_dense():
    # tt_operations.jl:343-360 (test_tt_operations.jl hadamard checks): element-wise product of the dense tensors
    rng = np.random.default_rng(9)
    x = o.rand_tt((2, 3, 2, 2), 3, rng=rng); y = o.rand_tt((2, 3, 2, 2), 2, rng=rng)
    z = o.hadamard(x, y)
    assert np.allclose(o.ttv_to_tensor(z), o.ttv_to_tensor(x) * o.ttv_to_tensor(y), rtol=1e-13, atol=1e-14)
    assert z.ttv_rks == [a * b for a, b in zip(x.ttv_rks, y.ttv_rks)]


def _dft_example():
    """examples/dft.jl:5-17 with NumPy's RNG in place of Julia's."""
    d, K, r = 10, 50, 12
    rng = np.random.default_rng(1234)
    coeffs = rng.standard_normal(r) + 1j * rng.standard_normal(r)
    f = lambda x: np.sum(coeffs * np.exp(2j * np.pi * np.arange(r) * x))
    return d, r, coeffs, o.fourier_qtto(d, K=K, sign=-1.0, normalize=True), o.function_to_qtt_uniform(f, d)


def test_dft_example_known_answer():
    # examples/dft.jl:17-25: QFT of a 12-mode signal through `A*x` + `tt_compress!`, spectrum read in bit-reversed order
    d, r, coeffs, F, x = _dft_example()
    y = o.tt_compress(o.apply(F, x), 100)
    spec = o.matricize(y, d)
    scale = np.sqrt(2.0 ** d)
    assert np.linalg.norm(spec[:r] - scale * coeffs) / (scale * np.linalg.norm(coeffs)) < 1e-8
    assert np.linalg.norm(spec[r:]) / np.linalg.norm(spec) < 1e-10
    # the operator is the DFT matrix with reversed output bits
    Fm = o.tto_to_matrix(o.fourier_qtto(4, K=25))
    n = 16
    rev = [int(format(i, "04b")[::-1], 2) for i in range(n)]
    W = np.exp(-2j * np.pi * np.outer(np.arange(n), np.arange(n)) / n) / np.sqrt(n)
    # tto_to_matrix is big-endian (site 1 = most significant bit): the input of the QFT operator is little-endian
    # (function_to_qtt_uniform), its output big-endian (matricize), so the columns are the bit-reversed ones
    assert np.allclose(Fm[:, rev], W, atol=1e-10)
