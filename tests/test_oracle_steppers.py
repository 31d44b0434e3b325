"""Oracle restatement of the time steppers (src/solvers/euler.jl) against dense linear algebra — ports of
test/test_euler.jl:5-135 (one step of each scheme on the 1-D QTT Laplacian, tolerances as in the reference)."""
import numpy as np

import ttn_oracle as o


def _setup(d=4, seed=0):
    h = 1.0 / d ** 2
    A = o.tto_scale(-h ** 2, o.toeplitz_to_qtto(-2.0, 1.0, 1.0, d))
    u0 = o.rand_tt((2,) * d, [1] + [2] * (d - 1) + [1], rng=np.random.default_rng(seed))
    Ad = o.tto_to_matrix(A)
    return A, u0, Ad, o.ttv_to_tensor(u0).reshape(-1)


def _vec(x):
    return o.ttv_to_tensor(x).reshape(-1)


def _rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def test_euler_single_step_vs_dense():
    # test_euler.jl:5-32
    A, u0, Ad, ud = _setup()
    sol = o.euler_method(A, u0, [0.05], normalize=False)
    assert _rel(_vec(sol), ud + 0.05 * (Ad @ ud)) < 1e-6


def test_implicit_euler_dmrg_vs_dense():
    # test_euler.jl:34-59
    A, u0, Ad, ud = _setup()
    sol = o.implicit_euler_method(A, u0, u0, [0.05], normalize=False, tt_solver="dmrg")
    ref = np.linalg.solve(np.eye(Ad.shape[0]) - 0.05 * Ad, ud)
    assert _rel(_vec(sol), ref) < 1e-5


def test_crank_nicholson_mals_vs_dense():
    # test_euler.jl:88-111
    A, u0, Ad, ud = _setup()
    sol = o.crank_nicholson_method(A, u0, u0, [0.05], normalize=False, tt_solver="mals")
    I = np.eye(Ad.shape[0])
    ref = np.linalg.solve(I - 0.025 * Ad, (I + 0.025 * Ad) @ ud)
    assert _rel(_vec(sol), ref) < 1e-5


def test_rk4_step_vs_dense_and_error_flag():
    A, u0, Ad, ud = _setup()
    h = 0.05
    sol, err = o.rk4_method(A, u0, [h], 16, normalize=False, return_error=True)
    k1 = Ad @ ud; k2 = Ad @ (ud + h / 2 * k1); k3 = Ad @ (ud + h / 2 * k2); k4 = Ad @ (ud + h * k3)
    ref = ud + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    assert _rel(_vec(sol), ref) < 1e-10
    assert err < 1e-10


def test_euler_normalize_and_return_error():
    A, u0, Ad, ud = _setup(seed=3)
    sol, err = o.euler_method(A, u0, [0.01, 0.01], normalize=True, return_error=True)
    v = ud
    for h in (0.01, 0.01):
        v = v + h * (Ad @ v)
        v = v / np.linalg.norm(v)
    assert _rel(_vec(sol), v) < 1e-10
    assert abs(np.linalg.norm(_vec(sol)) - 1.0) < 1e-12
    assert err < 1.0


def _grad(d):
    return o.tto_scale(0.1, o.toeplitz_to_qtto(1.0, 0.0, -1.0, d))     # 0.1 * nabla(d), tt_operators.jl:276-278


def test_implicit_euler_krylov_vs_dense():
    # test_euler.jl:61-86 (GMRES through the TT vector interface, tol 1e-12, rel. error < 1e-8)
    A, u0, Ad, ud = _setup()
    sol = o.implicit_euler_method(A, u0, u0, [0.05], normalize=False, tt_solver="krylov", tol=1e-12)
    ref = np.linalg.solve(np.eye(Ad.shape[0]) - 0.05 * Ad, ud)
    assert _rel(_vec(sol), ref) < 1e-8


def test_crank_nicholson_krylov_nonsymmetric_vs_dense():
    # test_euler.jl:139-166
    d = 4
    A = _grad(d)
    u0 = o.rand_tt((2,) * d, [1] + [2] * (d - 1) + [1], rng=np.random.default_rng(1))
    Ad = o.tto_to_matrix(A); ud = _vec(u0)
    assert not np.allclose(Ad, Ad.T)
    sol = o.crank_nicholson_method(A, u0, u0, [0.05], normalize=False, tt_solver="krylov", tol=1e-12)
    I = np.eye(Ad.shape[0])
    ref = np.linalg.solve(I - 0.025 * Ad, (I + 0.025 * Ad) @ ud)
    assert _rel(_vec(sol), ref) < 1e-8


def test_crank_nicholson_bounded_bicgstab_vs_dense():
    # test_euler.jl:168-200
    d, max_bond = 5, 8
    A = _grad(d)
    u0 = o.rand_tt((2,) * d, [1] + [2] * (d - 1) + [1], rng=np.random.default_rng(2))
    Ad = o.tto_to_matrix(A); ud = _vec(u0)
    sol = o.crank_nicholson_method(A, u0, u0, [0.05], normalize=False, tt_solver="krylov", max_bond=max_bond,
                                   krylov_solver=":bicgstab", maxiter=30, rtol=1e-10, atol=1e-12)
    I = np.eye(Ad.shape[0])
    ref = np.linalg.solve(I - 0.025 * Ad, (I + 0.025 * Ad) @ ud)
    assert _rel(_vec(sol), ref) < 1e-7
    assert max(sol.ttv_rks) <= max_bond


def test_krylov_linsolve_cg_spd():
    d = 5
    A = o.tto_add(o.laplace_dd(d), o.tto_scale(2.0, o.id_tto(d)))
    b = o.rand_tt((2,) * d, 2, rng=np.random.default_rng(4))
    x = o.krylov_linsolve(A, b, b, isposdef=True, issymmetric=True, krylovdim=10, maxiter=10, rtol=1e-12)
    ref = np.linalg.solve(o.tto_to_matrix(A), _vec(b))
    assert _rel(_vec(x), ref) < 1e-9


def test_krylov_cg_selection_and_unknown_solver():
    # test_euler.jl:203-236
    import pytest
    d = 3
    A = o.tto_scale(0.1, o.id_tto(d))
    u0 = o.rand_tt((2,) * d, [1] + [2] * (d - 1) + [1], rng=np.random.default_rng(5))
    sol = o.implicit_euler_method(A, u0, u0, [0.05], normalize=False, tt_solver="krylov", isposdef=True, issymmetric=True, tol=1e-12)
    Ad = o.tto_to_matrix(A)
    ref = np.linalg.solve(np.eye(Ad.shape[0]) - 0.05 * Ad, _vec(u0))
    assert _rel(_vec(sol), ref) < 1e-8
    with pytest.raises(ValueError):
        o.implicit_euler_method(A, u0, u0, [0.05], normalize=False, tt_solver="krylov", krylov_solver=":unknown")


def test_euler_family_normalize_and_return_error_options():
    # test_euler.jl:238-267
    d = 3
    h = 1.0 / d ** 2
    A = o.tto_scale(-h ** 2, o.toeplitz_to_qtto(-2.0, 1.0, 1.0, d))
    u0 = o.rand_tt((2,) * d, [1] + [2] * (d - 1) + [1], rng=np.random.default_rng(6))
    steps = [0.02]
    s1, e1 = o.euler_method(A, u0, steps, normalize=True, return_error=True)
    s2, e2 = o.implicit_euler_method(A, u0, u0, steps, normalize=True, return_error=True, tt_solver="krylov", tol=1e-10)
    s3, e3 = o.crank_nicholson_method(A, u0, u0, steps, normalize=True, return_error=True, tt_solver="krylov", tol=1e-10)
    s4 = o.rk4_method(A, u0, steps, 6, normalize=True)
    for s in (s1, s2, s3, s4):
        assert abs(o.norm(s) - 1.0) < 1e-10
    assert all(np.isfinite(e) for e in (e1, e2, e3))
