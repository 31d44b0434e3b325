"""GPU parity of the time steppers (src/solvers/euler.jl:76-222, compositions of hot-path calls on device-resident trains)
against the oracle and dense linear algebra (test/test_euler.jl patterns)."""
import numpy as np
import pytest

import ttn_oracle as o

pytestmark = pytest.mark.gpu


def _setup(d=4, seed=0):
    h = 1.0 / d ** 2
    A = o.tto_scale(-h ** 2, o.toeplitz_to_qtto(-2.0, 1.0, 1.0, d))
    u0 = o.rand_tt((2,) * d, [1] + [2] * (d - 1) + [1], rng=np.random.default_rng(seed))
    return A, u0, o.tto_to_matrix(A), o.ttv_to_tensor(u0).reshape(-1)


def _vec(x):
    return o.ttv_to_tensor(x).reshape(-1)


def _rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def test_operator_algebra_matches_oracle():
    import ttn_b200 as t
    A, _, Ad, _ = _setup()
    M = t.tto_add(t.id_tto(A.N), t.tto_scale(-0.3, A))
    assert np.allclose(o.tto_to_matrix(o.TToperator(M.N, M.tto_vec, M.tto_dims, M.tto_rks)), np.eye(Ad.shape[0]) - 0.3 * Ad)


@pytest.mark.parametrize("normalize", [False, True])
def test_euler_vs_oracle_and_dense(normalize):
    import ttn_b200 as t
    A, u0, Ad, ud = _setup()
    steps = [0.05, 0.02]
    sol, err = t.euler_method(A, u0, steps, normalize=normalize, return_error=True)
    solo, erro = o.euler_method(A, u0, steps, normalize=normalize, return_error=True)
    assert _rel(_vec(sol), _vec(solo)) < 1e-12
    assert abs(err - erro) < 1e-10
    v = ud
    for h in steps:
        v = v + h * (Ad @ v)
        if normalize:
            v = v / np.linalg.norm(v)
    assert _rel(_vec(sol), v) < 1e-10


@pytest.mark.parametrize("solver", ["mals", "als", "dmrg"])
def test_implicit_euler_vs_dense(solver):
    import ttn_b200 as t
    A, u0, Ad, ud = _setup()
    sol = t.implicit_euler_method(A, u0, u0, [0.05], normalize=False, tt_solver=solver)
    ref = np.linalg.solve(np.eye(Ad.shape[0]) - 0.05 * Ad, ud)
    # tolerance of test/test_euler.jl:34-59 (which uses dmrg); fixed-rank ALS after two half sweeps is only parity-checked
    assert _rel(_vec(sol), ref) < (1e-3 if solver == "als" else 1e-5)
    solo = o.implicit_euler_method(A, u0, u0, [0.05], normalize=False, tt_solver=solver)
    assert _rel(_vec(sol), _vec(solo)) < 1e-6


def test_crank_nicholson_vs_dense_and_oracle():
    import ttn_b200 as t
    A, u0, Ad, ud = _setup()
    sol, err = t.crank_nicholson_method(A, u0, u0, [0.05], normalize=False, tt_solver="mals", return_error=True)
    I = np.eye(Ad.shape[0])
    ref = np.linalg.solve(I - 0.025 * Ad, (I + 0.025 * Ad) @ ud)
    assert _rel(_vec(sol), ref) < 1e-5          # test/test_euler.jl:88-111
    solo = o.crank_nicholson_method(A, u0, u0, [0.05], normalize=False, tt_solver="mals")
    assert _rel(_vec(sol), _vec(solo)) < 1e-8
    assert err < 1e-5


def test_rk4_vs_dense_and_oracle():
    import ttn_b200 as t
    A, u0, Ad, ud = _setup(d=6, seed=5)
    h = 0.05
    sol, err = t.rk4_method(A, u0, [h, h], 16, normalize=True, return_error=True)
    solo = o.rk4_method(A, u0, [h, h], 16, normalize=True)
    assert _rel(_vec(sol), _vec(solo)) < 1e-10
    v = ud
    for _ in range(2):
        k1 = Ad @ v; k2 = Ad @ (v + h / 2 * k1); k3 = Ad @ (v + h / 2 * k2); k4 = Ad @ (v + h * k3)
        v = v + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        v = v / np.linalg.norm(v)
    assert _rel(_vec(sol), v) < 1e-10
    assert err < 1e-8


def _grad(d):
    return o.tto_scale(0.1, o.toeplitz_to_qtto(1.0, 0.0, -1.0, d))     # 0.1 * nabla(d), tt_operators.jl:276-278


def test_krylov_linsolve_solvers_vs_dense():
    """`krylov_linsolve` (euler.jl:34-74) on the device: GMRES (unbounded rank), bounded BiCGStab, CG on an SPD operator."""
    import ttn_b200 as t
    d = 5
    rng = np.random.default_rng(8)
    A = o.tto_add(o.laplace_dd(d), o.tto_scale(2.0, o.id_tto(d)))
    b = o.rand_tt((2,) * d, 2, rng=rng)
    ref = np.linalg.solve(o.tto_to_matrix(A), _vec(b))
    x = t.krylov_linsolve(A, b, b, krylov_solver=":gmres", krylovdim=12, maxiter=10, rtol=1e-12)
    assert _rel(_vec(x), ref) < 1e-9
    x = t.krylov_linsolve(A, b, b, isposdef=True, issymmetric=True, krylovdim=10, maxiter=10, rtol=1e-12)
    assert _rel(_vec(x), ref) < 1e-9
    x = t.krylov_linsolve(A, b, b, max_bond=16, maxiter=40, rtol=1e-11)          # :auto -> BiCGStab when max_bond > 0
    assert _rel(_vec(x), ref) < 1e-8 and max(x.ttv_rks) <= 16
    with pytest.raises(ValueError):
        t.krylov_linsolve(A, b, b, krylov_solver=":minres")


def test_crank_nicholson_krylov_nonsymmetric_vs_dense_and_oracle():
    # test/test_euler.jl:139-200
    import ttn_b200 as t
    d = 5
    A = _grad(d)
    u0 = o.rand_tt((2,) * d, [1] + [2] * (d - 1) + [1], rng=np.random.default_rng(2))
    Ad = o.tto_to_matrix(A); ud = _vec(u0)
    I = np.eye(Ad.shape[0])
    ref = np.linalg.solve(I - 0.025 * Ad, (I + 0.025 * Ad) @ ud)
    sol = t.crank_nicholson_method(A, u0, u0, [0.05], normalize=False, tt_solver="krylov", tol=1e-12)
    assert _rel(_vec(sol), ref) < 1e-8
    sol = t.crank_nicholson_method(A, u0, u0, [0.05], normalize=False, tt_solver="krylov", max_bond=8,
                                   krylov_solver=":bicgstab", maxiter=30, rtol=1e-10, atol=1e-12)
    assert _rel(_vec(sol), ref) < 1e-7 and max(sol.ttv_rks) <= 8
    solo = o.crank_nicholson_method(A, u0, u0, [0.05], normalize=False, tt_solver="krylov", max_bond=8,
                                    krylov_solver=":bicgstab", maxiter=30, rtol=1e-10, atol=1e-12)
    assert _rel(_vec(sol), _vec(solo)) < 1e-7


def test_krylov_cg_selection_unknown_solver_and_options():
    # test/test_euler.jl:203-267 on the device path
    import ttn_b200 as t
    d = 3
    A = o.tto_scale(0.1, o.id_tto(d))
    u0 = o.rand_tt((2,) * d, [1] + [2] * (d - 1) + [1], rng=np.random.default_rng(5))
    sol = t.implicit_euler_method(A, u0, u0, [0.05], normalize=False, tt_solver="krylov", isposdef=True, issymmetric=True, tol=1e-12)
    Ad = o.tto_to_matrix(A)
    ref = np.linalg.solve(np.eye(Ad.shape[0]) - 0.05 * Ad, _vec(u0))
    assert _rel(_vec(sol), ref) < 1e-8
    with pytest.raises(ValueError):
        t.implicit_euler_method(A, u0, u0, [0.05], normalize=False, tt_solver="krylov", krylov_solver=":unknown")
    h = 1.0 / d ** 2
    A = o.tto_scale(-h ** 2, o.toeplitz_to_qtto(-2.0, 1.0, 1.0, d))
    steps = [0.02]
    s1, e1 = t.euler_method(A, u0, steps, normalize=True, return_error=True)
    s2, e2 = t.implicit_euler_method(A, u0, u0, steps, normalize=True, return_error=True, tt_solver="krylov", tol=1e-10)
    s3, e3 = t.crank_nicholson_method(A, u0, u0, steps, normalize=True, return_error=True, tt_solver="krylov", tol=1e-10)
    s4 = t.rk4_method(A, u0, steps, 6, normalize=True)
    for s in (s1, s2, s3, s4):
        assert abs(o.norm(s) - 1.0) < 1e-10
    assert all(np.isfinite(e) for e in (e1, e2, e3))
