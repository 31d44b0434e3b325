"""GPU parity tests of the alternating solvers (ALS / MALS / DMRG / TDVP) through the C ABI, against the oracle and
dense ground truth.  Tolerance 1e-10 on solutions / energies of well-conditioned problems; local
Krylov solves are run to convergence because converged results are algorithm independent."""
import numpy as np
import pytest
import scipy.linalg as sla

import ttn_oracle as o

pytestmark = pytest.mark.gpu


def dv(x):
    return o.ttv_to_tensor(x).reshape(-1)


def relerr(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def spd_op(d, shift=3.0):
    return o.tto_add(o.laplace_dd(d), o.tto_scale(shift, o.id_tto(d)))


def test_cfg1_readme_quickstart():
    # README.md:82-102 (cfg1): als_linsolve(id_tto(6), qtt_sin(6, λ=π), rand x0; sweep_count=4) → rel. err ~ 4.6e-16
    import ttn_b200 as t
    d = 6
    A, b = o.id_tto(d), o.qtt_sin(d, lam=np.pi)
    x0 = o.rand_tt((2,) * d, b.ttv_rks, rng=np.random.default_rng(0))
    x = t.als_linsolve(A, b, x0, sweep_count=4)
    vb = o.qtt_to_vector(b)
    assert relerr(o.qtt_to_vector(x), vb) < 1e-12
    xo = o.als_linsolve(A, b, x0, sweep_count=4)
    assert relerr(o.qtt_to_vector(x), o.qtt_to_vector(xo)) < 1e-10
    assert x.ttv_rks == xo.ttv_rks


def test_als_linsolve_vs_dense_and_oracle():
    import ttn_b200 as t
    d = 5
    rng = np.random.default_rng(10)
    A = spd_op(d, 3.0)
    b = o.rand_tt((2,) * d, 2, rng=rng)
    x0 = o.rand_tt((2,) * d, 4, rng=rng)
    x, info = t.als_linsolve(A, b, x0, sweep_count=6, return_info=True)
    ref = np.linalg.solve(o.tto_to_matrix(A), dv(b))
    assert relerr(dv(x), ref) < 1e-10
    assert info["residual"] < 1e-7    # dot-based norm of the residual TT: sqrt(eps) floor, as in the reference
    # reduced rank: same fixed point as the oracle after the same number of half sweeps
    x0r = o.rand_tt((2,) * d, 2, rng=rng)
    xr = t.als_linsolve(A, b, x0r, sweep_count=8)
    xo = o.als_linsolve(A, b, x0r, sweep_count=8)
    assert relerr(dv(xr), dv(xo)) < 1e-9


def test_als_eigsolve_vs_dense():
    import ttn_b200 as t
    d = 5
    rng = np.random.default_rng(12)
    A = spd_op(d, 1.0)
    x0 = o.rand_tt((2,) * d, 4, rng=rng)
    E, x = t.als_eigsolve(A, x0, sweep_schedule=[6], rmax_schedule=[4], linsolv_tol=1e-13)
    lam = sla.eigvalsh(o.tto_to_matrix(A))[0]
    assert abs(E[-1] - lam) < 1e-10
    v = dv(x)
    assert abs(v @ o.tto_to_matrix(A) @ v / (v @ v) - lam) < 1e-10
    with pytest.raises(AssertionError):
        t.als_eigsolve(A, x0, sweep_schedule=[2, 3], rmax_schedule=[4])     # als.jl:263


def test_als_eigsolve_noise_schedule():
    # als.jl:289-291: a rank-1 start cannot leave its manifold under one-site ALS; the noisy rank increase of the second stage gives it
    # the directions the ground state needs.  E is the concatenation of the stages' histories: 2 (d-1) entries per sweep, als.jl:299-311
    import ttn_b200 as t
    d = 6
    A = spd_op(d, 1.0)
    x0 = o.rand_tt((2,) * d, 1, rng=np.random.default_rng(3))
    lam = sla.eigvalsh(o.tto_to_matrix(A))[0]
    E1, _ = t.als_eigsolve(A, x0, sweep_schedule=[8], rmax_schedule=[1], linsolv_tol=1e-13)
    E, x = t.als_eigsolve(A, x0, sweep_schedule=[3, 9], rmax_schedule=[1, 6], noise_schedule=[0.0, 1e-2], linsolv_tol=1e-13,
                          rng=np.random.default_rng(4))
    assert len(E) == 2 * (d - 1) * (9 - 1)                     # sweeps 1..8 of the reference's counter
    assert max(x.ttv_rks) == 6
    assert abs(E[-1] - lam) < 1e-9 < abs(E1[-1] - lam)         # rank 1 alone stays away from the ground state
    v = dv(x)
    assert abs(v @ o.tto_to_matrix(A) @ v / (v @ v) - lam) < 1e-9
    with pytest.raises(AssertionError):
        t.als_eigsolve(A, x0, sweep_schedule=[2, 3], rmax_schedule=[4], noise_schedule=[0.0, 0.1])     # als.jl:263


def test_mals_linsolve_vs_dense():
    import ttn_b200 as t
    d = 6
    rng = np.random.default_rng(13)
    A = spd_op(d, 3.0)
    b = o.rand_tt((2,) * d, 2, rng=rng)
    x0 = o.rand_tt((2,) * d, 2, rng=rng)
    x, info = t.mals_linsolve(A, b, x0, tol=1e-14, rmax=8, return_info=True)
    ref = np.linalg.solve(o.tto_to_matrix(A), dv(b))
    assert max(x.ttv_rks) <= 8
    assert relerr(dv(x), ref) < 1e-9
    xo = o.mals_linsolve(A, b, x0, tol=1e-14, rmax=8)
    assert relerr(dv(x), dv(xo)) < 1e-9
    x2 = t.mals_linsolve(A, b, x0, tol=1e-14, rmax=3)
    assert max(x2.ttv_rks) <= 3
    xo2 = o.mals_linsolve(A, b, x0, tol=1e-14, rmax=3)
    assert x2.ttv_rks == xo2.ttv_rks and relerr(dv(x2), dv(xo2)) < 1e-8


def test_mals_eigsolve_vs_dense():
    import ttn_b200 as t
    d = 6
    rng = np.random.default_rng(14)
    A = spd_op(d, 1.0)
    x0 = o.rand_tt((2,) * d, 2, rng=rng)
    E, x, rh = t.mals_eigsolve(A, x0, tol=1e-12, sweep_schedule=[4], rmax_schedule=[8], linsolv_tol=1e-13)
    lam = sla.eigvalsh(o.tto_to_matrix(A))[0]
    assert abs(E[-1] - lam) < 1e-9 and len(rh) == len(E)


@pytest.mark.parametrize("N", [1, 2])
def test_dmrg_linsolve_vs_dense(N):
    import ttn_b200 as t
    d = 5
    rng = np.random.default_rng(15)
    A = spd_op(d, 3.0)
    b = o.rand_tt((2,) * d, 2, rng=rng)
    x0 = o.rand_tt((2,) * d, 2, rng=rng)
    x, info = t.dmrg_linsolve(A, b, x0, N=N, sweep_schedule=[8], rmax_schedule=[4], linsolv_tol=1e-14, return_info=True)
    ref = np.linalg.solve(o.tto_to_matrix(A), dv(b))
    assert relerr(dv(x), ref) < 1e-9 and info["residual"] < 1e-7   # norm(A*x-b) via dot() bottoms out at sqrt(eps), as in the reference


@pytest.mark.parametrize("sym", [False, True])
def test_dmrg_eigsolve_heisenberg_vs_dense(sym):
    # examples/heisenberg_xyz_dmrg.jl:9-19 — DMRG energy vs eigvals(dense H)
    import ttn_b200 as t
    d = 10
    rng = np.random.default_rng(16)
    H = o.heisenberg_xyz_tto(d, jx=1.1, jy=0.8, jz=1.2)
    x0 = o.rand_tt((2,) * d, 4, rng=rng, normalise=True)
    E, x, rh = t.dmrg_eigsolve(H, x0, N=2, tol=1e-12, sweep_schedule=[5], rmax_schedule=[32], linsolv_tol=1e-12,
                               symmetrize=sym)
    lam = sla.eigvalsh(o.tto_to_matrix(H))[0]
    assert abs(E[-1] - lam) < 1e-10 * abs(lam)
    assert max(rh) <= 32 and len(rh) == len(E) == 2 * (d - 2) * 4 + 1
    v = dv(x)
    assert abs(v @ o.tto_to_matrix(H) @ v / (v @ v) - lam) < 1e-10 * abs(lam)
    with pytest.raises(AssertionError):
        t.dmrg_eigsolve(H, x0, sweep_schedule=[2, 3], rmax_schedule=[4])    # dmrg.jl:513


def test_dmrg_eigsolve_one_site():
    import ttn_b200 as t
    d = 6
    rng = np.random.default_rng(17)
    A = spd_op(d, 1.0)
    x0 = o.rand_tt((2,) * d, 8, rng=rng, normalise=True)
    E, x, rh = t.dmrg_eigsolve(A, x0, N=1, sweep_schedule=[6], rmax_schedule=[8], linsolv_tol=1e-13)
    lam = sla.eigvalsh(o.tto_to_matrix(A))[0]
    assert abs(E[-1] - lam) < 1e-9


@pytest.mark.parametrize("two_site", [False, True])
def test_tdvp_heat_eigenmode(two_site):
    # test/test_tdvp.jl:329-356 — imaginary-time evolution of an eigenmode follows exp(λ Σh) u0 (< 1e-8)
    import ttn_b200 as t
    d = 5
    A = o.tto_scale(-1.0, o.laplace_dd(d))
    w, v = sla.eigh(o.tto_to_matrix(A))
    u0d = v[:, -1]
    u0 = o.ttv_decomp(u0d.reshape((2,) * d), index=1, tol=1e-14)
    steps = [0.01] * 5
    if two_site:
        psi = t.tdvp2(A, u0, steps, normalize=False, imaginary_time=True, max_bond=8)
        ref_o = o.tdvp2(A, u0, steps, normalize=False, imaginary_time=True, max_bond=8)
    else:
        psi = t.tdvp(A, u0, steps, normalize=False, imaginary_time=True)
        ref_o = o.tdvp(A, u0, steps, normalize=False, imaginary_time=True)
    ref = np.exp(w[-1] * sum(steps)) * u0d
    assert relerr(dv(psi), ref) < 1e-8
    assert relerr(dv(psi), dv(ref_o)) < 1e-9
    assert psi.dtype == np.float64


@pytest.mark.parametrize("two_site", [False, True])
def test_tdvp_real_time_vs_oracle(two_site):
    # test/test_tdvp.jl:132-160,238-261: H = 0 keeps the state; real-time sweeps match the oracle's exact local exponentials
    import ttn_b200 as t
    d = 4
    rng = np.random.default_rng(19)
    u0 = o.rand_tt((2,) * d, 4, rng=rng, normalise=True)
    Z = o.tto_scale(0.0, o.id_tto(d))
    psi = t.tdvp(Z, u0, [0.1, 0.1], normalize=False)
    assert psi.dtype == np.complex128 and relerr(dv(psi), dv(u0)) < 1e-12
    H = o.heisenberg_xyz_tto(d, jx=1.0, jy=0.7, jz=0.3)
    steps = [0.05] * 3
    if two_site:
        got = t.tdvp2(H, u0, steps, normalize=True, max_bond=4)
        ref = o.tdvp2(H, u0, steps, normalize=True, max_bond=4)
    else:
        got = t.tdvp(H, u0, steps, normalize=True)
        ref = o.tdvp(H, u0, steps, normalize=True)
    assert relerr(dv(got), dv(ref)) < 1e-9
    assert abs(np.linalg.norm(dv(got)) - 1.0) < 1e-10


@pytest.mark.parametrize("bits", [3, 5, 6])
def test_cfg3_laplace2d_interleaved_mals_linsolve(bits):
    """cfg3 in miniature: interleaved 2-D Laplace QTT operator (tt_operators.jl:654-656 scaling),
    right-hand side sin(pi x) sin(pi y), `mals_linsolve` (mals.jl:240-309: exactly one forward + one backward two-site
    sweep from a random rank-4 start).  Parity is against the oracle running the same algorithm (dense local solves,
    mals.jl:148-169): one sweep with the squared-weight rule `sv_trunc` (mals.jl:42-56) is itself only accurate to
    ~1e-5 against the dense solution at 2 x 5 bits, so the dense check is loose and the oracle check is tight."""
    import ttn_b200 as t
    d = 2 * bits
    A = o.laplace2d_interleaved(bits)
    b = o.qtt_sin2d_interleaved(bits)
    x0 = o.rand_tt((2,) * d, 4, rng=np.random.default_rng(2), normalise=True)
    x, info = t.mals_linsolve(A, b, x0, tol=1e-12, rmax=32, return_info=True)
    xo = o.mals_linsolve(A, b, x0, tol=1e-12, rmax=32)
    assert x.ttv_rks == xo.ttv_rks
    assert relerr(dv(x), dv(xo)) < 1e-8
    Ad = o.tto_to_matrix(A)
    ref = np.linalg.solve(Ad, dv(b))
    assert relerr(dv(x), ref) < (1e-12 if bits == 3 else 1e-3)
    res_dense = np.linalg.norm(Ad @ dv(x) - dv(b)) / np.linalg.norm(dv(b))
    assert abs(info["residual"] - res_dense) < 1e-6 * max(1.0, res_dense)


def test_local_linsolve_dense_vs_gmres_paths():
    """als_linsolve with the dense direct local solve (default, K_full + `\\` of als.jl:58-70) and with it_solver=True
    (matrix-free GMRES) must agree with each other and with the dense solution."""
    import ttn_b200 as t
    d = 6
    rng = np.random.default_rng(21)
    A = spd_op(d, 2.0)
    b = o.rand_tt((2,) * d, 2, rng=rng)
    x0 = o.rand_tt((2,) * d, 4, rng=rng)
    ref = np.linalg.solve(o.tto_to_matrix(A), dv(b))
    xd = t.als_linsolve(A, b, x0, sweep_count=6)
    xg = t.als_linsolve(A, b, x0, sweep_count=6, it_solver=True)
    assert relerr(dv(xd), ref) < 1e-10
    assert relerr(dv(xg), ref) < 1e-9


def test_als_linsolve_complex_dense_local_solve():
    """ComplexF64 through the dense local solve (K assembled by the batched three-GEMM pass, complex Householder QR, complex
    back substitution): Hermitian positive definite operator with a complex right-hand side against the dense solution."""
    import ttn_b200 as t
    d = 5
    rng = np.random.default_rng(31)
    Ar = spd_op(d, 2.5)
    A = o.TToperator(Ar.N, [c.astype(np.complex128) for c in Ar.tto_vec], Ar.tto_dims, Ar.tto_rks)
    b = o.rand_tt((2,) * d, 2, rng=rng)
    b = o.TTvector(b.N, [c + 1j * rng.standard_normal(c.shape) for c in b.ttv_vec], b.ttv_dims, b.ttv_rks, b.ttv_ot)
    x0 = o.rand_tt((2,) * d, 4, rng=rng)
    x0 = o.TTvector(x0.N, [c.astype(np.complex128) for c in x0.ttv_vec], x0.ttv_dims, x0.ttv_rks, x0.ttv_ot)
    x = t.als_linsolve(A, b, x0, sweep_count=6)
    ref = np.linalg.solve(o.tto_to_matrix(A), dv(b))
    assert relerr(dv(x), ref) < 1e-10


def _spd_mass(d, eps=0.3):
    return o.tto_add(o.id_tto(d), o.tto_scale(-eps / 2.0, o.tto_add(o.laplace_dd(d), o.tto_scale(-2.0, o.id_tto(d)))))


@pytest.mark.gpu
def test_als_gen_eigsolv_vs_oracle_and_dense():
    """`als_gen_eigsolv` (als.jl:344-440; test/test_als.jl:153-197): the energies of every local solve equal the oracle's, the
    final Rayleigh quotient equals the lowest eigenvalue of the dense pencil, S = I reproduces `als_eigsolve`."""
    import ttn_b200 as t
    d = 5
    rng = np.random.default_rng(21)
    A = spd_op(d, 2.0)
    x0 = o.rand_tt((2,) * d, 4, rng=rng, normalise=True)
    S = _spd_mass(d)
    Am, Sm = o.tto_to_matrix(A), o.tto_to_matrix(S)
    lam = sla.eigh(Am, Sm, eigvals_only=True)[0]
    E, x = t.als_gen_eigsolv(A, S, x0, sweep_schedule=[6], rmax_schedule=[4])
    Eo, xo = o.als_gen_eigsolv(A, S, x0, sweep_schedule=[6], rmax_schedule=[4])
    assert len(E) == len(Eo) and abs(E[-1] - Eo[-1]) < 1e-9
    # step-by-step energies at rank 2, where every local eigenvector has a full-rank unfolding (with over-parametrised ranks the
    # QR of a rank-deficient unfolding completes the basis from rounding noise and the trajectory is not reproducible)
    x2 = o.rand_tt((2,) * d, 2, rng=rng, normalise=True)
    E2s, _ = t.als_gen_eigsolv(A, S, x2, sweep_schedule=[3], rmax_schedule=[2])
    E2o, _ = o.als_gen_eigsolv(A, S, x2, sweep_schedule=[3], rmax_schedule=[2])
    assert len(E2s) == len(E2o) and np.max(np.abs(E2s - E2o)) < 1e-9
    assert abs(E[-1] - lam) < 1e-9
    v = dv(x)
    assert abs((v @ Am @ v) / (v @ Sm @ v) - lam) < 1e-9
    vo = dv(xo)
    assert 1.0 - abs(v @ Sm @ vo) / np.sqrt((v @ Sm @ v) * (vo @ Sm @ vo)) < 1e-9
    E_gen, _ = t.als_gen_eigsolv(A, o.id_tto(d), x0, sweep_schedule=[4], rmax_schedule=[4])
    E_std, _ = t.als_eigsolve(A, x0, sweep_schedule=[4], rmax_schedule=[4], linsolv_tol=1e-13)
    assert abs(E_gen[-1] - E_std[-1]) < 1e-10
    # rank growth through the schedule (test_als.jl:184-197) and the schedule check
    E2, x2 = t.als_gen_eigsolv(spd_op(3, 2.0), o.id_tto(3), o.rand_tt((2,) * 3, 1, rng=rng), sweep_schedule=[1, 2], rmax_schedule=[1, 2])
    assert max(x2.ttv_rks) <= 2 and np.all(np.isfinite(E2))
    with pytest.raises(AssertionError):
        t.als_gen_eigsolv(A, S, x0, sweep_schedule=[2, 3], rmax_schedule=[4])


@pytest.mark.gpu
def test_als_gen_eigsolv_complex_follows_reference_transpose():
    """ComplexF64 pencil: `K_eiggenmin` builds the transposed local matrices (als.jl:91-92), so the local eigenvector is the
    conjugate of the mathematical one; the device path reproduces the reference's energies step by step."""
    import ttn_b200 as t
    d = 4
    rng = np.random.default_rng(22)
    def phased(Areal, phis):                        # D^H A D with D = kron_k diag(1, exp(i phi_k)): Hermitian, complex
        Ac = o.complex_tto(Areal)
        for k, phi in enumerate(phis):
            dk = np.array([1.0, np.exp(1j * phi)])
            Ac.tto_vec[k] = np.conj(dk)[:, None, None, None] * Ac.tto_vec[k] * dk[None, :, None, None]
        return Ac
    A = phased(spd_op(d, 2.0), [0.3, 1.1, -0.7, 2.0])
    S = phased(_spd_mass(d), [0.3, 1.1, -0.7, 2.0])
    x0 = o.rand_tt((2,) * d, 2, rng=rng, dtype=np.complex128, normalise=True)
    E, x = t.als_gen_eigsolv(A, S, x0, sweep_schedule=[3], rmax_schedule=[2])
    Eo, xo = o.als_gen_eigsolv(A, S, x0, sweep_schedule=[3], rmax_schedule=[2])
    assert len(E) == len(Eo) and np.max(np.abs(E - Eo)) < 1e-9


def test_mals_linsolve_matrix_free_gmres_saturated_ranks_vs_oracle():
    """`mals_linsolve` (mals.jl:240-309) in the regime where the dense local matrix of mals.jl:148-169 is replaced by the
    matrix-free GMRES solve (windows above 2048 unknowns: n^2 r_l r_r = 4 * 32 * 32 = 4096): d = 12, a right-hand side of rank 40
    and a start of rank 32 so that the two-site solves hit the cap rmax = 32 at the centre bonds.  The operator is a well
    conditioned SPD operator (Laplace + 3 I), so the comparison with the oracle (dense local solves, the reference algorithm)
    and with the dense solution is meaningful at 1e-8; for the Laplace operator of cfg3 itself (kappa ~ 1e12) parity is stated on
    residuals (test_cfg3_laplace2d_interleaved_mals_linsolve)."""
    import ttn_b200 as t
    d = 12
    rng = np.random.default_rng(91)
    A = spd_op(d, 3.0)
    b = o.rand_tt((2,) * d, 40, rng=rng, normalise=True)
    x0 = o.rand_tt((2,) * d, 32, rng=rng, normalise=True)
    t.reset_launch_count()
    x, info = t.mals_linsolve(A, b, x0, tol=1e-13, rmax=32, return_info=True, linsolv_maxiter=40, krylovdim=40)
    assert max(x.ttv_rks) == 32                                     # the cap is reached
    xo = o.mals_linsolve(A, b, x0, tol=1e-13, rmax=32)
    assert x.ttv_rks == xo.ttv_rks
    assert relerr(dv(x), dv(xo)) < 1e-8
    Ad = o.tto_to_matrix(A)
    res = np.linalg.norm(Ad @ dv(x) - dv(b)) / np.linalg.norm(dv(b))
    res_o = np.linalg.norm(Ad @ dv(xo) - dv(b)) / np.linalg.norm(dv(b))
    assert abs(res - res_o) < 1e-8 and abs(info["residual"] - res) < 1e-6
    # energy-norm distance to the oracle's result
    e = dv(x) - dv(xo)
    assert np.sqrt(e @ (Ad @ e)) < 1e-8 * np.sqrt(dv(xo) @ (Ad @ dv(xo)))
