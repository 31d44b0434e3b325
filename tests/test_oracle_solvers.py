"""Pins the solver part of the oracle (ALS / MALS / DMRG / TDVP) against dense ground truth and the
reference's own solver tests (CPU only).  Reference tests cited per test."""
import numpy as np
import pytest
import scipy.linalg as sla

import ttn_oracle as o


def dv(x):
    return o.ttv_to_tensor(x).reshape(-1)


def spd_op(d, shift=3.0):
    # test/test_dmrg.jl:18 / test/test_mals.jl:15 — Δ(d) + shift·I
    return o.tto_add(o.laplace_dd(d), o.tto_scale(shift, o.id_tto(d)))


def test_cut_off_index_kat():
    # test/test_dmrg.jl:20-25
    s = np.array([1.0, 1.0 - 5.0e-11, 0.1])
    tol = (1.0 - 2.0e-11) / np.linalg.norm(s)
    assert o.cut_off_index(s, tol) == 2


def test_sv_trunc_rule():
    # mals.jl:42-56 — the element that crosses the threshold is kept
    s = np.array([1.0, 0.1, 1e-3, 1e-7])
    assert len(o.sv_trunc(s, 0.0)) == 4
    assert len(o.sv_trunc(s, 1e-15)) == 4     # 1e-14 crosses 1e-15·Σs² immediately and is kept
    assert len(o.sv_trunc(s, 1e-12)) == 3     # 1e-14 < 1e-12·Σs² → dropped; 1e-6 crosses → 1e-3 kept
    assert len(o.sv_trunc(s, 1e-10)) == 3
    assert len(o.sv_trunc(s, 1e-3)) == 2


def test_als_linsolve_vs_dense():
    # test/test_als.jl:44-64 (loose) tightened: full-rank ALS on a well-conditioned SPD system reaches the dense solve
    d = 5
    rng = np.random.default_rng(10)
    A = spd_op(d, 3.0)
    b = o.rand_tt((2,) * d, 2, rng=rng)
    x0 = o.rand_tt((2,) * d, 4, rng=rng)       # ranks [1,2,4,4,2,1] = full → exact solution representable
    x, info = o.als_linsolve(A, b, x0, sweep_count=6, return_info=True)
    ref = np.linalg.solve(o.tto_to_matrix(A), dv(b))
    assert np.linalg.norm(dv(x) - ref) / np.linalg.norm(ref) < 1e-10
    assert info["residual"] < 1e-10


def test_als_identity_residual():
    # test/test_als.jl:63 — identity operator → residual < 0.05 (here: exact)
    d = 4
    rng = np.random.default_rng(11)
    b = o.rand_tt((2,) * d, 2, rng=rng)
    x0 = o.rand_tt((2,) * d, 2, rng=rng)
    x, info = o.als_linsolve(o.id_tto(d), b, x0, sweep_count=2, return_info=True)
    assert info["residual"] < 1e-12


def test_als_eigsolve_vs_dense():
    # test/test_als.jl:95-117 (Rayleigh ≈ λ at rtol 0.1; monotone energy) tightened to the dense eigenvalue
    d = 5
    rng = np.random.default_rng(12)
    A = spd_op(d, 1.0)
    x0 = o.rand_tt((2,) * d, 4, rng=rng)
    E, x = o.als_eigsolve(A, x0, sweep_schedule=[6], rmax_schedule=[4])
    lam = sla.eigvalsh(o.tto_to_matrix(A))[0]
    assert abs(E[-1] - lam) < 1e-10
    assert np.all(np.diff(E) < 1e-10)
    v = dv(x)
    assert abs(v @ o.tto_to_matrix(A) @ v / (v @ v) - lam) < 1e-10


def test_mals_linsolve_vs_dense():
    # test/test_mals.jl:33-77 tightened; rmax respected (:64)
    d = 6
    rng = np.random.default_rng(13)
    A = spd_op(d, 3.0)
    b = o.rand_tt((2,) * d, 2, rng=rng)
    x0 = o.rand_tt((2,) * d, 2, rng=rng)
    x, info = o.mals_linsolve(A, b, x0, tol=1e-14, rmax=8, return_info=True)
    ref = np.linalg.solve(o.tto_to_matrix(A), dv(b))
    assert max(x.ttv_rks) <= 8
    assert np.linalg.norm(dv(x) - ref) / np.linalg.norm(ref) < 1e-9
    x2 = o.mals_linsolve(A, b, x0, tol=1e-14, rmax=3)
    assert max(x2.ttv_rks) <= 3


def test_mals_eigsolve_vs_dense():
    # test/test_mals.jl:96-118 tightened
    d = 6
    rng = np.random.default_rng(14)
    A = spd_op(d, 1.0)
    x0 = o.rand_tt((2,) * d, 2, rng=rng)
    E, x, rh = o.mals_eigsolve(A, x0, tol=1e-12, sweep_schedule=[4], rmax_schedule=[8])
    lam = sla.eigvalsh(o.tto_to_matrix(A))[0]
    assert abs(E[-1] - lam) < 1e-9
    assert len(rh) == len(E)


@pytest.mark.parametrize("N", [1, 2])
def test_dmrg_linsolve_vs_dense(N):
    # test/test_dmrg.jl:43-75 tightened; test/test_euler.jl:34-59 uses dmrg_linsolve to 1e-5
    d = 5
    rng = np.random.default_rng(15)
    A = spd_op(d, 3.0)
    b = o.rand_tt((2,) * d, 2, rng=rng)
    x0 = o.rand_tt((2,) * d, 2, rng=rng)
    x, info = o.dmrg_linsolve(A, b, x0, N=N, sweep_schedule=[8], rmax_schedule=[4], return_info=True)
    ref = np.linalg.solve(o.tto_to_matrix(A), dv(b))
    assert np.linalg.norm(dv(x) - ref) / np.linalg.norm(ref) < 1e-9
    assert info["residual"] < 1e-9


def test_dmrg_eigsolve_heisenberg_vs_dense():
    # examples/heisenberg_xyz_dmrg.jl:9-19 — DMRG energy vs eigvals at d = 10 (here d = 8 to keep the CPU suite fast)
    d = 8
    rng = np.random.default_rng(16)
    H = o.heisenberg_xyz_tto(d, jx=1.1, jy=0.8, jz=1.2)
    x0 = o.rand_tt((2,) * d, 4, rng=rng, normalise=True)
    E, x, rh = o.dmrg_eigsolve(H, x0, N=2, tol=1e-12, sweep_schedule=[4], rmax_schedule=[16])
    lam = sla.eigvalsh(o.tto_to_matrix(H))[0]
    assert abs(E[-1] - lam) < 1e-10
    assert max(rh) <= 16
    v = dv(x)
    assert abs(v @ o.tto_to_matrix(H) @ v / (v @ v) - lam) < 1e-10


def test_dmrg_local_operator_is_projected_dense_operator():
    # K == PᴴAP for the orthogonal frame of the two-site window
    d = 5
    rng = np.random.default_rng(17)
    A = spd_op(d, 0.5)
    x = o.orthogonalize(o.rand_tt((2,) * d, 3, rng=rng), i=2)
    from ttn_oracle.dmrg import _init_H, K_full
    G = [np.ones((1, 1, 1))]
    G.append(o.dmrg_update_G(x.ttv_vec[0], A.tto_vec[0], G[0]))
    H = _init_H(x, A, 2)
    i = 2  # window sites (2,3), 1-based
    Am = o.amid(A, i, i + 1)
    K, dims = K_full(G[i - 1], H[i - 1], Am)
    # frame P: columns = basis of the window embedded in the full space
    left = x.ttv_vec[0][:, 0, :]                                   # (n1, r1)
    right = np.einsum("sab,tbc->astc", x.ttv_vec[3], x.ttv_vec[4])[..., 0]   # (r3, n4, n5)
    n = 2
    # explicit loop construction (clearer than a 9-index einsum)
    r1, r3 = left.shape[1], right.shape[0]
    P = np.zeros((2 ** d, r1 * n * n * r3))
    for a in range(r1):
        for s2 in range(n):
            for s3 in range(n):
                for c in range(r3):
                    col = a + r1 * (s2 + n * s3) + r1 * n * n * c
                    t = np.einsum("i,ml->iml", left[:, a], right[c])      # (n1, n4, n5)
                    full = np.zeros((n,) * d)
                    full[:, s2, s3, :, :] = t
                    P[:, col] = full.reshape(-1)
    Ad = o.tto_to_matrix(A)
    assert np.allclose(K, P.T @ Ad @ P, atol=1e-12)
    # and the matrix-free contraction equals K·v
    V = rng.standard_normal(dims)
    y = o.dmrg_matvec2(G[i - 1], Am, V, H[i - 1], symmetrize=False)
    assert np.allclose(y.reshape(-1, order="F"), K @ V.reshape(-1, order="F"))
    ys = o.dmrg_matvec2(G[i - 1], Am, V, H[i - 1], symmetrize=True)
    assert np.allclose(ys.reshape(-1, order="F"), 0.5 * (K + K.T) @ V.reshape(-1, order="F"))


def test_tdvp_local_maps_against_loops():
    # test/test_tdvp.jl:76-114 — explicit-loop oracles for _applyH1_lsr and _applyH0 (1e-12)
    rng = np.random.default_rng(18)
    Dl, d, Dr, w = 3, 2, 4, 2
    AC = rng.standard_normal((Dl, d, Dr))
    FL = rng.standard_normal((Dl, w, Dl))
    FR = rng.standard_normal((Dr, w, Dr))
    M = rng.standard_normal((w, d, w, d))
    ref = np.zeros((Dl, d, Dr))
    for al in range(Dl):
        for s in range(d):
            for be in range(Dr):
                acc = 0.0
                for a in range(w):
                    for alp in range(Dl):
                        for sp in range(d):
                            for bep in range(Dr):
                                for b in range(w):
                                    acc += FL[al, a, alp] * AC[alp, sp, bep] * M[a, s, b, sp] * FR[bep, b, be]
                ref[al, s, be] = acc
    assert np.abs(o.apply_H1_lsr(AC, FL, FR, M) - ref).max() < 1e-12
    C = rng.standard_normal((Dl, Dr))
    FRl = rng.standard_normal((Dr, w, Dr))
    ref0 = np.zeros((Dl, Dr))
    for al in range(Dl):
        for be in range(Dr):
            ref0[al, be] = sum(FL[al, a, alp] * C[alp, bep] * FRl[bep, a, be]
                               for a in range(w) for alp in range(Dl) for bep in range(Dr))
    assert np.abs(o.apply_H0(C, FL, FRl) - ref0).max() < 1e-12


@pytest.mark.parametrize("two_site", [False, True])
def test_tdvp_heat_eigenmode(two_site):
    # test/test_tdvp.jl:329-356 — imaginary-time evolution of an eigenmode of A follows exp(λ·Σh)·u0 (< 1e-8)
    d = 5
    A = o.tto_scale(-1.0, o.laplace_dd(d))
    Ad = o.tto_to_matrix(A)
    w, v = sla.eigh(Ad)
    u0d = v[:, -1]                               # slowest-decaying mode
    u0 = o.ttv_decomp(u0d.reshape((2,) * d), index=1, tol=1e-14)
    steps = [0.01] * 5
    f = o.tdvp2 if two_site else o.tdvp
    kw = dict(max_bond=8) if two_site else {}
    psi = f(A, u0, steps, normalize=False, imaginary_time=True, **kw)
    ref = np.exp(w[-1] * sum(steps)) * u0d
    assert np.linalg.norm(dv(psi) - ref) / np.linalg.norm(ref) < 1e-8


def test_tdvp_real_time_norm_and_exact_small():
    # test/test_tdvp.jl:132-160,238-261 — H = 0 leaves the state unchanged; full-rank real-time TDVP2 is exact
    d = 4
    rng = np.random.default_rng(19)
    u0 = o.rand_tt((2,) * d, 4, rng=rng, normalise=True)
    Z = o.tto_scale(0.0, o.id_tto(d))
    psi = o.tdvp(Z, u0, [0.1, 0.1], normalize=False)
    assert np.linalg.norm(dv(psi) - dv(u0)) < 1e-12
    H = o.heisenberg_xyz_tto(d, jx=1.0, jy=0.7, jz=0.3)
    psi2 = o.tdvp2(H, u0, [0.05] * 4, normalize=False, max_bond=4)
    ref = sla.expm(-1j * 0.2 * o.tto_to_matrix(H)) @ dv(u0)
    assert np.linalg.norm(dv(psi2) - ref) / np.linalg.norm(ref) < 1e-3   # Trotter error O(dt²) of the 2-site splitting
    assert abs(np.linalg.norm(dv(psi2)) - np.linalg.norm(dv(u0))) < 1e-10


def test_matvec_blas_restatement_equals_einsum():
    # the GEMM-lowered K_matfree (CPU baseline of bench.py) against the index-by-index contraction of dmrg.jl:239-244
    rng = np.random.default_rng(3)
    G = rng.standard_normal((3, 6, 6)); H = rng.standard_normal((4, 5, 5))
    Am = rng.standard_normal((3, 4, 4, 4)); V = rng.standard_normal((6, 4, 5))
    assert np.allclose(o.dmrg_matvec2_blas(G, Am, V, H), o.dmrg_matvec2(G, Am, V, H, symmetrize=False), rtol=1e-13, atol=1e-13)


def _spd_mass(d, eps=0.3):
    # a well-conditioned SPD "overlap" operator: I + eps * (shift + shift^T) as a rank-3 TTO sum
    S = o.tto_add(o.id_tto(d), o.tto_scale(-eps / 2.0, o.tto_add(o.laplace_dd(d), o.tto_scale(-2.0, o.id_tto(d)))))
    return S


def test_als_gen_eigsolv_vs_dense_and_reference_tests():
    # test/test_als.jl:153-197: structure, S = I agreement with als_eigsolve, rank growth; tightened against scipy eigh(A, S)
    d = 5
    rng = np.random.default_rng(21)
    A = spd_op(d, 2.0)
    x0 = o.rand_tt((2,) * d, 4, rng=rng, normalise=True)
    E_gen, xg = o.als_gen_eigsolv(A, o.id_tto(d), x0, sweep_schedule=[4], rmax_schedule=[4])
    E_std, _ = o.als_eigsolve(A, x0, sweep_schedule=[4], rmax_schedule=[4])
    assert abs(E_gen[-1] - E_std[-1]) < 1e-10                                   # test_als.jl:168-181 (rtol 0.05 there)
    S = _spd_mass(d)
    Am, Sm = o.tto_to_matrix(A), o.tto_to_matrix(S)
    assert np.all(sla.eigvalsh(Sm) > 0.2)
    lam = sla.eigh(Am, Sm, eigvals_only=True)[0]
    E, x = o.als_gen_eigsolv(A, S, x0, sweep_schedule=[6], rmax_schedule=[4])
    assert abs(E[-1] - lam) < 1e-9
    v = dv(x)
    assert abs((v @ Am @ v) / (v @ Sm @ v) - lam) < 1e-9
    E2, x2 = o.als_gen_eigsolv(spd_op(3, 2.0), o.id_tto(3), o.rand_tt((2,) * 3, 1, rng=rng), sweep_schedule=[1, 2], rmax_schedule=[1, 2])
    assert max(x2.ttv_rks) <= 2 and np.all(np.isfinite(E2))                     # test_als.jl:184-197


def test_reference_property_tests_of_the_solver_drivers():
    """Ports of the remaining loose property tests of the reference's solver test files, run on the oracle:
    test_mals.jl:55-77 (rmax respected; looser tol never needs more than 2 extra ranks), test_dmrg.jl:54-98 (two-stage
    schedule, identity operator, N = 1 with residual info), test_als.jl:66-78,109-117 (single half sweep; energies do not
    increase)."""
    rng = np.random.default_rng(31)
    d = 4
    A = spd_op(d, 5.0)
    b = o.rand_tt((2,) * d, 2, rng=rng)
    x = o.mals_linsolve(A, b, o.rand_tt((2,) * d, 1, rng=rng), tol=1e-10, rmax=4)
    assert max(x.ttv_rks) <= 4
    A3 = spd_op(d, 3.0)
    x0 = o.rand_tt((2,) * d, 2, rng=rng)
    x_loose = o.mals_linsolve(A3, b, x0, tol=1e-2, rmax=8)
    x_tight = o.mals_linsolve(A3, b, x0, tol=0.0, rmax=8)
    assert max(x_loose.ttv_rks) <= max(x_tight.ttv_rks) + 2
    y = o.dmrg_linsolve(A, b, o.rand_tt((2,) * d, 1, rng=rng), N=2, sweep_schedule=[2, 4], rmax_schedule=[2, 8])
    assert tuple(y.ttv_dims) == tuple(b.ttv_dims)
    b1 = o.rand_tt((2,) * d, 1, rng=rng)
    z = o.dmrg_linsolve(o.id_tto(d), b1, o.rand_tt((2,) * d, 1, rng=rng), N=2, sweep_schedule=[4], rmax_schedule=[4])
    assert np.linalg.norm(dv(z) - dv(b1)) / np.linalg.norm(dv(b1)) < 0.05
    A5 = spd_op(3, 5.0)
    w, info = o.dmrg_linsolve(A5, o.rand_tt((2,) * 3, 1, rng=rng), o.rand_tt((2,) * 3, 1, rng=rng), N=1, sweep_schedule=[2],
                              rmax_schedule=[2], return_info=True)
    assert np.isfinite(info["residual"])
    v = o.als_linsolve(A5, o.rand_tt((2,) * 3, 2, rng=rng), o.rand_tt((2,) * 3, 2, rng=rng), sweep_count=1)
    assert tuple(v.ttv_dims) == (2, 2, 2)
    E, _ = o.als_eigsolve(spd_op(d, 2.0), o.rand_tt((2,) * d, 2, rng=rng, normalise=True), sweep_schedule=[4], rmax_schedule=[2])
    assert E[-1] <= E[0] + 1e-8
