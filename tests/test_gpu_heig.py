"""GPU tests of the small Hermitian eigensolver behind the Gram path of tt_compress! (csrc/heig.cu) through the C ABI
(`ttn_heig_host`): eigenvalues 1e-13 |G|, residual |G U - U L| 1e-13 |G|, orthonormality 5e-12 (the accuracy the rank-cap path
needs: reconstructions 1e-10); clustered / degenerate / decaying spectra must either meet the same bar or raise
the fall-back flag."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def wishart(rng, n, q, cplx):
    th = rng.standard_normal((n, q)) + (1j * rng.standard_normal((n, q)) if cplx else 0)
    return th @ th.conj().T


def check(G, nev, lam, U, tol_orth=5e-12):
    w = np.linalg.eigvalsh(G)[::-1]
    nrm = max(abs(w[0]), 1e-300)
    assert np.max(np.abs(lam - w[:nev])) / nrm < 1e-13
    assert np.linalg.norm(U.conj().T @ U - np.eye(nev)) < tol_orth
    assert np.linalg.norm(G @ U - U * lam) / nrm < 1e-13 * max(4, G.shape[0]) ** 0.5 * 4


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("n,nev", [(128, 64), (128, 128), (64, 64), (64, 17), (96, 40), (33, 33), (16, 8), (5, 3), (2, 2), (2, 1), (1, 1)])
def test_heig_single(n, nev, cplx):
    import ttn_b200 as t
    rng = np.random.default_rng(1000 * n + nev + cplx)
    G = wishart(rng, n, 4 * n + 3, cplx)
    lam, U, flag = t.heig_top(G, nev)
    assert flag == 0
    check(G, nev, lam, U)


def test_heig_real_176():
    import ttn_b200 as t
    rng = np.random.default_rng(7)
    G = wishart(rng, 176, 700, False)
    lam, U, flag = t.heig_top(G, 90)
    assert flag == 0
    check(G, 90, lam, U)


@pytest.mark.parametrize("cplx", [False, True])
def test_heig_batched(cplx):
    import ttn_b200 as t
    rng = np.random.default_rng(11 + cplx)
    for n, nev, b in ((128, 64, 160), (64, 64, 300), (24, 10, 7)):
        G = np.stack([wishart(rng, n, 3 * n, cplx) for _ in range(b)])
        lam, U, flags = t.heig_top(G, nev)
        assert not flags.any()
        for k in (0, 1, b // 2, b - 1):
            check(G[k], nev, lam[k], U[k])


def test_heig_clusters_and_fallback_flags():
    import ttn_b200 as t
    rng = np.random.default_rng(3)
    n = 64
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    # close pairs inside the kept set are re-orthogonalised
    for gap in (1e-5, 1e-8):
        w = np.r_[np.linspace(1, 2, n - 2), 2.5, 2.5 + gap]
        G = (Q * w) @ Q.T
        lam, U, flag = t.heig_top(G, 32)
        assert flag == 0
        check(G, 32, lam, U)
    # exactly degenerate pair in the kept set / decaying spectrum: either accurate or flagged for the Jacobi path
    for w in (np.r_[np.linspace(1, 2, n - 2), 2.5, 2.5], 10.0 ** -np.arange(n, dtype=float), np.zeros(n)):
        G = (Q * w) @ Q.T
        lam, U, flag = t.heig_top(G, 32)
        if flag == 0:
            check(G, 32, lam, U)
    # a multiplet outside the kept set does not matter
    w = np.r_[np.ones(10), np.linspace(2, 3, n - 10)]
    G = (Q * w) @ Q.T
    lam, U, flag = t.heig_top(G, 40)
    assert flag == 0
    check(G, 40, lam, U)
