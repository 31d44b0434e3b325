"""GPU parity tests at the REAL shapes of BASELINE.json's configs ("configs with no parity test at their real
shapes"): the cfg5 bond shapes (ComplexF64, 128 x 512 L->R bonds, batched), a Heisenberg DMRG at chi up to 256 against a sparse
eigensolver with both large-SVD engines, the Gram path of tt_compress! against the Jacobi path, and its fall-back."""
import math

import numpy as np
import pytest

import ttn_oracle as o

pytestmark = pytest.mark.gpu


def cfg5_inputs(nvec, d=30, r=64, W=4, seed=7):
    rks = [min(2 ** k, 2 ** (d - k), r) for k in range(d + 1)]
    Rk = [min(4 ** k, 4 ** (d - k), W) for k in range(d + 1)]
    rng = np.random.default_rng(seed)
    A = o.TToperator(d, [np.asfortranarray((rng.standard_normal((2, 2, Rk[k], Rk[k + 1])) + 1j * rng.standard_normal((2, 2, Rk[k], Rk[k + 1])))
                                           / math.sqrt(2.0 * Rk[k + 1])) for k in range(d)], (2,) * d, Rk)
    xs = []
    for b in range(nvec):
        g = np.random.default_rng(100 + b)
        xs.append(o.TTvector(d, [np.asfortranarray((g.standard_normal((2, rks[k], rks[k + 1])) + 1j * g.standard_normal((2, rks[k], rks[k + 1])))
                                                  / math.sqrt(4.0 * rks[k + 1])) for k in range(d)], (2,) * d, rks, [0] * d))
    return A, xs


def test_cfg5_shape_batched_vs_oracle():
    """cfg5 at its real shape: ComplexF64, d = 30, rank 64, MPO rank 4, a batch of 8 trains through
    A*x and tt_compress!(y, 64).  Two of the trains are compared with the oracle's tt_compress(apply(A, x), 64) by TT distance
    and per-bond singular values.  Truncating a flat spectrum is ill conditioned (as in test_cfg2_shape_property_checks), so the
    TT distance is held to the oracle's own measured sensitivity, the singular values to 1e-8 and the error norm to 1e-10."""
    import ttn_b200 as t
    A, xs = cfg5_inputs(8)
    Ad = t.DeviceTTO.upload(A)
    xd = t.DeviceTT.upload(xs)
    yd, sig = t.tt_compress_(t.apply(Ad, xd), 64, return_sigma=True)
    assert max(yd.ttv_rks) == 64
    ys = yd.download()
    for b in (0, 5):
        sig_ref = []
        ax = o.apply(A, xs[b])
        ref = o.tt_compress(o.copy_tt(ax), 64, sigma_out=sig_ref)
        pert = o.copy_tt(ax)
        rng = np.random.default_rng(b)
        for k in range(pert.N):
            pert.ttv_vec[k] = pert.ttv_vec[k] * (1 + 1e-16 * rng.standard_normal(pert.ttv_vec[k].shape))
        sens = o.rel_distance(o.tt_compress(pert, 64), ref)
        assert ys[b].ttv_rks == ref.ttv_rks
        assert o.rel_distance(ys[b], ref) < max(1e-10, 40 * sens)
        assert abs(o.rel_distance(ys[b], ax) - o.rel_distance(ref, ax)) < 1e-10
        if b == 0:
            for a, c in zip(sig, sig_ref):
                assert len(a) == len(c) and np.abs(np.asarray(a) - np.asarray(c)).max() / c[0] < 1e-8
    # the same batch through the Jacobi path (Gram path off) agrees with the Gram path to the same tolerance
    t.set_option("gram_compress", 0)
    try:
        y2 = t.tt_compress_(t.apply(Ad, xd), 64).download()
    finally:
        t.set_option("gram_compress", 1)
    for b in (0, 3, 7):
        assert o.rel_distance(ys[b], y2[b]) < 1e-9


def _heisenberg_sparse(d, jx, jy, jz):
    import scipy.sparse as sp
    X = sp.csr_matrix(np.array([[0.0, 1.0], [1.0, 0.0]]))
    iY = sp.csr_matrix(np.array([[0.0, 1.0], [-1.0, 0.0]]))      # i * sigma_y (real); (i sy) x (i sy) = - sy x sy
    Z = sp.csr_matrix(np.array([[1.0, 0.0], [0.0, -1.0]]))
    H = sp.csr_matrix((2 ** d, 2 ** d))
    for i in range(d - 1):
        l, r = sp.identity(2 ** i, format="csr"), sp.identity(2 ** (d - i - 2), format="csr")
        H = H + jx * sp.kron(sp.kron(l, sp.kron(X, X)), r) - jy * sp.kron(sp.kron(l, sp.kron(iY, iY)), r) \
            + jz * sp.kron(sp.kron(l, sp.kron(Z, Z)), r)
    return H.tocsr()


@pytest.mark.parametrize("large_engine", ["gram_block", "default"])
def test_dmrg_heisenberg_d16_chi256_vs_sparse_eigsh(large_engine):
    """cfg4's parity run: Heisenberg XYZ chain, d = 16, two-site DMRG with bond caps 32 / 128 / 256
    (256 is exact at the centre bond, so the 512 x 512 two-site SVDs, the 128 x 128-tile GEMM and the large-matrix Jacobi
    engines all run inside the solver) against scipy.sparse.linalg.eigsh of the same Hamiltonian to 1e-10.  `gram_block`
    forces the Gram-block (DMMA) Jacobi from 128 columns up, `default` leaves the scalar block path below 640 columns."""
    import scipy.sparse.linalg as spla
    import ttn_b200 as t
    d = 16
    jx, jy, jz = 1.1, 0.8, 1.2
    H = o.heisenberg_xyz_tto(d, jx=jx, jy=jy, jz=jz, lam=0.0)
    # the sparse operator is the same matrix as the TT operator (checked densely at d = 6 on the same builder)
    H6 = np.real(o.tto_to_matrix(o.heisenberg_xyz_tto(6, jx=jx, jy=jy, jz=jz, lam=0.0)))
    S6 = _heisenberg_sparse(6, jx, jy, jz).toarray()
    scale = H6[0, 0] / S6[0, 0] if abs(S6[0, 0]) > 0 else 1.0
    assert np.allclose(H6, scale * S6, atol=1e-12)
    e0 = scale * spla.eigsh(_heisenberg_sparse(d, jx, jy, jz) * np.sign(scale), k=1, which="SA", tol=1e-13)[0][0] * np.sign(scale)
    x0 = o.rand_tt((2,) * d, 16, rng=np.random.default_rng(3), normalise=True)
    old = t.get_option("gram_jacobi_min")
    if large_engine == "gram_block":
        t.set_option("gram_jacobi_min", 128)
    try:
        E, x, rh = t.dmrg_eigsolve(H, x0, N=2, tol=1e-12, sweep_schedule=[2, 4, 6], rmax_schedule=[32, 128, 256],
                                   linsolv_tol=1e-12, linsolv_maxiter=400, krylovdim=24)
    finally:
        t.set_option("gram_jacobi_min", old)
    assert max(rh) > 64
    assert abs(E[-1] - e0) < 1e-10 * abs(e0)


def test_gram_path_falls_back_on_rank_deficient_and_decaying_inputs():
    """The Gram path of tt_compress! (truncerr = 0) must hand over to the Jacobi path whenever the kept spectrum is not safely
    inside the accuracy of the Gram matrix: (a) x + x (exactly rank-deficient bonds, max_bond above the true rank) and (b) a
    smooth function (singular values decaying to rounding level inside the kept set).  Both must match the oracle to 1e-10."""
    import ttn_b200 as t
    rng = np.random.default_rng(5)
    d = 12
    x = o.rand_tt((2,) * d, 12, rng=rng, normalise=True)
    s = o.add(x, x)
    got = t.tt_compress_(o.copy_tt(s), 40)
    ref = o.tt_compress(o.copy_tt(s), 40)
    assert got.ttv_rks == ref.ttv_rks
    assert o.rel_distance(got, ref) < 1e-10
    f = o.function_to_qtt_uniform(lambda u: np.exp(-3 * u) * np.cos(7 * u), 14)
    f2 = o.add(f, o.qtt_sin(14, lam=2.0))
    got = t.tt_compress_(o.copy_tt(f2), 10)
    ref = o.tt_compress(o.copy_tt(f2), 10)
    assert got.ttv_rks == ref.ttv_rks
    assert o.rel_distance(got, ref) < 1e-10


def test_gram_vs_jacobi_path_real_well_conditioned():
    """Float64, d = 14, rank 96 -> 24: the Gram path (tridiagonal eigensolver) and the QR + one-sided Jacobi path give the same
    train, the same per-bond singular values, and both match the oracle."""
    import ttn_b200 as t
    x = o.rand_tt((2,) * 14, 96, rng=np.random.default_rng(11), normalise=True)
    sig_ref = []
    ref = o.tt_compress(o.copy_tt(x), 24, sigma_out=sig_ref)
    a, sa = t.tt_compress_(o.copy_tt(x), 24, return_sigma=True)
    t.set_option("gram_compress", 0)
    try:
        b, sb = t.tt_compress_(o.copy_tt(x), 24, return_sigma=True)
    finally:
        t.set_option("gram_compress", 1)
    assert a.ttv_rks == b.ttv_rks == ref.ttv_rks
    pert = o.copy_tt(x)
    rng = np.random.default_rng(1)
    for k in range(pert.N):
        pert.ttv_vec[k] = pert.ttv_vec[k] * (1 + 1e-16 * rng.standard_normal(pert.ttv_vec[k].shape))
    sens = o.rel_distance(o.tt_compress(pert, 24), ref)
    tol = max(1e-10, 40 * sens)
    assert o.rel_distance(a, ref) < tol and o.rel_distance(b, ref) < tol
    for u, v, w in zip(sa, sb, sig_ref):
        assert np.abs(np.asarray(u) - np.asarray(w)).max() / w[0] < 1e-9
        assert np.abs(np.asarray(v) - np.asarray(w)).max() / w[0] < 1e-9


@pytest.mark.parametrize("cplx", [False, True])
def test_apply_compress_fused_equals_two_calls(cplx):
    """`ttn_apply_compress` (A*x folded into the two-site merges of the first tt_compress! pass) against apply followed by
    tt_compress!, and against the oracle: same ranks, TT distance at rounding level, same per-bond singular values; batched and
    single; a truncerr > 0 call (plain composition) as well."""
    import ttn_b200 as t
    d, r, W = 12, 16, 3
    rng = np.random.default_rng(21 + cplx)
    dt = np.complex128 if cplx else np.float64
    A = o.rand_tto((2,) * d, W, rng=rng, dtype=dt)
    xs = [o.rand_tt((2,) * d, r, rng=rng, dtype=dt, normalise=True) for _ in range(5)]
    Ad = t.DeviceTTO.upload(A)
    xd = t.DeviceTT.upload(xs)
    calls0, fb0 = t.get_option("gram_calls"), t.get_option("gram_fallbacks")
    yf, sf = t.apply_compress(Ad, xd, 10, return_sigma=True)
    assert t.get_option("gram_calls") == calls0 + 1 and t.get_option("gram_fallbacks") == fb0     # the fused fast path really ran
    y2, s2 = t.tt_compress_(t.apply(Ad, xd), 10, return_sigma=True)
    yf, y2 = yf.download(), y2.download()
    for b in range(5):
        ref = o.tt_compress(o.apply(A, xs[b]), 10)
        assert yf[b].ttv_rks == y2[b].ttv_rks == ref.ttv_rks
        assert o.rel_distance(yf[b], y2[b]) < 1e-10
        assert o.rel_distance(yf[b], ref) < 1e-9
    for u, v in zip(sf, s2):
        assert len(u) == len(v) and np.abs(np.asarray(u) - np.asarray(v)).max() < 1e-10 * max(v[0], 1e-300)
    # single train through the host API, and a tolerance-driven call (not served by the fast path: plain composition)
    y1 = t.apply_compress(A, xs[0], 10)
    assert o.rel_distance(y1, o.tt_compress(o.apply(A, xs[0]), 10)) < 1e-9
    y3 = t.apply_compress(A, xs[1], 64, truncerr=1e-3)
    ref3 = o.tt_compress(o.apply(A, xs[1]), 64, truncerr=1e-3)
    assert y3.ttv_rks == ref3.ttv_rks and o.rel_distance(y3, ref3) < 1e-10


def test_async_upload_download_pipeline_matches_sync():
    """`upload_batched(..., asynchronous=True)` / `download_into(..., asynchronous=True)` (copy stream overlapping the compute stream,
    the e2e pipeline of bench.py): three chunks through apply_compress, results identical to the synchronous calls."""
    import torch
    import ttn_b200 as t
    d, r, W, nb = 10, 8, 2, 6
    rng = np.random.default_rng(33)
    A = o.rand_tto((2,) * d, W, rng=rng, dtype=np.complex128)
    Ad = t.DeviceTTO.upload(A)
    rks = [min(2 ** k, 2 ** (d - k), r) for k in range(d + 1)]
    keep, chunks = [], []
    for c in range(3):
        cores = []
        for k in range(d):
            buf = torch.empty(2 * 2 * rks[k] * rks[k + 1] * nb, dtype=torch.float64).pin_memory()
            buf.normal_()
            keep.append(buf)
            cores.append(buf.numpy().view(np.complex128).reshape((2, rks[k], rks[k + 1], nb), order="F"))
        chunks.append(cores)
    ref = []
    for cores in chunks:
        y = t.apply_compress(Ad, t.DeviceTT.upload_batched(cores, (2,) * d, rks), 4)
        ref.append([np.array(a) for a in y.download_into([np.empty((2, min(rks[k], 4), min(rks[k + 1], 4), nb), dtype=np.complex128, order="F")
                                                          for k in range(d)])])
    outs = []
    nxt = t.DeviceTT.upload_batched(chunks[0], (2,) * d, rks, asynchronous=True)
    for i, cores in enumerate(chunks):
        xd = nxt
        if i + 1 < len(chunks):
            nxt = t.DeviceTT.upload_batched(chunks[i + 1], (2,) * d, rks, asynchronous=True)
        y = t.apply_compress(Ad, xd, 4)
        dst = []
        for k in range(d):
            buf = torch.empty(2 * 2 * min(rks[k], 4) * min(rks[k + 1], 4) * nb, dtype=torch.float64).pin_memory()
            keep.append(buf)
            dst.append(buf.numpy().view(np.complex128).reshape((2, min(rks[k], 4), min(rks[k + 1], 4), nb), order="F"))
        y.download_into(dst, asynchronous=True)
        xd.free(); y.free()
        outs.append(dst)
    t.copy_synchronize()
    for a, b in zip(outs, ref):
        for u, v in zip(a, b):
            assert np.array_equal(u, v)


def test_two_host_threads_two_contexts_match_serial():
    """One library context per host thread (include/ttn_b200.h): two threads compress different batches concurrently (their
    kernels overlap on the GPU); every result equals the single-threaded one bit for bit."""
    import threading
    import ttn_b200 as t
    d, r, W = 12, 16, 3
    rng = np.random.default_rng(77)
    A = o.rand_tto((2,) * d, W, rng=rng, dtype=np.complex128)
    Ad = t.DeviceTTO.upload(A)
    batches = [[o.rand_tt((2,) * d, r, rng=rng, dtype=np.complex128, normalise=True) for _ in range(4)] for _ in range(6)]
    serial = [[y.ttv_vec for y in t.apply_compress(Ad, t.DeviceTT.upload(b), 8).download()] for b in batches]
    t.synchronize()
    out, errs = [None] * len(batches), []

    def work(idx):
        try:
            for k in range(idx, len(batches), 2):
                for _ in range(3):
                    y = t.apply_compress(Ad, t.DeviceTT.upload(batches[k]), 8)
                out[k] = [v.ttv_vec for v in y.download()]
            t.synchronize()
        except Exception as e:   # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs, errs
    for a, b in zip(out, serial):
        for va, vb in zip(a, b):
            for ca, cb in zip(va, vb):
                assert np.array_equal(ca, cb)
