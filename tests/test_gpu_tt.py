"""GPU parity tests of the TT-level hot path (apply, dot, +, orthogonalize, tt_compress!) against the oracle and the
reference's own dense checks.  Gauge-invariant comparisons only (reconstructed tensors, singular values)."""
import numpy as np
import pytest

import ttn_oracle as o

pytestmark = pytest.mark.gpu


def dv(x):
    return o.ttv_to_tensor(x).reshape(-1)


def relerr(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_apply_vs_dense_and_oracle(dtype):
    # test/test_tt_tools.jl:345-358 (1e-10), test/test_tt_operations.jl:116-136 (1e-12)
    import ttn_b200 as t
    rng = np.random.default_rng(2)
    dims = (2, 3, 2, 2)
    A = o.rand_tto(dims, 3, rng=rng, dtype=dtype)
    x = o.rand_tt(dims, 4, rng=rng, dtype=dtype)
    y = t.apply(A, x)
    yo = o.apply(A, x)
    assert y.ttv_rks == yo.ttv_rks
    for a, b in zip(y.ttv_vec, yo.ttv_vec):
        assert relerr(a, b) < 1e-14          # same fused-bond order (MPO index fastest), element by element
    assert relerr(dv(y), o.tto_to_matrix(A) @ dv(x)) < 1e-12


def test_apply_mixed_and_errors():
    import ttn_b200 as t
    rng = np.random.default_rng(3)
    A = o.rand_tto((2, 2, 2), 2, rng=rng, dtype=np.complex128)
    x = o.rand_tt((2, 2, 2), 2, rng=rng)
    assert relerr(dv(t.apply(A, x)), o.tto_to_matrix(A) @ dv(x)) < 1e-12
    with pytest.raises(AssertionError):
        t.apply(o.rand_tto((2, 2), 2, rng=rng), x)      # tt_operations.jl:102 "Incompatible dimensions"


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_dot_norm_add_scale(dtype):
    import ttn_b200 as t
    rng = np.random.default_rng(4)
    dims = (2, 2, 3, 2)
    x = o.rand_tt(dims, 3, rng=rng, dtype=dtype)
    y = o.rand_tt(dims, 2, rng=rng, dtype=dtype)
    assert abs(t.dot(x, y) - np.vdot(dv(x), dv(y))) < 1e-12 * abs(np.vdot(dv(x), dv(y))) + 1e-13
    assert abs(t.norm(x) - np.linalg.norm(dv(x))) < 1e-12 * np.linalg.norm(dv(x))
    z = t.add(x, y)
    assert z.ttv_rks == o.add(x, y).ttv_rks
    assert relerr(dv(z), dv(x) + dv(y)) < 1e-13
    assert relerr(dv(t.scale(2.5, x)), 2.5 * dv(x)) < 1e-14
    assert relerr(dv(t.sub(x, y)), dv(x) - dv(y)) < 1e-13
    x.ttv_ot = [1, 0, -1, -1]
    s = t.scale(-3.0, x)
    assert relerr(s.ttv_vec[1], -3.0 * x.ttv_vec[1]) < 1e-15 and relerr(s.ttv_vec[0], x.ttv_vec[0]) == 0


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
@pytest.mark.parametrize("center", [1, 2, 3, 5])
def test_orthogonalize(center, dtype):
    # test/test_tt_tools.jl:981-1017 — reconstruction + orthonormality ≤ 1e-12, ot flags
    import ttn_b200 as t
    rng = np.random.default_rng(5)
    dims = (2, 3, 2, 2, 2)
    x = o.rand_tt(dims, 4, rng=rng, dtype=dtype)
    y = t.orthogonalize(x, i=center)
    yo = o.orthogonalize(x, i=center)
    assert y.ttv_rks == yo.ttv_rks and y.ttv_ot == yo.ttv_ot
    assert relerr(dv(y), dv(x)) < 1e-12
    for j in range(1, center):
        G = y.ttv_vec[j - 1]
        M = np.reshape(np.transpose(G, (1, 0, 2)), (-1, G.shape[2]), order="F")
        assert np.abs(M.conj().T @ M - np.eye(M.shape[1])).max() < 1e-12
    for j in range(center + 1, len(dims) + 1):
        G = y.ttv_vec[j - 1]
        M = np.reshape(np.transpose(G, (1, 2, 0)), (G.shape[1], -1), order="F")
        assert np.abs(M @ M.conj().T - np.eye(M.shape[0])).max() < 1e-12


def test_orthogonalize_errors_and_overfull_ranks():
    import ttn_b200 as t
    x = o.rand_tt((2, 2, 2), 2)
    with pytest.raises(ValueError):
        t.orthogonalize(x, i=0)                       # tt_tools.jl:513 DimensionMismatch
    with pytest.raises(ValueError):
        t.orthogonalize(x, i=4)
    rng = np.random.default_rng(6)
    xo = o.rand_tt((2, 2, 2, 2), [1, 5, 7, 5, 1], rng=rng)   # over-full ranks shrink (tt_tools.jl:522,533)
    y = t.orthogonalize(xo, i=2)
    assert y.ttv_rks == o.orthogonalize(xo, i=2).ttv_rks
    assert relerr(dv(y), dv(xo)) < 1e-12


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_tt_compress_vs_oracle(dtype):
    import ttn_b200 as t
    rng = np.random.default_rng(7)
    x = o.rand_tt((2,) * 10, 16, rng=rng, dtype=dtype, normalise=True)
    sig_ref = []
    ref = o.tt_compress(o.copy_tt(x), 6, sigma_out=sig_ref)
    xc = o.copy_tt(x)
    got, sig = t.tt_compress_(xc, 6, return_sigma=True)
    assert got is xc and got.ttv_rks == ref.ttv_rks
    assert o.rel_distance(got, ref) < 1e-10                      # reconstructed tensor (TT distance)
    assert relerr(dv(got), dv(ref)) < 1e-10
    assert len(sig) == len(sig_ref)
    for a, b in zip(sig, sig_ref):
        assert len(a) == len(b) and np.abs(a - b).max() / b[0] < 1e-10   # retained singular values per bond step


def test_tt_compress_behaviour_kats():
    # test/test_tt_tools.jl:500-574 and :433-497
    import ttn_b200 as t
    rng = np.random.default_rng(8)
    tt = o.rand_tt((2, 2, 2), [1, 2, 2, 1], rng=rng)
    ref = dv(tt).copy()
    before = list(tt.ttv_rks)
    y = t.tt_compress_(tt, 10, sweeps=1)
    assert y is tt and tt.ttv_rks == before and relerr(dv(tt), ref) < 1e-13
    tt4 = o.rand_tt((2, 2, 2, 2), [1, 2, 4, 2, 1], rng=rng)
    t.tt_compress_(tt4, 2)
    assert max(tt4.ttv_rks) <= 2
    for k in range(4):
        assert tt4.ttv_vec[k].shape == (2, tt4.ttv_rks[k], tt4.ttv_rks[k + 1])
    with pytest.raises(AssertionError):
        t.tt_compress_(tt4, 2, sweeps=0)                # tt_tools.jl:773
    # _tt_bond_truncate!: shapes, returned orthogonalised copy, exact rank-1, bad k
    tb = o.TTvector(3, [rng.standard_normal((2, 1, 4)), rng.standard_normal((2, 4, 4)), rng.standard_normal((2, 4, 1))],
                    (2, 2, 2), [1, 4, 4, 1], [0, 0, 0])
    yb = t.tt_bond_truncate_(tb, 1, max_bond=2)
    assert tb.ttv_rks[1] <= 2 and tb.ttv_vec[0].shape == (2, 1, tb.ttv_rks[1]) and tb.ttv_vec[1].shape == (2, tb.ttv_rks[1], 4)
    assert yb.ttv_rks[1] == tb.ttv_rks[1] and relerr(dv(yb), dv(tb)) < 1e-12
    u, v, p, q = np.array([1.2, -0.5]), np.array([0.7, 0.3]), np.array([2.0, 3.0]), np.array([4.0, 5.0])
    t2 = o.TTvector(2, [np.einsum("s,g->sg", u, p).reshape(2, 1, 2), np.einsum("s,g->sg", v, q).reshape(2, 2, 1)],
                    (2, 2), [1, 2, 1], [0, 0])
    ref2 = dv(t2).copy()
    t.tt_bond_truncate_(t2, 1, max_bond=1)
    assert t2.ttv_rks[1] == 1 and relerr(dv(t2), ref2) < 1e-13
    with pytest.raises(AssertionError):
        t.tt_bond_truncate_(tb, 0)
    with pytest.raises(AssertionError):
        t.tt_bond_truncate_(tb, tb.N)


def test_tt_compress_function_qtt_and_truncerr():
    # test/test_qtt_multidim.jl:577-614 pattern; rank-deficient two-site blocks; tail-norm rule on device == oracle
    import ttn_b200 as t
    d = 10
    s, c = o.qtt_sin(d, lam=1.0), o.qtt_cos(d, lam=2.0)
    f = o.add(o.add(s, c), s)
    ref = o.qtt_to_vector(f).copy()
    g = o.add(o.add(s, c), s)
    t.tt_compress_(f, 4)
    assert max(f.ttv_rks) <= 4 and relerr(o.qtt_to_vector(f), ref) < 1e-11
    go = o.tt_compress(o.add(o.add(s, c), s), 100, truncerr=1e-10)
    t.tt_compress_(g, 100, truncerr=1e-10)
    assert g.ttv_rks == go.ttv_rks
    assert relerr(o.qtt_to_vector(g), ref) < 1e-9
    rng = np.random.default_rng(9)
    x = o.rand_tt((2,) * 9, 12, rng=rng, normalise=True)
    for te in (1e-1, 1e-2, 1e-3):
        a = o.tt_compress(o.copy_tt(x), 100, truncerr=te)
        b = t.tt_compress_(o.copy_tt(x), 100, truncerr=te)
        assert a.ttv_rks == b.ttv_rks and o.rel_distance(b, a) < 1e-10


def test_rk4_like_apply_then_compress():
    # test/test_euler.jl:269-298 — an RK4 stage is `A*x` followed by `tt_compress!`
    import ttn_b200 as t
    d = 8
    A = o.laplace_dd(d)
    x = o.qtt_sin(d, lam=1.0)
    y = t.tt_compress_(t.apply(A, x), 4)
    ref = o.tto_to_matrix(A) @ o.qtt_to_vector(x)
    assert relerr(o.qtt_to_vector(y), ref) < 1e-10


def test_batched_apply_compress_matches_single():
    # cfg5 shape in miniature: a batch of independent ComplexF64 TTs through apply + tt_compress!
    import ttn_b200 as t
    rng = np.random.default_rng(10)
    d, B = 8, 5
    A = o.rand_tto((2,) * d, 3, rng=rng, dtype=np.complex128)
    xs = [o.rand_tt((2,) * d, 4, rng=np.random.default_rng(100 + b), dtype=np.complex128, normalise=True) for b in range(B)]
    dev = t.DeviceTT.upload(xs)
    assert dev.batch == B
    yb = t.tt_compress_(t.apply(A, dev), 6)
    outs = yb.download()
    nb = t.norm(yb)
    for b in range(B):
        ref = o.tt_compress(o.apply(A, xs[b]), 6)
        assert outs[b].ttv_rks == ref.ttv_rks
        assert o.rel_distance(outs[b], ref) < 1e-10
        assert abs(nb[b] - o.norm(ref)) < 1e-10 * o.norm(ref)


def test_cfg2_shape_property_checks():
    # full-size cfg2 (d=40, rank 512 -> 64): size-independent properties instead of a dense oracle
    import ttn_b200 as t
    rng = np.random.default_rng(1)
    d = 40
    x = o.rand_tt((2,) * d, 512, rng=rng, normalise=True)
    xd = t.DeviceTT.upload(x)
    n0 = t.norm(xd)
    yd = t.tt_compress_(xd.copy(), 64)
    assert max(yd.ttv_rks) == 64
    # idempotence: compressing again at the same bond changes nothing beyond rounding
    zd = t.tt_compress_(yd.copy(), 64)
    diff = t.orthogonalize(t.sub(zd, yd), 1).download()          # stable norm of the difference (centre core)
    assert np.linalg.norm(diff.ttv_vec[0]) / t.norm(yd) < 1e-10
    # projection property: <x, y> = <y, y> up to the truncation being (quasi-)optimal; error norm consistent
    err = t.norm(t.sub(xd, yd)) / n0
    assert 0.0 < err < 1.0
    # agreement with the CPU oracle on the same input.  Truncating a random TT with a flat spectrum is ill conditioned
    # (the gap sigma_64 - sigma_65 is ~1e-3 sigma_1 at every bond): the oracle itself moves by ~4e-10 when its input is
    # perturbed by 1e-16, so the comparison is made against that measured sensitivity, the retained singular values of
    # every bond step are compared to 1e-8, and the truncation error norms (well conditioned) to 1e-10.
    sig_ref = []
    ref = o.tt_compress(o.copy_tt(x), 64, sigma_out=sig_ref)
    pert = o.copy_tt(x)
    for k in range(d):
        pert.ttv_vec[k] = pert.ttv_vec[k] * (1 + 1e-16 * rng.standard_normal(pert.ttv_vec[k].shape))
    sens = o.rel_distance(o.tt_compress(pert, 64), ref)
    y2, sig = t.tt_compress_(xd.copy(), 64, return_sigma=True)
    got = y2.download()
    # 40 x: the 78 bond steps each carry GEMM rounding of ~sqrt(512) eps, i.e. a backward error of a few tens of eps in total
    assert o.rel_distance(got, ref) < max(1e-10, 40 * sens)
    for a, b in zip(sig, sig_ref):
        assert len(a) == len(b) and np.abs(a - b).max() / b[0] < 1e-8
    err_ref = o.rel_distance(ref, x)
    assert abs(err - err_ref) < 1e-10


@pytest.mark.gpu
def test_cuda_matches_golden():
    """The CUDA path against the committed vectors of tests/golden/hotpath_golden.npz (provenance in
    tests/golden/make_golden.py): cfg1 solution, tt_compress! result + per-bond singular values, orthogonalize, A*x,
    K_matfree, the Heisenberg ground-state energy (examples/heisenberg_xyz_dmrg.jl:9-19) and a prescribed-spectrum SVD."""
    import os
    import ttn_b200 as t
    import _golden
    g = _golden.load()     # Julia-made outputs (tests/golden/make_golden.jl) take precedence over the oracle-made ones
    print("golden provenance:", g["__provenance__"])

    def tt_from(prefix, d, dims=None):
        rks = [int(v) for v in g[prefix + "_rks"]]
        dims = (2,) * d if dims is None else tuple(int(v) for v in dims)
        return o.TTvector(d, [np.asfortranarray(g[f"{prefix}_core{k}"]) for k in range(d)], dims, rks, [0] * d)

    def dense(x):
        return o.ttv_to_tensor(x).reshape(-1)

    def rel(a, b):
        return np.linalg.norm(a - b) / np.linalg.norm(b)

    x = t.als_linsolve(o.id_tto(6), o.qtt_sin(6, lam=np.pi), tt_from("cfg1_x0", 6), sweep_count=4)
    assert rel(dense(x), g["cfg1_x"]) < 1e-10 and rel(dense(x), g["cfg1_b"]) < 1e-12
    y = tt_from("cmp_in", 8)
    z, sig = t.tt_compress_(o.copy_tt(y), 5, return_sigma=True)
    assert list(z.ttv_rks) == [int(v) for v in g["cmp_out_rks"]]
    assert rel(dense(z), g["cmp_out"]) < 1e-10
    gs = g["cmp_sigma"]
    for k in range(gs.shape[0]):
        n = min(len(sig[k]), gs.shape[1])
        assert np.abs(np.asarray(sig[k])[:n] - gs[k][:n]).max() < 1e-10 * gs[k][0]
    assert rel(dense(t.orthogonalize(y, 4)), g["orth_dense"]) < 1e-12
    assert rel(dense(t.apply(o.laplace_dd(8), y)), g["apply_dense"]) < 1e-12
    assert rel(t.matvec2(g["mv_G"], g["mv_Am"], g["mv_H"], g["mv_V"]), g["mv_Y"]) < 1e-13
    d = int(g["heis_d"])
    H = o.heisenberg_xyz_tto(d, jx=1.1, jy=0.8, jz=1.2, lam=0.0)
    x0 = o.rand_tt((2,) * d, 8, rng=np.random.default_rng(3), normalise=True)
    E, _, _ = t.dmrg_eigsolve(H, x0, N=2, tol=1e-12, sweep_schedule=[2, 4, 6], rmax_schedule=[8, 16, 32], linsolv_tol=1e-12,
                              linsolv_maxiter=20, krylovdim=20)
    assert abs(E[-1] - float(g["heis_e0"])) < 1e-9 * abs(float(g["heis_e0"]))
    U, s, Vt = t.svdtrunc(np.asfortranarray(g["svd_A"]))
    assert np.abs(s[:10] - g["svd_s"]).max() < 1e-12 and np.abs(s[10:]).max() < 1e-12
    # cases 8-12: hadamard_ttm, reorder, to_qtt, als_gen_eigsolv, the QFT example
    hz = t.hadamard_ttm(tt_from("had_x", 6, g["had_dims"]), tt_from("had_y", 6, g["had_dims"]), tol=1e-12)
    assert list(hz.ttv_rks) == [int(v) for v in g["had_out_rks"]] and rel(dense(hz), g["had_out"]) < 1e-10
    ry = t.reorder(tt_from("reo_x", 6), 2, 3, "serial", "interleaved")
    assert list(ry.ttv_rks) == [int(v) for v in g["reo_out_rks"]] and rel(dense(ry), g["reo_out"]) < 1e-10
    qq = t.to_qtt(tt_from("qtt_x", 3, g["qtt_dims"]), [[2, 2, 2], [4], [3, 2]])
    assert list(qq.ttv_rks) == [int(v) for v in g["qtt_out_rks"]] and rel(dense(qq), g["qtt_out"]) < 1e-10
    Ag = o.tto_add(o.laplace_dd(5), o.tto_scale(2.0, o.id_tto(5)))
    Sg = o.tto_add(o.id_tto(5), o.tto_scale(-0.15, o.tto_add(o.laplace_dd(5), o.tto_scale(-2.0, o.id_tto(5)))))
    Eg, _ = t.als_gen_eigsolv(Ag, Sg, tt_from("gen_x0", 5), sweep_schedule=[4], rmax_schedule=[2])
    assert len(Eg) == len(g["gen_E"]) and np.abs(Eg - g["gen_E"]).max() < 1e-9
    coeffs = g["dft_coeffs"]
    f = lambda xx: np.sum(coeffs * np.exp(2j * np.pi * np.arange(12) * xx))
    F, fx = o.fourier_qtto(10, K=50, sign=-1.0, normalize=True), o.function_to_qtt_uniform(f, 10)
    assert rel(o.matricize(t.tt_compress_(t.apply(F, fx), 100), 10), g["dft_spec"]) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("cplx", [False, True])
def test_hadamard_device_vs_oracle(cplx):
    """`hadamard` (tt_operations.jl:343-360) through the apply kernel: same tensor as the oracle's kron form and as the
    element-wise product of the dense tensors; followed by `tt_compress!` it stays within the truncation tolerance."""
    import ttn_b200 as t
    rng = np.random.default_rng(19)
    x = o.rand_tt((2,) * 7, 3, rng=rng); y = o.rand_tt((2,) * 7, 4, rng=rng)
    if cplx:
        x = o.complex_tt(x); y = o.complex_tt(y)
        y.ttv_vec[2] = y.ttv_vec[2] * (0.3 + 0.8j)
    z = t.hadamard(x, y)
    ref = o.ttv_to_tensor(x) * o.ttv_to_tensor(y)
    got = o.ttv_to_tensor(z)
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-13
    assert np.linalg.norm(got - o.ttv_to_tensor(o.hadamard(x, y))) / np.linalg.norm(ref) < 1e-13
    assert list(z.ttv_rks) == [a * b for a, b in zip(x.ttv_rks, y.ttv_rks)]
    zc = t.tt_compress_(z, 12)
    assert np.linalg.norm(o.ttv_to_tensor(zc) - ref) / np.linalg.norm(ref) < 1e-10


@pytest.mark.gpu
def test_dft_example_on_device():
    """examples/dft.jl:5-25 on the device: ComplexF64 `A*x` (MPO rank 51) + `tt_compress!(·, 100)`; the spectrum of a
    12-mode signal comes back to 1e-8 and the rest of the spectrum is below 1e-10, as the example asserts; the train also
    matches the oracle's to the tolerance of the path."""
    import ttn_b200 as t
    d, K, r = 10, 50, 12
    rng = np.random.default_rng(1234)
    coeffs = rng.standard_normal(r) + 1j * rng.standard_normal(r)
    f = lambda x: np.sum(coeffs * np.exp(2j * np.pi * np.arange(r) * x))
    F, x = o.fourier_qtto(d, K=K, sign=-1.0, normalize=True), o.function_to_qtt_uniform(f, d)
    y = t.tt_compress_(t.apply(F, x), 100)
    spec = o.matricize(y, d)
    scale = np.sqrt(2.0 ** d)
    assert np.linalg.norm(spec[:r] - scale * coeffs) / (scale * np.linalg.norm(coeffs)) < 1e-8
    assert np.linalg.norm(spec[r:]) / np.linalg.norm(spec) < 1e-10
    yo = o.tt_compress(o.apply(F, x), 100)
    assert np.linalg.norm(spec - o.matricize(yo, d)) / np.linalg.norm(spec) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("ordering", ["serial", "interleaved"])
def test_laplacian_2d_action_both_orderings(ordering):
    """test/test_qtt_multidim.jl:658-692: (Δ⊗I + I⊗Δ)/h² applied to sin(πx)sin(πy) on an 8 x 8 grid in serial and interleaved
    bit order against the dense Kronecker-sum matrix; the two orderings are tied together by `reorder`."""
    import ttn_b200 as t
    bits = 3
    n = 2 ** bits
    h = 1.0 / (n - 1)
    M1 = o.tto_to_matrix(o.laplace_dd(bits)) / h ** 2
    M2 = np.kron(M1, np.eye(n)) + np.kron(np.eye(n), M1)
    sx = o.qtt_sin(bits, lam=1.0)
    xs = o.qtt_to_vector(sx)
    ref = (M2 @ np.kron(xs, xs)).reshape(n, n)                              # [x, y]
    L, I = o.tto_scale(1.0 / h ** 2, o.laplace_dd(bits)), o.id_tto(bits)
    cat = lambda a, b: o.TToperator(2 * bits, a.tto_vec + b.tto_vec, (2,) * (2 * bits), list(a.tto_rks) + list(b.tto_rks)[1:])
    A_serial = o.tto_add(cat(L, I), cat(I, L))
    v_serial = o.TTvector(2 * bits, [c.copy() for c in sx.ttv_vec] * 2, (2,) * (2 * bits), list(sx.ttv_rks) + list(sx.ttv_rks)[1:],
                          [0] * (2 * bits))
    if ordering == "serial":
        Av = t.apply(A_serial, v_serial)
        got = o.ttv_to_tensor(Av).reshape(n, n)                             # C-order: x bits then y bits, MSB first
    else:
        A_il = o.laplace2d_interleaved(bits)
        v_il = t.reorder(v_serial, 2, bits, "serial", "interleaved")
        assert np.allclose(o.ttv_to_tensor(v_il), o.ttv_to_tensor(o.qtt_sin2d_interleaved(bits)), atol=1e-12)
        Av = t.reorder(t.apply(A_il, v_il), 2, bits, "interleaved", "serial", threshold=1e-14)
        got = o.ttv_to_tensor(Av).reshape(n, n)
    assert np.max(np.abs(got - ref)) < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("R", [51, 64, 96])
def test_apply_large_mpo_cores(R):
    """`A*x` with MPO cores beyond the default 48 KB of shared memory: staged with the opt-in limit (R = 51: 166 KB, the QFT
    operator of examples/dft.jl) or read from global memory (R = 64: 262 KB real would fit, ComplexF64 does not; R = 96)."""
    import ttn_b200 as t
    rng = np.random.default_rng(R)
    cplx = lambda shp: (rng.standard_normal(shp) + 1j * rng.standard_normal(shp)) / np.sqrt(R)
    A = o.TToperator(3, [cplx((2, 2, 1, R)), cplx((2, 2, R, R)), cplx((2, 2, R, 1))], (2, 2, 2), [1, R, R, 1])
    x = o.rand_tt((2, 2, 2), 2, rng=rng, dtype=np.complex128)
    y = t.apply(A, x)
    ref = o.ttv_to_tensor(o.apply(A, x))
    assert list(y.ttv_rks) == [1, 2 * R, 2 * R, 1]
    assert np.linalg.norm(o.ttv_to_tensor(y) - ref) / np.linalg.norm(ref) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("cplx", [False, True])
def test_distances_and_add_inplace(cplx):
    """`euclidean_distance`, `euclidean_distance_normalized` (tt_operations.jl:452-460) and `add!` (:36-66; test_tt_operations.jl
    :106-114) against the dense tensors."""
    import ttn_b200 as t
    rng = np.random.default_rng(41)
    dt = np.complex128 if cplx else np.float64
    a = o.rand_tt((2, 3, 2, 2), 3, rng=rng, dtype=dt); b = o.rand_tt((2, 3, 2, 2), 2, rng=rng, dtype=dt)
    A, B = o.ttv_to_tensor(a), o.ttv_to_tensor(b)
    assert abs(t.euclidean_distance(a, b) - np.linalg.norm(A - B)) < 1e-12 * np.linalg.norm(A)
    assert abs(t.euclidean_distance_normalized(a, b) - np.linalg.norm(A - B) / np.linalg.norm(B)) < 1e-12
    assert t.euclidean_distance(a, a) < 1e-7 * np.linalg.norm(A)           # cancellation floor sqrt(eps) of the formula
    x = o.copy_tt(a)
    y = t.add_(x, b)
    assert y is x and list(x.ttv_rks) == [1] + [p + q for p, q in zip(a.ttv_rks[1:-1], b.ttv_rks[1:-1])] + [1]
    assert all(v == 0 for v in x.ttv_ot)
    assert np.allclose(o.ttv_to_tensor(x), A + B, atol=1e-12)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_tto_round_trip(dtype):
    # TToperator storage (src/tt_tools.jl:48-54): upload -> metadata -> download is bit-exact, Base.complex promotes (tt_tools.jl:59-61)
    import ttn_b200 as t
    rng = np.random.default_rng(41)
    dims = (2, 3, 2, 4)
    A = o.rand_tto(dims, 3, rng=rng, dtype=dtype)
    Ad = t.DeviceTTO.upload(A)
    assert Ad.tto_dims == tuple(dims) and Ad.tto_rks == list(A.tto_rks)
    B = Ad.download()
    for a, b in zip(A.tto_vec, B.tto_vec):
        assert a.shape == b.shape and np.array_equal(np.asarray(a, dtype=dtype), b)
    C = Ad.complex().download()
    for a, c in zip(A.tto_vec, C.tto_vec):
        assert c.dtype == np.complex128 and np.array_equal(np.asarray(a, dtype=np.complex128), c)
