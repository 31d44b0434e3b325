"""Generates tests/golden/hotpath_golden.npz — small seeded input/output vectors of the hot path.

PROVENANCE: produced by the NumPy/SciPy oracle (oracle/ttn_oracle, a restatement of TensorTrainNumerics.jl v1.1.3), NOT by
the Julia reference itself: Julia is not installed in the build image, so raw reference outputs cannot be generated
("parity unpinned").  Every case is additionally tied to dense ground truth at generation time
(asserted below), so the file pins the *mathematics* of the reference path; it is a regression anchor for the oracle
(tests/test_oracle_tt.py::test_oracle_reproduces_golden) and for the CUDA path (tests/test_gpu_tt.py::test_cuda_matches_golden).

  python tests/golden/make_golden.py        # rewrites the .npz (deterministic: fixed seeds)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ttn_oracle as o  # noqa: E402


def dense(x):
    return o.ttv_to_tensor(x).reshape(-1)


def cores_to_dict(prefix, x, out):
    out[prefix + "_rks"] = np.array(x.ttv_rks)
    for k, c in enumerate(x.ttv_vec):
        out[f"{prefix}_core{k}"] = np.asarray(c)


def main():
    g = {}
    # 1. cfg1 (README.md:82-102): als_linsolve(id_tto(6), qtt_sin(6, lam=pi), random start, sweep_count=4)
    d = 6
    A, b = o.id_tto(d), o.qtt_sin(d, lam=np.pi)
    x0 = o.rand_tt((2,) * d, b.ttv_rks, rng=np.random.default_rng(0))
    x = o.als_linsolve(A, b, x0, sweep_count=4)
    assert np.linalg.norm(dense(x) - dense(b)) / np.linalg.norm(dense(b)) < 1e-12
    cores_to_dict("cfg1_x0", x0, g)
    g["cfg1_b"] = dense(b)
    g["cfg1_x"] = dense(x)
    # 2. tt_compress! (tt_tools.jl:772-789) of a seeded TT: d=8, rank 12 -> 5; per-bond singular values + result
    d = 8
    y = o.rand_tt((2,) * d, 12, rng=np.random.default_rng(1), normalise=True)
    sig = []
    z = o.tt_compress(o.copy_tt(y), 5, sigma_out=sig)
    cores_to_dict("cmp_in", y, g)
    g["cmp_out"] = dense(z)
    g["cmp_out_rks"] = np.array(z.ttv_rks)
    smax = max(len(s) for s in sig)
    g["cmp_sigma"] = np.array([list(s) + [0.0] * (smax - len(s)) for s in sig])
    # 3. orthogonalize (tt_tools.jl:511-543): same tensor, centre 4
    w = o.orthogonalize(y, i=4)
    assert np.linalg.norm(dense(w) - dense(y)) / np.linalg.norm(dense(y)) < 1e-12
    g["orth_dense"] = dense(w)
    # 4. A*x (tt_operations.jl:101-111) against the dense product
    Aop = o.laplace_dd(d)
    Ay = o.apply(Aop, y)
    ref = o.tto_to_matrix(Aop) @ dense(y)
    assert np.linalg.norm(dense(Ay) - ref) / np.linalg.norm(ref) < 1e-12
    g["apply_dense"] = dense(Ay)
    # 5. K_matfree (dmrg.jl:239-244), single application, seeded operands
    rng = np.random.default_rng(4)
    G = rng.standard_normal((3, 7, 7)); H = rng.standard_normal((4, 6, 6))
    Am = rng.standard_normal((3, 4, 4, 4)); V = rng.standard_normal((7, 4, 6))
    g["mv_G"], g["mv_H"], g["mv_Am"], g["mv_V"] = G, H, Am, V
    g["mv_Y"] = o.dmrg_matvec2(G, Am, V, H, symmetrize=False)
    # 6. examples/heisenberg_xyz_dmrg.jl:9-19: DMRG ground-state energy of the XYZ chain, d = 10, against eigvalsh
    d = 10
    Hh = o.heisenberg_xyz_tto(d, jx=1.1, jy=0.8, jz=1.2, lam=0.0)
    e0 = np.linalg.eigvalsh(o.tto_to_matrix(Hh).real)[0]
    g["heis_d"] = np.array(d)
    g["heis_e0"] = np.array(e0)
    # 7. `_svdtrunc` (tt_cross_interpolation.jl:149-166) on a prescribed spectrum
    rng = np.random.default_rng(7)
    U, _ = np.linalg.qr(rng.standard_normal((24, 10))); Vv, _ = np.linalg.qr(rng.standard_normal((40, 10)))
    s = np.logspace(0, -9, 10)
    g["svd_A"] = (U * s) @ Vv.T
    g["svd_s"] = s
    # 8. hadamard_ttm (tt_operations.jl:399-422), tol 1e-12: result against the element-wise product of the dense tensors
    rng = np.random.default_rng(8)
    hx = o.rand_tt((2, 3, 2, 2, 2, 2), 3, rng=rng); hy = o.rand_tt((2, 3, 2, 2, 2, 2), 2, rng=rng)
    hz = o.hadamard_ttm(hx, hy, tol=1e-12)
    href = o.ttv_to_tensor(hx) * o.ttv_to_tensor(hy)
    assert np.linalg.norm(o.ttv_to_tensor(hz) - href) / np.linalg.norm(href) < 1e-11
    cores_to_dict("had_x", hx, g); cores_to_dict("had_y", hy, g)
    g["had_dims"] = np.array(hx.ttv_dims)
    g["had_out"] = dense(hz)
    g["had_out_rks"] = np.array(hz.ttv_rks)
    # 9. reorder serial -> interleaved (qtt_tools.jl:731-774), 2 dims x 3 bits: the dense tensor with its axes permuted
    rx = o.rand_tt((2,) * 6, 4, rng=np.random.default_rng(9))
    ry = o.reorder(rx, 2, 3, "serial", "interleaved")
    axes = [0] * 6
    for src, tgt in enumerate(o.reorder_perm(2, 3, "serial")):
        axes[tgt] = src
    assert np.allclose(o.ttv_to_tensor(ry), np.transpose(o.ttv_to_tensor(rx), axes), atol=1e-12)
    cores_to_dict("reo_x", rx, g)
    g["reo_out"] = dense(ry)
    g["reo_out_rks"] = np.array(ry.ttv_rks)
    # 10. to_qtt (qtt_tools.jl:254-310): (8, 4, 6) -> (2, 2, 2, 4, 3, 2)
    qx = o.rand_tt((8, 4, 6), 3, rng=np.random.default_rng(10))
    qq = o.to_qtt(qx, [[2, 2, 2], [4], [3, 2]])
    assert np.allclose(o.ttv_to_tensor(qq), o.ttv_to_tensor(qx).reshape(qq.ttv_dims), atol=1e-12)
    cores_to_dict("qtt_x", qx, g)
    g["qtt_dims"] = np.array(qx.ttv_dims)
    g["qtt_out"] = dense(qq)
    g["qtt_out_rks"] = np.array(qq.ttv_rks)
    # 11. als_gen_eigsolv (als.jl:344-440): energies of every local solve, rank-2 start (full-rank unfoldings), and the
    #     lowest eigenvalue of the dense pencil the sweeps converge to
    import scipy.linalg as sla
    d = 5
    Ag = o.tto_add(o.laplace_dd(d), o.tto_scale(2.0, o.id_tto(d)))
    Sg = o.tto_add(o.id_tto(d), o.tto_scale(-0.15, o.tto_add(o.laplace_dd(d), o.tto_scale(-2.0, o.id_tto(d)))))
    gx0 = o.rand_tt((2,) * d, 2, rng=np.random.default_rng(11), normalise=True)
    Eg, xg = o.als_gen_eigsolv(Ag, Sg, gx0, sweep_schedule=[4], rmax_schedule=[2])
    lam = sla.eigh(o.tto_to_matrix(Ag), o.tto_to_matrix(Sg), eigvals_only=True)[0]
    assert Eg[-1] >= lam - 1e-12 and Eg[-1] - lam < 1e-3          # rank 2 is not exact for this pencil, but close
    cores_to_dict("gen_x0", gx0, g)
    g["gen_E"] = np.array(Eg)
    g["gen_lam"] = np.array(lam)
    # 12. examples/dft.jl:5-25: spectrum of a 12-mode signal through A*x + tt_compress!(., 100) with the rank-51 QFT operator
    d, K, r = 10, 50, 12
    rng = np.random.default_rng(1234)
    coeffs = rng.standard_normal(r) + 1j * rng.standard_normal(r)
    f = lambda x: np.sum(coeffs * np.exp(2j * np.pi * np.arange(r) * x))
    F, fx = o.fourier_qtto(d, K=K, sign=-1.0, normalize=True), o.function_to_qtt_uniform(f, d)
    spec = o.matricize(o.tt_compress(o.apply(F, fx), 100), d)
    assert np.linalg.norm(spec[:r] - 32.0 * coeffs) / (32.0 * np.linalg.norm(coeffs)) < 1e-8
    g["dft_coeffs"] = coeffs
    g["dft_spec"] = spec
    np.savez_compressed(os.path.join(HERE, "hotpath_golden.npz"), **g)
    print("wrote", os.path.join(HERE, "hotpath_golden.npz"), len(g), "arrays")


if __name__ == "__main__":
    main()
