"""Oracle checks for the site-surgery routines against dense ground truth, following the
reference's own tests: test/test_qtt_tools.jl (`reorder`, `to_qtt`) and test/test_tt_operations.jl (`hadamard_ttm`)."""
import numpy as np
import pytest

import ttn_oracle as o


def _axes_after_reorder(n_dims, bits, ordering):
    perm = o.reorder_perm(n_dims, bits, ordering)
    axes = [0] * len(perm)
    for src, tgt in enumerate(perm):
        axes[tgt] = src
    return axes


def test_bubble_sort_swaps_sorts():
    rng = np.random.default_rng(0)
    for _ in range(10):
        perm = list(rng.permutation(7))
        p = list(perm)
        for k in o.bubble_sort_swaps(perm):
            p[k - 1], p[k] = p[k], p[k - 1]
        assert p == sorted(perm)
    assert o.bubble_sort_swaps([0, 1, 2]) == []


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_swap_adjacent_sites_dense(dtype):
    rng = np.random.default_rng(1)
    x = o.rand_tt((2, 3, 2, 2), 3, rng=rng, dtype=dtype)
    A, B = o.swap_adjacent_sites(x.ttv_vec[1], x.ttv_vec[2])
    y = o.TTvector(4, [x.ttv_vec[0], A, B, x.ttv_vec[3]], (2, 2, 3, 2), [1, x.ttv_rks[1], A.shape[2], x.ttv_rks[3], 1], [0] * 4)
    assert np.allclose(o.ttv_to_tensor(y), np.transpose(o.ttv_to_tensor(x), (0, 2, 1, 3)), atol=1e-13)


@pytest.mark.parametrize("n_dims,bits", [(2, 3), (3, 2)])
def test_reorder_roundtrip_and_dense(n_dims, bits):
    rng = np.random.default_rng(2)
    x = o.rand_tt((2,) * (n_dims * bits), 4, rng=rng)
    y = o.reorder(x, n_dims, bits, "serial", "interleaved")
    assert np.allclose(o.ttv_to_tensor(y), np.transpose(o.ttv_to_tensor(x), _axes_after_reorder(n_dims, bits, "serial")), atol=1e-12)
    z = o.reorder(y, n_dims, bits, "interleaved", "serial")
    assert np.allclose(o.ttv_to_tensor(z), o.ttv_to_tensor(x), atol=1e-12)
    same = o.reorder(x, n_dims, bits, "serial", "serial")
    assert all(np.array_equal(a, b) for a, b in zip(same.ttv_vec, x.ttv_vec))


def test_hadamard_ttm_dense():
    rng = np.random.default_rng(3)
    x = o.rand_tt((2, 3, 2, 2, 2), 3, rng=rng); y = o.rand_tt((2, 3, 2, 2, 2), 2, rng=rng)
    ref = o.ttv_to_tensor(x) * o.ttv_to_tensor(y)
    z = o.hadamard_ttm(x, y)
    assert tuple(z.ttv_dims) == tuple(x.ttv_dims) and z.ttv_rks[0] == z.ttv_rks[-1] == 1
    assert np.linalg.norm(o.ttv_to_tensor(z) - ref) / np.linalg.norm(ref) < 1e-12
    zt = o.hadamard_ttm(x, y, tol=1e-12, rmax=3)
    assert max(zt.ttv_rks) <= 3


def test_to_qtt_dense():
    rng = np.random.default_rng(4)
    x = o.rand_tt((8, 4, 6), 3, rng=rng)
    q = o.to_qtt(x, [[2, 2, 2], [4], [3, 2]])
    assert tuple(q.ttv_dims) == (2, 2, 2, 4, 3, 2)
    full = o.ttv_to_tensor(x)
    # big-endian split: the first factor is the coarsest digit  ->  C-order reshape of each axis
    assert np.allclose(o.ttv_to_tensor(q), full.reshape(2, 2, 2, 4, 3, 2), atol=1e-12)


def _hadamard_cases(d=8):
    """test/test_tt_operations.jl:41-104 (the exp / sin / cos cases; λ = π means sin(π² x))."""
    xs = np.linspace(0.0, 1.0, 2 ** d)
    A1, A2, A3 = o.qtt_exp(d), o.qtt_sin(d, lam=np.pi), o.qtt_cos(d, lam=np.pi)
    return [(A2, A3, np.cos(np.pi ** 2 * xs) * np.sin(np.pi ** 2 * xs)), (A1, A2, np.exp(xs) * np.sin(np.pi ** 2 * xs))]


def test_hadamard_function_reconstruction():
    for x, y, expected in _hadamard_cases():
        assert np.allclose(o.qtt_to_vector(o.hadamard(x, y)), expected, atol=1e-12)            # :51-58
        z = o.hadamard_ttm(x, y)
        assert np.allclose(o.qtt_to_vector(z), expected, atol=1e-10)                            # :84-88
        h = o.hadamard(x, y)
        assert o.euclidean_distance(z, h) / o.norm(h) < 1e-5                                    # :98-99
    x, y, expected = _hadamard_cases()[1]
    assert np.allclose(o.qtt_to_vector(o.hadamard_ttm(x, y, tol=1e-8)), expected, atol=1e-3)    # :103-104
