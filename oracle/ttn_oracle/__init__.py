"""CPU oracle: a NumPy/SciPy restatement of TensorTrainNumerics.jl's core-contraction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: it may be imported
by ``tests/``, by ``__graft_entry__.smoke()`` and by ``bench.py``'s CPU-baseline / reference arm,
and only as the *checker* (or as the timed CPU baseline), never as the thing shipped.  The product
path (``tensortrainnumerics.jl_b200``) never imports this package and fails loudly when its CUDA
library is missing.

Pinning status: the reference is pure Julia and Julia is not installed in the build container, so
the reference itself cannot be executed here ("parity unpinned" for raw RNG-stream outputs).  The
restatement is pinned instead against every known-answer / dense-oracle test the reference's own
test-suite holds for this path (see ``tests/test_oracle_*.py`` — each test cites the reference test
it ports) and against dense ground truth (``numpy.linalg``) on small problems.

Every function cites the reference ``file:line`` (relative to the reference repository root) that
it follows.  Arrays use Julia's index order: TT core ``X[s, a, b]`` of shape ``(n, r_left, r_right)``,
MPO core ``A[i, j, a, b]`` of shape ``(n_out, n_in, R_left, R_right)``; Julia ``reshape`` is
``numpy.reshape(..., order="F")``.
"""
from .core import (TTvector, TToperator, zeros_tt, zeros_tto, rand_tt, rand_tto, r_and_d_to_rks,
                   ttv_to_tensor, tto_to_matrix, ttv_decomp, copy_tt, complex_tt, complex_tto,
                   increase_ranks)
from .generators import (toeplitz_to_qtto, laplace_dd, id_tto, heisenberg_xyz_tto, qtt_sin, qtt_cos,
                         qtt_to_vector, tto_add, tto_scale, laplace2d_interleaved, qtt_sin2d_interleaved,
                         shift_op, fourier_qtto, function_to_qtt_uniform, matricize,
                         qtt_exp)
from .ops import (apply, add, scale, sub, dot, norm, orthogonalize, svdtrunc, svdtrunc_abs,
                  tt_bond_truncate, tt_compress, euclidean_distance, rel_distance, norm_stable, hadamard)
from .als import als_linsolve, als_eigsolve, als_gen_eigsolv, K_eiggenmin
from .mals import mals_linsolve, mals_eigsolve, sv_trunc
from .dmrg import dmrg_linsolve, dmrg_eigsolve, cut_off_index, dmrg_matvec2, dmrg_matvec2_blas, dmrg_update_G, dmrg_update_H, amid
from .tdvp import tdvp, tdvp2, apply_H1_lsr, apply_H0, apply_H2_lsr, update_left_env, update_right_env
from .krylov_tt import krylov_linsolve
from .steppers import euler_method, implicit_euler_method, crank_nicholson_method, rk4_method
from .sites import (swap_adjacent_sites, bubble_sort_swaps, reorder_perm, reorder, ttm_swap, ttm_contract, hadamard_ttm,
                    to_qtt)
