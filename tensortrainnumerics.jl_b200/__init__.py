"""tensortrainnumerics.jl_b200 — B200 (sm_100a) implementation of TensorTrainNumerics.jl's core-contraction
hot path behind the reference's own API (host-side mirror in Python over the C ABI of libttn_b200.so).

The directory name contains a dot, so import it through the top-level shim:  ``import ttn_b200``.
"""
from .api import *  # noqa: F401,F403
from .api import __all__  # noqa: F401
from .krylov_tt import krylov_linsolve  # noqa: F401,E402
from .steppers import (euler_method, implicit_euler_method, crank_nicholson_method, rk4_method,  # noqa: F401,E402
                       id_tto, tto_add, tto_scale)
from .sites import (hadamard, ttv_to_diag_tto, hadamard_ttm, swap_adjacent_sites_, bubble_sort_swaps,  # noqa: F401,E402
                    reorder, to_qtt, QTTvector, dmrg_cross_superblock_split)
