import sys, time
sys.path.insert(0, '.')
import numpy as np, ttn_b200 as t
from ttn_b200 import _lib
lib = _lib.lib()
rng = np.random.default_rng(0)
shapes = [(256, 256), (512, 512), (1024, 1024), (2048, 2048), (1024, 300)] if len(sys.argv) < 2 else [tuple(map(int, a.split('x'))) for a in sys.argv[1:]]
for shape in shapes:
    A = np.asfortranarray(rng.standard_normal(shape))
    t.svdtrunc(A[:8, :8].copy(order='F'))
    t.profile(True)
    t0 = time.time(); U, s, Vt = t.svdtrunc(A); dt = time.time() - t0
    fam = t.profile_read(); t.profile(False)
    sref = np.linalg.svd(A, compute_uv=False)
    print(shape, "sweeps", lib.ttn_last_jacobi_sweeps(), "sigma err %.2e" % (np.abs(s - sref).max() / sref[0]),
          "orth %.2e" % np.abs(U.T @ U - np.eye(len(s))).max(), "recon %.2e" % (np.linalg.norm((U * s) @ Vt - A) / np.linalg.norm(A)),
          "wall ms %.1f" % (dt * 1e3), {k: round(v[0], 1) for k, v in fam.items() if v[0] > 0.05})
