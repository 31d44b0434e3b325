"""SVD probe: sweeps, accuracy and wall time of ttn_b200.svdtrunc on bond-sized matrices (run on the GPU box)."""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np, ttn_b200 as t
from ttn_b200 import _lib
lib = _lib.lib()
rng = np.random.default_rng(0)
shapes = [(128, 128), (128, 1024), (64, 64), (256, 256), (128, 512), (100, 77), (200, 130), (48, 48), (33, 200)]
for shape in shapes:
    A = np.asfortranarray(rng.standard_normal(shape))
    t.svdtrunc(A)
    t.profile(True)
    t0 = time.time(); U, s, Vt = t.svdtrunc(A); dt = time.time() - t0
    fam = t.profile_read(); t.profile(False)
    sref = np.linalg.svd(A, compute_uv=False)
    print(shape, "sweeps", lib.ttn_last_jacobi_sweeps(), "sigma err %.2e" % (np.abs(s - sref).max() / sref[0]),
          "orth %.2e" % np.abs(U.T @ U - np.eye(len(s))).max(), "recon %.2e" % (np.linalg.norm((U * s) @ Vt - A) / np.linalg.norm(A)),
          "wall ms %.2f" % (dt * 1e3), {k: round(v[0], 3) for k, v in fam.items() if v[0] > 0.01})
for shape in [(128, 512), (128, 128), (200, 96)]:
    Ac = np.asfortranarray(rng.standard_normal(shape) + 1j * rng.standard_normal(shape))
    t.profile(True)
    U, s, Vt = t.svdtrunc(Ac)
    fam = t.profile_read(); t.profile(False)
    print("complex", shape, "sweeps", lib.ttn_last_jacobi_sweeps(), "sigma err %.2e" % (np.abs(s - np.linalg.svd(Ac, compute_uv=False)).max() / s[0]),
          "recon %.2e" % (np.linalg.norm((U * s) @ Vt - Ac) / np.linalg.norm(Ac)), {k: round(v[0], 3) for k, v in fam.items() if v[0] > 0.01})
# rank-deficient product (the R->L bonds of tt_compress!: Theta = A B with inner dimension 64)
A = np.asfortranarray(rng.standard_normal((128, 64)) @ rng.standard_normal((64, 128)))
t.profile(True)
U, s, Vt = t.svdtrunc(A)
fam = t.profile_read(); t.profile(False)
sref = np.linalg.svd(A, compute_uv=False)
print("rank-64 128x128: sweeps", lib.ttn_last_jacobi_sweeps(), "sigma err %.2e" % (np.abs(s - sref).max() / sref[0]),
      "recon %.2e" % (np.linalg.norm((U * s) @ Vt - A) / np.linalg.norm(A)), "orth(64) %.2e" % np.abs(U[:, :64].T @ U[:, :64] - np.eye(64)).max(),
      {k: round(v[0], 3) for k, v in fam.items() if v[0] > 0.01})
