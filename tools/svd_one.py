"""One SVD of an m x n random matrix through the C ABI (profiling target).  usage: svd_one.py m n [complex]"""
import sys
sys.path.insert(0, '.')
import numpy as np, ttn_b200 as t
m, n = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(0)
A = rng.standard_normal((m, n))
if len(sys.argv) > 3:
    A = A + 1j * rng.standard_normal((m, n))
A = np.asfortranarray(A)
for _ in range(3):
    U, s, Vt = t.svdtrunc(A)
print("ok", s[0])
