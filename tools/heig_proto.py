"""NumPy prototype of the small Hermitian eigensolver behind the Gram path of tt_compress! (csrc/heig.cu).
Mirrors the kernel step by step: Householder tridiagonalisation (LAPACK zhetd2 'L' conventions), multisection on the
scaled division-free Sturm sequence, twisted-factorisation eigenvectors, MGS inside clusters, back-transformation."""
import numpy as np


def tridiag(A):
    A = A.astype(np.complex128 if np.iscomplexobj(A) else np.float64).copy()
    n = A.shape[0]
    d = np.zeros(n); e = np.zeros(max(n - 1, 0)); tau = np.zeros(max(n - 1, 0), dtype=A.dtype)
    for k in range(n - 1):
        x = A[k + 1:, k].copy()
        alpha = x[0]
        xn2 = np.sum(np.abs(x[1:]) ** 2)
        if xn2 == 0 and np.imag(alpha) == 0:
            t = 0.0; beta = np.real(alpha); v = x.copy(); v[0] = 1
        else:
            beta = -np.copysign(np.sqrt(np.abs(alpha) ** 2 + xn2), np.real(alpha))
            t = (beta - alpha) / beta          # LAPACK zlarfg: tau = ((beta-alphr)/beta, -alphi/beta)
            if np.iscomplexobj(A):
                t = complex((beta - alpha.real) / beta, -alpha.imag / beta)
            v = x / (alpha - beta); v[0] = 1
        e[k] = beta
        if t != 0:
            A22 = A[k + 1:, k + 1:]
            # hemv with the lower triangle only
            L = np.tril(A22); H = L + L.conj().T - np.diag(np.diag(L).real) if np.iscomplexobj(A) else L + L.T - np.diag(np.diag(L))
            p = t * (H @ v)
            a2 = -0.5 * t * np.vdot(p, v)
            w = p + a2 * v
            A22 -= np.outer(v, w.conj()) + np.outer(w, v.conj())
        d[k] = np.real(A[k, k])
        A[k + 1:, k] = v; tau[k] = t
    d[n - 1] = np.real(A[n - 1, n - 1])
    return d, e, tau, A   # reflector k: v = [1; A[k+2:,k]] acting on rows k+1..n-1


def sturm_count(d, e2, lam):
    """number of eigenvalues < lam; division-free scaled recurrence; d, e2 already scaled so that |T| <= 1"""
    n = len(d)
    cnt = 0
    pm1 = 1.0; p = d[0] - lam
    if p == 0.0: p = -1e-300
    if p < 0: cnt += 1
    for i in range(1, n):
        pn = (d[i] - lam) * p - e2[i - 1] * pm1
        if pn == 0.0: pn = -np.copysign(1e-300, p) if p != 0 else -1e-300
        if (pn < 0) != (p < 0): cnt += 1
        pm1, p = p, pn
        a = max(abs(p), abs(pm1))
        if (i & 7) == 7:
            if a < 1e-100: p *= 1e100; pm1 *= 1e100
            elif a > 1e100: p *= 1e-100; pm1 *= 1e-100
    return cnt


def bisect_top(d, e, nev, M=8, rounds=None):
    """top nev eigenvalues (descending) by multisection with M interior points"""
    n = len(d)
    ea = np.abs(np.concatenate([[0.0], e, [0.0]]))
    gl = np.min(d - ea[:-1] - ea[1:]); gu = np.max(d + ea[:-1] + ea[1:])
    tn = max(abs(gl), abs(gu))
    if tn == 0: return np.zeros(nev), 1.0
    sc = 1.0 / tn
    ds = d * sc; e2 = (e * sc) ** 2
    lo0 = gl * sc - 2e-16 * n - 1e-300; hi0 = gu * sc + 2e-16 * n + 1e-300
    if rounds is None: rounds = int(np.ceil(56 / np.log2(M + 1)))
    lam = np.zeros(nev)
    for j in range(nev):
        idx = n - 1 - j            # ascending index wanted
        lo, hi = lo0, hi0
        for _ in range(rounds):
            pts = lo + (hi - lo) * (np.arange(1, M + 1) / (M + 1))
            cnts = np.array([sturm_count(ds, e2, x) for x in pts])
            # eigenvalue idx in (lo,hi): count(lo) <= idx < count(hi)
            below = pts[cnts <= idx]; above = pts[cnts > idx]
            if len(below): lo = below.max()
            if len(above): hi = above.min()
        lam[j] = 0.5 * (lo + hi)
    return lam * tn, tn


def twisted_vec(d, e, lam, pivmin):
    """eigenvector of tridiag(d,e) for eigenvalue lam by the twisted factorisation (Parlett-Dhillon getvec)"""
    n = len(d)
    if n == 1: return np.ones(1), 0.0
    # stationary qd: L D L^T = T - lam: dplus_i, lplus_i; s_i
    s = np.zeros(n); lplus = np.zeros(n - 1)
    s[0] = d[0] - lam
    for i in range(n - 1):
        dp = s[i]
        if abs(dp) < pivmin: dp = -pivmin
        lplus[i] = e[i] / dp
        s[i + 1] = d[i + 1] - lam - lplus[i] * e[i]
    # progressive: U D U^T from the bottom: p_i
    p = np.zeros(n); uminus = np.zeros(n - 1)
    p[n - 1] = d[n - 1] - lam
    for i in range(n - 2, -1, -1):
        dm = p[i + 1]
        if abs(dm) < pivmin: dm = -pivmin
        uminus[i] = e[i] / dm
        p[i] = d[i] - lam - uminus[i] * e[i]
    gamma = s + p - (d - lam)
    r = int(np.argmin(np.abs(gamma)))
    z = np.zeros(n); z[r] = 1.0
    for i in range(r - 1, -1, -1):
        z[i] = -lplus[i] * z[i + 1]
    for i in range(r, n - 1):
        z[i + 1] = -uminus[i] * z[i]
    nz = np.linalg.norm(z)
    return z / nz, abs(gamma[r]) / nz


def heig_top(G, nev, M=8, ctol=1e-3, max_cluster=8):
    n = G.shape[0]
    d, e, tau, A = tridiag(G)
    lam, tn = bisect_top(d, e, nev, M)
    pivmin = max(2.2e-308, 1e-300) * max(1.0, np.max(e ** 2) if n > 1 else 1.0)
    pivmin = 1e-290 * tn if tn > 0 else 1e-290
    Z = np.zeros((n, nev)); res = np.zeros(nev)
    for j in range(nev):
        Z[:, j], res[j] = twisted_vec(d, e, lam[j], 2.3e-16 * tn * 1e-3 * 0 + pivmin)
    flag = 0
    # MGS inside clusters (descending order; cluster when gap < ctol*tn)
    start = 0
    for j in range(1, nev + 1):
        if j == nev or lam[j - 1] - lam[j] >= ctol * tn:
            c = j - start
            if c > max_cluster: flag |= 1
            for a in range(start + 1, j):
                for b in range(start, a):
                    Z[:, a] -= np.dot(Z[:, b], Z[:, a]) * Z[:, b]
                nz = np.linalg.norm(Z[:, a])
                if nz < 1e-3: flag |= 2
                Z[:, a] /= nz
            start = j
    if np.max(res) > 1e-10 * tn: flag |= 4
    # back-transform: U = H(0) ... H(n-2) Z
    U = Z.astype(G.dtype)
    for k in range(n - 2, -1, -1):
        v = np.concatenate([[1.0], A[k + 2:, k]])
        w = v.conj() @ U[k + 1:, :]
        U[k + 1:, :] -= tau[k] * np.outer(v, w)
    return lam, U, flag


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for cplx in (False, True):
        for n, q in ((128, 1024), (64, 128), (5, 9), (2, 3), (1, 4), (96, 300)):
            Th = rng.standard_normal((n, q)) + (1j * rng.standard_normal((n, q)) if cplx else 0)
            G = Th @ Th.conj().T
            nev = max(1, n // 2)
            lam, U, flag = heig_top(G, nev)
            wr, Vr = np.linalg.eigh(G)
            wr = wr[::-1][:nev]
            orth = np.linalg.norm(U.conj().T @ U - np.eye(nev))
            resid = np.linalg.norm(G @ U - U * lam) / np.linalg.norm(G)
            print(f"cplx={cplx} n={n}: lam err {np.max(np.abs(lam - wr)) / wr[0]:.1e} orth {orth:.1e} resid {resid:.1e} flag {flag}")
    # clustered / degenerate
    n = 64
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    for name, w in (("pair1e-6", np.r_[np.linspace(1, 2, n - 2), 2.5, 2.5 + 1e-6]), ("pair1e-11", np.r_[np.linspace(1, 2, n - 2), 2.5, 2.5 + 1e-11]),
                    ("decay", 10.0 ** -np.arange(n)), ("flatclusters", np.r_[np.ones(10), np.linspace(2, 3, n - 10)])):
        G = (Q * w) @ Q.T
        lam, U, flag = heig_top(G, 32)
        orth = np.linalg.norm(U.T @ U - np.eye(32)); resid = np.linalg.norm(G @ U - U * lam) / np.linalg.norm(G)
        print(f"{name}: orth {orth:.1e} resid {resid:.1e} flag {flag}")
