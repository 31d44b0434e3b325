// Krylov drivers around the effective-operator matvec (kernel family F8).
//
// They stand in for the third-party solvers the reference calls on the hot path (the sources are not part
// of the reference tree):
//   KrylovKit.eigsolve(..., 1, :SR)   src/solvers/dmrg.jl:245          -> lanczos_lowest
//   IterativeSolvers.lobpcg           src/solvers/als.jl:81, mals.jl:204 -> lanczos_lowest
//   `\` (LU) / KrylovKit.linsolve     src/solvers/als.jl:69, mals.jl:167, dmrg.jl:170,174 -> gmres_solve
//   KrylovKit.exponentiate            src/solvers/tdvp.jl:75,95,107,... -> lanczos_expm
// Converged results are algorithm independent; parity is therefore claimed on converged quantities.
// Krylov vectors live in HBM; orthogonalisation is two passes of classical Gram-Schmidt done by the fused
// multi_dot / multi_axpy kernels (one read of the basis per pass); the projected (<= krylovdim) problems are
// solved on the host.
#include "solvers.h"

namespace ttn {
namespace {

// cyclic Jacobi eigen-decomposition of a small real symmetric matrix (column-major n x n); ascending eigenvalues
void host_eigh(int n, std::vector<double> A, std::vector<double>& w, std::vector<double>& V) {
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) V[i + (size_t)i * n] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) (i == j ? diag : off) += A[i + (size_t)j * n] * A[i + (size_t)j * n];
    if (off <= 1e-32 * (diag + off) || off == 0.0) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[p + (size_t)q * n];
        if (apq == 0.0) continue;
        const double app = A[p + (size_t)p * n], aqq = A[q + (size_t)q * n];
        const double zeta = (aqq - app) / (2.0 * apq);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        for (int k = 0; k < n; ++k) {  // columns p,q
          const double akp = A[k + (size_t)p * n], akq = A[k + (size_t)q * n];
          A[k + (size_t)p * n] = c * akp - s * akq;
          A[k + (size_t)q * n] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {  // rows p,q
          const double apk = A[p + (size_t)k * n], aqk = A[q + (size_t)k * n];
          A[p + (size_t)k * n] = c * apk - s * aqk;
          A[q + (size_t)k * n] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[k + (size_t)p * n], vkq = V[k + (size_t)q * n];
          V[k + (size_t)p * n] = c * vkp - s * vkq;
          V[k + (size_t)q * n] = s * vkp + c * vkq;
        }
      }
  }
  std::vector<int> idx(n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return A[a + (size_t)a * n] < A[b + (size_t)b * n]; });
  w.resize(n);
  std::vector<double> Vs((size_t)n * n);
  for (int j = 0; j < n; ++j) {
    w[j] = A[idx[j] + (size_t)idx[j] * n];
    for (int i = 0; i < n; ++i) Vs[i + (size_t)j * n] = V[i + (size_t)idx[j] * n];
  }
  V.swap(Vs);
}

template <class T> struct HostScalar { typedef double type; };
template <> struct HostScalar<zc> { typedef std::complex<double> type; };
inline double hconj(double a) { return a; }
inline std::complex<double> hconj(std::complex<double> a) { return std::conj(a); }
inline double to_host(double a) { return a; }
inline std::complex<double> to_host(zc a) { return std::complex<double>(a.x, a.y); }
template <class T> T from_host(typename HostScalar<T>::type a);
template <> inline double from_host<double>(double a) { return a; }
template <> inline zc from_host<zc>(std::complex<double> a) { return make_cuDoubleComplex(a.real(), a.imag()); }

// one Lanczos/Arnoldi expansion step with two-pass classical Gram-Schmidt against Q[:,0..j]
// on return w is orthogonal to the basis, h[0..j] holds the projections, returns ||w||
template <class T>
double expand(LocalOp<T>& op, T* Q, int64_t n, int j, T* w, std::vector<T>& h) {
  op.apply(Q + (int64_t)j * n, w);
  h.assign(j + 1, t_zero<T>());
  std::vector<T> h1(j + 1);
  for (int pass = 0; pass < 2; ++pass) {
    multi_dot<T>(n, j + 1, Q, n, w, h1.data());
    multi_axpy<T>(n, j + 1, Q, n, h1.data(), w, -1.0);
    for (int i = 0; i <= j; ++i) h[i] = t_add(h[i], h1[i]);
  }
  return nrm2<T>(n, w);
}

template <class T>
void combine(T* Q, int64_t n, int k, const std::vector<T>& coef, T* x) {
  fill<T>(x, n, t_zero<T>());
  multi_axpy<T>(n, k, Q, n, coef.data(), x, 1.0);
}

}  // namespace

template <class T>
double lanczos_lowest(LocalOp<T>& op, T* x, int krylovdim, int maxiter, double tol, KrylovInfo* info) {
  const int64_t n = op.size();
  int m = (int)std::min<int64_t>(std::max(krylovdim, 2), n);
  DevBuf Q(sizeof(T) * (size_t)n * m), w(sizeof(T) * (size_t)n);
  double nrm = nrm2<T>(n, x);
  ttn_assert(nrm > 0.0 && std::isfinite(nrm), 5, "eigsolve: zero or non-finite start vector");
  scal<T>(n, t_from<T>(1.0 / nrm, 0.0), x);
  double theta = 0.0, resid = 0.0;
  bool conv = false;
  int matvecs = 0, restarts = 0;
  std::vector<T> h;
  for (int it = 0; it < std::max(1, maxiter) && !conv; ++it, ++restarts) {
    TTN_CUDA(cudaMemcpyAsync(Q.p, x, sizeof(T) * n, cudaMemcpyDeviceToDevice, ctx().stream));
    std::vector<double> alpha, beta;
    std::vector<double> sv;
    int k = 0;
    for (int j = 0; j < m; ++j) {
      const double b = expand<T>(op, Q.as<T>(), n, j, w.as<T>(), h);
      ++matvecs;
      alpha.push_back(t_real(h[j]));
      beta.push_back(b);
      k = j + 1;
      std::vector<double> Tm((size_t)k * k, 0.0), ev, evec;
      for (int i = 0; i < k; ++i) {
        Tm[i + (size_t)i * k] = alpha[i];
        if (i + 1 < k) { Tm[i + 1 + (size_t)i * k] = beta[i]; Tm[i + (size_t)(i + 1) * k] = beta[i]; }
      }
      host_eigh(k, Tm, ev, evec);
      theta = ev[0];
      sv.assign(evec.begin(), evec.begin() + k);
      resid = std::fabs(b * sv[k - 1]);
      const double scale = std::max(1.0, std::fabs(theta));
      if (resid <= tol || b <= 1e-14 * scale) { conv = true; break; }
      if (j + 1 < m) {
        TTN_CUDA(cudaMemcpyAsync(Q.as<T>() + (int64_t)(j + 1) * n, w.p, sizeof(T) * n, cudaMemcpyDeviceToDevice, ctx().stream));
        scal<T>(n, t_from<T>(1.0 / b, 0.0), Q.as<T>() + (int64_t)(j + 1) * n);
      }
    }
    std::vector<T> coef(k);
    for (int i = 0; i < k; ++i) coef[i] = t_from<T>(sv[i], 0.0);
    combine<T>(Q.as<T>(), n, k, coef, x);
    const double xn = nrm2<T>(n, x);
    if (xn > 0) scal<T>(n, t_from<T>(1.0 / xn, 0.0), x);
    if (m >= n) conv = true;  // full space: the Ritz pair is exact
  }
  if (info) { info->matvecs = matvecs; info->restarts = restarts; info->resid = resid; info->converged = conv; }
  return theta;
}

namespace {
// back substitution R x = y for the upper triangle of a column-major N x N matrix; one CTA, x staged in shared memory
template <class T>
__global__ void __launch_bounds__(256) trsv_upper_kernel(const T* __restrict__ R, int N, int64_t ld, T* __restrict__ y) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* x = reinterpret_cast<T*>(smem_raw);
  __shared__ T s_xj;
  const int tid = threadIdx.x;
  for (int i = tid; i < N; i += 256) x[i] = y[i];
  __syncthreads();
  for (int j = N - 1; j >= 0; --j) {
    if (tid == 0) {
      const T d = R[j + (int64_t)j * ld];
      const double den = t_abs2(d);
      const T v = x[j];
      // x_j = v / d (complex-safe); a zero pivot propagates inf/nan exactly like LAPACK's `\` on a singular K
      const T q = t_scale(t_mul(v, t_conj(d)), 1.0 / den);
      x[j] = q;
      s_xj = q;
    }
    __syncthreads();
    const T xj = s_xj;
    const T* col = R + (int64_t)j * ld;
    for (int i = tid; i < j; i += 256) x[i] = t_sub(x[i], t_mul(col[i], xj));
    __syncthreads();
  }
  for (int i = tid; i < N; i += 256) y[i] = x[i];
}
}  // namespace

template <class T>
void dense_solve(LocalOp<T>& op, const T* rhs, T* x) {
  const int64_t n64 = op.size();
  ttn_assert(n64 <= 8192, 7, "dense_solve: window too large");
  const int N = (int)n64;
  DevBuf Id(sizeof(T) * (size_t)N * N), K(sizeof(T) * (size_t)N * N), tau(sizeof(T) * (size_t)N);
  set_identity<T>(Id.as<T>(), N, N, N);
  op.apply_batch(Id.as<T>(), K.as<T>(), N);          // column j of K = K e_j
  Id.release();
  qr_factor<T>(K.as<T>(), N, N, N, tau.as<T>());
  TTN_CUDA(cudaMemcpyAsync(x, rhs, sizeof(T) * (size_t)N, cudaMemcpyDeviceToDevice, ctx().stream));
  qr_apply<T>(K.as<T>(), N, N, N, tau.as<T>(), x, 1, N, true);
  static int attr_dev = -1;   // function attributes are per device (ttn_init may re-bind)
  if (attr_dev != ctx().device) {
    TTN_CUDA(cudaFuncSetAttribute(trsv_upper_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * (int)sizeof(T)));
    attr_dev = ctx().device;
  }
  trsv_upper_kernel<T><<<1, 256, sizeof(T) * (size_t)N, ctx().stream>>>(K.as<T>(), N, N, x);
  TTN_CHECK_LAUNCH();
  ctx().launches++;
  TTN_CUDA(cudaStreamSynchronize(ctx().stream));     // K / tau are released when this scope ends
}

// Lowest eigenpair of the Hermitian-definite pencil (K, M):  K v = lambda M v.  Dense, like the reference's `K_eiggenmin`
// (src/solvers/als.jl:89-102), which assembles both local matrices in full before `eigen(K, S)` / `lobpcg`:
//   M = U Sigma U^H (the SVD of a Hermitian positive definite matrix is its eigen-decomposition; Jacobi SVD of svd.cu),
//   W = U Sigma^{-1/2}  so that  W^H M W = I,   C = W^H K W (two DMMA GEMMs),  lowest pair (theta, y) of C by Lanczos,  v = W y.
template <class T>
double gen_eig_lowest(LocalOp<T>& opK, LocalOp<T>& opM, T* v, int krylovdim, int maxiter, double tol) {
  const int64_t n64 = opK.size();
  ttn_assert(n64 == opM.size(), 1, "Incompatible dimensions");
  ttn_assert(n64 <= 4096, 7, "als_gen_eigsolv: local problem too large for the dense generalized eigen-solve");
  const int N = (int)n64;
  const size_t NN = (size_t)N * N;
  DevBuf K(sizeof(T) * NN), M(sizeof(T) * NN);
  {
    DevBuf Id(sizeof(T) * NN);
    set_identity<T>(Id.as<T>(), N, N, N);
    opK.apply_batch(Id.as<T>(), K.as<T>(), N);          // column j = K e_j
    opM.apply_batch(Id.as<T>(), M.as<T>(), N);
  }
  SvdLeft sv;
  svd_left<T>(M.as<T>(), N, N, 1, N, false, sv, 1, 0);
  std::vector<double> sc(N);
  std::vector<int> perm(sv.perm.begin(), sv.perm.begin() + N);
  for (int j = 0; j < N; ++j) {
    const double sg = sv.sigma[j];
    ttn_assert(sg > sv.sigma[0] * 1e-14 && sg > 0.0, 5, "als_gen_eigsolv: S is numerically singular on the local subspace");
    sc[j] = 1.0 / (sg * std::sqrt(sg));                   // X_j = u_j sigma_j  ->  u_j sigma_j^{-1/2}
  }
  DevBuf dperm(sizeof(int) * N), dsc(sizeof(double) * N), W(sizeof(T) * NN), T1(sizeof(T) * NN), Cm(sizeof(T) * NN);
  TTN_CUDA(cudaMemcpyAsync(dperm.p, perm.data(), sizeof(int) * N, cudaMemcpyHostToDevice, ctx().stream));
  TTN_CUDA(cudaMemcpyAsync(dsc.p, sc.data(), sizeof(double) * N, cudaMemcpyHostToDevice, ctx().stream));
  gather_cols<T>(sv.X.as<T>(), N, N, dperm.as<int>(), dsc.as<double>(), N, W.as<T>(), 1, N);
  auto mm = [&](const T* A, bool hA, const T* B, int ncol, T* Cc) {   // Cc (N x ncol) = op(A) B,  A: N x N
    GemmArgs g;
    g.M = N; g.N = ncol; g.K = N;
    g.A = A; g.sAm = hA ? N : 1; g.sAk = hA ? 1 : N; g.conjA = hA;
    g.B = B; g.sBk = 1; g.sBn = N;
    g.C = Cc; g.sCm = 1; g.sCn = N;
    gemm<T>(g);
  };
  mm(K.as<T>(), false, W.as<T>(), N, T1.as<T>());
  mm(W.as<T>(), true, T1.as<T>(), N, Cm.as<T>());
  DevBuf t(sizeof(T) * N), y(sizeof(T) * N);
  mm(M.as<T>(), false, v, 1, t.as<T>());                 // y0 = W^-1 v = W^H M v
  mm(W.as<T>(), true, t.as<T>(), 1, y.as<T>());
  if (!(nrm2<T>(N, y.as<T>()) > 0.0)) fill<T>(y.as<T>(), N, t_one<T>());
  LocalOp<T> opC;
  opC.chi_l = N; opC.nn = 1; opC.chi_r = 1;
  const T* Cp = Cm.as<T>();
  opC.ext_apply = [&mm, Cp](const T* a, T* b) { mm(Cp, false, a, 1, b); };
  const double theta = lanczos_lowest<T>(opC, y.as<T>(), krylovdim, maxiter, tol, nullptr);
  mm(W.as<T>(), false, y.as<T>(), 1, v);
  TTN_CUDA(cudaStreamSynchronize(ctx().stream));
  return theta;
}

template <class T>
void local_linsolve(LocalOp<T>& op, const T* rhs, T* x, int krylovdim, int maxiter, double tol, const ttn_solver_params& p) {
  const int64_t dense_max = std::max<int64_t>(p.itslv_thresh, 2048);
  if (!p.it_solver && op.size() <= dense_max) dense_solve<T>(op, rhs, x);
  else gmres_solve<T>(op, rhs, x, krylovdim, maxiter, tol, nullptr);
}

template <class T>
void gmres_solve(LocalOp<T>& op, const T* rhs, T* x, int krylovdim, int maxiter, double tol, KrylovInfo* info) {
  typedef typename HostScalar<T>::type S;
  const int64_t n = op.size();
  const int m = (int)std::min<int64_t>(std::max(krylovdim, 2), n);
  DevBuf Q(sizeof(T) * (size_t)n * (m + 1)), w(sizeof(T) * (size_t)n);
  const double bnorm = nrm2<T>(n, rhs);
  int matvecs = 0, restarts = 0;
  double resid = 0.0;
  bool conv = false;
  if (bnorm == 0.0) {
    fill<T>(x, n, t_zero<T>());
    if (info) { info->converged = true; }
    return;
  }
  std::vector<T> h;
  for (int it = 0; it < std::max(1, maxiter) && !conv; ++it, ++restarts) {
    // r = rhs - K x
    op.apply(x, w.as<T>());
    ++matvecs;
    T* q0 = Q.as<T>();
    TTN_CUDA(cudaMemcpyAsync(q0, rhs, sizeof(T) * n, cudaMemcpyDeviceToDevice, ctx().stream));
    axpy<T>(n, t_from<T>(-1.0, 0.0), w.as<T>(), q0);
    const double beta = nrm2<T>(n, q0);
    resid = beta / bnorm;
    if (beta <= tol * bnorm) { conv = true; break; }
    scal<T>(n, t_from<T>(1.0 / beta, 0.0), q0);
    std::vector<S> H((size_t)(m + 1) * m, S(0)), g(m + 1, S(0)), cs(m, S(0)), sn(m, S(0));
    g[0] = S(beta);
    int k = 0;
    for (int j = 0; j < m; ++j) {
      const double hn = expand<T>(op, Q.as<T>(), n, j, w.as<T>(), h);
      ++matvecs;
      for (int i = 0; i <= j; ++i) H[i + (size_t)j * (m + 1)] = to_host(h[i]);
      H[j + 1 + (size_t)j * (m + 1)] = S(hn);
      // apply the previous rotations to the new column
      for (int i = 0; i < j; ++i) {
        const S a = H[i + (size_t)j * (m + 1)], b = H[i + 1 + (size_t)j * (m + 1)];
        H[i + (size_t)j * (m + 1)] = hconj(cs[i]) * a + hconj(sn[i]) * b;
        H[i + 1 + (size_t)j * (m + 1)] = -sn[i] * a + cs[i] * b;
      }
      {
        const S a = H[j + (size_t)j * (m + 1)], b = H[j + 1 + (size_t)j * (m + 1)];
        const double den = std::sqrt(std::norm(a) + std::norm(b));
        if (den == 0.0) { cs[j] = S(1); sn[j] = S(0); }
        else { cs[j] = a / den; sn[j] = b / den; }
        H[j + (size_t)j * (m + 1)] = hconj(cs[j]) * a + hconj(sn[j]) * b;
        H[j + 1 + (size_t)j * (m + 1)] = S(0);
        const S g0 = g[j];
        g[j] = hconj(cs[j]) * g0;
        g[j + 1] = -sn[j] * g0;
      }
      k = j + 1;
      resid = std::abs(g[j + 1]) / bnorm;
      const bool breakdown = hn <= 1e-14 * beta;
      if (resid <= tol || breakdown) break;
      {
        T* qn = Q.as<T>() + (int64_t)(j + 1) * n;
        TTN_CUDA(cudaMemcpyAsync(qn, w.p, sizeof(T) * n, cudaMemcpyDeviceToDevice, ctx().stream));
        scal<T>(n, t_from<T>(1.0 / hn, 0.0), qn);
      }
    }
    // back substitution  H(0:k,0:k) y = g(0:k)
    std::vector<S> y(k);
    for (int i = k - 1; i >= 0; --i) {
      S s = g[i];
      for (int l = i + 1; l < k; ++l) s -= H[i + (size_t)l * (m + 1)] * y[l];
      const S d = H[i + (size_t)i * (m + 1)];
      y[i] = (std::abs(d) > 0.0) ? s / d : S(0);
    }
    std::vector<T> coef(k);
    for (int i = 0; i < k; ++i) coef[i] = from_host<T>(y[i]);
    multi_axpy<T>(n, k, Q.as<T>(), n, coef.data(), x, 1.0);
    if (resid <= tol) conv = true;
    if (m >= n && k == m) {
      // full-space Arnoldi is exact up to rounding; one more outer pass refines, then stop
      if (it >= 1) conv = true;
    }
  }
  if (info) { info->matvecs = matvecs; info->restarts = restarts; info->resid = resid; info->converged = conv; }
}

template <class T>
void lanczos_expm(LocalOp<T>& op, T* x, double tre, double tim, int krylovdim, int maxiter, double tol, KrylovInfo* info) {
  typedef std::complex<double> C;
  const int64_t n = op.size();
  ttn_assert(is_cplx<T>::value || tim == 0.0, 2, "expm: complex time step on a real state");
  const int m = (int)std::min<int64_t>(std::max(krylovdim, 2), n);
  DevBuf Q(sizeof(T) * (size_t)n * m), w(sizeof(T) * (size_t)n);
  const C t(tre, tim);
  double left = 1.0;
  int matvecs = 0, restarts = 0;
  std::vector<T> h;
  double err = 0.0;
  for (int it = 0; it < std::max(1, maxiter) && left > 0.0; ++it, ++restarts) {
    const double beta0 = nrm2<T>(n, x);
    if (beta0 == 0.0) break;
    TTN_CUDA(cudaMemcpyAsync(Q.p, x, sizeof(T) * n, cudaMemcpyDeviceToDevice, ctx().stream));
    scal<T>(n, t_from<T>(1.0 / beta0, 0.0), Q.as<T>());
    std::vector<double> alpha, beta;
    int k = 0;
    bool exact = false;
    std::vector<double> ev, evec;
    std::vector<C> c;
    double dt = left;
    for (int j = 0; j < m; ++j) {
      const double b = expand<T>(op, Q.as<T>(), n, j, w.as<T>(), h);
      ++matvecs;
      alpha.push_back(t_real(h[j]));
      beta.push_back(b);
      k = j + 1;
      std::vector<double> Tm((size_t)k * k, 0.0);
      for (int i = 0; i < k; ++i) {
        Tm[i + (size_t)i * k] = alpha[i];
        if (i + 1 < k) { Tm[i + 1 + (size_t)i * k] = beta[i]; Tm[i + (size_t)(i + 1) * k] = beta[i]; }
      }
      host_eigh(k, Tm, ev, evec);
      auto coeffs = [&](double frac) {
        c.assign(k, C(0));
        for (int l = 0; l < k; ++l) {
          const C e = std::exp(t * frac * ev[l]) * evec[0 + (size_t)l * k];
          for (int i = 0; i < k; ++i) c[i] += e * evec[i + (size_t)l * k];
        }
      };
      coeffs(dt);
      double cn = 0.0;
      for (int i = 0; i < k; ++i) cn += std::norm(c[i]);
      err = b * std::abs(c[k - 1]);
      const double anorm = std::max(std::fabs(ev[0]), std::fabs(ev[k - 1]));
      if (b <= 1e-14 * std::max(1.0, anorm) || k >= n) { exact = true; break; }
      if (err <= tol * std::sqrt(cn)) break;
      if (j + 1 < m) {
        T* qn = Q.as<T>() + (int64_t)(j + 1) * n;
        TTN_CUDA(cudaMemcpyAsync(qn, w.p, sizeof(T) * n, cudaMemcpyDeviceToDevice, ctx().stream));
        scal<T>(n, t_from<T>(1.0 / b, 0.0), qn);
      } else {
        // subspace exhausted: shrink the time step until the error estimate passes
        for (int tries = 0; tries < 40 && err > tol; ++tries) {
          dt *= 0.5;
          c.assign(k, C(0));
          for (int l = 0; l < k; ++l) {
            const C e = std::exp(t * dt * ev[l]) * evec[0 + (size_t)l * k];
            for (int i = 0; i < k; ++i) c[i] += e * evec[i + (size_t)l * k];
          }
          err = b * std::abs(c[k - 1]);
        }
      }
    }
    (void)exact;
    std::vector<T> coef(k);
    for (int i = 0; i < k; ++i) coef[i] = t_from<T>(c[i].real() * beta0, c[i].imag() * beta0);
    combine<T>(Q.as<T>(), n, k, coef, x);
    left -= dt;
    if (left < 1e-15) left = 0.0;
  }
  ttn_assert(left == 0.0, 5, "exponentiate: time step not completed within maxiter restarts");
  if (info) { info->matvecs = matvecs; info->restarts = restarts; info->resid = err; info->converged = true; }
}

#define INST(T)                                                                                         \
  template double lanczos_lowest<T>(LocalOp<T>&, T*, int, int, double, KrylovInfo*);                    \
  template void gmres_solve<T>(LocalOp<T>&, const T*, T*, int, int, double, KrylovInfo*);               \
  template void dense_solve<T>(LocalOp<T>&, const T*, T*);                                                \
  template double gen_eig_lowest<T>(LocalOp<T>&, LocalOp<T>&, T*, int, int, double);                      \
  template void local_linsolve<T>(LocalOp<T>&, const T*, T*, int, int, double, const ttn_solver_params&); \
  template void lanczos_expm<T>(LocalOp<T>&, T*, double, double, int, int, double, KrylovInfo*);
INST(double)
INST(zc)
#undef INST

}  // namespace ttn
