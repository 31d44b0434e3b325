// Top eigenpairs of small Hermitian positive semi-definite matrices (n <= 128 ComplexF64, n <= 176 Float64), batched:
// the SVD engine of the Gram path of `tt_compress!` (src/tt_tools.jl:743-789, `_svdtrunc` of
// src/tt_cross_interpolation.jl:149-166 with truncerr = 0): G = Theta Theta^H = U diag(sigma^2) U^H.
//
// One-sided Jacobi on a 128-column bond matrix is a chain of ~1 300 dependent rotation steps (csrc/jacobi_cluster.cu:
// 620 us on an 8-SM cluster).  A rank cap only needs the leading invariant subspace to the accuracy the problem itself is
// conditioned to, which the Gram matrix delivers (error eps*sigma_1/sigma_r on the reconstruction, accepted while
// sigma_r >= 1e-5 sigma_1; everything else falls back to the Jacobi path).  So the chain is replaced by
//   K1  Householder tridiagonalisation, one CTA per matrix, packed lower triangle in shared memory   (n-1 dependent steps)
//   K2  eigenvalues by multisection on the division-free scaled Sturm sequence (M threads per eigenvalue), eigenvectors of
//       the tridiagonal matrix by the twisted factorisation (one thread per vector), Gram-Schmidt inside clusters
//   K3  back-transformation U = H_0 ... H_{n-2} Z, columns spread over lane groups, reflectors staged in shared memory.
// LAPACK conventions (zhetd2 'L', zlarfg); tools/heig_proto.py is the NumPy twin the kernels were checked against.
#include "ttn_internal.h"
#include "tma.h"

namespace ttn {
namespace {

template <class T> struct HT;
template <> struct HT<double> { static constexpr int R = 16; };   // rows per shared-memory access phase (half warp x 8 B)
template <> struct HT<zc> { static constexpr int R = 8; };        // quarter warp x 16 B

__device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }

__device__ __forceinline__ double shfl_xor_t(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
__device__ __forceinline__ zc shfl_xor_t(zc v, int o) {
  return make_cuDoubleComplex(__shfl_xor_sync(0xffffffffu, v.x, o), __shfl_xor_sync(0xffffffffu, v.y, o));
}
__device__ __forceinline__ double t_realpart(double a) { return a; }
__device__ __forceinline__ zc t_realpart(zc a) { return make_cuDoubleComplex(a.x, 0.0); }

// ---------------------------------------------------------------------------------------------------------------------
// K1: A = Q T Q^H.  Thread (row i, residue lq): rows of a warp are R consecutive rows (aligned), so that both the row
// access L[tri(i)+j] (triangular numbers are a permutation mod R over R aligned consecutive i) and the column access
// L[tri(j)+i] (contiguous in i) are bank-conflict free per access phase.
// One step k (three block barriers):
//   A  y2_i = sum_{j>=k+2} H(i,j) x_j (x = column k, unscaled), y1_i = H(i,k+1); block sums of |x_i|^2, conj(y1_i) x_i,
//      conj(y2_i) x_i over the rows below k+1 — everything that does not need the reflector scalars;
//   B  scalars once (rsqrt only): beta, tau, s = 1/(alpha-beta); v = [1; s x], p = tau (y1 + s y2),
//      p^H v = conj(tau) [conj(y1_0) + conj(s) conj(y2_0) + s a + |s|^2 b],  w = p - (tau/2)(p^H v) v;
//   C  A22 -= v w^H + w v^H on the lower triangle; the thread that owns column k+1 hands the next x over.
//   G is read as sum_{s<nsplit} Gp[s*sG + ...] (split-K partial Gram matrices are summed on the fly), upper triangle.
//   V_out: reflector k occupies rows k+1..n-1 (unit first element stored) at offset k(n-1) - k(k-1)/2.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int refl_off(int k, int n) { return k * (n - 1) - ((k * (k - 1)) >> 1); }

template <class T>
__global__ void __launch_bounds__(512) heig_tridiag_kernel(const T* __restrict__ Gp, int n, int64_t ldg, int64_t bG, int nsplit,
                                                            int64_t sG, double* __restrict__ d_out, double* __restrict__ e_out,
                                                            T* __restrict__ tau_out, T* __restrict__ V_out, int64_t bV) {
  constexpr int R = HT<T>::R, g = 32 / R;
  constexpr int NC = is_cplx<T>::value ? 5 : 3;           // block-reduced quantities per step
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int npad = ((n + R - 1) / R) * R;
  T* L = reinterpret_cast<T*>(smem_raw);      // packed lower triangle, row i at tri(i)
  T* xs = L + tri(n) + (tri(n) & 1);          // column k below the diagonal (unscaled)
  T* vs = xs + npad;
  T* ws = vs + npad;
  double* red = reinterpret_cast<double*>(ws + npad);   // [NC][32] warp partials
  T* y0s = reinterpret_cast<T*>(red + NC * 32);          // y1, y2 of row k+1

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwarps = blockDim.x >> 5;
  const int ir = lane % R, lq = lane / R;
  const int i = warp * R + ir;
  const int ti = tri(i);
  T* __restrict__ Lr = L + ti;
  const int64_t mb = blockIdx.x;
  const T* Gb = Gp + mb * bG;
  d_out += mb * n; e_out += mb * n; tau_out += mb * n; V_out += mb * bV;

  // lower (ii, j) = conj(upper (j, ii)): column ii of G is contiguous in j
  for (int ii = warp; ii < n; ii += nwarps)
    for (int j = lane; j <= ii; j += 32) {
      T a = t_zero<T>();
      for (int s = 0; s < nsplit; ++s) a = t_add(a, Gb[s * sG + j + (int64_t)ii * ldg]);
      L[tri(ii) + j] = (j == ii) ? t_realpart(a) : t_conj(a);
    }
  for (int idx = tid; idx < NC * 32; idx += blockDim.x) red[idx] = 0.0;
  __syncthreads();
  if (lq == 0 && i > 0 && i < n) xs[i] = Lr[0];
  __syncthreads();

  for (int k = 0; k < n - 1; ++k) {
    const int k1 = k + 1;
    const bool rowact = (i > k) && (i < n);
    // ---- A ----
    T y2 = t_zero<T>();
    if (rowact) {
      T a0 = t_zero<T>(), a1 = t_zero<T>(), a2 = t_zero<T>(), a3 = t_zero<T>();
      int j = k + 2 + ((lq - (k + 2)) & (g - 1));
      const int je = i < n ? i : n;
      for (; j + 3 * g < je; j += 4 * g) {
        const T l0 = Lr[j], l1 = Lr[j + g], l2 = Lr[j + 2 * g], l3 = Lr[j + 3 * g];
        const T x0 = xs[j], x1 = xs[j + g], x2 = xs[j + 2 * g], x3 = xs[j + 3 * g];
        t_fma(a0, l0, x0); t_fma(a1, l1, x1); t_fma(a2, l2, x2); t_fma(a3, l3, x3);
      }
      for (; j < je; j += g) t_fma(a0, Lr[j], xs[j]);
      if (j == i) { t_fma(a1, t_realpart(Lr[i]), xs[i]); j += g; }
      for (; j + 3 * g < n; j += 4 * g) {
        const T l0 = L[tri(j) + i], l1 = L[tri(j + g) + i], l2 = L[tri(j + 2 * g) + i], l3 = L[tri(j + 3 * g) + i];
        const T x0 = xs[j], x1 = xs[j + g], x2 = xs[j + 2 * g], x3 = xs[j + 3 * g];
        t_fma(a0, t_conj(l0), x0); t_fma(a1, t_conj(l1), x1); t_fma(a2, t_conj(l2), x2); t_fma(a3, t_conj(l3), x3);
      }
      for (; j < n; j += g) t_fma(a0, t_conj(L[tri(j) + i]), xs[j]);
      y2 = t_add(t_add(a0, a1), t_add(a2, a3));
    }
#pragma unroll
    for (int o = R; o < 32; o <<= 1) y2 = t_add(y2, shfl_xor_t(y2, o));
    T y1 = t_zero<T>(), xi = t_zero<T>();
    double q[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) q[c] = 0.0;
    if (rowact && lq == 0) {
      if (i == k1) {
        y1 = t_realpart(Lr[i]);
        y0s[0] = y1; y0s[1] = y2;
      } else {
        y1 = Lr[k1];
        xi = xs[i];
        q[0] = t_abs2(xi);
        const T c1 = t_mul(t_conj(y1), xi), c2 = t_mul(t_conj(y2), xi);
        q[1] = t_real(c1); q[2] = t_real(c2);
        if (is_cplx<T>::value) { q[NC - 2] = t_imag(c1); q[NC - 1] = t_imag(c2); }
      }
    }
#pragma unroll
    for (int o = R / 2; o > 0; o >>= 1) {
#pragma unroll
      for (int c = 0; c < NC; ++c) q[c] += __shfl_xor_sync(0xffffffffu, q[c], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < NC; ++c) red[c * 32 + warp] = q[c];
    }
    __syncthreads();   // 1
    // ---- B ----
#pragma unroll
    for (int c = 0; c < NC; ++c) q[c] = red[c * 32 + lane];     // entries >= nwarps stay zero
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int c = 0; c < NC; ++c) q[c] += __shfl_xor_sync(0xffffffffu, q[c], o);
    }
    const double xn2 = q[0];
    const T alpha = xs[k1];
    const double ar = t_real(alpha), ai = t_imag(alpha);
    const bool have = !(xn2 == 0.0 && ai == 0.0);
    double beta = ar;
    T tau = t_zero<T>();
    if (have) {
      const double s2 = ar * ar + ai * ai + xn2;
      const double inv = rsqrt(s2), nrm = s2 * inv;
      const double sg = ar >= 0.0 ? 1.0 : -1.0;
      beta = -sg * nrm;
      tau = t_from<T>(1.0 + fabs(ar) * inv, sg * ai * inv);
      const double dr = sg * (fabs(ar) + nrm), di = ai;          // alpha - beta
      const double r1 = rsqrt(dr * dr + di * di), idn = r1 * r1;
      const T sc = t_from<T>(dr * idn, -di * idn);               // 1 / (alpha - beta)
      if (rowact && lq == 0) {
        const T a = t_from<T>(q[1], is_cplx<T>::value ? q[NC - 2] : 0.0), b = t_from<T>(q[2], is_cplx<T>::value ? q[NC - 1] : 0.0);
        const T y10 = y0s[0], y20 = y0s[1];
        // p^H v = conj(tau) [conj(y1_0) + conj(s) conj(y2_0) + s a + |s|^2 b]
        T pv = t_add(t_conj(y10), t_mul(t_conj(sc), t_conj(y20)));
        pv = t_add(pv, t_add(t_mul(sc, a), t_scale(b, t_abs2(sc))));
        pv = t_mul(t_conj(tau), pv);
        const T a2 = t_scale(t_mul(tau, pv), -0.5);
        const T vi = (i == k1) ? t_one<T>() : t_mul(sc, xi);
        const T p = t_mul(tau, t_add(y1, t_mul(sc, y2)));
        vs[i] = vi;
        ws[i] = t_add(p, t_mul(a2, vi));
        V_out[refl_off(k, n) + (i - k1)] = vi;
      }
    } else if (rowact && lq == 0) {
      V_out[refl_off(k, n) + (i - k1)] = (i == k1) ? t_one<T>() : t_zero<T>();
    }
    if (tid == 0) { d_out[k] = t_real(L[tri(k) + k]); e_out[k] = beta; tau_out[k] = tau; }
    __syncthreads();   // 2
    // ---- C ----
    if (rowact) {
      int j = k1 + ((lq - k1) & (g - 1));
      if (have) {
        const T vi = vs[i], wi = ws[i];
        if (j == k1 && i > k1) {       // next pivot column: hand x over
          const T a = t_sub(Lr[j], t_add(t_mul(vi, t_conj(ws[j])), t_mul(wi, t_conj(vs[j]))));
          Lr[j] = a;
          xs[i] = a;
          j += g;
        }
        for (; j + 3 * g < i; j += 4 * g) {
          const T l0 = Lr[j], l1 = Lr[j + g], l2 = Lr[j + 2 * g], l3 = Lr[j + 3 * g];
          const T w0 = ws[j], w1 = ws[j + g], w2 = ws[j + 2 * g], w3 = ws[j + 3 * g];
          const T v0 = vs[j], v1 = vs[j + g], v2 = vs[j + 2 * g], v3 = vs[j + 3 * g];
          Lr[j] = t_sub(l0, t_add(t_mul(vi, t_conj(w0)), t_mul(wi, t_conj(v0))));
          Lr[j + g] = t_sub(l1, t_add(t_mul(vi, t_conj(w1)), t_mul(wi, t_conj(v1))));
          Lr[j + 2 * g] = t_sub(l2, t_add(t_mul(vi, t_conj(w2)), t_mul(wi, t_conj(v2))));
          Lr[j + 3 * g] = t_sub(l3, t_add(t_mul(vi, t_conj(w3)), t_mul(wi, t_conj(v3))));
        }
        for (; j <= i; j += g) {
          T a = t_sub(Lr[j], t_add(t_mul(vi, t_conj(ws[j])), t_mul(wi, t_conj(vs[j]))));
          if (j == i) a = t_realpart(a);
          Lr[j] = a;
        }
      } else if (j == k1 && i > k1) {
        xs[i] = Lr[j];
      }
    }
    __syncthreads();   // 3
  }
  if (tid == 0) { d_out[n - 1] = t_real(L[tri(n - 1) + n - 1]); e_out[n - 1] = 0.0; tau_out[n - 1] = t_zero<T>(); }
}

// ---------------------------------------------------------------------------------------------------------------------
// K1 (Float64, latency path): the same tridiagonalisation with the whole symmetric matrix in REGISTERS.
// Warp w owns the rows w, w + NW, w + 2 NW, ... (8 of them: cyclic, so the shrinking trailing block stays balanced over the
// warps); lane l owns the columns l + 32 c, c < NC.  A step touches every trailing element once in the matrix-vector product
// and once in the rank-2 update, both from registers; per step a lane loads only ITS NC entries of x and of (v, w) from shared
// memory (conflict-free, one wavefront per 16 lanes) plus one (v_i, w_i) pair per row, and the 8 row sums of the product are
// reduced over the 32 lanes by a transposed butterfly (9 exchanges instead of 40).  The shared-memory pipe, not the FP64
// pipe, bounds this kernel (measured on the first version: 4 lanes per row and broadcast loads: 2.4 us per step), hence the layout.
// Rows k+1 and k+2 are published to shared memory after every update: by symmetry they are the next pivot column x and the
// next y1 = A(:, k+2).  The reflector scalars are computed by warp 0 alone and broadcast through shared memory.
// ---------------------------------------------------------------------------------------------------------------------
template <int NW, int NC>
__global__ void __launch_bounds__(NW * 32) heig_tridiag_reg_kernel(const double* __restrict__ Gp, int n, int64_t ldg, int64_t bG,
                                                                   int nsplit, int64_t sG, double* __restrict__ d_out,
                                                                   double* __restrict__ e_out, double* __restrict__ tau_out,
                                                                   double* __restrict__ V_out, int64_t bV) {
  constexpr int NP = 32 * NC;     // padded order (columns); rows: 8 * NW >= n
  constexpr int NRW = 8;
  __shared__ __align__(16) double xs[NP];
  __shared__ __align__(16) double ys[NP];
  __shared__ __align__(16) double2 vw[NP > 8 * NW ? NP : 8 * NW];
  __shared__ double red[3 * 32];
  __shared__ double y0s[2];
  __shared__ double scal[4];      // sc, tau, a2-prefactor pieces, have
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int64_t mb = blockIdx.x;
  const double* Gb = Gp + mb * bG;
  d_out += mb * n; e_out += mb * n; tau_out += mb * n; V_out += mb * bV;

  // the upper triangle of G is read (G is symmetric up to the rounding of the split-K sums); split loop outermost so that the
  // 8 * NC loads of a partial matrix are independent and in flight together (the profile of the first version showed 10 % of
  // the kernel in this prologue, latency bound)
  double a[NRW][NC];
#pragma unroll
  for (int rr = 0; rr < NRW; ++rr)
#pragma unroll
    for (int c = 0; c < NC; ++c) a[rr][c] = 0.0;
  for (int s = 0; s < nsplit; ++s) {
    const double* Gs = Gb + s * sG;
#pragma unroll
    for (int h = 0; h < 2; ++h) {          // two batches of 4 rows: 4 NC loads in flight without spilling the 8 NC accumulators
      double v[NRW / 2][NC];
#pragma unroll
      for (int r2 = 0; r2 < NRW / 2; ++r2) {
        const int i = w + NW * (r2 + h * (NRW / 2));
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int j = lane + 32 * c;
          const int lo = j <= i ? j : i, hi = j <= i ? i : j;
          v[r2][c] = (i < n && j < n) ? __ldg(Gs + lo + (int64_t)hi * ldg) : 0.0;
        }
      }
#pragma unroll
      for (int r2 = 0; r2 < NRW / 2; ++r2)
#pragma unroll
        for (int c = 0; c < NC; ++c) a[r2 + h * (NRW / 2)][c] += v[r2][c];
    }
  }
  for (int idx = tid; idx < NP; idx += blockDim.x) { xs[idx] = 0.0; ys[idx] = 0.0; }
  for (int idx = tid; idx < (NP > 8 * NW ? NP : 8 * NW); idx += blockDim.x) vw[idx] = make_double2(0.0, 0.0);
  for (int idx = tid; idx < 96; idx += blockDim.x) red[idx] = 0.0;
  __syncthreads();
  // rows 0 and 1 -> xs, ys  (row i lives in warp i % NW, register row i / NW)
#pragma unroll
  for (int rr = 0; rr < NRW; ++rr) {
    const int i = w + NW * rr;
    if (i == 0 || i == 1) {
      double* dst = i == 0 ? xs : ys;
#pragma unroll
      for (int c = 0; c < NC; ++c) dst[lane + 32 * c] = a[rr][c];
    }
  }
  __syncthreads();

  const int myrr = (lane >> 2) & 7;          // the row whose sum this lane ends up with
  const int myrow = w + NW * myrr;
  const bool owner = (lane & 3) == 0;

  for (int k = 0; k < n - 1; ++k) {
    const int k1 = k + 1;
    // ---- A: y2_i = sum_{j >= k+2} A(i,j) x_j for the warp's rows ----
    double xr[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int j = lane + 32 * c;
      xr[c] = (j >= k + 2) ? xs[j] : 0.0;
    }
    double p[NRW];
#pragma unroll
    for (int rr = 0; rr < NRW; ++rr) {
      double acc = 0.0;
      if (w + NW * rr > k) {                 // warp-uniform
#pragma unroll
        for (int c = 0; c < NC; ++c) acc = fma(a[rr][c], xr[c], acc);
      }
      p[rr] = acc;
    }
    // transposed butterfly: 8 sums over 32 lanes; lane l ends with the total of row (l >> 2) & 7
    double y2;
    {
      const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
      double p4[4], p2[2];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double send = h16 ? p[q] : p[q + 4], keep = h16 ? p[q + 4] : p[q];
        p4[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const double send = h8 ? p4[q] : p4[q + 2], keep = h8 ? p4[q + 2] : p4[q];
        p2[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      {
        const double send = h4 ? p2[0] : p2[1], keep = h4 ? p2[1] : p2[0];
        y2 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      y2 += __shfl_xor_sync(0xffffffffu, y2, 2);
      y2 += __shfl_xor_sync(0xffffffffu, y2, 1);
    }
    const bool rowact = owner && myrow > k && myrow < n;
    double y1 = 0.0, xi = 0.0, q0 = 0.0, q1 = 0.0, q2 = 0.0;
    if (rowact) {
      y1 = ys[myrow];
      if (myrow == k1) { y0s[0] = y1; y0s[1] = y2; }
      else {
        xi = xs[myrow];
        q0 = xi * xi; q1 = y1 * xi; q2 = y2 * xi;
      }
    }
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      q0 += __shfl_xor_sync(0xffffffffu, q0, o);
      q1 += __shfl_xor_sync(0xffffffffu, q1, o);
      q2 += __shfl_xor_sync(0xffffffffu, q2, o);
    }
    if (lane == 0) { red[w] = q0; red[32 + w] = q1; red[64 + w] = q2; }
    __syncthreads();   // 1
    // ---- B: scalars (warp 0), broadcast ----
    if (w == 0) {
      double s0 = red[lane], s1 = red[32 + lane], s2 = red[64 + lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if (lane == 0) {
        const double ar = xs[k1];
        const bool have = s0 != 0.0;
        double beta = ar, tau = 0.0, sc = 0.0, a2 = 0.0;
        if (have) {
          const double sq = fma(ar, ar, s0);
          const double inv = rsqrt(sq), nrm = sq * inv;
          const double sg = ar >= 0.0 ? 1.0 : -1.0;
          beta = -sg * nrm;
          tau = fma(fabs(ar), inv, 1.0);
          // 1 / (alpha - beta) = sg / (|ar| + nrm) = sg inv / tau with tau in [1, 2]: reciprocal by Newton from a float seed
          // (two steps: 2^-23 -> 2^-46 -> below 2^-53), shorter than a second rsqrt on the serial path of the step
          double rt = (double)__frcp_rn((float)tau);
          rt = rt * fma(-tau, rt, 2.0);
          rt = rt * fma(-tau, rt, 2.0);
          sc = sg * inv * rt;
          // p^T v = tau [y1_0 + s y2_0 + s a + s^2 b],  a2 = -tau/2 (p^T v)
          const double pv = tau * (y0s[0] + sc * (y0s[1] + s1 + sc * s2));
          a2 = -0.5 * tau * pv;
        }
        scal[0] = sc; scal[1] = tau; scal[2] = a2; scal[3] = have ? 1.0 : 0.0;
        d_out[k] = xs[k]; e_out[k] = beta; tau_out[k] = tau;
      }
    }
    __syncthreads();   // 2
    const bool have = scal[3] != 0.0;
    if (rowact) {
      double vi = (myrow == k1) ? 1.0 : 0.0;
      if (have) {
        const double sc = scal[0], tau = scal[1], a2 = scal[2];
        if (myrow != k1) vi = sc * xi;
        const double pp = tau * fma(sc, y2, y1);
        vw[myrow] = make_double2(vi, fma(a2, vi, pp));
      }
      V_out[refl_off(k, n) + (myrow - k1)] = vi;
    }
    __syncthreads();   // 3
    // ---- C: A22 -= v w^T + w v^T (whole rows, registers), hand rows k+1, k+2 over ----
    if (have) {
      double2 cj[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) cj[c] = vw[lane + 32 * c];     // zero beyond n; columns <= k are dead
#pragma unroll
      for (int rr = 0; rr < NRW; ++rr) {
        const int i = w + NW * rr;
        if (i > k && i < n) {                  // warp-uniform
          const double2 me = vw[i];
#pragma unroll
          for (int c = 0; c < NC; ++c) a[rr][c] = fma(-me.x, cj[c].y, fma(-me.y, cj[c].x, a[rr][c]));
        }
      }
    }
#pragma unroll
    for (int rr = 0; rr < NRW; ++rr) {
      const int i = w + NW * rr;
      if (i == k1 || i == k1 + 1) {            // warp-uniform
        double* dst = i == k1 ? xs : ys;
#pragma unroll
        for (int c = 0; c < NC; ++c) dst[lane + 32 * c] = a[rr][c];
      }
    }
    __syncthreads();   // 4
  }
  if (tid == 0) { d_out[n - 1] = xs[n - 1]; e_out[n - 1] = 0.0; tau_out[n - 1] = 0.0; }
}

// ---------------------------------------------------------------------------------------------------------------------
// K2a: eigenvalues n-1-j (ascending index), j < nev, of tridiag(d, e) by multisection on T / |T|: M lanes (aligned
// segment of a warp) per eigenvalue, grid.x CTAs share the eigenvalues of one matrix.  lam_s: scaled eigenvalues
// (descending in j), tn_out: |T| bound (Gershgorin); tn = 0 marks a zero / non-finite matrix.
// ---------------------------------------------------------------------------------------------------------------------
// number of eigenvalues below x (sign changes of the scaled Sturm sequence).  An exact zero p_i counts correctly by its
// sign bit alone (p_{i+1} = -e^2 p_{i-1}: one change across the triple) unless the next e is zero too, so the zero test is
// kept off the dependency chain: it only raises `zero`, and the rare evaluation that saw one is redone by the guarded loop.
__device__ __forceinline__ int sturm_count_safe(const double* __restrict__ ds, const double* __restrict__ e2, int n, double x) {
  int cnt;
  double pm1 = 1.0, p = ds[0] - x;
  if (p == 0.0) p = -1e-300;
  cnt = p < 0.0;
  for (int i = 1; i < n; ++i) {
    double pn = (ds[i] - x) * p - e2[i - 1] * pm1;
    if (pn == 0.0) pn = p < 0.0 ? 1e-300 : -1e-300;
    cnt += (unsigned)(__double2hiint(pn) ^ __double2hiint(p)) >> 31;
    pm1 = p; p = pn;
    const double a = fmax(fabs(p), fabs(pm1));
    if (a < 1e-100) { p *= 1e100; pm1 *= 1e100; }
    else if (a > 1e100) { p *= 1e-100; pm1 *= 1e-100; }
  }
  return cnt;
}

__device__ __forceinline__ int sturm_count(const double* __restrict__ ds, const double* __restrict__ e2, int n, double x) {
  int cnt;
  double pm1 = 1.0, p = ds[0] - x;
  bool zero = (p == 0.0);
  cnt = p < 0.0;
  int i = 1;
  for (; i + 7 < n; i += 8) {
    double dv[8], ev[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { dv[u] = ds[i + u] - x; ev[u] = e2[i + u - 1]; }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const double pn = fma(dv[u], p, -(ev[u] * pm1));
      zero |= (pn == 0.0);
      cnt += (unsigned)(__double2hiint(pn) ^ __double2hiint(p)) >> 31;
      pm1 = p; p = pn;
    }
    const double a = fmax(fabs(p), fabs(pm1));
    if (a < 1e-100) { p *= 1e100; pm1 *= 1e100; }
    else if (a > 1e100) { p *= 1e-100; pm1 *= 1e-100; }
  }
  for (; i < n; ++i) {
    const double pn = fma(ds[i] - x, p, -(e2[i - 1] * pm1));
    zero |= (pn == 0.0);
    cnt += (unsigned)(__double2hiint(pn) ^ __double2hiint(p)) >> 31;
    pm1 = p; p = pn;
  }
  if (zero) return sturm_count_safe(ds, e2, n, x);
  return cnt;
}

__global__ void __launch_bounds__(1024) heig_bisect_kernel(const double* __restrict__ d_in, const double* __restrict__ e_in, int n,
                                                            int nev, int epc, int M, int rounds, double* __restrict__ lam_s,
                                                            double* __restrict__ tn_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* ds = reinterpret_cast<double*>(smem_raw);   // [n]
  double* e2 = ds + n;                                // [n]
  __shared__ double red[64];
  __shared__ double s_tn, s_lo, s_hi;
  __shared__ int s_bad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int64_t mb = blockIdx.y;
  d_in += mb * n; e_in += mb * n; lam_s += mb * nev;

  double glo = 1e300, ghi = -1e300;
  int bad = 0;
  for (int i = tid; i < n; i += blockDim.x) {
    const double di = d_in[i];
    const double el = i > 0 ? fabs(e_in[i - 1]) : 0.0, er = i < n - 1 ? fabs(e_in[i]) : 0.0;
    if (!isfinite(di) || !isfinite(el) || !isfinite(er)) bad = 1;
    glo = fmin(glo, di - el - er);
    ghi = fmax(ghi, di + el + er);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    glo = fmin(glo, __shfl_xor_sync(0xffffffffu, glo, o));
    ghi = fmax(ghi, __shfl_xor_sync(0xffffffffu, ghi, o));
    bad |= __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if (tid == 0) s_bad = 0;
  if (lane == 0) { red[warp] = glo; red[32 + warp] = ghi; }
  __syncthreads();
  if (bad && lane == 0) atomicOr(&s_bad, 1);
  __syncthreads();
  if (tid == 0) {
    double a = 1e300, b = -1e300;
    for (int w = 0; w < nwarps; ++w) { a = fmin(a, red[w]); b = fmax(b, red[32 + w]); }
    double tn = fmax(fabs(a), fabs(b));
    if (s_bad || !isfinite(tn)) tn = 0.0;
    s_tn = tn;
    const double isc = tn > 0.0 ? 1.0 / tn : 0.0;
    s_lo = a * isc - 4e-16 * n - 1e-300;
    s_hi = b * isc + 4e-16 * n + 1e-300;
    if (blockIdx.x == 0) tn_out[mb] = tn;
  }
  __syncthreads();
  const double tn = s_tn;
  const int j_lo = blockIdx.x * epc;
  if (!(tn > 0.0)) {
    for (int j = j_lo + tid; j < min(nev, j_lo + epc); j += blockDim.x) lam_s[j] = 0.0;
    return;
  }
  const double isc = 1.0 / tn;
  for (int i = tid; i < n; i += blockDim.x) {
    ds[i] = d_in[i] * isc;
    const double ev = i < n - 1 ? e_in[i] * isc : 0.0;
    e2[i] = ev * ev;
  }
  __syncthreads();
  const int groups = blockDim.x / M;
  const int gi = tid / M, m = tid % M;
  const double invM1 = 1.0 / (double)(M + 1);
  const unsigned gmask = (M == 32) ? 0xffffffffu : (((1u << M) - 1u) << ((lane / M) * M));
  for (int j0 = 0; j0 < epc; j0 += groups) {
    const int jj = j_lo + j0 + gi;
    const bool act = (j0 + gi < epc) && jj < nev;
    const int idx = n - 1 - jj;
    double lo = s_lo, hi = s_hi;
    for (int r = 0; r < rounds; ++r) {
      const double w = hi - lo;
      const double pt = lo + w * ((double)(m + 1) * invM1);
      const int cnt = act ? sturm_count(ds, e2, n, pt) : 0;
      const unsigned b = __ballot_sync(0xffffffffu, act && cnt <= idx) & gmask;
      const int c = __popc(b);
      const double nlo = c > 0 ? lo + w * ((double)c * invM1) : lo;
      const double nhi = c < M ? lo + w * ((double)(c + 1) * invM1) : hi;
      lo = nlo; hi = nhi;
    }
    if (act && m == 0) lam_s[jj] = 0.5 * (lo + hi);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// K2b: eigenvectors of tridiag(d, e) for the eigenvalues of K2a by the twisted factorisation (Parlett-Dhillon) written on
// the division-free Sturm polynomials: forward phi_{i+1} = (d_i - l) phi_i - e_{i-1}^2 phi_{i-1}, backward psi_i likewise;
// the pivots s_i = phi_{i+1}/phi_i and p_i = psi_i/psi_{i+1} (all divisions off the dependency chains), gamma_i = s_i + p_i
// - (d_i - l), twist at r = argmin |gamma_i|, z_i = -(e_i/s_i) z_{i+1} below r, z_{i+1} = -(e_i/p_{i+1}) z_i above.
// Modified Gram-Schmidt inside clusters.  One CTA per matrix, NSEG threads per vector.
// flags: 1 cluster larger than 8, 2 parallel vectors inside a cluster, 4 twisted residual too large / non-finite, 16 bad input
// ---------------------------------------------------------------------------------------------------------------------
constexpr int NSEG = 8;

__global__ void __launch_bounds__(1024) heig_vec_kernel(const double* __restrict__ d_in, const double* __restrict__ e_in,
                                                         const double* __restrict__ lam_in, const double* __restrict__ tn_in, int n,
                                                         int nev, int C, double ctol, double* __restrict__ lam_out,
                                                         double* __restrict__ Z, int* __restrict__ flags) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int pitch = C | 1;                            // odd row pitch: conflict-free along i as well as along the vector index
  double* ds = reinterpret_cast<double*>(smem_raw);   // [n]
  double* es = ds + n;                                // [n]
  double* lam_s = es + n;                             // [nev]
  double* red = lam_s + nev;                          // [64]
  double* gm = red + 64;                              // [NSEG][C] segment minima of |gamma|
  int* gr = reinterpret_cast<int*>(gm + NSEG * C);    // [NSEG][C] their positions
  int* rtw = gr + NSEG * C;                           // [C] twist index
  double* zn = reinterpret_cast<double*>(rtw + C + (C & 1));   // [2][C] partial squared norms
  double* PH = zn + 2 * C;                            // [n+1][pitch]  phi, then left factors / z
  double* PS = PH + (size_t)(n + 1) * pitch;          // [n+2][pitch]  psi, then right factors / z
  __shared__ double s_dot;
  __shared__ int s_flag;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int64_t mb = blockIdx.x;
  d_in += mb * n; e_in += mb * n; lam_in += mb * nev; lam_out += mb * nev; Z += mb * (int64_t)n * nev;
  const double tn = tn_in[mb];
  if (tid == 0) s_flag = 0;
  if (!(tn > 0.0)) {
    // zero (or non-finite) matrix: lambda = 0, unit vectors; the caller's acceptance test sends it to the Jacobi path
    for (int idx = tid; idx < n * nev; idx += blockDim.x) Z[idx] = (idx % n) == (idx / n) ? 1.0 : 0.0;
    for (int j = tid; j < nev; j += blockDim.x) lam_out[j] = 0.0;
    if (tid == 0) atomicOr(&flags[mb], 16);
    return;
  }
  const double isc = 1.0 / tn;
  for (int i = tid; i < n; i += blockDim.x) {
    ds[i] = d_in[i] * isc;
    es[i] = i < n - 1 ? e_in[i] * isc : 0.0;
  }
  for (int j = tid; j < nev; j += blockDim.x) { const double l = lam_in[j]; lam_s[j] = l; lam_out[j] = l * tn; }
  __syncthreads();
  const int seglen = (n + NSEG - 1) / NSEG;

  for (int c0 = 0; c0 < nev; c0 += C) {
    const int nc = min(C, nev - c0);
    // (1) Sturm polynomial chains: thread (vector, direction)
    if (tid < 2 * nc) {
      const int c = tid >> 1, dir = tid & 1;
      const double lam = lam_s[c0 + c];
      if (dir == 0) {
        double pm1 = 1.0, p = ds[0] - lam;
        PH[c] = 1.0; PH[pitch + c] = p;
        for (int i = 1; i < n; ++i) {
          const double pn = (ds[i] - lam) * p - (es[i - 1] * es[i - 1]) * pm1;
          pm1 = p; p = pn;
          PH[(size_t)(i + 1) * pitch + c] = pn;
        }
      } else {
        double pp1 = 1.0, p = ds[n - 1] - lam;
        PS[(size_t)(n + 1) * pitch + c] = 1.0;      // psi_{n+1} (only as a safe denominator slot)
        PS[(size_t)n * pitch + c] = 1.0; PS[(size_t)(n - 1) * pitch + c] = p;
        for (int i = n - 2; i >= 0; --i) {
          const double pn = (ds[i] - lam) * p - (es[i] * es[i]) * pp1;
          pp1 = p; p = pn;
          PS[(size_t)i * pitch + c] = pn;
        }
      }
    }
    __syncthreads();
    // (2) pivots, gamma, chain factors: thread (vector c, segment sg) owns rows [i0, i1).  In place: a thread overwrites row i
    // only after it has read rows i and i+1; the one row it reads from its neighbour's segment (i1) is taken before the
    // barrier.  Two reciprocals per row (1/phi_{i+1}, 1/psi_{i+1}, each reused by the next row) instead of four divisions.
    {
      const int c = tid % C, sg = tid / C;
      const bool act = sg < NSEG && c < nc;
      const int i0 = sg * seglen, i1 = min(n, i0 + seglen);
      double ph_end = 1.0, ps_end = 1.0;
      if (act && i0 < i1) { ph_end = PH[(size_t)i1 * pitch + c]; ps_end = PS[(size_t)i1 * pitch + c]; }
      __syncthreads();
      double gbest = 1e300;
      int rbest = 0;
      if (act && i0 < i1) {
        const double lam = lam_s[c0 + c];
        double ph = PH[(size_t)i0 * pitch + c], ps = PS[(size_t)i0 * pitch + c];
        double rph = 1.0 / ph, rps = 1.0 / ps;
        for (int i = i0; i < i1; ++i) {
          const bool last = i + 1 == i1;
          const double ph1 = last ? ph_end : PH[(size_t)(i + 1) * pitch + c];
          const double ps1 = last ? ps_end : PS[(size_t)(i + 1) * pitch + c];
          const double rph1 = 1.0 / ph1, rps1 = 1.0 / ps1;
          const double gam = fabs(ph1 * rph + ps * rps1 - (ds[i] - lam));      // s_i + p_i - (d_i - lambda)
          if (gam < gbest) { gbest = gam; rbest = i; }
          PH[(size_t)i * pitch + c] = -es[i] * (ph * rph1);                    // z_i = fl_i z_{i+1}:  -(e_i / s_i)
          PS[(size_t)i * pitch + c] = i > 0 ? -es[i - 1] * (ps1 * rps) : 0.0;  // z_i = fr_i z_{i-1}:  -(e_{i-1} / p_i)
          ph = ph1; ps = ps1; rph = rph1; rps = rps1;
        }
      }
      if (act) {
        gm[sg * C + c] = gbest;
        gr[sg * C + c] = rbest;
      }
    }
    __syncthreads();
    if (tid < nc) {
      double gb = gm[tid];
      int r = gr[tid];
      for (int s2 = 1; s2 < NSEG; ++s2)
        if (s2 * seglen < n && gm[s2 * C + tid] < gb) { gb = gm[s2 * C + tid]; r = gr[s2 * C + tid]; }
      rtw[tid] = r;
      gm[tid] = gb;
    }
    __syncthreads();
    // (3) z chains: thread (vector, side); z overwrites the factors (PH below r, PS above r, z_r = 1 in PH)
    if (tid < 2 * nc) {
      const int c = tid >> 1, dir = tid & 1;
      const int r = rtw[c];
      double z = 1.0, nrm2 = 0.0;
      if (dir == 0) {
        nrm2 = 1.0;
        for (int i = r - 1; i >= 0; --i) {
          z *= PH[(size_t)i * pitch + c];
          PH[(size_t)i * pitch + c] = z;
          nrm2 += z * z;
        }
        PH[(size_t)r * pitch + c] = 1.0;
      } else {
        for (int i = r + 1; i < n; ++i) {
          z *= PS[(size_t)i * pitch + c];
          PS[(size_t)i * pitch + c] = z;
          nrm2 += z * z;
        }
      }
      zn[dir * C + c] = nrm2;
    }
    __syncthreads();
    // (4) normalise and store (coalesced along i); one rsqrt per vector
    if (tid < nc) {
      const double nrm2 = zn[tid] + zn[C + tid];
      const double inv = rsqrt(nrm2);
      zn[tid] = inv;
      if (!(gm[tid] * inv <= 1e-9) || !isfinite(nrm2)) atomicOr(&s_flag, 4);
    }
    __syncthreads();
    for (int idx = tid; idx < nc * n; idx += blockDim.x) {
      const int c = idx / n, i = idx - c * n;
      Z[(int64_t)(c0 + c) * n + i] = (i <= rtw[c] ? PH[(size_t)i * pitch + c] : PS[(size_t)i * pitch + c]) * zn[c];
    }
    __syncthreads();
  }

  // modified Gram-Schmidt inside clusters (eigenvalues closer than ctol |T|); block-cooperative, rare
  int cstart = 0;
  for (int j = 1; j <= nev; ++j) {
    // cluster: the gap does not separate the two vectors to the accuracy the reconstruction needs.  The twisted vectors of
    // eigenvalues gap apart are orthogonal to eps/gap (scaled units) and enter the reconstruction with weight sigma =
    // sqrt(lambda), so the gap that matters shrinks with sqrt(lambda): small but well separated eigenvalues of a full
    // spectrum are NOT a cluster.
    const bool brk = (j == nev) || (lam_s[j - 1] - lam_s[j] >= ctol * sqrt(fmax(lam_s[j - 1], 0.0)) + 1e-14);
    if (!brk) continue;
    const int csize = j - cstart;
    if (csize > 8) {
      if (tid == 0) atomicOr(&s_flag, 1);
    } else if (csize > 1) {
      for (int a = cstart + 1; a < j; ++a) {
        double* Za = Z + (int64_t)a * n;
        for (int b = cstart; b <= a; ++b) {
          // b < a: remove the component along z_b;  b == a: normalise
          const double* Zb = Z + (int64_t)b * n;
          double v = 0.0;
          for (int i2 = tid; i2 < n; i2 += blockDim.x) v += Zb[i2] * Za[i2];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0) red[warp] = v;
          __syncthreads();
          if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < nwarps; ++w) t += red[w];
            s_dot = t;
          }
          __syncthreads();
          const double dot = s_dot;
          if (b < a) {
            for (int i2 = tid; i2 < n; i2 += blockDim.x) Za[i2] -= dot * Zb[i2];
          } else {
            if (!(dot > 1e-6)) { if (tid == 0) atomicOr(&s_flag, 2); }
            const double inv = dot > 0.0 ? rsqrt(dot) : 0.0;
            for (int i2 = tid; i2 < n; i2 += blockDim.x) Za[i2] *= inv;
          }
          __syncthreads();
        }
      }
    }
    cstart = j;
  }
  __syncthreads();
  if (tid == 0 && s_flag) atomicOr(&flags[mb], s_flag);
}

// ---------------------------------------------------------------------------------------------------------------------
// K3: U = H_0 H_1 ... H_{n-2} Z.  8 lanes per column, rows l + 8 m in registers, reflectors in shared memory.
// ---------------------------------------------------------------------------------------------------------------------
template <class T, int NR, int CPG>
__global__ void __launch_bounds__(256) heig_back_kernel(const T* __restrict__ Vp, int64_t bV, const T* __restrict__ tau_in,
                                                        const double* __restrict__ Z, int n, int nev, T* __restrict__ U,
                                                        int64_t ldu, int64_t bU) {
  // CPG columns per 8-lane group: every reflector element fetched from shared memory serves CPG dot products / updates
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Vs = reinterpret_cast<T*>(smem_raw);
  const int nv = (n * (n - 1)) >> 1;
  T* taus = Vs + nv + (nv & 1);
  const int tid = threadIdx.x;
  const int64_t mb = blockIdx.y;
  Vp += mb * bV; tau_in += mb * n; Z += mb * (int64_t)n * nev; U += mb * bU;
  // stage the reflectors (one contiguous run of up to 130 KB): a single bulk copy of the copy engine (TMA, 1-D) completing
  // on an mbarrier when the run is 16-byte aligned, a cooperative loop otherwise
  __shared__ __align__(8) uint64_t bar;
  const size_t vbytes = sizeof(T) * (size_t)nv;
  const bool bulk = nv > 0 && (vbytes & 15) == 0 && ((reinterpret_cast<uintptr_t>(Vp) & 15) == 0);
  if (bulk) {
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (tid == 0) {
      mbar_arrive_expect_tx(&bar, (uint32_t)vbytes);
      bulk_copy_g2s(Vs, Vp, (uint32_t)vbytes, &bar);
    }
  } else {
    for (int idx = tid; idx < nv; idx += blockDim.x) Vs[idx] = Vp[idx];
  }
  for (int idx = tid; idx < n; idx += blockDim.x) taus[idx] = tau_in[idx];
  if (bulk) mbar_wait(&bar, 0);
  __syncthreads();
  const int cpb = (blockDim.x >> 3) * CPG;      // columns per CTA pass
  const int l = tid & 7;
  for (int c0 = blockIdx.x * cpb + (tid >> 3) * CPG; c0 < ((nev + cpb - 1) / cpb) * cpb; c0 += gridDim.x * cpb) {
    T z[CPG][NR];
#pragma unroll
    for (int c = 0; c < CPG; ++c)
#pragma unroll
      for (int m = 0; m < NR; ++m) {
        const int i = l + 8 * m;
        z[c][m] = (c0 + c < nev && i < n) ? t_from<T>(Z[(int64_t)(c0 + c) * n + i], 0.0) : t_zero<T>();
      }
    for (int k = n - 2; k >= 0; --k) {
      const T tk = taus[k];
      if (t_abs2(tk) == 0.0) continue;
      const T* vk = Vs + refl_off(k, n) - (k + 1);   // vk[i], i >= k+1
      // rows <= k are untouched by reflector k: whole register rows (8 m + 7 <= k, uniform over the lanes) are skipped
      T v[NR];
#pragma unroll
      for (int m = 0; m < NR; ++m) {
        const int i = l + 8 * m;
        v[m] = (8 * m + 7 > k && i > k && i < n) ? vk[i] : t_zero<T>();
      }
      T acc[CPG];
#pragma unroll
      for (int c = 0; c < CPG; ++c) {
        T a0 = t_zero<T>(), a1 = t_zero<T>();
#pragma unroll
        for (int m = 0; m < NR; ++m) {
          if (8 * m + 7 > k) {
            if (m & 1) t_fma(a1, t_conj(v[m]), z[c][m]);
            else t_fma(a0, t_conj(v[m]), z[c][m]);
          }
        }
        acc[c] = t_add(a0, a1);
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1)
#pragma unroll
        for (int c = 0; c < CPG; ++c) acc[c] = t_add(acc[c], shfl_xor_t(acc[c], o));
#pragma unroll
      for (int c = 0; c < CPG; ++c) {
        const T w = t_mul(tk, acc[c]);
#pragma unroll
        for (int m = 0; m < NR; ++m)
          if (8 * m + 7 > k) z[c][m] = t_sub(z[c][m], t_mul(w, v[m]));
      }
    }
#pragma unroll
    for (int c = 0; c < CPG; ++c)
      if (c0 + c < nev) {
#pragma unroll
        for (int m = 0; m < NR; ++m) {
          const int i = l + 8 * m;
          if (i < n) U[(int64_t)(c0 + c) * ldu + i] = z[c][m];
        }
      }
  }
}

// out1[i,j] = rs[i] * U[i,j] * sigma_j^{1/2},  out2[i,j] = U[i,j] * sigma_j^{-1/2},  sigma_j = sqrt(max(lambda_j, 0));
// rs[i] = w_in[i]^{-1/2} when w_in is given.  Also writes sigma (device), and raises flag 8 when the kept spectrum is not
// safely inside the accuracy of the Gram matrix (lambda_r < 1e-10 lambda_1) or is not positive.
template <class T>
__global__ void heig_finalize_kernel(const T* __restrict__ U, int n, int nev, int64_t bU, const double* __restrict__ lam,
                                     const double* __restrict__ w_in, int64_t bw, T* __restrict__ out1, int64_t ld1, int64_t b1,
                                     T* __restrict__ out2, int64_t ld2, int64_t b2, double* __restrict__ sig_out, int64_t bsig,
                                     double* __restrict__ sig0, int* __restrict__ flags) {
  const int64_t mb = blockIdx.y;
  U += mb * bU; lam += mb * nev;
  const double l0 = lam[0], lr = lam[nev - 1];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (!(l0 > 0.0) || !(lr >= 1e-10 * l0) || !isfinite(l0)) atomicOr(&flags[mb], 8);
  }
  if (blockIdx.x == 0 && (sig_out || (sig0 && mb == 0)))
    for (int j = threadIdx.x; j < nev; j += blockDim.x) {
      const double sj = sqrt(fmax(lam[j], 0.0));
      if (sig_out) sig_out[mb * bsig + j] = sj;
      if (sig0 && mb == 0) sig0[j] = sj;
    }
  const int total = n * nev;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int i = idx % n, j = idx / n;
    const double lj = fmax(lam[j], 0.0);
    const double s = sqrt(lj);
    const bool ok = s > 1e-290 && lj > l0 * 1e-280;
    const double a = ok ? sqrt(s) : 0.0, b = ok ? 1.0 / sqrt(s) : 0.0;
    double rs = 1.0;
    if (w_in) {
      const double wi = w_in[mb * bw + i];
      rs = wi > 0.0 ? 1.0 / sqrt(wi) : 0.0;   // w^{-1/2}
    }
    const T u = U[(int64_t)j * n + i];
    if (out1) out1[mb * b1 + (int64_t)j * ld1 + i] = t_scale(u, a * rs);
    if (out2) out2[mb * b2 + (int64_t)j * ld2 + i] = t_scale(u, b);
  }
}

template <class T> size_t tridiag_smem(int n) {
  constexpr int R = HT<T>::R;
  constexpr int NC = is_cplx<T>::value ? 5 : 3;
  const int npad = ((n + R - 1) / R) * R;
  const size_t t = (size_t)n * (n + 1) / 2;
  return sizeof(T) * (t + (t & 1) + 3 * (size_t)npad + 2) + sizeof(double) * (NC * 32);
}

}  // namespace

template <class T> int heig_max_n() { return is_cplx<T>::value ? 128 : 176; }

// Top `nev` eigenpairs (descending) of `batch` Hermitian matrices G (n x n, upper triangle read, leading dimension ldg,
// batch stride bG; G = sum of nsplit partial matrices sG apart).  lam: nev per matrix (device), U: n x nev (ld n) per
// matrix (device), flags: one int per matrix, OR-ed (must be zeroed by the caller).  Returns false if the shape is not served.
template <class T>
bool heig_top(const T* G, int n, int64_t ldg, int64_t bG, int nsplit, int64_t sG, int nev, int batch, double* lam, T* U,
              int* flags) {
  if (n < 1 || n > heig_max_n<T>() || nev < 1 || nev > n || batch < 1) return false;
  constexpr int R = HT<T>::R, g = 32 / R;
  const int npad = ((n + R - 1) / R) * R;
  const int64_t bV = std::max<int64_t>(1, (int64_t)n * (n - 1) / 2);
  DevBuf d(sizeof(double) * (size_t)n * batch), e(sizeof(double) * (size_t)n * batch), tau(sizeof(T) * (size_t)n * batch);
  DevBuf V(sizeof(T) * (size_t)bV * batch), Z(sizeof(double) * (size_t)n * nev * batch);
  DevBuf lam_s(sizeof(double) * (size_t)nev * batch), tn(sizeof(double) * (size_t)batch);
  ProfScope prof_scope_(KF_JACOBI);
  // LAPACK counts: tridiagonalisation 4/3 n^3, back-transformation of nev vectors 2 n^2 nev (complex: x4); the O(n^2) stages are not counted
  ctx().flops_heig += (double)batch * (is_cplx<T>::value ? 4.0 : 1.0) * (4.0 / 3.0 * n * (double)n * n + 2.0 * n * (double)n * nev);
  bool done_tridiag = false;
  if (!is_cplx<T>::value && n <= 128 && n >= 2) {
    // Float64: whole matrix in registers (latency path of the single-train rounding chain; also the fastest batched form)
    const double* Gd = reinterpret_cast<const double*>(G);
    double* taud = reinterpret_cast<double*>(tau.p);
    double* Vd = reinterpret_cast<double*>(V.p);
    if (n <= 32) heig_tridiag_reg_kernel<4, 1><<<batch, 128, 0, ctx().stream>>>(Gd, n, ldg, bG, nsplit, sG, d.as<double>(), e.as<double>(), taud, Vd, bV);
    else if (n <= 64) heig_tridiag_reg_kernel<8, 2><<<batch, 256, 0, ctx().stream>>>(Gd, n, ldg, bG, nsplit, sG, d.as<double>(), e.as<double>(), taud, Vd, bV);
    else heig_tridiag_reg_kernel<16, 4><<<batch, 512, 0, ctx().stream>>>(Gd, n, ldg, bG, nsplit, sG, d.as<double>(), e.as<double>(), taud, Vd, bV);
    TTN_CHECK_LAUNCH();
    ctx().launches++;
    done_tridiag = true;
  }
  if (!done_tridiag) {
    auto kern = heig_tridiag_kernel<T>;
    const size_t smem = tridiag_smem<T>(n);
    static int attr_dev = -1;
    if (attr_dev != ctx().device) {
      TTN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_dev = ctx().device;
    }
    ttn_assert(smem <= 200 * 1024, 7, "heig: shared memory");
    kern<<<batch, npad * g, smem, ctx().stream>>>(G, n, ldg, bG, nsplit, sG, d.as<double>(), e.as<double>(), tau.as<T>(),
                                                  V.as<T>(), bV);
    TTN_CHECK_LAUNCH();
    ctx().launches++;
  }
  {
    // eigenvalues: a single matrix spreads its eigenvalues over many CTAs with 32 lanes each (latency), a batch runs one CTA
    // per matrix with few lanes per eigenvalue (multisection does M evaluations for log2(M+1) bits)
    int M, epc;
    if (batch * 16 <= ctx().sm_count) { M = 32; epc = 4; }
    else if (batch * 4 <= ctx().sm_count) { M = 8; epc = 16; }
    else { M = nev * 4 <= 256 ? 4 : 2; epc = nev; }
    epc = std::min(epc, nev);
    const int gx = (nev + epc - 1) / epc;
    const int nt = std::min(1024, std::max(32, ((epc * M + 31) / 32) * 32));
    const int rounds = (int)std::ceil(56.0 / std::log2((double)M + 1.0));
    const size_t smem = sizeof(double) * (size_t)2 * n;
    heig_bisect_kernel<<<dim3(gx, batch), nt, smem, ctx().stream>>>(d.as<double>(), e.as<double>(), n, nev, epc, M, rounds,
                                                                   lam_s.as<double>(), tn.as<double>());
    TTN_CHECK_LAUNCH();
    ctx().launches++;
  }
  {
    int C = std::min(nev, 64);
    auto vec_smem = [&](int c) {
      const int pitch = c | 1;
      return sizeof(double) * ((size_t)2 * n + nev + 64 + (size_t)NSEG * c + 2 * (size_t)c + (size_t)(2 * n + 3) * pitch) +
             sizeof(int) * ((size_t)NSEG * c + c + 2);
    };
    while (vec_smem(C) > 190 * 1024 && C > 8) C = (C + 1) / 2;
    const int nt = std::min(1024, std::max(64, ((C * NSEG + 31) / 32) * 32));
    auto kern = heig_vec_kernel;
    static int attr_dev = -1;
    if (attr_dev != ctx().device) {
      TTN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_dev = ctx().device;
    }
    kern<<<batch, nt, vec_smem(C), ctx().stream>>>(d.as<double>(), e.as<double>(), lam_s.as<double>(), tn.as<double>(), n, nev, C,
                                                   1e-3, lam, Z.as<double>(), flags);
    TTN_CHECK_LAUNCH();
    ctx().launches++;
  }
  {
    // few matrices: one column per lane group and many CTAs (latency); batches: two columns per group, one CTA per matrix
    const bool few = batch * 8 <= ctx().sm_count;
    const int cpg = 1;   // two columns per group (CPG = 2) measured no faster: the kernel is FP64-issue bound, not load bound
    int groups = few ? 8 : 32;                     // 8-lane groups per CTA
    groups = std::min(groups, ((((nev + cpg - 1) / cpg) + 3) / 4) * 4);
    const int nt = groups * 8;
    const int cpb = groups * cpg;
    const int nv = n * (n - 1) / 2;
    const size_t smem = sizeof(T) * ((size_t)nv + (nv & 1) + n);
    dim3 grid(few ? (nev + cpb - 1) / cpb : 1, batch);
    const int nr = (n + 7) / 8;
#define TTN_BACK2(NR, CPG)                                                                                        \
    {                                                                                                             \
      auto kern = heig_back_kernel<T, NR, CPG>;                                                                   \
      static int attr_dev = -1;                                                                                   \
      if (attr_dev != ctx().device) {                                                                             \
        TTN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));            \
        attr_dev = ctx().device;                                                                                  \
      }                                                                                                           \
      kern<<<grid, nt, smem, ctx().stream>>>(V.as<T>(), bV, tau.as<T>(), Z.as<double>(), n, nev, U, n, (int64_t)n * nev); \
    }
#define TTN_BACK(NR) TTN_BACK2(NR, 1)
    if (nr <= 4) TTN_BACK(4)
    else if (nr <= 8) TTN_BACK(8)
    else if (nr <= 16) TTN_BACK(16)
    else TTN_BACK(22)
#undef TTN_BACK2
#undef TTN_BACK
    TTN_CHECK_LAUNCH();
    ctx().launches++;
  }
  return true;
}

template <class T>
void heig_finalize(const T* U, int n, int nev, int batch, const double* lam, const double* w_in, T* out1, int64_t ld1, int64_t b1,
                   T* out2, int64_t ld2, int64_t b2, double* sig_out, int64_t bsig, double* sig0, int* flags) {
  dim3 grid(std::max(1, std::min(8, (n * nev + 255) / 256)), batch);
  heig_finalize_kernel<T><<<grid, 256, 0, ctx().stream>>>(U, n, nev, (int64_t)n * nev, lam, w_in, n, out1, ld1, b1, out2, ld2, b2,
                                                         sig_out, bsig, sig0, flags);
  TTN_CHECK_LAUNCH();
  ctx().launches++;
}

template int heig_max_n<double>();
template int heig_max_n<zc>();
template bool heig_top<double>(const double*, int, int64_t, int64_t, int, int64_t, int, int, double*, double*, int*);
template bool heig_top<zc>(const zc*, int, int64_t, int64_t, int, int64_t, int, int, double*, zc*, int*);
template void heig_finalize<double>(const double*, int, int, int, const double*, const double*, double*, int64_t, int64_t, double*,
                                    int64_t, int64_t, double*, int64_t, double*, int*);
template void heig_finalize<zc>(const zc*, int, int, int, const double*, const double*, zc*, int64_t, int64_t, zc*, int64_t, int64_t,
                                double*, int64_t, double*, int*);

}  // namespace ttn
